#!/bin/bash
# GPU box (>= 2 GPUs): compare block assignments of the sharded chain at N = 2 (device-resident value and e2e).
cd "$(dirname "$0")/.."
N=${1:-2}
IFS=";" read -ra ARR <<< "${CFGS:-0 4;0 8;1 4;1 8}"
for cfg in "${ARR[@]}"; do
  IFS=" "
  set -- $cfg
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 \
     --mirror $1 --blocks-per-rank $2 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); pk=d['roofline']['per_kernel']
print('N=$N mirror=$1 bpr=$2 | ms', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], '| rank0 K1', round(pk['K1_ssgi_trace']['ms_per_launch'],3), 'K2', round(pk['K2_temporal_reproject']['ms_per_launch'],3), 'K3p1', round(pk['K3_poisson_pass1plus']['ms_per_launch'],3), '| per-rank kernel ms', d['config']['multi_gpu']['per_rank_kernel_ms_per_frame'], 'K1', d['config']['multi_gpu']['per_rank_K1_ms'])"
done
