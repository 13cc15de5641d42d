#!/bin/bash
# GPU box, N GPUs (default 2): the row-sharded group (peer-mapped history) bit-exactness tests, then bench lines at N
# (peer reads, and the replicated-exchange fallback for comparison)
cd "$(dirname "$0")/.."
N=${1:-2}
nvidia-smi -L | head -8
if [ "$N" = "2" ]; then timeout 1200 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -25; fi
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 60 --warmup 10 $2 > gpurun_out/$1.json 2> gpurun_out/$1.err
  python - gpurun_out/$1.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
except Exception as e:
    print(sys.argv[1], 'NO JSON', e); sys.exit(0)
m=d['config']['multi_gpu']
print(sys.argv[1], 'N', d['n_gpus'], 'ms/frame', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], 'bit_exact', d['multi_gpu_bit_exact'])
print('   bounds', m['bounds_during_timed_frames'], 'per-rank ms', m['per_rank_kernel_ms_per_frame'], 'launches/frame', m['launches_per_frame_per_rank'])
print('   exchange', m['exchange'][:90]); print('   c5', d.get('c5_8k'))
PY
  tail -4 gpurun_out/$1.err; }
run r02_n${N}_bench ""
RFX_GROUP_BARRIER=flags run r02_n${N}_bench_flags "--no-c5"
RFX_GROUP_EXCHANGE=allgather run r02_n${N}_bench_allgather "--no-c5"
