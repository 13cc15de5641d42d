#!/bin/bash
# First GPU calls of the next round: measure what round 1 left unmeasured (it ran out of GPU minutes), cheapest first.
#   tools/next_round_sweeps.sh k1      (1 GPU)  RFX_K1_SIMPLE_MARCH=1: parity tests + K1 time, fused and split
#   tools/next_round_sweeps.sh n8      (8 GPUs) exchange variants at N = 8: default | dual communicator | equal bands + in-place all-gather
# Every line printed is one configuration; nothing here changes a default.
cd "$(dirname "$0")/.."
line() { python -c "
import json,sys
d=json.loads(sys.stdin.read()); pk=d['roofline']['per_kernel']; m=d['config'].get('multi_gpu', {})
k1=pk['K1_ssgi_trace']; print('$1 | frame', d['ms_per_step'], 'ms', d['value'], 'Mpx/s | K1 per frame', round(k1['ms_per_launch']*k1['launches']/d['steps'],3), '| e2e', d['e2e']['value'], '| per-rank kernel ms', m.get('per_rank_kernel_ms_per_frame'))"; }
case "$1" in
k1)
  python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | line "default march            "
  rm -f realism_effects_b200/csrc/build/k_ssgi.o
  RFX_NVCC_EXTRA="-DRFX_K1_SIMPLE_MARCH=1" python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
  python -m pytest tests -m gpu -x -q 2>&1 | tail -2          # parity with the rewritten loop (expected: bit-identical K1)
  python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | line "simple march             "
  python bench.py --steps 8 --warmup 3 --no-cpu-baseline --split-parts 2>/dev/null | tail -1 | line "simple march, split K1   "
  rm -f realism_effects_b200/csrc/build/k_ssgi.o
  RFX_NVCC_EXTRA="-DRFX_K1_SIMPLE_MARCH=1 -DRFX_K1_FAST_TRIG=1" python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
  python -m pytest tests -m gpu -x -q 2>&1 | tail -2          # polynomial atan2 / acos in the fast env lookup: within the 1e-3 bars?
  python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | line "simple march + poly trig "
  rm -f realism_effects_b200/csrc/build/k_ssgi.o; python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
  ;;
n8)
  N=${2:-8}
  run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 10 --warmup 3 $1 2>/dev/null | tail -1 | line "$2"; }
  run ""                                             "adaptive bands, send/recv   "
  run "--dual-comm 1"                                "  + dnB on a 2nd communicator"
  run "--balance static --blocks-per-rank 1"         "equal bands, all-gather     "
  run "--balance static --blocks-per-rank 4"         "cyclic x4, all-gather       "
  run "--split-k1 0"                                 "adaptive, fused K1          "
  ;;
*) echo "usage: $0 k1 | n8 [N]"; exit 2 ;;
esac
