#!/bin/bash
# GPU box: cost of the split K1 on one GPU, and occupancy of the march-only kernel.
cd "$(dirname "$0")/.."
run() { python bench.py --steps 8 --warmup 3 --no-cpu-baseline $1 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); pk=d['roofline']['per_kernel']; print('$2 | K1 total per frame', round(pk['K1_ssgi_trace']['ms_per_launch']*pk['K1_ssgi_trace']['launches']/d['steps'],3), 'frame', d['ms_per_step'])"; }
run "" "fused"
for mb in ${MBS:-4 5 6}; do
  rm -f realism_effects_b200/csrc/build/k_ssgi.o
  RFX_NVCC_EXTRA="-DRFX_K1_MARCH_MIN_BLOCKS=$mb" python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
  run "--split-parts" "split, march minblocks=$mb"
done
rm -f realism_effects_b200/csrc/build/k_ssgi.o
