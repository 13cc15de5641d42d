#!/bin/bash
# Occupancy sweep (run on the GPU box): rebuild K2 / K3 with different __launch_bounds__ minimum-blocks targets and time the chain.
cd "$(dirname "$0")/.."
for cfg in "3 4" "4 4" "3 5" "3 6" "4 5" "2 4"; do
  set -- $cfg
  rm -f realism_effects_b200/csrc/build/k_temporal.o realism_effects_b200/csrc/build/k_denoise.o
  RFX_NVCC_EXTRA="-DRFX_K2_MIN_BLOCKS=$1 -DRFX_K3_MIN_BLOCKS=$2" python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
  python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); pk=d['roofline']['per_kernel']; print('K2 minblocks=$1 K3 minblocks=$2 | K2', round(pk['K2_temporal_reproject']['ms_per_launch'],3), 'K3p0', round(pk['K3_poisson_pass0']['ms_per_launch'],3), 'K3p1+', round(pk['K3_poisson_pass1plus']['ms_per_launch'],3), 'frame', d['ms_per_step'])"
done
rm -f realism_effects_b200/csrc/build/k_temporal.o realism_effects_b200/csrc/build/k_denoise.o
