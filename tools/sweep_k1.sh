#!/bin/bash
# K1 tuning sweep (run on the GPU box): rebuild k_ssgi.cu with different march batch sizes / occupancy targets and time the chain
cd "$(dirname "$0")/.."
for cfg in "3 2" "4 1" "5 2" "4 3" "6 2"; do
  set -- $cfg
  rm -f realism_effects_b200/csrc/build/k_ssgi.o
  RFX_NVCC_EXTRA="-DRFX_K1_MIN_BLOCKS=$1 -DRFX_MARCH_BATCH=$2" python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
  regs=$(grep -A1 "ssgi_kernelILi0ELb1ELb1ELb1" realism_effects_b200/csrc/build/k_ssgi.cu.log | grep -o "Used [0-9]* registers" | head -1)
  spill=$(grep -B1 -A2 "ssgi_kernelILi0ELb1ELb1ELb1" realism_effects_b200/csrc/build/k_ssgi.cu.log | grep -o "[0-9]* bytes spill stores" | head -1)
  python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('minblocks=$1 batch=$2 | $regs | $spill | K1', round(d['roofline']['per_kernel']['K1_ssgi_trace']['ms_per_launch'],3), 'frame', d['ms_per_step'])"
done
rm -f realism_effects_b200/csrc/build/k_ssgi.o
