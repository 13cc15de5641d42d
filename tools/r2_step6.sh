#!/bin/bash
# GPU box (1 GPU): K1 tiled-viewZ A/B (correctness first, then timing), and ncu captures of the round-2 streaming kernels.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
RFX_K1_VZ_TILED=1 python -m pytest tests/test_gpu_chain.py -m gpu -x -q -k "chain_parity or matches or CASES" > gpurun_out/r02_s6_tiled_tests.log 2>&1; tail -2 gpurun_out/r02_s6_tiled_tests.log
for t in 0 1; do
  RFX_K1_VZ_TILED=$t python bench.py --steps 60 --warmup 10 --no-cpu-baseline --no-configs 2>/dev/null | tail -1 > gpurun_out/r02_s6_bench_tiled$t.json
  python - <<PY
import json
d = json.load(open("gpurun_out/r02_s6_bench_tiled$t.json"))
pk = d["roofline"]["per_kernel"]
print("tiled=$t frame", d["ms_per_step"], {k: v["ms_per_launch"] for k, v in pk.items()})
PY
done
ncu --set full --clock-control none --import-source on -k 'regex:gbuffer_ingest|effects_kernel|taa_kernel' --launch-skip 6 --launch-count 3 -f -o gpurun_out/r02_new_kernels python tools/run_new_kernels.py > gpurun_out/r02_s6_ncu_new.log 2>&1
RFX_K1_VZ_TILED=1 ncu --set full --clock-control none --import-source on -k 'regex:ssgi_fast_kernel' --launch-skip 6 --launch-count 1 -f -o gpurun_out/r02_s6_k1_tiled python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/r02_s6_ncu_k1.log 2>&1
ls -la gpurun_out | tail -6
