#!/bin/bash
# GPU box (1 GPU): the in-process N-band test (also once with RFX_DEBUG_NO_A_CARRY=1 = the pre-fix single-buffered A target, for comparison),
# the rest of the chain tests, then step 6 (tiled-viewZ A/B + ncu of the new kernels).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
RFX_DEBUG_NO_A_CARRY=1 python -m pytest tests/test_gpu_chain.py -m gpu -q -k inprocess > gpurun_out/r02_s7_inprocess_without_fix.log 2>&1; tail -3 gpurun_out/r02_s7_inprocess_without_fix.log
python -m pytest tests/test_gpu_chain.py -m gpu -x -q > gpurun_out/r02_s7_chain_tests.log 2>&1; tail -3 gpurun_out/r02_s7_chain_tests.log
bash tools/r2_step6.sh
