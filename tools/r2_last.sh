#!/bin/bash
# GPU box (2 GPUs), last call of the round: the multi-process group path after the A-target change (IPC export of 7 allocations, the
# replicated fallback's extra exchange), the three 1-GPU tests touched after the final full run, and a short N = 2 bench (bit-exact self-check at 4K).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest "tests/test_gpu_multi.py::test_sharded_chain_equals_single_gpu_bit_exact[case0]" "tests/test_gpu_multi.py::test_sharded_chain_equals_single_gpu_bit_exact[case2]" \
  "tests/test_gpu_multi.py::test_sharded_chain_equals_single_gpu_bit_exact[case4]" tests/test_gpu_passes.py::test_abi_error_behaviour \
  tests/test_gpu_effects.py::test_cosmetic_effects_tail_kernel_and_taa tests/test_gpu_effects.py::test_effect_pass_and_taa_pass_host_classes -m gpu -q > gpurun_out/r02_last_tests.log 2>&1
tail -3 gpurun_out/r02_last_tests.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 3 --no-c5 > gpurun_out/r02_last_bench_n2.json 2> gpurun_out/r02_last_bench_n2.err
python - <<'PY'
import json
for ln in open("gpurun_out/r02_last_bench_n2.json"):
    if ln.startswith("{"):
        d = json.loads(ln)
        print("N=2", d["ms_per_step"], d["value"], "bit_exact", d["multi_gpu_bit_exact"], d["config"]["multi_gpu"]["bounds_during_timed_frames"])
PY
tail -2 gpurun_out/r02_last_bench_n2.err | cut -c1-300
