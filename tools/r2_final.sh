#!/bin/bash
# GPU box (1 GPU): the final evidence of the round - whole GPU suite, bench line, reference arm.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -q) > gpurun_out/r02_final_gputest.log 2>&1; tail -4 gpurun_out/r02_final_gputest.log
python bench.py > gpurun_out/r02_final_bench_n1.json 2> gpurun_out/r02_final_bench_n1.err; tail -c 600 gpurun_out/r02_final_bench_n1.json; tail -2 gpurun_out/r02_final_bench_n1.err
python bench.py --impl reference > gpurun_out/r02_final_reference_arm.json 2> gpurun_out/r02_final_reference_arm.err; cut -c1-400 gpurun_out/r02_final_reference_arm.json
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_final_smoke.log 2>&1; tail -2 gpurun_out/r02_final_smoke.log | cut -c1-400
