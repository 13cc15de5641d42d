#!/bin/bash
# GPU box (1 GPU): the evidence set for profiles/ - bench line, reference-arm line, ncu launch list of the bench command, and one
# `--set full` capture of a steady-state frame (9 launches).  usage: tools/profile_round.sh r01_v4
cd "$(dirname "$0")/.."
tag=${1:-rXX}
K='regex:ssgi_kernel|temporal_kernel|poisson|gi_compose|viewz|gbuffer_decode'
python bench.py --steps 20 --warmup 5 2>gpurun_out/${tag}_bench.err | tail -1 > gpurun_out/${tag}_bench.json
python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > gpurun_out/${tag}_reference_arm.json
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 90 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k "$K" --launch-skip 27 --launch-count 9 -f -o gpurun_out/${tag}_frame \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_full.log 2>&1
ls -la gpurun_out/ | tail -8
