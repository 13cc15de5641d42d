#!/bin/bash
# GPU box: first check of the fast chain (k_chain.cu): chain parity tests, parity at 4K, bench line
cd "$(dirname "$0")/.."
python -m pytest tests/test_gpu_chain.py -x -q 2>&1 | tail -15
python tools/parity_at_size.py C3 --frames 3 --fast-only --out gpurun_out/parity_step1.json 2>&1 | tail -22
python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/step1_bench.err | tail -1 > gpurun_out/step1_bench.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/step1_bench.json'))
print('ms/frame', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], d['e2e'].get('sync_call_ms'))
for k,v in d['roofline']['per_kernel'].items(): print(' ', k, round(v['ms_per_launch'],4), 'x', v['launches'])
PY
tail -5 gpurun_out/step1_bench.err
