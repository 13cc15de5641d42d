#!/bin/bash
# GPU box, 1 GPU: the whole -m gpu suite (incl. parity at the BASELINE sizes), then the default bench line and the reference arm
cd "$(dirname "$0")/.."
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25
python tools/parity_at_size.py C3 --frames 3 --fast-only --out gpurun_out/parity_step4.json 2>&1 | grep -E "worst|composed|ssgi" | head -12
( time python bench.py > gpurun_out/step4_bench.json 2> gpurun_out/step4_bench.err ) 2>&1 | tail -3
python - <<'PY'
import json
d=json.loads(open('gpurun_out/step4_bench.json').read().strip().splitlines()[-1])
print('ms/frame', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], d['e2e'].get('sync_call_ms'), 'launches', d['gpu_launches'])
for k,v in d['roofline']['per_kernel'].items(): print('   ', k, round(v['ms_per_launch'],4), 'x', v['launches'])
print('parity', d['parity']); print('cpu', d['cpu_baseline']); print('clocks', d['clocks'])
for k,v in (d['configs'] or {}).items(): print('  ', k, v)
PY
tail -5 gpurun_out/step4_bench.err
python bench.py --impl reference 2>/dev/null | tail -1 | cut -c1-600
