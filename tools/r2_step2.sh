#!/bin/bash
# GPU box: fast chain v2 (packed K3, compacted-ray K1): parity tests, parity at 4K, bench A/B against the round-1 K1, ncu of one frame
cd "$(dirname "$0")/.."
python -m pytest tests/test_gpu_chain.py -x -q 2>&1 | tail -15
python tools/parity_at_size.py C3 --frames 3 --fast-only --out gpurun_out/parity_step2.json 2>&1 | tail -20
show() { python - "$1" <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print(sys.argv[1], 'ms/frame', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], d['e2e'].get('sync_call_ms'))
for k,v in d['roofline']['per_kernel'].items(): print('   ', k, round(v['ms_per_launch'],4), 'x', v['launches'])
PY
}
python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/step2_bench.err | tail -1 > gpurun_out/step2_bench.json; show gpurun_out/step2_bench.json
RFX_LEGACY_K1=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/step2_bench_legacyk1.json; show gpurun_out/step2_bench_legacyk1.json
tail -3 gpurun_out/step2_bench.err
K='regex:ssgi_fast|ssgi_kernel|ctemporal|cpoisson|viewz|cdecode'
ncu --set full --clock-control none --import-source on -k "$K" --launch-skip 24 --launch-count 8 -f -o gpurun_out/r02_s2_frame \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r02_s2_ncu.log 2>&1
tail -3 gpurun_out/r02_s2_ncu.log
ls -la gpurun_out | tail -5
