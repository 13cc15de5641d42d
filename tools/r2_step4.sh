#!/bin/bash
# GPU box, 1 GPU: the whole -m gpu suite (incl. parity at the BASELINE sizes), the default bench line, the reference arm,
# the TMA-staging A/B of the Poisson passes, and one ncu --set full capture of a steady-state frame.
cd "$(dirname "$0")/.."
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25
python tools/parity_at_size.py C3 --frames 3 --fast-only --out gpurun_out/parity_step4.json 2>&1 | grep -E "worst|composed|ssgi|dn1" | head -12
( time python bench.py > gpurun_out/step4_bench.json 2> gpurun_out/step4_bench.err ) 2>&1 | tail -3
show() { python - "$1" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], 'ms/frame', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], d['e2e'].get('sync_call_ms'), 'launches', d['gpu_launches'])
for k,v in d['roofline']['per_kernel'].items(): print('   ', k, round(v['ms_per_launch'],4), 'x', v['launches'])
if d.get('parity'): print('parity', d['parity'])
if d.get('cpu_baseline'): print('cpu', d['cpu_baseline'])
print('clocks', d['clocks'])
for k,v in (d.get('configs') or {}).items(): print('  ', k, v)
PY
}
show gpurun_out/step4_bench.json
tail -5 gpurun_out/step4_bench.err
python bench.py --impl reference 2>/dev/null | tail -1 | cut -c1-700 | tee gpurun_out/step4_reference.json
# TMA experiment: same tests (bytes must match: the staged path is bit-identical), then the A/B timing
RFX_K3_TMA=1 timeout 600 python -m pytest tests/test_gpu_chain.py -x -q 2>&1 | tail -3
RFX_K3_TMA=1 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-configs 2>gpurun_out/step4_tma.err | tail -1 > gpurun_out/step4_bench_tma.json; show gpurun_out/step4_bench_tma.json; tail -3 gpurun_out/step4_tma.err
RFX_LEGACY_K1=1 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-configs 2>/dev/null | tail -1 > gpurun_out/step4_bench_legacyk1.json; show gpurun_out/step4_bench_legacyk1.json
for b in 1 4; do RFX_K1_BATCH=$b timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-configs 2>/dev/null | tail -1 > gpurun_out/step4_bench_k1batch$b.json; show gpurun_out/step4_bench_k1batch$b.json | head -2; done
K='regex:ssgi_fast|ctemporal|cpoisson|viewz|cdecode'
timeout 600 ncu --set full --clock-control none --import-source on -k "$K" --launch-skip 80 --launch-count 8 -f -o gpurun_out/r02_s4_frame \
    python bench.py --steps 4 --warmup 10 --no-cpu-baseline --no-configs > gpurun_out/r02_s4_ncu.log 2>&1
RFX_K3_TMA=1 timeout 600 ncu --set full --clock-control none --import-source on -k "regex:cpoisson_tma" --launch-skip 30 --launch-count 2 -f -o gpurun_out/r02_s4_tma \
    python bench.py --steps 4 --warmup 10 --no-cpu-baseline --no-configs > gpurun_out/r02_s4_ncu_tma.log 2>&1
ls -la gpurun_out | tail -8
