#!/bin/bash
# GPU box, 1 GPU: full -m gpu suite; fused-K4 arithmetic modes (parity at 4K + time); default bench; TMA off A/B; K1 occupancy sweep; ncu
cd "$(dirname "$0")/.."
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
show() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
except Exception as e:
    print(sys.argv[1], 'NO JSON', e); sys.exit(0)
print(sys.argv[1], 'ms/frame', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], d['e2e'].get('sync_call_ms'), 'launches', d['gpu_launches'])
for k,v in d['roofline']['per_kernel'].items(): print('   ', k, round(v['ms_per_launch'],4), 'x', v['launches'])
if d.get('parity'): print('   parity', d['parity'])
if d.get('cpu_baseline'): print('   cpu', d['cpu_baseline'])
print('   clocks', d['clocks'], 'roofline', d['roofline']['kernel'], d['roofline']['frac'], 'chain', d['roofline']['chain_frac'])
for k,v in (d.get('configs') or {}).items(): print('   ', k, v)
PY
}
for m in 0 1 2; do
  echo "== compose mode $m"
  RFX_COMPOSE_MODE=$m python tools/parity_at_size.py C3 --frames 3 --fast-only --out gpurun_out/parity_compose$m.json 2>&1 | grep -E "worst|composed|f2.ssgi" | head -5
  RFX_COMPOSE_MODE=$m timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-configs 2>/dev/null | tail -1 > gpurun_out/step5_compose$m.json; show gpurun_out/step5_compose$m.json | head -6
done
( time python bench.py > gpurun_out/step5_bench.json 2> gpurun_out/step5_bench.err ) 2>&1 | tail -3
show gpurun_out/step5_bench.json; tail -5 gpurun_out/step5_bench.err
python bench.py --impl reference 2>/dev/null | tail -1 > gpurun_out/step5_reference.json; cut -c1-400 gpurun_out/step5_reference.json
RFX_K3_TMA=0 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-configs 2>/dev/null | tail -1 > gpurun_out/step5_notma.json; show gpurun_out/step5_notma.json | head -6
K='regex:ssgi_fast|ctemporal|cpoisson|viewz|cdecode'
timeout 600 ncu --set full --clock-control none --import-source on -k "$K" --launch-skip 80 --launch-count 8 -f -o gpurun_out/r02_s5_frame \
    python bench.py --steps 4 --warmup 10 --no-cpu-baseline --no-configs > gpurun_out/r02_s5_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 120 --csv --log-file gpurun_out/r02_s5_launches.csv \
    python bench.py --steps 4 --warmup 10 --no-cpu-baseline --no-configs > gpurun_out/r02_s5_ncu_list.log 2>&1
timeout 300 ncu --set full --clock-control none -k "regex:hbao_kernel|motion_blur_kernel|ao_compose|poisson_fast" --launch-count 6 -f -o gpurun_out/r02_s5_post \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_s5_ncu_post.log 2>&1
for mb in 5 6; do
  rm -f realism_effects_b200/csrc/build/k_ssgi.o
  RFX_NVCC_EXTRA="-DRFX_K1_MIN_BLOCKS=$mb" python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
  echo "== K1 min blocks $mb"; timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-configs 2>/dev/null | tail -1 > gpurun_out/step5_k1mb$mb.json; show gpurun_out/step5_k1mb$mb.json | head -2
done
ls -la gpurun_out | tail -6
