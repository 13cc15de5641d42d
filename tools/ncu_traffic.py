"""DRAM traffic per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the chain kernels from an `ncu --set full` report,
written as the JSON bench.py quotes in roofline.traffic.  usage: python tools/ncu_traffic.py rep.ncu-rep profiles/traffic.json"""
import csv
import json
import subprocess
import sys

KEY = [("ssgi_fast_kernel", "K1_ssgi_trace"), ("ssgi_kernel", "K1_ssgi_trace"), ("ctemporal_kernel", "K2_temporal_reproject"), ("temporal_kernel", "K2_temporal_reproject"),
       ("cpoisson_kernel<1,", "K3_poisson_pass0"), ("cpoisson_kernel<0,", "K3_poisson_pass1plus"), ("cpoisson_tma_kernel", "K3_poisson_pass1plus"),
       ("poisson_fast_kernel<2, 0>", "K3_poisson_pass0"), ("poisson_fast_kernel<2, 1>", "K3_poisson_pass1plus"), ("gi_compose_kernel", "K4_gi_compose"),
       ("viewz_kernel", "prepass_viewz"), ("cdecode_kernel", "prepass_gbuffer_decode"), ("gbuffer_decode_kernel", "prepass_gbuffer_decode")]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
idx = {h: i for i, h in enumerate(rows[0])}
acc = {}
for r in rows[2:]:
    name = r[idx["Kernel Name"]]
    key = next((k for pat, k in KEY if pat in name), None)
    if key is None:
        continue
    b = sum(float(r[idx[m]]) * UNIT[rows[1][idx[m]]] for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    acc.setdefault(key, []).append(b)
res = {"source": f"ncu --set full --clock-control none, one steady-state 3840x2160 frame ({sys.argv[1].split('/')[-1]})",
       "bytes_per_launch": {k: round(sum(v) / len(v)) for k, v in acc.items()}}
json.dump(res, open(sys.argv[2], "w"), indent=1)
print(json.dumps(res, indent=1))
