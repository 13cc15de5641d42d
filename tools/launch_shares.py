"""Per-kernel share of an ncu launch list (gpu__time_duration.sum csv).  usage: python tools/launch_shares.py launches.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[start]
idx = {h: i for i, h in enumerate(hdr)}
agg = collections.OrderedDict()
for r in rows[start + 1:]:
    if len(r) < len(hdr) or r[idx["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "")
    v, u = float(r[idx["Metric Value"]]), r[idx["Metric Unit"]]
    v = v / 1e3 if u.startswith("us") else (v / 1e6 if u.startswith("ns") else v)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(v[1] for v in agg.values())
print("launch list: ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES, not absolutes)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:60]:60s} n={v[0]:3d} total={v[1]:8.3f} ms  mean={v[1]/v[0]:7.3f} ms  share={100*v[1]/tot:5.1f}%")
