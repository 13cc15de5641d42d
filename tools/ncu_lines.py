"""Per-CUDA-source-line executed instruction counts from an ncu report.
usage: ncu -i rep.ncu-rep --page source --csv --print-source cuda,sass | python tools/ncu_lines.py [top_n] [kernel_index]"""
import csv
import sys

top = int(sys.argv[1]) if len(sys.argv) > 1 else 40
want_k = int(sys.argv[2]) if len(sys.argv) > 2 else 0
fname, kidx, cols, agg, kname = None, -1, None, {}, None
last_file = None
for r in csv.reader(sys.stdin):
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        if kname != r[1]:
            kname = r[1]
            kidx += 1
        continue
    if r[0] == "Line No":
        cols = {h: i for i, h in enumerate(r)}
        ie = [i for i, h in enumerate(r) if h == "Instructions Executed"][0]
        st = [i for i, h in enumerate(r) if h == "# Samples"][0]
        continue
    if cols is None or kidx != want_k or not r[0].isdigit():
        continue
    try:
        n, s = int(r[ie]), int(r[st])
    except ValueError:
        continue
    key = (fname, int(r[0]))
    a = agg.setdefault(key, [0, 0, r[1].strip()[:110]])
    a[0] += n
    a[1] += s
tot = sum(a[0] for a in agg.values()) or 1
tots = sum(a[1] for a in agg.values()) or 1
print(f"kernel #{want_k}: {kname}\n total {tot/1e6:.1f} M warp-instructions, {tots} stall samples")
for (f, ln), (n, s, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{n/1e6:8.1f}M {100*n/tot:5.1f}%  samples {100*s/tots:5.1f}%  {f}:{ln}  {src}")
