"""Aggregate an ncu source page (SASS) into executed-instruction counts per opcode, per kernel.
usage: ncu -i rep.ncu-rep --page source --csv | python tools/ncu_opcodes.py [top_n]"""
import collections
import csv
import sys

top = int(sys.argv[1]) if len(sys.argv) > 1 else 25
kern, cols, agg = None, None, {}
for row in csv.reader(sys.stdin):
    if not row:
        continue
    if row[0] == "Kernel Name":
        kern = row[1]
        agg.setdefault(kern, collections.Counter())
        cols = None
        continue
    if row[0] == "Address":
        cols = {h: i for i, h in enumerate(row)}
        continue
    if cols is None or kern is None:
        continue
    sass = row[cols["Source"]].strip()
    op = sass.split()[0] if sass and not sass.startswith("@") else (sass.split()[1] if len(sass.split()) > 1 else sass)
    op = op.split(".")[0] + ("." + op.split(".")[1] if op.startswith("MUFU") and "." in op else "")
    try:
        n = int(row[cols["Instructions Executed"]])
    except ValueError:
        continue
    agg[kern][op] += n
for k, c in agg.items():
    tot = sum(c.values())
    print(f"== {k}: {tot/1e6:.1f} M warp-instructions")
    for op, n in c.most_common(top):
        print(f"   {op:14s} {n/1e6:9.1f} M  {100*n/tot:5.1f}%")
