"""Warp-stall samples of an ncu report aggregated by CUDA source line.
usage: ncu -i x.ncu-rep --page source --csv --print-source cuda,sass > dump.csv; python tools/ncu_lines.py dump.csv [N]"""
import csv
import collections
import os
import sys

rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
cur, hdr, col = None, None, {}
agg = collections.Counter()
why = collections.defaultdict(collections.Counter)
text = {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur, hdr = os.path.basename(r[1]), None
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        col = {h: i for i, h in enumerate(hdr)}
        continue
    if hdr is None or len(r) < len(hdr) or not r[0].strip():
        continue
    try:
        line, n = int(r[0]), int(r[col["# Samples"]])
    except ValueError:
        continue
    agg[(cur, line)] += n
    text[(cur, line)] = r[1].strip()
    for h in hdr:
        if h.startswith("stall_") and "Not Issued" not in h:
            try:
                why[(cur, line)][h] += int(r[col[h]])
            except ValueError:
                pass
tot = sum(agg.values())
print("total samples", tot)
for (f, l), n in agg.most_common(top_n):
    top = ", ".join(f"{k[6:]}={v}" for k, v in why[(f, l)].most_common(2))
    print(f"{n:6d} {100 * n / max(tot, 1):5.1f}% {f}:{l:4d} [{top}] {text[(f, l)][:90]}")
