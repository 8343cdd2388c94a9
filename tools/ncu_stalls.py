"""Summarise an `ncu --page source --csv` dump: total warp-stall samples by reason, and the
instructions with the most samples (SASS view)."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = collections.Counter()
items = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    try:
        n = int(r[col["# Samples"]])
    except ValueError:
        continue
    for s in stalls:
        try:
            tot[s] += int(r[col[s]])
        except ValueError:
            pass
    items.append((n, r[col["Source"]].strip(), {s: int(r[col[s]] or 0) for s in stalls if (r[col[s]] or "0") != "0"}))
allsamp = sum(tot.values())
print("total samples", allsamp)
for s, n in tot.most_common(12):
    print(f"  {s:28s} {n:8d} {100.0*n/allsamp:5.1f}%")
print("top instructions:")
for n, src, st in sorted(items, key=lambda x: -x[0])[: int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    top = sorted(st.items(), key=lambda x: -x[1])[:2]
    print(f"  {n:6d}  {src[:70]:70s} {top}")
# instruction mix
mix = collections.Counter()
for r in rows[2:]:
    if len(r) < len(hdr): continue
    try: ex = int(r[col["Instructions Executed"]])
    except ValueError: continue
    op = r[col["Source"]].strip().split()[0] if r[col["Source"]].strip() else "?"
    if op.startswith("@"): op = r[col["Source"]].strip().split()[1]
    mix[op.split(".")[0]] += ex
tot_i = sum(mix.values())
print("instruction mix (warp-level):", tot_i)
for op, n in mix.most_common(18):
    print(f"  {op:10s} {n:10d} {100.0*n/tot_i:5.1f}%")
