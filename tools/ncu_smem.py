"""Shared-memory instructions of an `ncu --page source --csv` dump with their wavefront counts
(actual vs ideal) -- finds bank-conflicted accesses."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
tot = ideal = 0
out = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    try:
        w = int(r[col["L1 Wavefronts Shared"]] or 0)
        wi = int(r[col["L1 Wavefronts Shared Ideal"]] or 0)
        ex = int(r[col["Instructions Executed"]] or 0)
    except ValueError:
        continue
    if w == 0:
        continue
    tot += w
    ideal += wi
    out.append((w - wi, w, wi, ex, r[col["Source"]].strip()))
print("shared wavefronts", tot, "ideal", ideal)
agg = {}
for d, w, wi, ex, s in out:
    op = s.split()[0] if not s.startswith("@") else s.split()[1]
    a = agg.setdefault(op, [0, 0, 0])
    a[0] += w; a[1] += wi; a[2] += ex
for op, (w, wi, ex) in sorted(agg.items(), key=lambda x: -x[1][0]):
    print(f"  {op:12s} wavefronts {w:10d} ideal {wi:10d} executed {ex:9d}  wf/inst {w/max(ex,1):.2f} (ideal {wi/max(ex,1):.2f})")
print("worst instructions:")
for d, w, wi, ex, s in sorted(out, key=lambda x: -x[0])[: int(sys.argv[2]) if len(sys.argv) > 2 else 12]:
    print(f"  +{d:8d}  wf {w:8d} ideal {wi:8d} exec {ex:7d}  {s[:60]}")
