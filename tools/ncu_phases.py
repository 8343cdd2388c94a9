"""Split an `ncu --page source --csv` SASS dump into phases at barrier/global-memory landmarks and
print stall samples and instruction counts per phase (address order = program order inside the loop).
usage: python tools/ncu_phases.py src.csv"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
def kind(src):
    s = src.split()
    if not s: return "?"
    op = s[1] if s[0].startswith("@") else s[0]
    return op.split(".")[0]
phases = []
cur = dict(start=None, n=0, inst=0, st=collections.Counter(), ops=collections.Counter(), first="")
LAND = {"BAR", "LDGSTS", "STG", "SYNCS", "LDGDEPBAR", "DEPBAR", "RED", "ATOMG", "LDG"}
prev_land = None
for r in rows[2:]:
    if len(r) < len(hdr): continue
    try: n = int(r[col["# Samples"]]); ex = int(r[col["Instructions Executed"]])
    except ValueError: continue
    src = r[col["Source"]].strip()
    k = kind(src)
    land = k if k in LAND else None
    if land != prev_land and land is not None and cur["inst"] > 0:
        phases.append(cur)
        cur = dict(start=r[col["Address"]], n=0, inst=0, st=collections.Counter(), ops=collections.Counter(), first=src)
    prev_land = land if land is not None else prev_land
    if land is None and prev_land is not None:
        # leaving a landmark run: start a new phase
        phases.append(cur)
        cur = dict(start=r[col["Address"]], n=0, inst=0, st=collections.Counter(), ops=collections.Counter(), first=src)
        prev_land = None
    cur["n"] += n; cur["inst"] += ex; cur["ops"][k] += ex
    for s in stalls:
        try: cur["st"][s] += int(r[col[s]])
        except ValueError: pass
phases.append(cur)
tot = sum(p["n"] for p in phases)
# merge tiny phases into neighbours for readability
out = []
for p in phases:
    if out and (p["n"] < 0.004 * tot and p["inst"] < 20000):
        q = out[-1]; q["n"] += p["n"]; q["inst"] += p["inst"]; q["st"] += p["st"]; q["ops"] += p["ops"]
    else:
        out.append(p)
print("total samples", tot)
for p in out:
    top = ", ".join(f"{k[6:]}:{v}" for k, v in p["st"].most_common(3))
    ops = ", ".join(f"{k}:{v//1000}k" for k, v in p["ops"].most_common(4))
    print(f"{100.0*p['n']/tot:5.1f}%  inst {p['inst']//1000:7d}k  [{ops}]  stalls {top}   | {p['first'][:50]}")
