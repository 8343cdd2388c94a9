import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d["roofline"]
        print(f"{f}: {d['value']:.0f} Ms/s  ms/step {d['ms_per_step']:.4f}  fft1 ms {r['kernel_ms']:.4f}  frac {r['frac']:.3f}  whole {r['whole_step_frac']:.3f}  e2e {d['e2e']['value'] if d.get('e2e') else None}")
    except Exception as e:
        print(f, "ERR", e)
