"""Generates the golden fixtures under tests/golden/ from the reference itself.

Run in the build container (needs /root/reference and oracle/_ref/libref_oracle.so, built by
`make -C oracle`):   python tests/golden/make_golden.py
Each fixture holds the synthetic timf1 input and what the reference's own compiled C path
(fft1_b -> fft1_c -> fft1_waterfall -> fft1_mix1_fixed, driven by oracle/ref_driver.c) produced
for it.  The reference ships no golden vectors of its own; these files are what
pins oracle/port.py and the CUDA path on machines where /root/reference does not exist.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from linrad_b200 import sizing  # noqa: E402
from linrad_b200.synth import make_timf1  # noqa: E402
from oracle import refwrap  # noqa: E402
from tests.helpers import run_reference  # noqa: E402

IQ, DW, TWO = sizing.IQ_DATA, sizing.DWORD_INPUT, sizing.TWO_CHANNELS

CASES = {
    # name: (PathSetup kwargs, reference fft_cntrl row, selections in bins, blocks)
    "iq16_n512_sin2": (dict(input_mode=IQ, rf_channels=1, ad_speed=96000, fft1_n=9, mix1_red_n=3), 6, [180.37, 60.0], 12),
    "iq32x2_n256_sin2": (dict(input_mode=IQ | DW | TWO, rf_channels=2, ad_speed=192000, fft1_n=8, mix1_red_n=3), 7, [90.74], 11),
    "real16_n512_sin2": (dict(input_mode=0, rf_channels=1, ad_speed=48000, fft1_n=9, mix1_red_n=3), 2, [100.3], 12),
    "real32x2_n256_sin2": (dict(input_mode=DW | TWO, rf_channels=2, ad_speed=48000, fft1_n=8, mix1_red_n=3), 2, [40.5], 10),
    "iq16_n512_sin3_rev": (dict(input_mode=IQ, rf_channels=1, ad_speed=96000, fft1_n=9, mix1_red_n=3, sinpow=3, direction=-1), 6, [200.4], 12),
    "iq16_n256_nowin": (dict(input_mode=IQ, rf_channels=1, ad_speed=96000, fft1_n=8, mix1_red_n=3, sinpow=0), 7, [100.25], 9),
    "iq16_n512_range": (dict(input_mode=IQ, rf_channels=1, ad_speed=96000, fft1_n=9, mix1_red_n=3, first_xpoint=70, xpoints=330), 6, [372.6, -1], 10),
}


def main():
    assert refwrap.available(), "build oracle/_ref first (make -C oracle)"
    for name, (kw, ver, sel, nb) in CASES.items():
        s = sizing.PathSetup(**kw)
        raw = make_timf1(s.input_mode, s.rf_channels, s.fft1_size, nb, s.fft1_new_points, seed=11)
        ref = run_reference(dict(kw, version=ver), raw, sel, nb, want_raw=True)
        r = ref["ref"]
        st = ref["states"]
        out = dict(
            raw_input=raw, fft1_raw=ref["raw"].astype(np.float32), fft1=ref["fft1"].astype(np.float32),
            timf3=ref["timf3"].astype(np.float32) if ref["timf3"] is not None else np.zeros(0, np.float32),
            sumsq=ref["sumsq"].astype(np.float32), sumsq_pa=np.int64(ref["sumsq_pa"]),
            sumsq_counter=np.int64(ref["sumsq_counter"]), slowsum=r.slowsum(), waterf=r.waterf(),
            waterf_ptr=np.int64(r.waterf_ptr()), waterf_yfac=r.waterf_yfac(), waterf_sum=r.waterf_sum(),
            window=r.window(r.fft1_size + 1 if not (s.input_mode & IQ) else r.fft1_size), filtercorr=r.filtercorr(),
            mix1_fqwin=r.mix1_fqwin(), mix1_window=r.mix1_window(), mix1_cos2win=r.mix1_cos2win(), mix1_sin2win=r.mix1_sin2win(),
            mix1_crossover=np.int64(r.mix1_crossover), timf3_size=np.int64(r.timf3_size), timf3_pa=np.int64(ref["timf3_pa"]),
            sel_point=np.array([x["point"] for x in st], np.int64), sel_old_point=np.array([x["old_point"] for x in st], np.int64),
            sel_phase=np.array([x["phase"] for x in st], np.float32), sel_phase_step=np.array([x["phase_step"] for x in st], np.float32),
            sel_phase_rot=np.array([x["phase_rot"] for x in st], np.float32), sel_old_phase=np.array([x["old_phase"] for x in st], np.float32),
            timf3_ring=np.stack(ref["timf3_ring"]) if sel else np.zeros(0, np.float32),
            version=np.int64(ver), selbins=np.array(sel, np.float64), nblocks=np.int64(nb),
            kw_keys=np.array(list(kw.keys())), kw_vals=np.array([int(v) for v in kw.values()], np.int64),
        )
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(name, os.path.getsize(path) >> 10, "KiB")
    # table builders: make_window for every layout/kind the path uses (fft0.c:812-921)
    r = refwrap.RefOracle(input_mode=IQ, rf_channels=1, ad_speed=96000, fft1_n=8, fft1_version=6)
    tabs = {}
    for mo, sz, n in [(4, 64, 2), (4, 256, 1), (4, 256, 3), (4, 128, 7), (4, 256, 8), (4, 256, 9), (1, 64, 2), (1, 256, 4),
                      (2, 64, 2), (2, 128, 3), (3, 64, 3), (3, 128, 8), (5, 64, 4), (5, 512, 4)]:
        cnt = sz if mo in (1, 4) else (sz + 1 if mo == 2 else sz // 2 + 1)
        tabs[f"win_{mo}_{sz}_{n}"] = r.make_window(mo, sz, n, count=cnt)
    rng = np.random.default_rng(5)
    x = (rng.standard_normal(64) + 1j * rng.standard_normal(64)).astype(np.complex64)
    tabs["fftback_in"] = x
    tabs["fftback_out"] = r.fftback(x)
    np.savez_compressed(os.path.join(HERE, "tables.npz"), **tabs)
    print("tables ok")


if __name__ == "__main__":
    main()
