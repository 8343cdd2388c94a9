"""Shared drivers for the parity tests: run the same synthetic timf1 through the compiled
reference (oracle/_ref), and through liblinrad_b200.so via its C ABI, with identical ring
sizes and the reference's own index bookkeeping (wcw.c:1036-1047, fft1.c:4507-4523,
mix1.c:1038-1040)."""
import numpy as np

from linrad_b200 import api, sizing
from linrad_b200.synth import make_timf1

IQ_DATA, DWORD_INPUT, TWO_CHANNELS = sizing.IQ_DATA, sizing.DWORD_INPUT, sizing.TWO_CHANNELS

# BASELINE.json configurations; version = reference fft_cntrl row used as oracle
CONFIGS = {
    "cfg1": dict(input_mode=IQ_DATA, rf_channels=1, ad_speed=96000, fft1_n=13, mix1_red_n=4, version=6),
    "cfg2": dict(input_mode=IQ_DATA | DWORD_INPUT | TWO_CHANNELS, rf_channels=2, ad_speed=192000, fft1_n=14,
                 mix1_red_n=4, version=7),
    "cfg3": dict(input_mode=0, rf_channels=1, ad_speed=2400000, fft1_n=15, mix1_red_n=5, version=2),
    "cfg4": dict(input_mode=IQ_DATA, rf_channels=1, ad_speed=20000000, fft1_n=18, mix1_red_n=6, version=20),
}


def rel_rms(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.sqrt(((a - b) ** 2).sum() / max((b ** 2).sum(), 1e-300)))


def pow2_at_least(x):
    p = 1
    while p < x:
        p *= 2
    return p


def run_reference(kw, raw, selbins, nblocks, want_raw=False, foldcorr=None, pg_ch2=None, **extra):
    from oracle.refwrap import RefOracle
    kw = dict(kw)
    version = kw.pop("version")
    r = RefOracle(fft1_version=version, n_sel=len(selbins), max_fft1n=8, **kw, **extra)
    if foldcorr is not None:
        r.set_foldcorr(foldcorr)
    if pg_ch2 is not None:
        r.set_ch2_phasing(*pg_ch2)
    hz = kw["ad_speed"] / (1 << kw["fft1_n"]) / (1 if kw["input_mode"] & IQ_DATA else 2)
    for i, fb in enumerate(selbins):
        r.set_selfreq(i, fb * hz if fb >= 0 else -1.0)
    rawb = np.ascontiguousarray(raw).view(np.uint8).reshape(-1)
    out = r.process(rawb[: nblocks * r.timf1_blockbytes], want_raw=want_raw)
    out["sumsq"] = r.sumsq()
    out["sumsq_pa"] = r.sumsq_pa()
    out["sumsq_counter"] = r.sumsq_counter()
    out["states"] = [r.sel_state(i) for i in range(len(selbins))]
    out["timf3_ring"] = [r.timf3_ring(i) for i in range(len(selbins))]
    out["timf3_pa"] = r.timf3_pa()
    out["timf3_size"] = r.timf3_size
    out["ref"] = r
    return out


class CudaStream:
    """Host-ring driver for the C ABI: keeps Linrad-style rings and indices in numpy and
    forwards to lb200_fft1 / lb200_mix1 exactly where the reference calls fft1_b+fft1_c and
    fft1_mix1_fixed."""

    def __init__(self, setup, selbins, window=None, filtercorr=None, max_fft1n=8, sumsq_rows=16,
                 timf3_size=None, timf1_bytes=None, foldcorr=None, sample_shift=0, pg_ch2=(1.0, 0.0), correlation=0):
        self.s = setup
        self.plan = api.Plan(setup, window=window, filtercorr=filtercorr, foldcorr=foldcorr, sample_shift=sample_shift,
                             pg_ch2=pg_ch2)
        N = setup.fft1_size
        # 8 transforms' worth of input (a real-input transform covers 2N frames)
        self.timf1_bytes = timf1_bytes or pow2_at_least(8 * N * setup.frame_bytes * (1 if setup.input_mode & IQ_DATA else 2))
        self.timf1 = np.zeros(self.timf1_bytes, np.uint8)
        self.fft1 = np.zeros(max_fft1n * setup.fft1_block, np.float32)
        self.sumsq = np.zeros(sumsq_rows * N, np.float32)
        self.corrsum = np.zeros(2 * sumsq_rows * N, np.float32) if correlation == 1 else None
        self.timf3_size = timf3_size or 16 * setup.mix1_size * 2 * setup.rf_channels
        self.nsel = len(selbins)
        self.timf3 = np.zeros(max(self.nsel, 1) * 2 * self.timf3_size, np.float32)
        hz = setup.ad_speed / N / (1 if setup.input_mode & IQ_DATA else 2)
        self.states = api.new_states([fb * hz if fb >= 0 else -1.0 for fb in selbins])
        self.pa = self.px = 0
        self.fft1_pa = self.fft1_px = 0
        self.sumsq_pa = self.sumsq_counter = 0
        self.timf3_pa = 0

    def process(self, raw, nblocks, chunk=4, apply_fc=True, mix=True):
        """raw: integer frames for nblocks transforms; returns per-transform copies like the
        oracle driver does."""
        s = self.s
        rawb = np.ascontiguousarray(raw).view(np.uint8).reshape(-1)
        fft1_out = np.zeros((nblocks, s.fft1_block), np.float32)
        t3_out = np.zeros((nblocks, max(self.nsel, 1), s.timf3_block), np.float32)
        done = 0
        while done < nblocks:
            nb = min(chunk, nblocks - done)
            nbytes = nb * s.timf1_blockbytes
            src = rawb[done * s.timf1_blockbytes: done * s.timf1_blockbytes + nbytes]
            idx = (self.pa + np.arange(nbytes)) & (self.timf1_bytes - 1)
            self.timf1[idx] = src
            self.pa = (self.pa + nbytes) & (self.timf1_bytes - 1)
            self.plan.fft1_host(timf1=self.timf1, ref=self.px, nblocks=nb, fft1=self.fft1, fft1_pa=self.fft1_pa,
                                apply_fc=apply_fc, sumsq=self.sumsq if apply_fc else None,
                                sumsq_pa=self.sumsq_pa, counter=self.sumsq_counter,
                                corrsum=self.corrsum if apply_fc else None)
            for b in range(nb):
                at = (self.fft1_pa + b * s.fft1_block) & (self.fft1.size - 1)
                fft1_out[done + b] = self.fft1[at: at + s.fft1_block]
            self.px = (self.px + nbytes) & (self.timf1_bytes - 1)
            if apply_fc:
                tot = self.sumsq_counter + nb
                self.sumsq_pa = (self.sumsq_pa + (tot // s.avg1num) * s.fft1_size) & (self.sumsq.size - 1)
                self.sumsq_counter = tot % s.avg1num
            if mix and self.nsel:
                self.plan.mix1_host(fft1=self.fft1, fft1_px=self.fft1_px, nblocks=nb, states=self.states,
                                    timf3=self.timf3, timf3_floats=self.timf3_size, timf3_pa=self.timf3_pa)
                for b in range(nb):
                    pa = (self.timf3_pa + b * s.timf3_block)
                    for ss in range(self.nsel):
                        ring = self.timf3[ss * 2 * self.timf3_size: ss * 2 * self.timf3_size + self.timf3_size]
                        t3_out[done + b, ss] = ring[(pa + np.arange(s.timf3_block)) & (self.timf3_size - 1)]
                self.timf3_pa = (self.timf3_pa + nb * s.timf3_block) & (self.timf3_size - 1)
            self.fft1_pa = (self.fft1_pa + nb * s.fft1_block) & (self.fft1.size - 1)
            self.fft1_px = self.fft1_pa
            done += nb
        return dict(fft1=fft1_out, timf3=t3_out)

    def close(self):
        self.plan.close()


# ---------------------------------------------------------------------------------------------
# Parity report: every comparison against the reference also records its PLAIN figures (no noise
# floor allowance), so that the widened tolerances of the tests can be judged:
#   worst_plain  = max over bins of |got - ref| / (1e-4 * ref)      (north_star bound = 1.0)
#   frac_over    = fraction of bins with |got - ref| > 1e-4 * ref
# and, where the reference has two float versions for the set-up (fft_cntrl rows 6 and 7), the same
# two figures between the reference's OWN versions on the same input.
# Lines go to $LB200_PARITY_REPORT (default gpurun_out/parity_report.jsonl when that directory exists).
def power_plain_figures(got, ref, tol=1e-4):
    got = np.asarray(got, np.float64)
    ref = np.asarray(ref, np.float64)
    ok = ref > 0
    if not ok.any():
        return 0.0, 0.0
    r = np.abs(got - ref)[ok] / (tol * ref[ok])
    return float(r.max()), float((r > 1.0).mean())


def parity_record(**fields):
    import json
    import os
    path = os.environ.get("LB200_PARITY_REPORT")
    if not path:
        d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
        if not os.path.isdir(d):
            return
        path = os.path.join(d, "parity_report.jsonl")
    fields["test"] = os.environ.get("PYTEST_CURRENT_TEST", "?").split(" ")[0]
    try:
        with open(path, "a") as f:
            f.write(json.dumps(fields) + "\n")
    except OSError:
        pass
