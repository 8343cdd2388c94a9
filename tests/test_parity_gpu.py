"""Parity of the CUDA path (through the C ABI) with the reference's own compiled C path
(oracle/_ref) on identical synthetic inputs.  Tolerances are BASELINE.json's north_star:
fft1_float relative RMS <= 1e-5, averaged power <= 1e-4 per bin, bin selection bit-exact;
timf3 (no number given there) relative RMS <= 2e-5."""
import numpy as np
import pytest

from linrad_b200 import sizing
from linrad_b200.synth import make_timf1
from oracle import refwrap
from tests.helpers import CONFIGS, CudaStream, rel_rms, run_reference, IQ_DATA, DWORD_INPUT, TWO_CHANNELS

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not refwrap.available(), reason="oracle/_ref not built")]

TOL_FFT1 = 1e-5
TOL_POWER = 1e-4      # per bin, for every bin within 50 dB of the strongest bin of the row
TOL_TIMF3 = 2e-5
# Bins far below the strongest signal carry the float32 rounding noise of the transform itself:
# the reference's own two C implementations (fft_cntrl rows 6 and 7) differ there by up to 3e-4
# per bin on the cfg1 signal (77 dB of dynamic range).  For those bins the CUDA result must stay
# within POWER_SPREAD_FACTOR times the reference-vs-reference spread measured on the same input.
POWER_SPREAD_FACTOR = 3.0


def _setup(kw, **over):
    k = {a: b for a, b in kw.items() if a != "version"}
    k.update(over)
    return sizing.PathSetup(**k)


def _compare(kw, nblocks, selbins, chunk, seed=1, natural_window=True, **over):
    s = _setup(kw, **over)
    raw = make_timf1(s.input_mode, s.rf_channels, s.fft1_size, nblocks, s.fft1_new_points, seed=seed)
    kwr = dict(kw)
    kwr.update(over)
    ref = run_reference(kwr, raw, selbins, nblocks, want_raw=True)
    alt_sumsq = None
    if kwr["version"] in (6, 7):
        kwa = dict(kwr, version=13 - kwr["version"])          # 6 <-> 7
        alt_sumsq = run_reference(kwa, raw, [], nblocks)["sumsq"]
        ref = run_reference(kwr, raw, selbins, nblocks, want_raw=True)   # the oracle keeps one state
    cs = CudaStream(s, selbins)
    try:
        got = cs.process(raw, nblocks, chunk=chunk)
        # fft1_float
        e = rel_rms(got["fft1"], ref["fft1"])
        assert e <= TOL_FFT1, f"fft1_float rel rms {e}"
        # fft1_sumsq: every completed row, per bin
        N, lo, hi = s.fft1_size, s.fft1_first_point, s.fft1_last_point
        rows = (nblocks // s.avg1num)
        assert cs.sumsq_pa == ref["sumsq_pa"] and cs.sumsq_counter == ref["sumsq_counter"]
        for r in range(min(rows, 8)):
            a = cs.sumsq[r * N + lo: r * N + hi + 1]
            b = ref["sumsq"][r * N + lo: r * N + hi + 1]
            err = np.abs(a - b) / np.maximum(np.abs(b), 1e-30)
            strong = b >= 1e-5 * b.max()
            assert err[strong].max() <= TOL_POWER, f"sumsq row {r} strong-bin max rel err {err[strong].max()}"
            limit = TOL_POWER
            if alt_sumsq is not None:
                c = alt_sumsq[r * N + lo: r * N + hi + 1]
                spread = (np.abs(c - b) / np.maximum(np.abs(b), 1e-30)).max()
                limit = max(TOL_POWER, POWER_SPREAD_FACTOR * spread)
            assert err.max() <= limit, f"sumsq row {r} max rel err {err.max()} limit {limit}"
        # mix1
        for ss in range(len(selbins)):
            st = ref["states"][ss]
            assert cs.states[ss].mix1_point == st["point"]
            assert cs.states[ss].mix1_phase == float(st["phase"])
            e3 = rel_rms(got["timf3"][:, ss], ref["timf3"][:, ss])
            assert e3 <= TOL_TIMF3, f"timf3 sel {ss} rel rms {e3}"
            # parked tail + whole ring identical up to tolerance
            ring = cs.timf3[ss * 2 * cs.timf3_size: ss * 2 * cs.timf3_size + cs.timf3_size]
            assert rel_rms(ring, ref["timf3_ring"][ss]) <= TOL_TIMF3
        return e
    finally:
        cs.close()


def test_cfg1_iq16_8192():
    _compare(CONFIGS["cfg1"], 12, [3000.37, 1234.0], chunk=5)


def test_cfg2_iq32_2ch_16384():
    _compare(CONFIGS["cfg2"], 11, [6000.74], chunk=4)


@pytest.mark.parametrize("n", [7, 8, 9, 10, 11, 12])
def test_sizes_iq16(n):
    kw = dict(input_mode=IQ_DATA, rf_channels=1, ad_speed=96000, fft1_n=n, mix1_red_n=3, version=6)
    N = 1 << n
    _compare(kw, 13, [0.3663 * N + 0.37], chunk=3)


@pytest.mark.parametrize("mode,ch,ver", [(IQ_DATA | TWO_CHANNELS, 2, 7), (IQ_DATA | DWORD_INPUT, 1, 6),
                                         (IQ_DATA | DWORD_INPUT | TWO_CHANNELS, 2, 7)])
def test_formats(mode, ch, ver):
    kw = dict(input_mode=mode, rf_channels=ch, ad_speed=96000, fft1_n=11, mix1_red_n=3, version=ver)
    _compare(kw, 9, [700.25], chunk=2)


@pytest.mark.parametrize("sinpow", [0, 1, 3, 4, 8, 9])
def test_window_kinds(sinpow):
    """rectangular, sin^1/3/4, Gaussian, erfc: the crossover overlap scheme of do_mix1"""
    kw = dict(input_mode=IQ_DATA, rf_channels=1, ad_speed=96000, fft1_n=11, mix1_red_n=3, version=6)
    _compare(kw, 14, [812.4], chunk=4, sinpow=sinpow)


def test_direction_reversed():
    kw = dict(input_mode=IQ_DATA, rf_channels=1, ad_speed=96000, fft1_n=11, mix1_red_n=3, version=6)
    _compare(kw, 8, [900.3], chunk=8, direction=-1)


def test_limited_display_range():
    """fft1_first_point/last_point clamps of fft1_c and of the mix1 gather (mix1.c:1020-1030)"""
    kw = dict(input_mode=IQ_DATA, rf_channels=1, ad_speed=96000, fft1_n=11, mix1_red_n=3, version=6)
    # selection close to the upper edge so that part of the M bins fall outside
    _compare(kw, 8, [1480.6], chunk=3, first_xpoint=300, xpoints=1300)


def test_unselected_channel_is_cleared():
    kw = dict(input_mode=IQ_DATA, rf_channels=1, ad_speed=96000, fft1_n=10, mix1_red_n=3, version=6)
    _compare(kw, 6, [400.2, -1], chunk=2)


def test_avg1num_variants():
    kw = dict(input_mode=IQ_DATA, rf_channels=1, ad_speed=96000, fft1_n=10, mix1_red_n=3, version=6)
    for avg in (1, 3, 9):
        _compare(kw, 19, [], chunk=7, avg1num=avg)
