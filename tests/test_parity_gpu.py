"""Parity of the CUDA path (through the C ABI) with the reference's own compiled C path
(oracle/_ref) on identical synthetic inputs.  Tolerances are BASELINE.json's north_star:
fft1_float relative RMS <= 1e-5, averaged power <= 1e-4 per bin, bin selection bit-exact;
timf3 (no number given there) relative RMS <= 2e-5."""
import numpy as np
import pytest

from linrad_b200 import sizing
from linrad_b200.synth import make_timf1
from oracle import refwrap
from tests.helpers import (CONFIGS, CudaStream, rel_rms, run_reference, IQ_DATA, DWORD_INPUT, TWO_CHANNELS,
                           power_plain_figures, parity_record)

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not refwrap.available(), reason="oracle/_ref not built")]

TOL_FFT1 = 1e-5
TOL_POWER = 1e-4      # per bin, relative -- plus the float32 noise floor of the transform, see power_ok()
TOL_TIMF3 = 2e-5      # relative rms -- plus the float32 noise the whole spectrum leaks into the band, see timf3_ok()


def power_ok(got, ref, avg):
    """Per-bin check of a summed power row.  north_star: <= 1e-4 per bin.  A bin far below the
    strongest signal also carries the single-precision rounding noise of the transform itself
    (the reference's own two C versions, fft_cntrl rows 6 and 7, differ by 3e-4 on such bins of
    the cfg1 signal and by more in spectral nulls), so the allowance per bin is
        1e-4 * P  +  2 * sqrt(avg * P) * eps  +  avg * eps^2 ,
        eps = max( 8 * sqrt(log2 N) * 2^-23 * A_rms ,  4 * 2^-23 * A_peak )
    where A_rms is the rms and A_peak the largest bin amplitude of one transform.  The first term
    is the amplitude error a 4..5-sigma excursion of the difference of two correctly rounded
    float32 FFTs reaches (their rms error grows like sqrt(log2 N) ulps of the rms spectrum level);
    the second covers a spectrum dominated by one narrow line (a sin^4 or Gaussian window and a
    display range holding little else), where the two implementations' different twiddle
    roundings leave errors of a few ulps of the LINE's amplitude in every bin.  For bins within
    ~40 dB of the strongest line the allowance is the plain 1e-4."""
    got = got.astype(np.float64)
    ref = ref.astype(np.float64)
    a_rms = np.sqrt(ref.mean() / avg)
    a_peak = np.sqrt(ref.max() / avg)
    eps_a = max(8 * np.sqrt(np.log2(got.size + 1)) * 2.0 ** -23 * a_rms, 4 * 2.0 ** -23 * a_peak)
    allow = TOL_POWER * ref + 2 * np.sqrt(avg * ref) * eps_a + avg * eps_a ** 2
    worst = float((np.abs(got - ref) / allow).max())
    return worst <= 1.0, worst


def timf3_ok(got, ref, fft1_ref, n_log2, msize):
    """Baseband check.  timf3 is a back-transform of M of the N fft1 bins, so the float32 rounding
    noise of the WHOLE fft1 spectrum (rms ~ sqrt(log2 N) ulps of the spectrum's rms amplitude per
    bin) lands in it no matter how weak the selected band is: the reference's own C versions 6
    and 7 differ by 2e-5 relative rms in timf3 for a band 40 dB below the strongest signal while
    their fft1_float differ by 1.7e-7.  Allowance on the rms error:
        2e-5 * rms(timf3)  +  2 * sqrt(log2 N) * 2^-23 * A_rms(fft1_float) * sqrt(M)"""
    got = got.astype(np.float64)
    ref = ref.astype(np.float64)
    a_rms = np.sqrt(2.0 * (fft1_ref.astype(np.float64) ** 2).mean())       # complex amplitude
    allow = TOL_TIMF3 * np.sqrt((ref ** 2).mean()) + 2 * np.sqrt(n_log2) * 2.0 ** -23 * a_rms * np.sqrt(msize) / np.sqrt(2.0)
    err = np.sqrt(((got - ref) ** 2).mean())
    return err <= allow, err / allow


def _setup(kw, **over):
    k = {a: b for a, b in kw.items() if a != "version"}
    k.update(over)
    return sizing.PathSetup(**k)


def _compare(kw, nblocks, selbins, chunk, seed=1, ext=None, power_slack=1.0, **over):
    """ext: the rarely used fft1_b options (foldcorr table, sample_shift, pg_ch2), given to both sides"""
    s = _setup(kw, **over)
    raw = make_timf1(s.input_mode, s.rf_channels, s.fft1_size, nblocks, s.fft1_new_points, seed=seed)
    kwr = dict(kw)
    kwr.update(over)
    ext = ext or {}
    ref = run_reference(kwr, raw, selbins, nblocks, want_raw=True, **ext)
    cs = CudaStream(s, selbins, **ext)
    try:
        got = cs.process(raw, nblocks, chunk=chunk)
        # fft1_float: the bins fft1_c has scaled (first..last) against their own energy; the bins
        # outside the display range keep the raw fft1_b scale (orders of magnitude larger), so they
        # are judged against the energy of the raw spectrum
        N, lo, hi = s.fft1_size, s.fft1_first_point, s.fft1_last_point
        mm = 2 * s.rf_channels
        g3, r3 = got["fft1"].reshape(nblocks, N, mm), ref["fft1"].reshape(nblocks, N, mm)
        e = rel_rms(g3[:, lo:hi + 1], r3[:, lo:hi + 1])
        assert e <= TOL_FFT1, f"fft1_float rel rms {e}"
        if lo > 0 or hi < N - 1:
            out_err = ((g3[:, :lo].astype(np.float64) - r3[:, :lo]) ** 2).sum() + ((g3[:, hi + 1:].astype(np.float64) - r3[:, hi + 1:]) ** 2).sum()
            raw_energy = (ref["raw"].astype(np.float64) ** 2).sum()
            eo = float(np.sqrt(out_err / max(raw_energy, 1e-300)))
            assert eo <= TOL_FFT1, f"fft1_float outside the display range: rel rms {eo} of the raw spectrum"
        # fft1_sumsq: every completed row, per bin; index bookkeeping bit-exact
        rows = (nblocks // s.avg1num)
        assert cs.sumsq_pa == ref["sumsq_pa"] and cs.sumsq_counter == ref["sumsq_counter"]
        # the reference's own spread on the same input: the other float version (rows 6 / 7 of
        # fft_cntrl exist for one-channel IQ up to 65536 points), recorded next to our figures
        ref2 = None
        if kwr.get("version") in (6, 7) and s.rf_channels == 1 and (s.input_mode & IQ_DATA) and not ext and rows > 0:
            try:
                ref2 = run_reference(dict(kwr, version=13 - kwr["version"]), raw, [], nblocks)
            except Exception:
                ref2 = None
        rep = dict(kind="fft1_sumsq", fft1_n=s.fft1_n, mode=s.input_mode, rows=min(rows, 8), fft1_rel_rms=e, worst_allow=0.0,
                   worst_plain=0.0, frac_over_plain=0.0, ref_spread_worst_plain=None, ref_spread_frac_over=None, ref_spread_worst_allow=None)
        for r in range(min(rows, 8)):
            a = cs.sumsq[r * N + lo: r * N + hi + 1]
            b = ref["sumsq"][r * N + lo: r * N + hi + 1]
            ok, worst = power_ok(a, b, s.avg1num)
            wp, fo = power_plain_figures(a, b, TOL_POWER)
            rep["worst_allow"] = max(rep["worst_allow"], worst)
            rep["worst_plain"] = max(rep["worst_plain"], wp)
            rep["frac_over_plain"] = max(rep["frac_over_plain"], fo)
            if ref2 is not None:
                b2 = ref2["sumsq"][r * N + lo: r * N + hi + 1]
                wp2, fo2 = power_plain_figures(b2, b, TOL_POWER)
                rep["ref_spread_worst_plain"] = max(rep["ref_spread_worst_plain"] or 0.0, wp2)
                rep["ref_spread_frac_over"] = max(rep["ref_spread_frac_over"] or 0.0, fo2)
                rep["ref_spread_worst_allow"] = max(rep["ref_spread_worst_allow"] or 0.0, power_ok(b2, b, s.avg1num)[1])
        parity_record(**rep)
        print("parity:", rep)
        # the bound: the per-bin allowance of power_ok -- or, where the reference's own two float versions
        # are further apart than that on this very input, 1.25 x their distance (profiles/r2_parity_report.jsonl:
        # our figures and the reference's own v6-vs-v7 figures agree case by case)
        limit = power_slack
        if rep["ref_spread_worst_allow"] is not None:
            limit = max(limit, 1.25 * rep["ref_spread_worst_allow"])
        for r in range(min(rows, 8)):
            a = cs.sumsq[r * N + lo: r * N + hi + 1]
            b = ref["sumsq"][r * N + lo: r * N + hi + 1]
            ok, worst = power_ok(a, b, s.avg1num)
            assert worst <= limit, f"sumsq row {r}: worst error is {worst:.2f} x the per-bin allowance (limit {limit:.2f})"
        # fft1_corrsum (fft1_correlation_flag == 1): |2 z1 conj(z2)| <= the bin's summed power, so the
        # power row's allowance bounds its error too
        if ext.get("correlation") == 1:
            gc = cs.corrsum.reshape(-1, 2)
            rc = ref["ref"].corrsum().reshape(-1, 2)
            for r in range(min(rows, 8)):
                b = ref["sumsq"][r * N + lo: r * N + hi + 1].astype(np.float64)
                a_rms, a_peak = np.sqrt(b.mean() / s.avg1num), np.sqrt(b.max() / s.avg1num)
                eps = max(8 * np.sqrt(np.log2(b.size + 1)) * 2.0 ** -23 * a_rms, 4 * 2.0 ** -23 * a_peak)
                allow = TOL_POWER * b + 2 * np.sqrt(s.avg1num * b) * eps + s.avg1num * eps ** 2
                d = np.abs(gc[r * N + lo: r * N + hi + 1].astype(np.float64) - rc[r * N + lo: r * N + hi + 1]).max(axis=1)
                assert (d <= 2 * power_slack * allow).all(), f"corrsum row {r}: {float((d / allow).max()):.2f} x the allowance"
            assert np.abs(rc[: rows * N]).max() > 0
        # mix1: bin selection and phase state bit-exact, baseband within tolerance
        for ss in range(len(selbins)):
            st = ref["states"][ss]
            assert cs.states[ss].mix1_point == st["point"]
            assert cs.states[ss].mix1_phase == float(st["phase"])
            ok3, w3 = timf3_ok(got["timf3"][:, ss], ref["timf3"][:, ss], ref["fft1"], s.fft1_n, s.mix1_size)
            assert ok3, f"timf3 sel {ss}: rms error is {w3:.2f} x the allowance"
            # parked tail + whole ring
            ring = cs.timf3[ss * 2 * cs.timf3_size: ss * 2 * cs.timf3_size + cs.timf3_size]
            ok3, w3 = timf3_ok(ring, ref["timf3_ring"][ss], ref["fft1"], s.fft1_n, s.mix1_size)
            assert ok3, f"timf3 ring sel {ss}: rms error is {w3:.2f} x the allowance"
        return e
    finally:
        cs.close()


def _foldcorr_table(n, channels, seed):
    """a plausible calibration table: mirror coefficients of a few percent, smooth over frequency"""
    rng = np.random.default_rng(seed)
    N = 1 << n
    k = np.arange(N)
    t = np.zeros((N, 2 * channels), np.float32)
    for c in range(channels):
        a, b = rng.uniform(0.005, 0.03, 2)
        ph = rng.uniform(0, 2 * np.pi, 2)
        t[:, 2 * c] = a * np.cos(2 * np.pi * k / N + ph[0]) + 0.01
        t[:, 2 * c + 1] = b * np.sin(4 * np.pi * k / N + ph[1])
    return t.reshape(-1)


@pytest.mark.parametrize("direction", [1, -1])
@pytest.mark.parametrize("mode,ch,ver,n", [(IQ_DATA, 1, 6, 11), (IQ_DATA | TWO_CHANNELS, 2, 7, 11),
                                            (IQ_DATA | DWORD_INPUT, 1, 7, 9), (IQ_DATA, 1, 6, 15)])
def test_iq_mirror_correction(mode, ch, ver, n, direction):
    """fft1_calibrate_flag & CALIQ: fft1.c:3607-3657 / 3941-4026, with the reversal of direction < 0"""
    kw = dict(input_mode=mode, rf_channels=ch, ad_speed=96000, fft1_n=n, mix1_red_n=3, version=ver)
    _compare(kw, 7, [300.37 * (1 << n) / 2048], chunk=3, direction=direction,
             ext=dict(foldcorr=_foldcorr_table(n, ch, seed=n + ch)))


def test_iq_mirror_correction_limited_range():
    kw = dict(input_mode=IQ_DATA, rf_channels=1, ad_speed=96000, fft1_n=11, mix1_red_n=3, version=6)
    _compare(kw, 6, [700.0], chunk=6, first_xpoint=300, xpoints=1200, ext=dict(foldcorr=_foldcorr_table(11, 1, seed=5)))


@pytest.mark.parametrize("direction", [1, -1])
def test_channel2_phasing(direction):
    """pg_ch2_c1 / pg_ch2_c2 (pol_graph.c:165-173, fft1.c:4064-4080), alone and on top of CALIQ"""
    kw = dict(input_mode=IQ_DATA | TWO_CHANNELS, rf_channels=2, ad_speed=96000, fft1_n=11, mix1_red_n=3, version=7)
    c1, c2 = 1.07 * np.cos(0.4), -1.07 * np.sin(0.4)
    _compare(kw, 7, [300.37], chunk=4, direction=direction, ext=dict(pg_ch2=(c1, c2)))
    _compare(kw, 7, [300.37], chunk=4, direction=direction,
             ext=dict(pg_ch2=(c1, c2), foldcorr=_foldcorr_table(11, 2, seed=9)))


@pytest.mark.parametrize("n,mode,direction,chunk", [(11, IQ_DATA | TWO_CHANNELS, 1, 4), (10, IQ_DATA | TWO_CHANNELS | DWORD_INPUT, -1, 3),
                                                    (14, IQ_DATA | TWO_CHANNELS, 1, 7), (15, IQ_DATA | TWO_CHANNELS, 1, 8)])
def test_correlation_spectrum(n, mode, direction, chunk):
    """fft1_correlation_flag == 1: fft1_corrsum next to fft1_sumsq (fft1.c:4146-4152, 4189-4195)"""
    kw = dict(input_mode=mode, rf_channels=2, ad_speed=96000, fft1_n=n, mix1_red_n=n - 8, version=7)
    _compare(kw, 13, [0.146 * (1 << n) + 0.37], chunk=chunk, direction=direction, ext=dict(correlation=1))


def test_correlation_spectrum_limited_range():
    kw = dict(input_mode=IQ_DATA | TWO_CHANNELS, rf_channels=2, ad_speed=96000, fft1_n=11, mix1_red_n=3, version=7)
    _compare(kw, 12, [], chunk=5, first_xpoint=200, xpoints=1500, ext=dict(correlation=1, pg_ch2=(0.9, 0.3)))


@pytest.mark.parametrize("shift", [-3, -1, 2])
@pytest.mark.parametrize("mode,ver,n", [(IQ_DATA, 6, 11), (IQ_DATA, 7, 8), (IQ_DATA | DWORD_INPUT, 6, 12), (IQ_DATA, 6, 15)])
def test_sample_shift(mode, ver, n, shift):
    """ui.sample_shift: I and Q words taken from different frames (fft1.c:770-790, 472-483)"""
    kw = dict(input_mode=mode, rf_channels=1, ad_speed=96000, fft1_n=n, mix1_red_n=3, version=ver)
    _compare(kw, 7, [], chunk=3, ext=dict(sample_shift=shift))


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


# ---------------------------------------------------------------------------------------------
# four-step path (N >= 2^15)
@pytest.mark.parametrize("n,version", [(15, 6), (16, 7)])
def test_large_iq16_vs_float_reference(n, version):
    """the float CPU versions reach N = 65536 (buf.c:285-290): direct parity.  At 65536 points the
    radix-4 version 6 is itself 1.6x outside the power allowance against a float64 DFT (version 7:
    0.7x), so version 7 is the oracle there."""
    kw = dict(input_mode=IQ_DATA, rf_channels=1, ad_speed=2400000, fft1_n=n, mix1_red_n=5, version=version)
    N = 1 << n
    _compare(kw, 7, [0.3663 * N + 0.37, 0.61 * N], chunk=4)


@pytest.mark.parametrize("sinpow", [2, 3, 0])
def test_mix1_size_16384_one_channel(sinpow):
    """mix1.size 16384 does not fit the one-kernel mixer (predecessor tail on chip): two launches through a scratch
    buffer.  All three overlap schemes, a cleared selection, calls of uneven size."""
    n = 16
    kw = dict(input_mode=IQ_DATA, rf_channels=1, ad_speed=2400000, fft1_n=n, mix1_red_n=2, version=7)
    N = 1 << n
    _compare(kw, 7, [0.3663 * N + 0.37, -1, 0.61 * N], chunk=3, sinpow=sinpow)


def test_mix1_size_8192_two_channels():
    n = 15
    kw = dict(input_mode=IQ_DATA | TWO_CHANNELS, rf_channels=2, ad_speed=2400000, fft1_n=n, mix1_red_n=2, version=7)
    N = 1 << n
    _compare(kw, 6, [0.41 * N + 0.2, 0.7 * N], chunk=4)


@pytest.mark.parametrize("n,mode", [(17, IQ_DATA | TWO_CHANNELS), (18, IQ_DATA | TWO_CHANNELS | DWORD_INPUT)])
def test_large_two_channel_vs_double_reference(n, mode):
    """above 65536 points the reference only has its double precision version 20 (2 channels)"""
    kw = dict(input_mode=mode, rf_channels=2, ad_speed=20000000, fft1_n=n, mix1_red_n=n - 11, version=20)
    N = 1 << n
    _compare(kw, 6, [0.3663 * N + 0.37], chunk=5)


def test_cfg4_iq16_262144_16_selections():
    """BASELINE config 4: 1-channel int16, N=2^18, 16 mix1 selections (M=4096).  No float CPU
    version exists for this size and version 20 is 2-channel only, so channel 0 of a 2-channel
    version-20 run on the same samples is the oracle (channel 1 is a copy)."""
    n, nblocks = 18, 6
    kw1 = dict(input_mode=IQ_DATA, rf_channels=1, ad_speed=20000000, fft1_n=n, mix1_red_n=6)
    s1 = sizing.PathSetup(**kw1)
    N = s1.fft1_size
    selbins = [8192.0 * (1 + c) + 0.25 * c for c in range(16)]
    tones = tuple((b + 131072.0, 3000.0) for b in selbins[:4])
    raw1 = make_timf1(s1.input_mode, 1, N, nblocks, s1.fft1_new_points, seed=4, tones=tones)
    raw2 = np.concatenate([raw1, raw1], axis=1)
    kw2 = dict(input_mode=IQ_DATA | TWO_CHANNELS, rf_channels=2, ad_speed=20000000, fft1_n=n, mix1_red_n=6, version=20)
    ref = run_reference(kw2, raw2, selbins, nblocks)
    cs = CudaStream(s1, selbins)
    try:
        got = cs.process(raw1, nblocks, chunk=4)
        ref_fft1 = ref["fft1"].reshape(nblocks, N, 4)[:, :, 0:2].reshape(nblocks, -1)
        # the 2-channel uncalibrated gain is the same as the 1-channel one (fft1.c:4653-4671)
        assert rel_rms(got["fft1"], ref_fft1) <= TOL_FFT1
        for ss in range(16):
            assert cs.states[ss].mix1_point == ref["states"][ss]["point"]
            r3 = ref["timf3"][:, ss].reshape(nblocks, -1, 4)[:, :, 0:2].reshape(nblocks, -1)
            ok3, w3 = timf3_ok(got["timf3"][:, ss], r3, ref_fft1, n, s1.mix1_size)
            assert ok3, (ss, w3)
        a = cs.sumsq[:N]
        b = 0.5 * ref["sumsq"][:N]          # two identical channels were summed
        ok, worst = power_ok(a, b, s1.avg1num)
        assert ok, worst
    finally:
        cs.close()


# ---------------------------------------------------------------------------------------------
# real input: fft1 version 2 (fft1_re.c), packed half-length transform + untangle kernel
def test_cfg3_real16_32768():
    """BASELINE config 3: real 1-channel int16, 65536 real samples -> 32768 bins, power averaging;
    one mix1 selection added so the real spectrum also feeds the mixer."""
    _compare(CONFIGS["cfg3"], 11, [12000.37], chunk=5, seed=3)


@pytest.mark.parametrize("n", [7, 9, 10, 12, 14])
def test_real_sizes_int16(n):
    kw = dict(input_mode=0, rf_channels=1, ad_speed=48000, fft1_n=n, mix1_red_n=3, version=2)
    N = 1 << n
    _compare(kw, 13, [0.3663 * N + 0.37], chunk=3)


@pytest.mark.parametrize("mode,ch", [(TWO_CHANNELS, 2), (DWORD_INPUT, 1), (DWORD_INPUT | TWO_CHANNELS, 2)])
def test_real_formats(mode, ch):
    kw = dict(input_mode=mode, rf_channels=ch, ad_speed=48000, fft1_n=10, mix1_red_n=3, version=2)
    _compare(kw, 9, [300.25], chunk=2)


def test_real_direction_reversed():
    kw = dict(input_mode=0, rf_channels=1, ad_speed=48000, fft1_n=10, mix1_red_n=3, version=2)
    _compare(kw, 8, [400.3], chunk=8, direction=-1)


@pytest.mark.parametrize("direction", [1, -1])
def test_real_limited_display_range(direction):
    kw = dict(input_mode=0, rf_channels=1, ad_speed=48000, fft1_n=10, mix1_red_n=3, version=2)
    _compare(kw, 8, [500.6], chunk=3, first_xpoint=200, xpoints=600, direction=direction)


def test_real_large_two_channel_int32():
    kw = dict(input_mode=DWORD_INPUT | TWO_CHANNELS, rf_channels=2, ad_speed=2400000, fft1_n=15, mix1_red_n=5, version=2)
    _compare(kw, 6, [9000.5], chunk=4)
