"""make_timf2 (the front end of the second FFT, timf2.c:31-208) through the C ABI against the
reference's own timf2.c compiled into oracle/_ref: strong/weak split by liminfo, back transform,
fft1back_fp_finish in its three window cases, one and two channels, calls of several transforms and
the half parked in the ring between calls.  Tolerance: the split and the ring bookkeeping are exact
(zeros where the reference has zeros, same ring positions); the samples are a float32 transform of
different structure than the reference's radix-4 one: relative RMS <= 1e-5 like fft1_float."""
import numpy as np
import pytest

from linrad_b200 import api, sizing
from linrad_b200.synth import make_timf1
from oracle import refwrap
from tests.helpers import rel_rms, IQ_DATA, TWO_CHANNELS

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not refwrap.available(), reason="oracle/_ref not built")]


def _case(n, sinpow, mode, ch, version, att_n=0, chunks=(3, 2, 1), seed=3, first_xpoint=0, xpoints=-1):
    from oracle.refwrap import RefOracle
    kw = dict(input_mode=mode, rf_channels=ch, ad_speed=96000, fft1_n=n, mix1_red_n=3, sinpow=sinpow)
    s = sizing.PathSetup(**kw, first_xpoint=first_xpoint, xpoints=xpoints)
    extra = dict(first_xpoint=first_xpoint)
    if xpoints > 0:
        extra["xpoints"] = xpoints
    r = RefOracle(fft1_version=version, n_sel=0, max_fft1n=8, **kw, **extra)
    nblocks = sum(chunks)
    assert nblocks <= 8
    raw = make_timf1(mode, ch, s.fft1_size, nblocks, s.fft1_new_points, seed=seed)
    rawb = np.ascontiguousarray(raw).view(np.uint8).reshape(-1)[: nblocks * r.timf1_blockbytes]
    out = r.process(rawb)                                   # fft1_float blocks 0..nblocks-1 of the reference's ring
    N = s.fft1_size
    rng = np.random.default_rng(seed)
    liminfo = np.zeros(N, np.float32)                       # a few strong regions, as fft1_update_liminfo leaves them
    for _ in range(6):
        a = int(rng.integers(0, N - 40))
        liminfo[a: a + int(rng.integers(1, 40))] = float(rng.uniform(0.01, 1.0))
    liminfo[int(rng.integers(0, N))] = -1.0
    pow_size = 1 << 15
    r.timf2_setup(att_n, pow_size)
    # reference: the same call pattern (liminfo fixed, blocks in order)
    ref = r.make_timf2(liminfo, 0, nblocks)
    sf = 4 * ch
    timf2 = np.zeros(sf * pow_size, np.float32)
    pwr = np.full(pow_size, 0.5, np.float32)
    fft1 = np.zeros(8 * s.fft1_block, np.float32)
    fft1[: nblocks * s.fft1_block] = out["fft1"].reshape(-1)
    plan = api.Plan(s, inverted_window=r.inverted_window())
    try:
        done, pa = 0, 0
        for nb in chunks:
            low = api.make_timf2_host(plan, fft1=fft1, fft1_px=done * s.fft1_block, nblocks=nb, liminfo=liminfo, timf2=timf2,
                                      timf2_pwr=pwr, timf2_pa=pa, att_n=att_n)
            assert low == ref["lowlevel_points"]
            pa = (pa + nb * ref["input_block"]) & (timf2.size - 1)
            done += nb
        assert pa == ref["timf2_pa"]
    finally:
        plan.close()
    newp = s.fft1_new_points
    used = sf * (nblocks * newp + (N // 2 if s.fft1_interleave_points == N // 2 else 0))
    a, b = timf2[:used], ref["timf2"][:used]
    e = rel_rms(a, b)
    assert e <= 1e-5, f"timf2_float rel rms {e}"
    w = a.reshape(-1, sf)
    wr = b.reshape(-1, sf)
    # weak and strong halves separately (the strong one is small when few bins are strong)
    for lo, hi in ((0, 2 * ch), (2 * ch, 4 * ch)):
        if np.abs(wr[:, lo:hi]).max() > 0:
            assert rel_rms(w[:, lo:hi], wr[:, lo:hi]) <= 1e-5
    assert not timf2[used:].any() and not ref["timf2"][used:].any()
    ps = nblocks * newp
    pa_, pb_ = pwr[:ps].astype(np.float64), ref["pwr"][:ps].astype(np.float64)
    assert np.abs(pa_ - pb_).max() <= 2e-5 * np.abs(pb_).max() + 1e-30
    assert np.array_equal(pwr[ps:], ref["pwr"][ps:])        # untouched: still the initial 0.5
    return e


@pytest.mark.parametrize("n", [7, 9, 10, 11, 13, 14])
def test_timf2_sin2_window_sizes(n):
    _case(n, 2, IQ_DATA, 1, 6)


@pytest.mark.parametrize("sinpow", [0, 1, 3, 4])
def test_timf2_window_kinds(sinpow):
    _case(10, sinpow, IQ_DATA, 1, 6, chunks=(2, 3) if sinpow else (1, 2))


@pytest.mark.parametrize("sinpow", [2, 3, 0])
def test_timf2_two_channels(sinpow):
    _case(10, sinpow, IQ_DATA | TWO_CHANNELS, 2, 7, chunks=(2, 2) if sinpow else (1, 2))


def test_timf2_attenuation_and_limited_range():
    _case(11, 2, IQ_DATA, 1, 7, att_n=3, first_xpoint=300, xpoints=1200)


def test_timf2_one_call_equals_single_blocks():
    _case(10, 2, IQ_DATA, 1, 6, chunks=(1, 1, 1, 1, 1, 1))
    _case(10, 2, IQ_DATA, 1, 6, chunks=(6,))
