"""The transform half of make_fft3_all (fft3.c:215-470; SURVEY 8(f) rank 4, the step right after timf3): a plan
created with LB200_FLOAT_INPUT takes Linrad's timf3_float ring as its input ring and fft3 as its output ring;
lb200_fft1 without filter correction is then the window + DIF transform + permute of make_fft3_all.  Checked against
the reference's own fft3.c compiled into oracle/_ref (tables as baseb_graph.c:3679-3680 builds them), one and two
channels, several sizes and windows, ring wrap, single and multi-block calls.  Tolerance: relative RMS <= 1e-5 as
for fft1_float (measured ~2e-7)."""
import numpy as np
import pytest

from linrad_b200 import api, sizing
from oracle import refwrap
from tests.helpers import rel_rms, IQ_DATA, DWORD_INPUT, TWO_CHANNELS

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not refwrap.available(), reason="oracle/_ref not built")]


def _case(n, ch, sinpow, nblocks=3, seed=2):
    from oracle.refwrap import RefOracle
    two = TWO_CHANNELS if ch == 2 else 0
    r = RefOracle(input_mode=IQ_DATA | two, rf_channels=ch, ad_speed=96000, fft1_n=10, fft1_version=6 if ch == 1 else 7, n_sel=0)
    N = 1 << n
    ring_floats = 8 * N * 2 * ch
    r.fft3_setup(n, sinpow, ring_floats)
    s = sizing.PathSetup(input_mode=IQ_DATA | DWORD_INPUT | sizing.FLOAT_INPUT | two, rf_channels=ch, ad_speed=96000, fft1_n=n,
                         mix1_red_n=3, sinpow=sinpow)
    assert s.frame_bytes == 8 * ch
    # the same table the reference built (make_window(1, ...) in natural order)
    w1 = r.fft3_window()
    wnat = np.zeros(N, np.float32)
    wnat[: N // 2], wnat[N // 2:] = w1[0::2], w1[1::2]
    assert np.allclose(wnat, s.window, rtol=1e-6, atol=1e-7)
    rng = np.random.default_rng(seed)
    ring = (rng.standard_normal(ring_floats) * 3000.0).astype(np.float32)
    step = 2 * ch * s.fft1_new_points                  # fft3.c:768: timf3_px advances by 2*fft3_new_points*channels
    px0 = ring_floats - 3 * N * ch                     # the second block wraps the ring
    want = np.stack([r.make_fft3(ring, (px0 + b * step) % ring_floats) for b in range(nblocks)])
    plan = api.Plan(s)
    try:
        out = np.zeros(8 * s.fft1_block, np.float32)
        # lb200_fft1 transforms the N frames that start fft1_interleave_points frames before `ref`
        ref_bytes = (px0 * 4 + s.fft1_interleave_points * s.frame_bytes) % (ring_floats * 4)
        plan.fft1_host(timf1=ring.view(np.uint8), ref=ref_bytes, nblocks=nblocks, fft1=out, fft1_pa=0, apply_fc=False)
        got = out[: nblocks * s.fft1_block].reshape(nblocks, -1)
        e = rel_rms(got, want)
        assert e <= 1e-5, e
        # block by block gives the same bits as one call
        out1 = np.zeros_like(out)
        for b in range(nblocks):
            plan.fft1_host(timf1=ring.view(np.uint8), ref=(ref_bytes + b * s.timf1_blockbytes) % (ring_floats * 4), nblocks=1,
                           fft1=out1, fft1_pa=b * s.fft1_block, apply_fc=False)
        assert np.array_equal(out1[: nblocks * s.fft1_block], out[: nblocks * s.fft1_block])
    finally:
        plan.close()
    return e


@pytest.mark.parametrize("n", [7, 9, 10, 12, 14])
def test_fft3_one_channel_sizes(n):
    _case(n, 1, 2)


@pytest.mark.parametrize("sinpow", [1, 3, 4])          # make_window leaves the table empty for 0: not a third-FFT setting
def test_fft3_window_kinds(sinpow):
    _case(10, 1, sinpow)


@pytest.mark.parametrize("n,sinpow", [(8, 2), (11, 2), (13, 3)])
def test_fft3_two_channels(n, sinpow):
    _case(n, 2, sinpow)


def test_float_input_needs_iq_dword_and_a_single_cta_size():
    s = sizing.PathSetup(input_mode=IQ_DATA | DWORD_INPUT | sizing.FLOAT_INPUT, rf_channels=1, ad_speed=96000, fft1_n=15, mix1_red_n=5)
    with pytest.raises(api.Lb200Error):
        api.Plan(s)
