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


def test_fft3_all_selections_in_one_call():
    """no_of_rings: the loop over ss of make_fft3_all as one launch (timf3 selections 2*timf3_size floats apart, their fft3
    blocks side by side) gives the same bits as one call per selection"""
    import torch
    n, ch, nsel, nblocks = 10, 1, 5, 7
    s = sizing.PathSetup(input_mode=IQ_DATA | DWORD_INPUT | sizing.FLOAT_INPUT, rf_channels=ch, ad_speed=96000, fft1_n=n,
                         mix1_red_n=3, sinpow=2)
    N = s.fft1_size
    timf3_size = 16 * N * 2 * ch
    rng = np.random.default_rng(4)
    host = (rng.standard_normal(nsel * 2 * timf3_size) * 1000.0).astype(np.float32)
    dev = torch.device("cuda", 0)
    timf3 = torch.from_numpy(host).to(dev)
    stride_out = nblocks * s.fft1_block
    out_floats = 1
    while out_floats < nsel * stride_out:
        out_floats *= 2
    a = torch.zeros(out_floats, dtype=torch.float32, device=dev)
    b = torch.zeros(out_floats, dtype=torch.float32, device=dev)
    plan = api.Plan(s)
    try:
        ref = s.fft1_interleave_points * s.frame_bytes + 4 * 2 * ch * 37
        plan.fft1_dev(timf1=timf3.data_ptr(), timf1_bytes=timf3_size * 4, ref=ref, nblocks=nblocks, fft1=a.data_ptr(),
                      fft1_floats=out_floats, fft1_pa=0, apply_fc=False, rings=nsel, ring_stride=2 * timf3_size * 4,
                      pa_stride=stride_out)
        for ss in range(nsel):
            plan.fft1_dev(timf1=timf3.data_ptr() + ss * 2 * timf3_size * 4, timf1_bytes=timf3_size * 4, ref=ref, nblocks=nblocks,
                          fft1=b.data_ptr(), fft1_floats=out_floats, fft1_pa=ss * stride_out, apply_fc=False)
        plan.synchronize()
        got, want = a.cpu().numpy(), b.cpu().numpy()
        assert np.abs(want[: nsel * stride_out]).max() > 0
        assert np.array_equal(got, want)
        # and a plan without float input refuses it
    finally:
        plan.close()
    s2 = sizing.PathSetup(input_mode=IQ_DATA | DWORD_INPUT, rf_channels=1, ad_speed=96000, fft1_n=n, mix1_red_n=3)
    plan2 = api.Plan(s2)
    try:
        with pytest.raises(api.Lb200Error):
            plan2.fft1_dev(timf1=timf3.data_ptr(), timf1_bytes=timf3_size * 4, ref=0, nblocks=1, fft1=a.data_ptr(), fft1_floats=out_floats,
                           fft1_pa=0, apply_fc=False, rings=2, ring_stride=2 * timf3_size * 4, pa_stride=stride_out)
    finally:
        plan2.close()
