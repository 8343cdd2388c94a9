"""Parity of the wide-graph consumers (update_fft1_slowsum + new_fft1_averages, fft1_waterfall +
update_wg_waterf) and of the input codecs with the reference / the oracle, through the C ABI.

The consumers are checked in isolation: the reference's OWN fft1_sumsq ring is the input of
both sides, so the only float differences left are the reference's -ffast-math reassociations
(slowsum: 1e-6 relative) and a last-bit log10 (waterfall: +-1 count of 0.1 dB on a handful of
pixels).  Scalar state (line pointer, counters, recalc window) must be identical."""
import ctypes as C

import numpy as np
import pytest

from linrad_b200 import api, sizing
from linrad_b200.synth import make_timf1
from oracle import port, refwrap
from tests.helpers import run_reference, IQ_DATA

pytestmark = [pytest.mark.gpu]


def _wg_case(xpp, ppx, first_xpoint=0, xpoints=None, nblocks=47, change_flag=0, chunks=(3, 1, 5)):
    kw = dict(input_mode=IQ_DATA, rf_channels=1, ad_speed=96000, fft1_n=10, mix1_red_n=3, version=6)
    extra = dict(xpoints_per_pixel=xpp, pixels_per_xpoint=ppx, wf_lines=6, first_xpoint=first_xpoint)
    if xpoints is not None:
        extra["xpoints"] = xpoints
    k = {a: b for a, b in kw.items() if a != "version"}
    s = sizing.PathSetup(**k, first_xpoint=first_xpoint, xpoints=xpoints if xpoints is not None else -1)
    raw = make_timf1(s.input_mode, 1, s.fft1_size, nblocks, s.fft1_new_points, seed=11)
    N = s.fft1_size
    # reference: block by block, so that intermediate states can be sampled at row boundaries
    from oracle.refwrap import RefOracle
    r = RefOracle(fft1_version=6, n_sel=0, max_fft1n=8, **k, **extra)
    if change_flag:
        r.set_change_fft1_flag(1)
    st0 = r.wg_state()
    wgc = api.WgConfig(s.avg2num, s.waterfall_avgnum, first_xpoint, s.xpoints, st0["wg_first_point"], st0["wg_last_point"],
                       r.wg_xpixels, xpp, ppx, st0["first_fft_bandwidth"])
    rawb = np.ascontiguousarray(raw).view(np.uint8).reshape(-1)
    r.process(rawb[: nblocks * r.timf1_blockbytes])
    ref_sumsq = r.sumsq()
    rows_total = nblocks // s.avg1num
    assert rows_total <= 16 - s.avg2num - 1
    # our side: same sumsq ring as input, rows handed over in uneven chunks
    plan = api.Plan(s)
    try:
        state = api.WgState(0, st0["fft1_sumsq_recalc"], change_flag, 0, 0, st0["latest_wg_spectrum"])
        slowsum = np.zeros(N, np.float32)
        wsum = np.full(N, 0.00001, np.float32)
        yfac = r.waterf_yfac()
        wsize = r.lib.ref_waterf_size()
        waterf = np.full(wsize + r.wg_xpixels + 64, -32768, np.int16)
        done = 0
        ci = 0
        while done < rows_total:
            n = min(chunks[ci % len(chunks)], rows_total - done)
            ci += 1
            api.wide_graph_host(plan, wgc, state, sumsq=ref_sumsq, sumsq_pa=done * N, nrows=n, slowsum=slowsum, wsum=wsum,
                                yfac=yfac, waterf=waterf, waterf_size=wsize)
            done += n
        st1 = r.wg_state()
        assert state.fft1_sumsq_recalc == st1["fft1_sumsq_recalc"]
        assert state.wg_waterf_ptr == st1["wg_waterf_ptr"]
        assert state.wg_waterf_sum_counter == st1["wg_waterf_sum_counter"]
        assert state.fft1_sumsq_pwg == st1["fft1_sumsq_pwg"]
        assert state.latest_wg_spectrum == st1["latest_wg_spectrum"]
        lo, hi = s.fft1_first_point, s.fft1_last_point
        a, b = slowsum[lo:hi + 1].astype(np.float64), r.slowsum()[lo:hi + 1].astype(np.float64)
        assert np.abs(a - b).max() <= 2e-6 * np.abs(b).max(), np.abs(a - b).max() / np.abs(b).max()
        assert np.all(np.abs(a - b) <= 1e-5 * np.abs(b) + 1e-6 * np.abs(b).mean())
        wa, wb = wsum.astype(np.float64), r.waterf_sum().astype(np.float64)
        assert np.all(np.abs(wa - wb) <= 1e-6 * np.abs(wb) + 1e-12)
        got, want = waterf[:wsize].astype(np.int32), r.waterf().astype(np.int32)
        d = np.abs(got - want)
        assert d.max() <= 1, (d.max(), np.nonzero(d > 1)[0][:10], got[np.nonzero(d > 1)[0][:10]], want[np.nonzero(d > 1)[0][:10]])
        assert (d > 0).mean() <= 2e-3, (d > 0).mean()
        assert (want != -32768).any()          # at least one line was written
    finally:
        plan.close()


def test_wide_graph_one_to_one():
    _wg_case(1, 1)


def test_wide_graph_change_flag_full_recompute():
    _wg_case(1, 1, change_flag=1, chunks=(2, 7))


def test_wide_graph_limited_range():
    _wg_case(1, 1, first_xpoint=100, xpoints=600)


def test_wide_graph_max_of_group():
    _wg_case(4, 0, chunks=(4, 1))


def test_wide_graph_interpolated():
    _wg_case(0, 3, first_xpoint=100, xpoints=300)


def test_wide_graph_interpolated_full():
    _wg_case(0, 2, chunks=(9,))


@pytest.mark.parametrize("xpp,ppx,first_xpoint,xpoints,change_flag,waterfall_avgnum,odd_ring", [
    (1, 1, 0, None, 0, 10, 0), (1, 1, 100, 600, 1, 10, 0), (4, 0, 0, None, 0, 15, 0), (0, 3, 100, 300, 0, 10, 0),
    (0, 2, 0, None, 1, 5, 0), (1, 1, 0, None, 0, 3, 0), (0, 3, 100, 300, 0, 5, 0)])
def test_wide_graph_many_rows_in_one_call_equal_row_by_row(xpp, ppx, first_xpoint, xpoints, change_flag, waterfall_avgnum, odd_ring):
    """A batch call hands over hundreds of rows at once (bench: 1184 at configs[0]): the kernels then skip to the
    last from-scratch recalculation of a bin (slowsum) and write all waterfall lines of the call side by side.  That
    has to be the sequential walk bit for bit: 57 rows in one call, in uneven chunks and one at a time, on a random
    ring, for the three pixel mappings, with and without the change flag and with lines every 1, 2 and 3 rows.
    (The waterfall ring is a whole number of lines, wide_graph.c:1374; anything else is refused.)"""
    s = sizing.PathSetup(input_mode=IQ_DATA, rf_channels=1, ad_speed=96000, fft1_n=10, mix1_red_n=3,
                         first_xpoint=first_xpoint, xpoints=xpoints if xpoints is not None else -1)
    N = s.fft1_size
    rows_total, ring_rows = 57, 64
    rng = np.random.default_rng(3)
    ring = (rng.random(ring_rows * N, dtype=np.float32) * 1e6 + 1.0).astype(np.float32)
    wg_first = s.first_xpoint
    wg_last = min(s.first_xpoint + s.xpoints, N - 1)
    if xpp > 1:
        xpix = s.xpoints // xpp
    elif ppx > 1:
        xpix = s.xpoints * ppx
    else:
        xpix = min(s.xpoints, N - s.first_xpoint)
    wgc = api.WgConfig(s.avg2num, waterfall_avgnum, s.first_xpoint, s.xpoints, wg_first, wg_last, xpix, xpp, ppx, 100)
    yfac = (rng.random(N, dtype=np.float32) * 1e-3 + 1e-4).astype(np.float32)
    wsize = xpix * 5 + odd_ring                        # the line pointer wraps many times
    results = []
    plan = api.Plan(s)
    try:
        for chunks in ((rows_total,), (1,), (7, 2, 13)):
            state = api.WgState(0, s.fft1_first_point, change_flag, 0, 0, 0)
            slowsum = np.zeros(N, np.float32)
            wsum = np.full(N, 0.00001, np.float32)
            waterf = np.full(wsize + xpix + 64, -32768, np.int16)
            done = ci = 0
            while done < rows_total:
                n = min(chunks[ci % len(chunks)], rows_total - done)
                ci += 1
                api.wide_graph_host(plan, wgc, state, sumsq=ring, sumsq_pa=(done * N) % ring.size, nrows=n, slowsum=slowsum,
                                    wsum=wsum, yfac=yfac, waterf=waterf, waterf_size=wsize)
                done += n
            results.append((slowsum, wsum, waterf, (state.fft1_sumsq_recalc, state.wg_waterf_ptr, state.wg_waterf_sum_counter,
                                                    state.fft1_sumsq_pwg, state.latest_wg_spectrum, state.change_fft1_flag)))
    finally:
        plan.close()
    ref = results[1]                                   # row by row
    assert (ref[2] != -32768).sum() >= min(wsize, xpix)
    for got in (results[0], results[2]):
        assert got[3] == ref[3]
        assert np.array_equal(got[0].view(np.uint32), ref[0].view(np.uint32))
        assert np.array_equal(got[1].view(np.uint32), ref[1].view(np.uint32))
        assert np.array_equal(got[2], ref[2])


# ---------------------------------------------------------------------------------------------
def test_expand_rawdat_bit_exact():
    s = sizing.PathSetup(input_mode=IQ_DATA, rf_channels=1, ad_speed=96000, fft1_n=10, mix1_red_n=3)
    plan = api.Plan(s)
    try:
        rng = np.random.default_rng(5)
        for groups in (1, 255, 256, 257, 10007):
            packed = rng.integers(0, 256, 9 * groups, dtype=np.uint8)
            want = port.expand_rawdat(packed, 16 * groups)
            assert np.array_equal(want, port.expand_rawdat_numpy(packed, 16 * groups))
            got = api.expand_rawdat_host(plan, packed, 16 * groups)
            assert np.array_equal(got, want)
        # round trip through the reference's packing (getiq64.s:39-96): 18 significant bits survive
        words = (rng.integers(-2**31, 2**31, 4 * 4096, dtype=np.int64) & ~0x3fff).astype(np.int32)
        got = api.expand_rawdat_host(plan, port.compress_rawdat(words), words.nbytes)
        assert np.array_equal(got, (words.view(np.uint32) + np.uint32(0x2000)).view(np.int32))
    finally:
        plan.close()


def test_widen_24bit_bit_exact():
    s = sizing.PathSetup(input_mode=IQ_DATA, rf_channels=1, ad_speed=96000, fft1_n=10, mix1_red_n=3)
    plan = api.Plan(s)
    try:
        rng = np.random.default_rng(6)
        for n in (4, 1024, 4 * 3333):
            pcm = rng.integers(0, 256, 3 * n, dtype=np.uint8)
            assert np.array_equal(api.widen_24bit_host(plan, pcm), port.widen_24bit(pcm))
    finally:
        plan.close()


def test_widen_8bit_and_float_bit_exact():
    """the other two file formats of rx_file_input (rxin.c:1573-1583, 1624-1634) against the C restatement
    run on the host CPU (whose float -> int conversion is the reference's)"""
    s = sizing.PathSetup(input_mode=IQ_DATA, rf_channels=1, ad_speed=96000, fft1_n=10, mix1_red_n=3)
    plan = api.Plan(s)
    try:
        rng = np.random.default_rng(16)
        allb = np.arange(256, dtype=np.uint8)
        assert np.array_equal(api.widen_8bit_host(plan, allb), port.widen_8bit(allb))
        for n in (4, 4096, 4 * 3333):
            pcm = rng.integers(0, 256, n, dtype=np.uint8)
            assert np.array_equal(api.widen_8bit_host(plan, pcm), port.widen_8bit(pcm))
            z = rng.uniform(-1.2, 1.2, n).astype(np.float32)
            z[::97] = np.array([1.0, -1.0, np.nan, np.inf, -np.inf, 0.99999994, -0.99999994, 1e-12], np.float32)[np.arange(z[::97].size) % 8]
            assert np.array_equal(api.float_to_int32_host(plan, z), port.float_to_int32(z))
    finally:
        plan.close()


def test_playback_chain_18bit_to_spectrum():
    """the playback front end end to end: 18-bit packed .raw payload -> expand_rawdat (GPU) ->
    timf1 -> fft1; equals fft1 of the oracle-expanded samples bit for bit"""
    kw = dict(input_mode=IQ_DATA | sizing.DWORD_INPUT, rf_channels=1, ad_speed=96000, fft1_n=11, mix1_red_n=3)
    s = sizing.PathSetup(**kw)
    from tests.helpers import CudaStream
    nblocks = 6
    raw = make_timf1(s.input_mode, 1, s.fft1_size, nblocks, s.fft1_new_points, seed=8)
    packed = port.compress_rawdat(raw.reshape(-1))
    want_words = port.expand_rawdat(packed, raw.nbytes)
    cs1, cs2 = CudaStream(s, []), CudaStream(s, [])
    try:
        got_words = api.expand_rawdat_host(cs1.plan, packed, raw.nbytes)
        assert np.array_equal(got_words, want_words)
        a = cs1.process(got_words.reshape(raw.shape), nblocks, chunk=3, mix=False)["fft1"]
        b = cs2.process(want_words.reshape(raw.shape), nblocks, chunk=3, mix=False)["fft1"]
        assert np.array_equal(a, b)
    finally:
        cs1.close()
        cs2.close()
