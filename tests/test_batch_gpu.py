"""Bench-shaped batches on the GPU: thousands of transforms per call, so that every CTA of the
persistent kernels walks through several work items (the small parity cases never get past the
first one).  The oracle cannot run these sizes in seconds, so the checks are the size-independent
properties the path offers:
  * one call of B transforms == the same input fed in calls of 7 (different work decomposition,
    same arithmetic): fft1_float bit-exact, fft1_sumsq / timf3 to rounding;
  * a periodic input gives periodic spectra, bit for bit;
  * the first blocks agree with the compiled reference (oracle/_ref) within the parity tolerances.
"""
import numpy as np
import pytest

from linrad_b200 import api, sizing
from linrad_b200.synth import make_timf1
from oracle import refwrap
from tests.helpers import CONFIGS, rel_rms, run_reference, pow2_at_least, IQ_DATA

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not refwrap.available(), reason="oracle/_ref not built")]


def _rings(s, nblocks, nsel):
    timf1 = np.zeros(pow2_at_least((nblocks + 2) * s.timf1_blockbytes), np.uint8)
    fft1 = np.zeros(pow2_at_least(nblocks * s.fft1_block), np.float32)
    sumsq = np.zeros(pow2_at_least((nblocks // s.avg1num + 2) * s.fft1_size), np.float32)
    t3size = pow2_at_least((nblocks + 2) * s.timf3_block + 2 * s.rf_channels * s.mix1_size)
    timf3 = np.zeros(max(nsel, 1) * 2 * t3size, np.float32)
    return timf1, fft1, sumsq, timf3, t3size


def _run(s, rawb, nblocks, selbins, chunk):
    """Feed nblocks transforms in calls of `chunk`, with the reference's index bookkeeping."""
    timf1, fft1, sumsq, timf3, t3size = _rings(s, nblocks, len(selbins))
    timf1[: rawb.size] = rawb
    plan = api.Plan(s)
    hz = s.ad_speed / s.fft1_size / (1 if s.input_mode & IQ_DATA else 2)
    states = api.new_states([b * hz for b in selbins])
    try:
        done, pa, counter, t3pa = 0, 0, 0, 0
        while done < nblocks:
            nb = min(chunk, nblocks - done)
            plan.fft1_host(timf1=timf1, ref=done * s.timf1_blockbytes, nblocks=nb, fft1=fft1, fft1_pa=done * s.fft1_block,
                           apply_fc=True, sumsq=sumsq, sumsq_pa=pa, counter=counter)
            tot = counter + nb
            pa += (tot // s.avg1num) * s.fft1_size
            counter = tot % s.avg1num
            if selbins:
                plan.mix1_host(fft1=fft1, fft1_px=done * s.fft1_block, nblocks=nb, states=states, timf3=timf3,
                               timf3_floats=t3size, timf3_pa=t3pa)
                t3pa += nb * s.timf3_block
            done += nb
        plan.synchronize()
    finally:
        plan.close()
    rows = nblocks // s.avg1num
    t3 = [timf3[i * 2 * t3size: i * 2 * t3size + nblocks * s.timf3_block].copy() for i in range(len(selbins))]
    return fft1[: nblocks * s.fft1_block].reshape(nblocks, -1), sumsq[: rows * s.fft1_size].reshape(rows, -1), t3


def _check(name, nblocks, selbins, period, nref, over=None):
    kw = dict(CONFIGS[name])
    if over:
        kw.update(over)
    s = sizing.PathSetup(**{a: b for a, b in kw.items() if a != "version"})
    raw = make_timf1(s.input_mode, s.rf_channels, s.fft1_size, period, s.fft1_new_points, seed=11)
    one = np.ascontiguousarray(raw).view(np.uint8).reshape(-1)[: period * s.timf1_blockbytes]
    rawb = np.tile(one, (nblocks + period - 1) // period)[: nblocks * s.timf1_blockbytes]
    f_big, p_big, t_big = _run(s, rawb, nblocks, selbins, chunk=nblocks)
    f_small, p_small, t_small = _run(s, rawb, nblocks, selbins, chunk=7)
    # same arithmetic whatever the work decomposition
    assert np.array_equal(f_big, f_small), "fft1_float differs between one call and calls of 7"
    assert np.allclose(p_big, p_small, rtol=2e-6, atol=0), "fft1_sumsq differs between one call and calls of 7"
    for a, b in zip(t_big, t_small):
        assert rel_rms(a, b) <= 1e-6, "timf3 differs between one call and calls of 7"
    # periodic input -> periodic spectra (transform 0 sees the empty ring in its overlap half)
    assert np.array_equal(f_big[1: nblocks - period], f_big[1 + period:]), "spectra of a periodic input are not periodic"
    assert np.isfinite(f_big).all() and np.isfinite(p_big).all()
    # the head of the batch against the compiled reference
    ref = run_reference(kw, rawb[: nref * s.timf1_blockbytes], selbins, nref)
    e = rel_rms(f_big[:nref], ref["fft1"])
    assert e <= 1e-5, f"fft1_float rel rms {e}"
    rows = nref // s.avg1num
    N = s.fft1_size
    a = p_big[:rows].astype(np.float64)
    b = ref["sumsq"][: rows * N].reshape(rows, N).astype(np.float64)
    strong = b > 1e-4 * b.max()
    assert (np.abs(a - b)[strong] <= 1e-4 * b[strong]).all(), "fft1_sumsq of strong bins"
    for ss in range(len(selbins)):
        got = t_big[ss][: nref * s.timf3_block].reshape(nref, -1)
        e3 = rel_rms(got[1:], ref["timf3"][1:, ss])
        assert e3 <= 1e-4, f"timf3 sel {ss} rel rms {e3}"


def test_batch_cfg1_1600_transforms():
    # 320 averaging groups on 296 resident CTAs: the TMA-store kernel crosses work items
    _check("cfg1", 1600, [3000.37], period=16, nref=20)


def test_batch_cfg2_800_transforms():
    # 160 groups x 2 channels = 320 work items on 148 CTAs
    _check("cfg2", 800, [6000.74], period=16, nref=10)


def test_batch_four_step_200_transforms():
    # N = 2^15: two L2-sized sub-batches, (transform, tile) work items adding into shared rows
    _check("cfg1", 200, [], period=8, nref=10, over=dict(fft1_n=15, mix1_red_n=5))


def test_batch_real_input_300_transforms():
    _check("cfg3", 300, [], period=8, nref=10, over=dict(fft1_n=13))


def test_spectrum_stays_on_device_flag():
    """LB200_FFT1_SPECTRUM_STAYS_ON_DEVICE: the host fft1_float ring is not written, lb200_mix1 reads
    the transforms from the plan's device mirror; fft1_sumsq and timf3 are the same as without it."""
    kw = dict(CONFIGS["cfg2"])
    s = sizing.PathSetup(**{a: b for a, b in kw.items() if a != "version"})
    nblocks, sel = 23, [6000.74]
    raw = make_timf1(s.input_mode, s.rf_channels, s.fft1_size, nblocks, s.fft1_new_points, seed=3)
    rawb = np.ascontiguousarray(raw).view(np.uint8).reshape(-1)[: nblocks * s.timf1_blockbytes]
    hz = s.ad_speed / s.fft1_size
    res = []
    for keep in (False, True):
        timf1, fft1, sumsq, timf3, t3size = _rings(s, nblocks, 1)
        timf1[: rawb.size] = rawb
        plan = api.Plan(s)
        states = api.new_states([sel[0] * hz])
        try:
            done, pa, counter, t3pa = 0, 0, 0, 0
            for nb in (9, 14):
                plan.fft1_host(timf1=timf1, ref=done * s.timf1_blockbytes, nblocks=nb, fft1=fft1, fft1_pa=done * s.fft1_block,
                               sumsq=sumsq, sumsq_pa=pa, counter=counter, keep_on_device=keep)
                tot = counter + nb
                pa += (tot // s.avg1num) * s.fft1_size
                counter = tot % s.avg1num
                plan.mix1_host(fft1=fft1, fft1_px=done * s.fft1_block, nblocks=nb, states=states, timf3=timf3,
                               timf3_floats=t3size, timf3_pa=t3pa)
                t3pa += nb * s.timf3_block
                done += nb
            plan.synchronize()
            d2h = plan.d2h_bytes()
        finally:
            plan.close()
        res.append((fft1.copy(), sumsq.copy(), timf3.copy(), d2h))
    (f0, p0, t0, d0), (f1, p1, t1, d1) = res
    assert np.abs(f0).max() > 0 and not f1.any()                  # host ring untouched
    assert np.allclose(p0, p1, rtol=2e-6, atol=0)
    assert np.array_equal(t0, t1)
    assert d0 - d1 == nblocks * s.fft1_block * 4                  # exactly the spectrum stayed behind


@pytest.mark.parametrize("world", [2, 3])
def test_time_block_sharding_equals_one_pass(world):
    """SURVEY.md 8(e), secondary partitioning: one stream cut into per-rank time-block ranges on
    averaging-group borders (shard.block_ranges), each rank re-reading the overlap halo, stepping
    the mixer state to its start and running one warm-up transform for the mix1 seam.  Ranks are
    run one after the other on this GPU; together they must reproduce the one-pass result."""
    from linrad_b200 import shard
    kw = dict(CONFIGS["cfg1"], fft1_n=11, mix1_red_n=3)
    s = sizing.PathSetup(**{a: b for a, b in kw.items() if a != "version"})
    nblocks, sel = 38, [300.37, 1234.0]
    raw = make_timf1(s.input_mode, s.rf_channels, s.fft1_size, nblocks, s.fft1_new_points, seed=5)
    rawb = np.ascontiguousarray(raw).view(np.uint8).reshape(-1)[: nblocks * s.timf1_blockbytes]
    f_one, p_one, t_one = _run(s, rawb, nblocks, sel, chunk=nblocks)
    hz = s.ad_speed / s.fft1_size
    f_all = np.zeros_like(f_one)
    p_all = np.zeros_like(p_one)
    t_all = [np.zeros_like(t) for t in t_one]
    for br in shard.block_ranges(nblocks, world, avg1num=s.avg1num):
        if br.count == 0:
            continue
        timf1, fft1, sumsq, timf3, t3size = _rings(s, nblocks, len(sel))
        timf1[: rawb.size] = rawb                       # a rank reads only [first - warmup - halo, first + count)
        plan = api.Plan(s)
        try:
            states = api.new_states([b * hz for b in sel])
            start = br.first - br.warmup
            api.advance_mix1_states(plan, states, start)
            if br.warmup:                               # warm-up transform: spectrum only, its power belongs to the previous rank's row
                plan.fft1_host(timf1=timf1, ref=start * s.timf1_blockbytes, nblocks=br.warmup, fft1=fft1,
                               fft1_pa=start * s.fft1_block, apply_fc=True, sumsq=None)
            plan.fft1_host(timf1=timf1, ref=br.first * s.timf1_blockbytes, nblocks=br.count, fft1=fft1,
                           fft1_pa=br.first * s.fft1_block, apply_fc=True, sumsq=sumsq,
                           sumsq_pa=(br.first // s.avg1num) * s.fft1_size, counter=0)
            plan.mix1_host(fft1=fft1, fft1_px=start * s.fft1_block, nblocks=br.warmup + br.count, states=states,
                           timf3=timf3, timf3_floats=t3size, timf3_pa=start * s.timf3_block)
            plan.synchronize()
        finally:
            plan.close()
        sl = slice(br.first, br.first + br.count)
        f_all[sl] = fft1[: nblocks * s.fft1_block].reshape(nblocks, -1)[sl]
        r0, r1 = br.first // s.avg1num, min((br.first + br.count) // s.avg1num, p_one.shape[0])
        p_all[r0:r1] = sumsq[: p_one.size].reshape(p_one.shape)[r0:r1]
        for i in range(len(sel)):
            ring = timf3[i * 2 * t3size: i * 2 * t3size + nblocks * s.timf3_block].reshape(nblocks, -1)
            t_all[i].reshape(nblocks, -1)[sl] = ring[sl]
    assert np.array_equal(f_all, f_one)
    assert np.array_equal(p_all, p_one)
    for a, b in zip(t_all, t_one):
        assert rel_rms(a, b) <= 1e-6


def test_concurrent_fft1b_workers():
    """Linrad farms fft1_b out to up to six worker threads, each with its own handle, on distinct
    time blocks of the SAME timf1 / fft1_float rings (wcw.c:476-513, 974-1033): the entry point must
    be re-entrant across plans.  Three threads, one plan each, one transform per call, against one
    plan doing the same blocks in order."""
    import threading
    kw = dict(CONFIGS["cfg1"], fft1_n=12, mix1_red_n=4)
    s = sizing.PathSetup(**{a: b for a, b in kw.items() if a != "version"})
    nblocks, workers = 30, 3
    raw = make_timf1(s.input_mode, s.rf_channels, s.fft1_size, nblocks, s.fft1_new_points, seed=6)
    rawb = np.ascontiguousarray(raw).view(np.uint8).reshape(-1)[: nblocks * s.timf1_blockbytes]

    def run(nthreads):
        timf1, fft1, _, _, _ = _rings(s, nblocks, 0)
        timf1[: rawb.size] = rawb
        power = np.zeros((nblocks, s.fft1_size), np.float32)
        plans = [api.Plan(s) for _ in range(nthreads)]
        errors = []

        def work(k):
            try:
                for b in range(k, nblocks, nthreads):
                    plans[k].fft1_host(timf1=timf1, ref=b * s.timf1_blockbytes, nblocks=1, fft1=fft1,
                                       fft1_pa=b * s.fft1_block, apply_fc=True, power=power[b])
            except Exception as e:              # surfaced below: a worker must not die silently
                errors.append(e)
        try:
            th = [threading.Thread(target=work, args=(k,)) for k in range(nthreads)]
            for t in th:
                t.start()
            for t in th:
                t.join()
            for p in plans:
                p.synchronize()
        finally:
            for p in plans:
                p.close()
        assert not errors, errors
        return fft1[: nblocks * s.fft1_block].copy(), power

    f1, p1 = run(1)
    f3, p3 = run(workers)
    assert np.array_equal(f1, f3) and np.array_equal(p1, p3)
    assert np.abs(f1).max() > 0 and p1.min() >= 0
