"""The drop-in boundary, exercised the way Linrad would: oracle/_ref/libref_shim.so is the compiled
reference (its own globals, tables, ring indices, fft1_waterfall, update_fft1_slowsum ...) with the
three hot-path calls -- fft1_b, fft1_c, fft1_mix1_fixed -- replaced by linrad_b200/host/lb200_shim.c,
which forwards to liblinrad_b200.so.  Same input through the untouched reference
(libref_oracle.so): spectra, power sums, baseband and the mixer's state must agree."""
import numpy as np
import pytest

from linrad_b200.synth import make_timf1
from oracle import refwrap
from tests.helpers import parity_record, rel_rms, run_reference, IQ_DATA, DWORD_INPUT, TWO_CHANNELS

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not (refwrap.available() and refwrap.shim_available()), reason="oracle/_ref not built")]


@pytest.mark.parametrize("mode,ch,ver,n,sinpow", [
    (IQ_DATA, 1, 6, 11, 2),                                   # radix-4 DIT tables (window layout mo=4)
    (IQ_DATA, 1, 7, 10, 2),                                   # radix-2 DIF tables (interleaved window, mo=1)
    (IQ_DATA | TWO_CHANNELS | DWORD_INPUT, 2, 7, 10, 2),
    (IQ_DATA | TWO_CHANNELS, 2, 6, 10, 3),                    # crossover-window mixer
    (0, 1, 2, 10, 2),                                         # real input, fft1_re.c
])
def test_shim_inside_the_reference(mode, ch, ver, n, sinpow):
    kw = dict(input_mode=mode, rf_channels=ch, ad_speed=96000, fft1_n=n, mix1_red_n=3, version=ver, sinpow=sinpow)
    N = 1 << n
    nblocks = 12
    P = None
    a = run_reference(kw, np.zeros(16, np.int16), [], 0)      # sizes only
    P = a["ref"].lib.ref_new_points()
    raw = make_timf1(mode, ch, N, nblocks, P, seed=7)
    sel = [N * 0.146 + 0.37]
    ref = run_reference(kw, raw, sel, nblocks)
    got = run_reference(kw, raw, sel, nblocks, through_shim=True)
    assert got["ref"].lib.ref_uses_shim() == 1 and ref["ref"].lib.ref_uses_shim() == 0
    assert rel_rms(got["fft1"], ref["fft1"]) <= 1e-5
    rows = nblocks // 5
    a, b = got["sumsq"][: rows * N].astype(np.float64), ref["sumsq"][: rows * N].astype(np.float64)
    strong = b > 1e-4 * b.max()
    assert (np.abs(a - b)[strong] <= 1e-4 * b[strong]).all()
    assert (got["sumsq_pa"], got["sumsq_counter"]) == (ref["sumsq_pa"], ref["sumsq_counter"])
    assert got["timf3_pa"] == ref["timf3_pa"]
    sg, sr = got["states"][0], ref["states"][0]
    assert sg["point"] == sr["point"] and float(sg["phase"]) == float(sr["phase"])
    e3_all, e3_rest = rel_rms(got["timf3"][:, 0], ref["timf3"][:, 0]), rel_rms(got["timf3"][1:, 0], ref["timf3"][1:, 0])
    parity_record(kind="shim_timf3", fft1_n=n, mode=mode, timf3_rel_rms_all_blocks=e3_all, timf3_rel_rms_without_block0=e3_rest,
                  fft1_rel_rms=rel_rms(got["fft1"], ref["fft1"]))
    # all blocks, the first one included; measured 0.4e-5 ... 2.3e-5 (profiles/r2_parity_report.jsonl).  The reference's own CUDA
    # row against its CPU row on the same input: 3.3e-5 (tests/test_reference_cufft_gpu.py)
    assert e3_all <= 3e-5 and e3_rest <= 3e-5, (e3_all, e3_rest)
    # the reference's own consumers ran on the shim's output: slowsum / waterfall stay consistent
    assert np.allclose(got["ref"].slowsum(), ref["ref"].slowsum(), rtol=2e-4, atol=1e-3 * float(np.abs(ref["ref"].slowsum()).max()))


def test_shim_correlation_spectrum():
    """fft1_correlation_flag == 1 inside the reference: the shim's fft1_c fills fft1_corrsum from the
    library's per-transform cross-spectrum rows, the reference's own update_fft1_slowsum turns it
    into fft1_slowcorr / fft1_slowcorr_tot"""
    mode, n = IQ_DATA | TWO_CHANNELS, 10
    kw = dict(input_mode=mode, rf_channels=2, ad_speed=96000, fft1_n=n, mix1_red_n=3, version=7)
    N = 1 << n
    nblocks = 27
    P = run_reference(kw, np.zeros(16, np.int16), [], 0)["ref"].lib.ref_new_points()
    raw = make_timf1(mode, 2, N, nblocks, P, seed=8)
    ref = run_reference(kw, raw, [], nblocks, correlation=1)
    got = run_reference(kw, raw, [], nblocks, through_shim=True, correlation=1)
    rows = nblocks // 5
    a, b = got["ref"].corrsum()[: 2 * rows * N], ref["ref"].corrsum()[: 2 * rows * N]
    scale = float(np.abs(b).max())
    assert scale > 0 and np.abs(a - b).max() <= 2e-5 * scale
    sa, sb = got["ref"].slowcorr(), ref["ref"].slowcorr()
    assert np.abs(sa - sb).max() <= 2e-5 * float(np.abs(sb).max())
    ta, tb = got["ref"].slowcorr_tot(), ref["ref"].slowcorr_tot()
    assert np.abs(ta - tb).max() <= 2e-5 * float(np.abs(tb).max())
    assert got["ref"].lib.ref_slowcorr_tot_avgnum() == ref["ref"].lib.ref_slowcorr_tot_avgnum() > 0


@pytest.mark.parametrize("mode,ch,ver", [(IQ_DATA, 1, 6), (IQ_DATA | TWO_CHANNELS, 2, 7)])
def test_shim_afc_mode_fft1_c(mode, ch, ver):
    """fft1afc_flag > 0 (AFC run from fft1, no spur elimination): fft1_c also keeps the per-transform
    powers -- fft1_power for one channel, fft1_xypower {x2, y2, im_xy, re_xy} for two
    (fft1.c:4203-4426).  The shim fills them from the library's power / xypower rows."""
    n = 10
    kw = dict(input_mode=mode, rf_channels=ch, ad_speed=96000, fft1_n=n, mix1_red_n=3, version=ver)
    N = 1 << n
    nblocks = 11
    P = run_reference(kw, np.zeros(16, np.int16), [], 0)["ref"].lib.ref_new_points()
    raw = make_timf1(mode, ch, N, nblocks, P, seed=9)
    ext = dict(afc=1, correlation=1 if ch == 2 else 0)
    ref = run_reference(kw, raw, [], nblocks, **ext)
    got = run_reference(kw, raw, [], nblocks, through_shim=True, **ext)
    assert rel_rms(got["fft1"], ref["fft1"]) <= 1e-5
    rows = nblocks // 5
    a, b = got["sumsq"][: rows * N].astype(np.float64), ref["sumsq"][: rows * N].astype(np.float64)
    strong = b > 1e-4 * b.max()
    assert (np.abs(a - b)[strong] <= 1e-4 * b[strong]).all()
    if ch == 1:
        pa, pb = got["ref"].fft1_power(), ref["ref"].fft1_power()
    else:
        pa, pb = got["ref"].fft1_xypower(), ref["ref"].fft1_xypower()
        ca, cb = got["ref"].corrsum()[: 2 * rows * N], ref["ref"].corrsum()[: 2 * rows * N]
        assert np.abs(ca - cb).max() <= 2e-5 * float(np.abs(cb).max())
    assert float(np.abs(pb).max()) > 0
    assert np.abs(pa - pb).max() <= 2e-5 * float(np.abs(pb).max())


@pytest.mark.parametrize("mode,ch,ver,sinpow", [(IQ_DATA, 1, 6, 2), (IQ_DATA | TWO_CHANNELS, 2, 7, 2), (IQ_DATA, 1, 7, 3)])
def test_shim_mix1_afc(mode, ch, ver, sinpow):
    """fft1_mix1_afc (mix1.c:1044-1096): the mixer frequency of every transform comes from the AFC
    track mix1_fq_mid[], which do_mix1_afc keeps bending (mix1.c:648-768).  The harness feeds both
    sides the same synthetic track; the shim mixes on the GPU and leaves the table maintenance to
    the reference's own code (here through a stand-in for the split of do_mix1_afc that
    lb200_shim.c asks of mix1.c)."""
    n = 10
    kw = dict(input_mode=mode, rf_channels=ch, ad_speed=96000, fft1_n=n, mix1_red_n=3, version=ver, sinpow=sinpow)
    N = 1 << n
    nblocks = 17
    P = run_reference(kw, np.zeros(16, np.int16), [], 0)["ref"].lib.ref_new_points()
    raw = make_timf1(mode, ch, N, nblocks, P, seed=10)
    sel = [N * 0.146 + 0.37]
    ref = run_reference(kw, raw, sel, nblocks, afc_mix=1)
    got = run_reference(kw, raw, sel, nblocks, through_shim=True, afc_mix=1)
    fixed = run_reference(kw, raw, sel, nblocks)
    assert np.abs(ref["timf3"] - fixed["timf3"]).max() > 1e-3 * np.abs(fixed["timf3"]).max()     # the track really moves the mixer
    sg, sr = got["states"][0], ref["states"][0]
    assert sg["point"] == sr["point"] and float(sg["phase"]) == float(sr["phase"])
    assert float(sg["phase_rot"]) == float(sr["phase_rot"]) and float(sg["phase_step"]) == float(sr["phase_step"])
    e3_all, e3_rest = rel_rms(got["timf3"][:, 0], ref["timf3"][:, 0]), rel_rms(got["timf3"][1:, 0], ref["timf3"][1:, 0])
    parity_record(kind="shim_timf3", fft1_n=n, mode=mode, timf3_rel_rms_all_blocks=e3_all, timf3_rel_rms_without_block0=e3_rest,
                  fft1_rel_rms=rel_rms(got["fft1"], ref["fft1"]))
    # all blocks, the first one included; measured 0.4e-5 ... 2.3e-5 (profiles/r2_parity_report.jsonl).  The reference's own CUDA
    # row against its CPU row on the same input: 3.3e-5 (tests/test_reference_cufft_gpu.py)
    assert e3_all <= 3e-5 and e3_rest <= 3e-5, (e3_all, e3_rest)
    assert got["timf3_pa"] == ref["timf3_pa"]


@pytest.mark.parametrize("sinpow,ch,mode,ver", [(2, 1, IQ_DATA, 6), (3, 1, IQ_DATA, 6), (2, 2, IQ_DATA | TWO_CHANNELS, 7)])
def test_shim_make_timf2(sinpow, ch, mode, ver):
    """make_timf2 inside the compiled reference through lb200_shim_make_timf2: timf2_float, timf2_pwr_float,
    timf2_pa, fft1_lowlevel_points and fft1_lowlevel_fraction against the reference's own timf2.c"""
    from oracle.refwrap import RefOracle
    n, nblocks = 10, 6
    N = 1 << n
    kw = dict(input_mode=mode, rf_channels=ch, ad_speed=96000, fft1_n=n, mix1_red_n=3, sinpow=sinpow)
    rng = np.random.default_rng(5)
    liminfo = np.zeros(N, np.float32)
    for _ in range(5):
        a = int(rng.integers(0, N - 30))
        liminfo[a: a + int(rng.integers(1, 30))] = float(rng.uniform(0.01, 1.0))
    res = []
    for shim in (False, True):
        r = RefOracle(fft1_version=ver, n_sel=0, max_fft1n=8, through_shim=shim, **kw)
        P = r.lib.ref_new_points()
        raw = make_timf1(mode, ch, N, nblocks, P, seed=7)
        rawb = np.ascontiguousarray(raw).view(np.uint8).reshape(-1)[: nblocks * r.timf1_blockbytes]
        r.process(rawb)
        r.timf2_setup(1, 1 << 14)
        res.append(r.make_timf2(liminfo, 0, nblocks))
    ref, got = res
    assert got["timf2_pa"] == ref["timf2_pa"] and got["lowlevel_points"] == ref["lowlevel_points"]
    assert got["lowlevel_fraction"] == ref["lowlevel_fraction"]
    assert rel_rms(got["timf2"], ref["timf2"]) <= 1e-5
    assert np.abs(got["pwr"].astype(np.float64) - ref["pwr"]).max() <= 2e-5 * np.abs(ref["pwr"]).max()
    assert np.array_equal(got["timf2"] == 0, ref["timf2"] == 0)



@pytest.mark.parametrize("n,ch,sinpow,new_points", [(10, 1, 2, 0), (12, 1, 3, 1536), (9, 2, 2, 0)])
def test_shim_fft3_transforms(n, ch, sinpow, new_points):
    """the transform half of make_fft3_all inside the compiled reference through lb200_shim_fft3_transforms (a second
    plan on the timf3 ring, LB200_FLOAT_INPUT) against the reference's own make_fft3_all on the same ring"""
    from oracle.refwrap import RefOracle
    mode = IQ_DATA | (TWO_CHANNELS if ch == 2 else 0)
    kw = dict(input_mode=mode, rf_channels=ch, ad_speed=96000, fft1_n=10, fft1_version=6 if ch == 1 else 7, n_sel=0)
    N = 1 << n
    ring_floats = 8 * N * 2 * ch
    rng = np.random.default_rng(n)
    ring = (rng.standard_normal(ring_floats) * 2000.0).astype(np.float32)
    a = RefOracle(**kw)
    a.fft3_setup(n, sinpow, ring_floats, new_points)
    b = RefOracle(through_shim=True, **kw)
    b.fft3_setup(n, sinpow, ring_floats, new_points)
    for px in (0, 2 * ch * 100, ring_floats - 2 * ch * (N // 2)):          # the last one wraps
        want = a.make_fft3(ring, px)
        got = b.make_fft3(ring, px)
        assert rel_rms(got, want) <= 1e-5
