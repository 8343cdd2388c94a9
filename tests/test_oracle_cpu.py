"""CPU suite (-m "not gpu"): pins the oracle, the host-side logic and the C-ABI surface.

  * oracle/port.py (numpy restatement) against the golden fixtures under tests/golden/ (outputs
    of the reference's own compiled C path, made by tests/golden/make_golden.py) and, when
    oracle/_ref is present, against the compiled reference directly;
  * linrad_b200/sizing.py tables against the reference's own make_window / clear_fft1_filtercorr;
  * the scalar host logic of the C ABI (lb200_set_mix1_phases, lb200_phase_advance,
    lb200_window_to_natural) -- these run without a GPU;
  * liblinrad_b200.so loads and exports every symbol include/linrad_b200.h declares;
  * the multi-GPU sharding helper on 2 gloo ranks.
"""
import ctypes as C
import glob
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from linrad_b200 import sizing, shard
from oracle import port, refwrap
from tests.helpers import rel_rms

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN = sorted(glob.glob(os.path.join(HERE, "golden", "*.npz")))
CASES = [g for g in GOLDEN if not g.endswith("tables.npz")]


def load_case(path):
    z = np.load(path)
    kw = {str(k): int(v) for k, v in zip(z["kw_keys"], z["kw_vals"])}
    return z, kw, sizing.PathSetup(**kw)


# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[:-4] for p in CASES])
def test_port_matches_golden(path):
    z, kw, s = load_case(path)
    nb = int(z["nblocks"])
    sel = list(z["selbins"])
    got = port.run_path(s, z["raw_input"], sel, nb)
    assert rel_rms(got["raw"], z["fft1_raw"]) <= 1e-6
    assert rel_rms(got["fft1"], z["fft1"]) <= 1e-6
    assert got["sumsq_pa"] == int(z["sumsq_pa"]) and got["sumsq_counter"] == int(z["sumsq_counter"])
    N = s.fft1_size
    rows = nb // s.avg1num
    assert rel_rms(got["sumsq"][: rows * N], z["sumsq"][: rows * N]) <= 1e-6
    for i, fb in enumerate(sel):
        st = got["states"][i]
        if fb < 0:
            assert np.all(got["timf3"][:, i] == 0) and np.all(z["timf3"][:, i] == 0)
            continue
        # bin selection and the float phase state are bit-exact
        assert st["point"] == int(z["sel_point"][i])
        assert np.float32(st["phase"]) == z["sel_phase"][i]
        assert np.float32(st["phase_rot"]) == z["sel_phase_rot"][i]
        assert np.float32(st["phase_step"]) == z["sel_phase_step"][i]
        assert rel_rms(got["timf3"][:, i], z["timf3"][:, i]) <= 3e-5
        assert rel_rms(got["timf3_ring"][i], z["timf3_ring"][i]) <= 3e-5


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[:-4] for p in CASES])
def test_sizing_tables_match_golden(path):
    """window / filtercorr / mix1 tables of linrad_b200.sizing are the reference's, bit for bit."""
    z, kw, s = load_case(path)
    if s.window is not None:
        if s.input_mode & sizing.IQ_DATA:
            mo = 4 if int(z["version"]) == 6 else 1
            nat = np.zeros(s.fft1_size, np.float32)
            _window_to_natural(mo, s.fft1_size, z["window"], nat)
        else:
            nat = np.zeros(2 * s.fft1_size, np.float32)
            _window_to_natural(2, s.fft1_size, z["window"], nat)
        assert np.array_equal(nat, s.window)
    assert np.array_equal(z["filtercorr"], s.filtercorr)
    assert np.array_equal(z["mix1_fqwin"], s.mix1_fqwin)
    assert int(z["mix1_crossover"]) == s.mix1_crossover_points
    if s.mix1_crossover_points:
        h = s.mix1_size // 2 + 1
        assert np.array_equal(z["mix1_window"][:h], s.mix1_window[:h])
        assert np.array_equal(z["mix1_cos2win"], s.mix1_cos2win)
        assert np.array_equal(z["mix1_sin2win"], s.mix1_sin2win)
    assert np.array_equal(z["waterf_yfac"], s.waterfall_yfac())


def test_make_window_tables():
    z = np.load(os.path.join(HERE, "golden", "tables.npz"))
    for key in z.files:
        if not key.startswith("win_"):
            continue
        mo, sz, n = (int(v) for v in key.split("_")[1:])
        ref = z[key]
        if mo == 1:
            nat = np.zeros(sz, np.float32)
            _window_to_natural(1, sz, ref, nat)
            assert np.array_equal(nat, sizing.make_window(4, sz, n)), key
        else:
            got = sizing.make_window(mo, sz, n)
            assert np.array_equal(got[: ref.size], ref), key
    # fftback convention: unnormalised sum_k x_k exp(-2 pi i n k / M)  (fft0.c:481-533)
    assert rel_rms(np.fft.fft(z["fftback_in"].astype(np.complex128)).view(np.float64),
                   z["fftback_out"].astype(np.complex128).view(np.float64)) <= 1e-6


# ------------------------------------------------------------------------------------------
# 18-bit .raw codec: getiq64.s:158-220 (asm only; pinned by hand-derived vectors + round trip)
def test_expand_rawdat_known_answers():
    packed = np.array([0x34, 0x12, 0xff, 0xff, 0x00, 0x80, 0xcd, 0xab, 0b11100100], np.uint8)
    out = port.expand_rawdat(packed, 16)
    # word i = (bytes 2i,2i+1) << 16 | 2-bit field i of byte 8 << 14, + 0x2000
    want = [(0x1234 << 16) | (0 << 14), (0xffff << 16) | (1 << 14), (0x8000 << 16) | (2 << 14), (0xabcd << 16) | (3 << 14)]
    want = (np.array(want, np.uint64) + 0x2000).astype(np.uint32).view(np.int32)
    assert np.array_equal(out, want)
    assert np.array_equal(port.expand_rawdat_numpy(packed, 16), want)


def test_rawdat_round_trip():
    rng = np.random.default_rng(7)
    words = rng.integers(-2**31, 2**31, 4096, dtype=np.int64).astype(np.int32)
    packed = port.compress_rawdat(words)
    assert packed.size == words.size // 4 * 9
    back = port.expand_rawdat(packed, words.nbytes)
    # the top 18 bits survive, bit 13 carries the half-LSB offset, the rest is zero
    assert np.array_equal(back.view(np.uint32) >> 14, words.view(np.uint32) >> 14)
    assert np.all((back.view(np.uint32) & 0x3fff) == 0x2000)
    assert np.array_equal(back, port.expand_rawdat_numpy(packed, words.nbytes))
    assert np.array_equal(port.compress_rawdat(back), packed)
    # empty and ragged input: whole 16-byte groups only
    assert port.expand_rawdat(packed[:0], 0).size == 0
    assert np.array_equal(port.expand_rawdat(packed, 40)[:8], back[:8])


def test_widen_8bit_known_answers():
    """rxin.c:1573-1583: (0,255) -> (-32640, 32640), hand-computed"""
    b = np.array([0, 1, 127, 128, 254, 255], np.uint8)
    assert list(port.widen_8bit(b)) == [-32640, -32384, -128, 128, 32384, 32640]
    allb = np.arange(256, dtype=np.uint8)
    assert np.array_equal(port.widen_8bit(allb), (allb.astype(np.int32) * 256 - 32640).astype(np.int16))


def test_float_to_int32_known_answers():
    """rxin.c:1624-1634: 0x7fffffff*z through float; full scale and beyond become 0x80000000"""
    z = np.array([0.0, 0.5, -0.5, 0.25, 1.0, -1.0, 2.0, -3.0, 1e-10, 0.99999994, np.nan, np.inf, -np.inf], np.float32)
    got = port.float_to_int32(z)
    want = [0, 2 ** 30, -2 ** 30, 2 ** 29, -2 ** 31, -2 ** 31, -2 ** 31, -2 ** 31, 0, 2147483520, -2 ** 31, -2 ** 31, -2 ** 31]
    assert list(got) == want


def test_widen_24bit():
    b = np.array([0x01, 0x02, 0x03, 0xff, 0xff, 0xff, 0x00, 0x00, 0x80], np.uint8)
    assert list(port.widen_24bit(b)) == [0x03020100, -256, -2**31]


# ------------------------------------------------------------------------------------------
# C ABI surface (no GPU needed)
def _lib():
    from linrad_b200 import api
    return api.load_library(), api


def _window_to_natural(mo, size, win, out):
    lib, _ = _lib()
    win = np.ascontiguousarray(win, np.float32)
    lib.lb200_window_to_natural(mo, size, win.ctypes.data, out.ctypes.data)


def test_library_exports_every_declared_symbol():
    lib, api = _lib()
    hdr = open(os.path.join(ROOT, "include", "linrad_b200.h")).read()
    declared = set(re.findall(r"\b(lb200_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no prototypes found"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/linrad_b200.h but not exported"
    assert declared == set(api.EXPORTS)
    assert lib.lb200_abi_version() == api.LB200_ABI_VERSION
    assert b"CUDA device" in lib.lb200_strerror(3100)


def test_struct_layouts_match_header():
    """ctypes mirrors vs the C compiler's view of include/linrad_b200.h"""
    _, api = _lib()
    src = r'''
#include <stdio.h>
#include <stddef.h>
#include "linrad_b200.h"
int main(void){
  printf("%zu %zu %zu %zu %zu\n", sizeof(lb200_config), sizeof(lb200_ring), sizeof(lb200_fft1_args), sizeof(lb200_mix1_state), sizeof(lb200_mix1_args));
  printf("%zu %zu %zu %zu %zu\n", sizeof(lb200_wg_config), sizeof(lb200_wg_state), sizeof(lb200_wg_args), offsetof(lb200_wg_args, wg_waterf_size), offsetof(lb200_wg_args, state));
  printf("%zu %zu %zu %zu\n", offsetof(lb200_config, fft1_window), offsetof(lb200_config, max_batch), offsetof(lb200_fft1_args, power_rows), offsetof(lb200_mix1_args, timf3_pa));
  return 0; }'''
    exe = os.path.join(HERE, "golden", "_abi_probe")
    r = subprocess.run(["gcc", "-x", "c", "-", "-I", os.path.join(ROOT, "include"), "-o", exe], input=src, text=True,
                       capture_output=True)
    assert r.returncode == 0, r.stderr
    try:
        out = subprocess.run([exe], capture_output=True, text=True).stdout.split()
    finally:
        os.remove(exe)
    sizes = [C.sizeof(api.Config), C.sizeof(api.Ring), C.sizeof(api.Fft1Args), C.sizeof(api.Mix1State), C.sizeof(api.Mix1Args)]
    offs = [api.Config.fft1_window.offset, api.Config.max_batch.offset, api.Fft1Args.power_rows.offset, api.Mix1Args.timf3_pa.offset]
    wg = [C.sizeof(api.WgConfig), C.sizeof(api.WgState), C.sizeof(api.WgArgs), api.WgArgs.wg_waterf_size.offset, api.WgArgs.state.offset]
    got = [int(v) for v in out]
    assert got[:5] == sizes and got[5:10] == wg and got[10:] == offs


def test_18bit_codec_port_equals_the_assembled_reference():
    """getiq64.s is NASM source and NASM is not in the image; oracle/nasm2gas.py rewrites it mechanically into GNU
    assembler syntax and gcc assembles it (oracle/_ref/getiq_check).  The C / numpy restatements that the GPU kernel is
    checked against must be that code bit for bit: expand_rawdat and both compress_rawdat variants, random data,
    group counts around the kernel's tile size."""
    from oracle import refwrap
    if not refwrap.getiq_available():
        pytest.skip("oracle/_ref/getiq_check not built")
    rng = np.random.default_rng(18)
    for groups in (1, 2, 255, 256, 257, 4099):
        packed = rng.integers(0, 256, 9 * groups, dtype=np.uint8)
        want = refwrap.ref_expand_rawdat(packed)
        assert np.array_equal(port.expand_rawdat(packed, 16 * groups), want)
        assert np.array_equal(port.expand_rawdat_numpy(packed, 16 * groups), want)
        words = rng.integers(-2**31, 2**31, 4 * groups, dtype=np.int64).astype(np.int32)
        assert np.array_equal(port.compress_rawdat(words), refwrap.ref_compress_rawdat(words))
        assert np.array_equal(port.compress_rawdat(words), refwrap.ref_compress_rawdat(words, net=True))
    # extreme words
    words = np.array([0x7fffffff, -0x80000000, -1, 0, 0x00003fff, 0x00004000, -0x4000, 0x12345678], np.int32)
    assert np.array_equal(port.compress_rawdat(words), refwrap.ref_compress_rawdat(words))


def test_prebuilt_shim_harness_was_compiled_against_the_current_header():
    """oracle/_ref/libref_shim.so travels prebuilt to the GPU box (the reference's sources do not): a header change without
    `make -C oracle` would leave its shim object with stale argument structures"""
    from oracle import refwrap
    if not refwrap.shim_available():
        pytest.skip("oracle/_ref/libref_shim.so not built")
    _, api = _lib()
    shim = C.CDLL(refwrap.SHIM_SO)
    want = [C.sizeof(api.Config), C.sizeof(api.Fft1Args), C.sizeof(api.Mix1Args), C.sizeof(api.Timf2Args)]
    assert [shim.lb200_shim_sizeof_args(i) for i in range(4)] == want


def test_create_fails_loudly_without_gpu():
    """no CPU fallback: without a CUDA device lb200_create returns LB200_ERR_NO_DEVICE"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    _, api = _lib()
    s = sizing.PathSetup(input_mode=sizing.IQ_DATA, rf_channels=1, ad_speed=96000, fft1_n=9, mix1_red_n=3)
    with pytest.raises(api.Lb200Error) as e:
        api.Plan(s)
    assert e.value.code == 3100


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[:-4] for p in CASES])
def test_host_set_mix1_phases_bit_exact(path):
    """lb200_set_mix1_phases + lb200_phase_advance walk the reference's per-selection state
    (mix1.c:781-861 and the running sum of do_mix1) bit-exactly over the fixture's blocks."""
    lib, api = _lib()
    z, kw, s = load_case(path)
    cfg, keep = api.make_config(s)
    for i, fb in enumerate(z["selbins"]):
        if fb < 0:
            continue
        hz = s.ad_speed / s.fft1_size / (1 if s.input_mode & sizing.IQ_DATA else 2)
        st = api.new_states([fb * hz])
        for _ in range(int(z["nblocks"])):
            rc = lib.lb200_set_mix1_phases(C.byref(cfg), st, C.c_float(st[0].mix1_selfreq))
            assert rc == 0
            st[0].mix1_phase = lib.lb200_phase_advance(st[0].mix1_phase, st[0].mix1_phase_rot, s.mix1_new_points if s.mix1_interleave_points else s.mix1_size)
        assert st[0].mix1_point == int(z["sel_point"][i])
        assert st[0].mix1_old_point == int(z["sel_old_point"][i])
        assert np.float32(st[0].mix1_phase) == z["sel_phase"][i]
        assert np.float32(st[0].mix1_phase_step) == z["sel_phase_step"][i]
        assert np.float32(st[0].mix1_phase_rot) == z["sel_phase_rot"][i]
        assert np.float32(st[0].mix1_old_phase) == z["sel_old_phase"][i]
    del keep


def test_set_mix1_phases_range_errors():
    lib, api = _lib()
    s = sizing.PathSetup(input_mode=sizing.IQ_DATA, rf_channels=1, ad_speed=96000, fft1_n=9, mix1_red_n=3)
    cfg, keep = api.make_config(s)
    st = api.new_states([0.0])
    assert lib.lb200_set_mix1_phases(C.byref(cfg), st, C.c_float(0.0)) == 1211          # mix1.c:787-791
    assert lib.lb200_set_mix1_phases(C.byref(cfg), st, C.c_float(1e9)) == 1212          # mix1.c:792-796
    del keep


def test_phase_advance_equals_running_float_sum():
    lib, _ = _lib()
    rng = np.random.default_rng(3)
    cases = [(0.0, 0.0, 100), (1.0, 2.0 ** -24, 1000), (3.0, -0.37, 5000), (-7.9, 0.011, 3000), (6.2831855, 1e-9, 50),
             (0.5, 0.0061359233, 4096), (1e-30, 1e-31, 77), (8388607.5, 0.5, 9), (-2.0, 2.0 ** -23, 5000)]
    for _ in range(300):
        cases.append((float(np.float32(rng.uniform(-10, 10))), float(np.float32(rng.uniform(-0.8, 0.8) * 10.0 ** rng.integers(-6, 1))),
                      int(rng.integers(0, 3000))))
    for x, d, n in cases:
        want, _ = port.phase_chain(x, d, n)
        got = np.float32(lib.lb200_phase_advance(x, d, n))
        assert got == want or (np.isnan(got) and np.isnan(want)), (x, d, n, got, want)


# ------------------------------------------------------------------------------------------
@pytest.mark.skipif(not refwrap.available(), reason="oracle/_ref not built (no /root/reference here)")
def test_port_matches_compiled_reference_other_windows():
    """windows the fixtures do not hold: sin^1, sin^4, Gaussian, erfc (crossover overlap of do_mix1)"""
    from linrad_b200.synth import make_timf1
    from tests.helpers import run_reference
    for sinpow in (1, 4, 8, 9):
        kw = dict(input_mode=sizing.IQ_DATA, rf_channels=1, ad_speed=96000, fft1_n=8, mix1_red_n=2, sinpow=sinpow)
        s = sizing.PathSetup(**kw)
        raw = make_timf1(s.input_mode, 1, s.fft1_size, 9, s.fft1_new_points, seed=sinpow)
        ref = run_reference(dict(kw, version=6), raw, [101.3], 9)
        got = port.run_path(s, raw, [101.3], 9)
        assert rel_rms(got["fft1"], ref["fft1"]) <= 1e-6
        assert rel_rms(got["timf3"], ref["timf3"]) <= 3e-5
        assert got["states"][0]["point"] == ref["states"][0]["point"]


@pytest.mark.skipif(not refwrap.available(), reason="oracle/_ref not built")
def test_port_fft1_b_options_match_compiled_reference():
    """the rarely used fft1_b / fft1_c options in the restatement: ui.sample_shift, CALIQ foldcorr
    with both directions and a display range that leaves fft1_first_sym_point > 1, channel-2
    phasing, the cross spectrum of fft1_correlation_flag == 1"""
    from linrad_b200.synth import make_timf1
    from tests.helpers import run_reference
    rng = np.random.default_rng(12)
    cases = [
        (dict(input_mode=sizing.IQ_DATA, rf_channels=1, fft1_n=8, version=6), dict(sample_shift=-3), {}),
        (dict(input_mode=sizing.IQ_DATA | sizing.DWORD_INPUT, rf_channels=1, fft1_n=8, version=7), dict(sample_shift=2), dict(direction=-1)),
        (dict(input_mode=sizing.IQ_DATA, rf_channels=1, fft1_n=9, version=6), dict(foldcorr=True), dict(direction=-1, first_xpoint=60, xpoints=300)),
        (dict(input_mode=sizing.IQ_DATA, rf_channels=1, fft1_n=8, version=7), {}, dict(direction=-1, first_xpoint=90, xpoints=120)),
        (dict(input_mode=sizing.IQ_DATA | sizing.TWO_CHANNELS, rf_channels=2, fft1_n=8, version=7), dict(foldcorr=True, pg_ch2=(0.8, -0.5), correlation=1), {}),
        (dict(input_mode=sizing.IQ_DATA | sizing.TWO_CHANNELS, rf_channels=2, fft1_n=8, version=7), dict(pg_ch2=(1.02, 0.2), correlation=1), dict(direction=-1, first_xpoint=20, xpoints=200)),
    ]
    for kw, ext, over in cases:
        kw = dict(kw, ad_speed=96000, mix1_red_n=2)
        N, ch = 1 << kw["fft1_n"], kw["rf_channels"]
        ext = dict(ext)
        if ext.get("foldcorr"):
            ext["foldcorr"] = (0.03 * rng.standard_normal(2 * ch * N)).astype(np.float32)
        s = sizing.PathSetup(**{k: v for k, v in kw.items() if k != "version"}, **over)
        raw = make_timf1(s.input_mode, ch, N, 7, s.fft1_new_points, seed=3)
        ref = run_reference(dict(kw, **over), raw, [], 7, want_raw=True, **ext)
        got = port.run_path(s, raw, [], 7, **ext)
        scale = float(np.sqrt((ref["raw"].astype(np.float64) ** 2).mean()))
        assert np.abs(got["raw"] - ref["raw"]).max() <= 2e-5 * scale * np.sqrt(N), (kw, ext.keys(), over)
        rows = 7 // s.avg1num
        assert rel_rms(got["sumsq"][: rows * N], ref["sumsq"][: rows * N]) <= 1e-5
        if ext.get("correlation") == 1:
            assert rel_rms(got["corrsum"][: 2 * rows * N], ref["ref"].corrsum()[: 2 * rows * N]) <= 1e-5


# ------------------------------------------------------------------------------------------
def test_stream_and_block_sharding():
    a = shard.stream_assignment(64, 8)
    assert [len(r) for r in a] == [8] * 8 and a[3].start == 24
    a = shard.stream_assignment(5, 2)
    assert [list(r) for r in a] == [[0, 1, 2], [3, 4]]
    assert [len(r) for r in shard.stream_assignment(1, 4)] == [1, 0, 0, 0]
    br = shard.block_ranges(23, 2, avg1num=5)
    assert (br[0].first, br[0].count, br[0].warmup) == (0, 15, 0)
    assert (br[1].first, br[1].count, br[1].warmup) == (15, 8, 1)
    assert sum(b.count for b in shard.block_ranges(7, 4, avg1num=3)) == 7


_GLOO_WORKER = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from linrad_b200 import sizing, shard
from linrad_b200.synth import make_timf1
from oracle import port
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
s = sizing.PathSetup(input_mode=sizing.IQ_DATA, rf_channels=1, ad_speed=96000, fft1_n=7, mix1_red_n=3)
nstreams, nb = 5, 10
mine = shard.stream_assignment(nstreams, world)[rank]
rows = np.zeros((nb // s.avg1num) * s.fft1_size)
for st in mine:
    raw = make_timf1(s.input_mode, 1, s.fft1_size, nb, s.fft1_new_points, seed=100 + st)
    rows += port.run_path(s, raw, [], nb)["sumsq"][: rows.size]
t = torch.from_numpy(rows.astype(np.float32))
shard.reduce_sumsq(t)
if rank == 0:
    np.save(sys.argv[2], t.numpy())
dist.destroy_process_group()
'''


def test_two_rank_gloo_power_reduce(tmp_path):
    """world_size 2 on gloo: streams dealt to ranks, averaged power all-reduced = single-process sum"""
    from linrad_b200.synth import make_timf1
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER)
    out = tmp_path / "rows.npy"
    port_no = 29500 + (os.getpid() % 2000)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port_no))
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT, str(out)], env=env))
    for p in procs:
        assert p.wait(timeout=300) == 0
    got = np.load(out)
    s = sizing.PathSetup(input_mode=sizing.IQ_DATA, rf_channels=1, ad_speed=96000, fft1_n=7, mix1_red_n=3)
    want = np.zeros_like(got, dtype=np.float64)
    for st in range(5):
        raw = make_timf1(s.input_mode, 1, s.fft1_size, 10, s.fft1_new_points, seed=100 + st)
        want += port.run_path(s, raw, [], 10)["sumsq"][: want.size]
    assert rel_rms(got, want) <= 1e-6


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference headers not mounted")
def test_shim_compiles_against_reference_headers():
    """linrad_b200/host/lb200_shim.c is the binding a Linrad maintainer adds:
    it must compile against Linrad's own headers with every identifier declared."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run(["gcc", "-fsyntax-only", "-Werror=implicit-function-declaration", "-w", "-DOSNUM=1", "-DCPU=1",
                        "-DIA64=1", "-DHAVE_CUFFT=0", "-DOPENCL_PRESENT=0", "-I/root/reference",
                        "-I" + os.path.join(root, "include"), os.path.join(root, "linrad_b200", "host", "lb200_shim.c")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
