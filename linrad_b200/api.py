"""ctypes binding of liblinrad_b200.so (include/linrad_b200.h) for the test and bench harness.

The product is the C-ABI library; this module only marshals arguments.  It never computes
samples itself and there is no fallback: if the CUDA library is missing the import fails.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LB200_LIB") or os.path.join(_HERE, "liblinrad_b200.so")   # LB200_LIB: A/B builds of the same ABI

LB200_ABI_VERSION = 3
ERR = {0: "OK", 3100: "NO_DEVICE", 3101: "CUDA", 3102: "BAD_CONFIG", 3103: "UNSUPPORTED", 3104: "BAD_ARG",
       1211: "MIX1_RANGE_LOW", 1212: "MIX1_RANGE_HIGH"}

# every symbol include/linrad_b200.h declares
EXPORTS = ["lb200_create", "lb200_destroy", "lb200_strerror", "lb200_abi_version", "lb200_stream",
           "lb200_synchronize", "lb200_launch_count", "lb200_h2d_bytes", "lb200_d2h_bytes",
           "lb200_fft1_dev", "lb200_fft1", "lb200_mix1_dev", "lb200_mix1", "lb200_set_mix1_phases",
           "lb200_phase_advance", "lb200_window_to_natural",
           "lb200_update_fft1_slowsum_dev", "lb200_update_fft1_slowsum", "lb200_fft1_waterfall_dev",
           "lb200_fft1_waterfall", "lb200_expand_rawdat_dev", "lb200_expand_rawdat", "lb200_widen_24bit_dev",
           "lb200_widen_24bit", "lb200_raw_header_parse", "lb200_raw_block_bytes",
           "lb200_widen_8bit_dev", "lb200_widen_8bit", "lb200_float_to_int32_dev", "lb200_float_to_int32",
           "lb200_make_timf2_dev", "lb200_make_timf2",
           "lb200_reduce_create", "lb200_reduce_export", "lb200_reduce_connect", "lb200_reduce_push",
           "lb200_reduce_rows_released", "lb200_reduce_sum", "lb200_reduce_result_ready", "lb200_reduce_synchronize",
           "lb200_reduce_destroy"]


class Lb200Error(RuntimeError):
    def __init__(self, code, what=""):
        self.code = code
        super().__init__(f"lb200 error {code} ({ERR.get(code, '?')}) {what}")


class Config(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int), ("device", C.c_int),
        ("rx_input_mode", C.c_int), ("rx_rf_channels", C.c_int), ("sample_shift", C.c_int),
        ("fft1_n", C.c_int), ("fft1_interleave_points", C.c_int), ("fft1_direction", C.c_int),
        ("fft1_first_point", C.c_int), ("fft1_last_point", C.c_int),
        ("fft1_window", C.c_void_p), ("fft1_filtercorr", C.c_void_p), ("fft1_foldcorr", C.c_void_p),
        ("fft_avg1num", C.c_int),
        ("mix1_n", C.c_int), ("mix1_interleave_points", C.c_int), ("mix1_crossover_points", C.c_int),
        ("mix1_fqwin", C.c_void_p), ("mix1_window", C.c_void_p), ("mix1_cos2win", C.c_void_p),
        ("mix1_sin2win", C.c_void_p),
        ("fftx_points_per_hz", C.c_float), ("mix1_lowest_fq", C.c_float), ("mix1_highest_fq", C.c_float),
        ("max_batch", C.c_int),
        ("pg_ch2_c1", C.c_float), ("pg_ch2_c2", C.c_float),
        ("fft1_inverted_window", C.c_void_p),
    ]


class RawHeader(C.Structure):
    _fields_ = [
        ("remember_tag", C.c_int), ("chunk_size", C.c_int), ("chunk_offset", C.c_uint64),
        ("diskread_time", C.c_double), ("passband_center", C.c_double), ("passband_direction", C.c_int),
        ("rx_input_mode", C.c_int), ("rx_rf_channels", C.c_int), ("rx_ad_channels", C.c_int),
        ("rx_ad_speed", C.c_int), ("save_init_flag", C.c_int), ("payload_offset", C.c_uint64),
    ]


class Ring(C.Structure):
    _fields_ = [("base", C.c_void_p), ("size", C.c_size_t)]


class Fft1Args(C.Structure):
    _fields_ = [
        ("timf1", Ring), ("timf1p_ref", C.c_uint32), ("nblocks", C.c_int),
        ("fft1_float", Ring), ("fft1_pa", C.c_uint32), ("apply_filtercorr", C.c_int),
        ("fft1_sumsq", Ring), ("fft1_sumsq_pa", C.c_uint32), ("fft1_sumsq_counter", C.c_int),
        ("power_rows", C.c_void_p), ("flags", C.c_int),
        ("fft1_corrsum", Ring), ("corr_rows", C.c_void_p), ("xypower_rows", C.c_void_p),
        ("no_of_rings", C.c_int), ("timf1_ring_stride", C.c_size_t), ("fft1_pa_stride", C.c_uint32),
    ]


FFT1_SPECTRUM_STAYS_ON_DEVICE = 1


class Mix1State(C.Structure):
    _fields_ = [
        ("mix1_selfreq", C.c_double), ("mix1_phase", C.c_float), ("mix1_phase_step", C.c_float),
        ("mix1_phase_rot", C.c_float), ("mix1_old_phase", C.c_float),
        ("mix1_point", C.c_int), ("mix1_old_point", C.c_int),
    ]


class Mix1Args(C.Structure):
    _fields_ = [
        ("fft1_float", Ring), ("fft1_px", C.c_uint32), ("nblocks", C.c_int), ("no_of_channels", C.c_int),
        ("state", C.POINTER(Mix1State)), ("timf3_float", Ring), ("timf3_pa", C.c_uint32),
    ]


class Timf2Args(C.Structure):
    _fields_ = [
        ("fft1_float", Ring), ("fft1_px", C.c_uint32), ("nblocks", C.c_int), ("liminfo", C.c_void_p),
        ("timf2_float", Ring), ("timf2_pwr_float", C.c_void_p), ("timf2_pa", C.c_uint32),
        ("first_bckfft_att_n", C.c_int), ("fft1_lowlevel_points", C.POINTER(C.c_int)),
    ]


class WgConfig(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "wg_fft_avg2num", "waterfall_avgnum", "first_xpoint", "xpoints", "wg_first_point", "wg_last_point",
        "wg_xpixels", "xpoints_per_pixel", "pixels_per_xpoint", "first_fft_bandwidth")]


class WgState(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "fft1_sumsq_pwg", "fft1_sumsq_recalc", "change_fft1_flag", "wg_waterf_sum_counter", "wg_waterf_ptr",
        "latest_wg_spectrum")]


class WgArgs(C.Structure):
    _fields_ = [
        ("fft1_sumsq", Ring), ("fft1_sumsq_pa", C.c_uint32), ("nrows", C.c_int),
        ("fft1_slowsum", C.c_void_p), ("wg_waterf_sum", C.c_void_p), ("wg_waterf_yfac", C.c_void_p),
        ("wg_waterf", C.c_void_p), ("wg_waterf_size", C.c_int), ("state", C.POINTER(WgState)),
    ]


_lib = None


def load_library():
    """Loads liblinrad_b200.so; raises if it has not been built (no CPU fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(the CUDA library is the only implementation of this path)")
    lib = C.CDLL(LIB_PATH)
    lib.lb200_create.argtypes = [C.POINTER(Config), C.POINTER(C.c_void_p)]
    lib.lb200_destroy.argtypes = [C.c_void_p]
    lib.lb200_destroy.restype = None
    lib.lb200_strerror.argtypes = [C.c_int]
    lib.lb200_strerror.restype = C.c_char_p
    lib.lb200_stream.argtypes = [C.c_void_p]
    lib.lb200_stream.restype = C.c_void_p
    lib.lb200_synchronize.argtypes = [C.c_void_p]
    for f in ("lb200_launch_count", "lb200_h2d_bytes", "lb200_d2h_bytes"):
        getattr(lib, f).argtypes = [C.c_void_p]
        getattr(lib, f).restype = C.c_uint64
    for f in ("lb200_fft1_dev", "lb200_fft1"):
        getattr(lib, f).argtypes = [C.c_void_p, C.POINTER(Fft1Args)]
    for f in ("lb200_mix1_dev", "lb200_mix1"):
        getattr(lib, f).argtypes = [C.c_void_p, C.POINTER(Mix1Args)]
    lib.lb200_set_mix1_phases.argtypes = [C.POINTER(Config), C.POINTER(Mix1State), C.c_float]
    lib.lb200_phase_advance.argtypes = [C.c_float, C.c_float, C.c_int]
    lib.lb200_phase_advance.restype = C.c_float
    lib.lb200_window_to_natural.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    lib.lb200_window_to_natural.restype = None
    for f in ("lb200_update_fft1_slowsum_dev", "lb200_update_fft1_slowsum", "lb200_fft1_waterfall_dev",
              "lb200_fft1_waterfall"):
        getattr(lib, f).argtypes = [C.c_void_p, C.POINTER(WgConfig), C.POINTER(WgArgs)]
    for f in ("lb200_expand_rawdat_dev", "lb200_expand_rawdat", "lb200_widen_24bit_dev", "lb200_widen_24bit"):
        getattr(lib, f).argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
    lib.lb200_raw_header_parse.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(RawHeader)]
    lib.lb200_raw_block_bytes.argtypes = [C.POINTER(RawHeader), C.c_size_t]
    lib.lb200_raw_block_bytes.restype = C.c_size_t
    _lib = lib
    return lib


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, np.float32)


def _ptr(a):
    return None if a is None else a.ctypes.data


def make_config(setup, device=0, window=None, filtercorr=None, max_batch=0, foldcorr=None, sample_shift=0,
                pg_ch2=(1.0, 0.0), inverted_window=None):
    """Build an lb200_config from a sizing.PathSetup.  `window`/`filtercorr` override the tables
    (e.g. with the reference's own, taken from the oracle in the parity tests).  Returns
    (Config, keepalive) -- keepalive holds the numpy tables until lb200_create has copied them."""
    keep = dict(
        window=_f32(window if window is not None else setup.window),
        filtercorr=_f32(filtercorr if filtercorr is not None else setup.filtercorr),
        foldcorr=_f32(foldcorr), invwin=_f32(inverted_window),
        fqwin=_f32(setup.mix1_fqwin), mwin=_f32(setup.mix1_window),
        cos2=_f32(setup.mix1_cos2win), sin2=_f32(setup.mix1_sin2win))
    cfg = Config()
    cfg.abi_version = LB200_ABI_VERSION
    cfg.device = device
    cfg.rx_input_mode = setup.input_mode
    cfg.rx_rf_channels = setup.rf_channels
    cfg.sample_shift = sample_shift
    cfg.fft1_n = setup.fft1_n
    cfg.fft1_interleave_points = setup.fft1_interleave_points
    cfg.fft1_direction = setup.direction
    cfg.fft1_first_point = setup.fft1_first_point
    cfg.fft1_last_point = setup.fft1_last_point
    cfg.fft1_window = _ptr(keep["window"])
    cfg.fft1_filtercorr = _ptr(keep["filtercorr"])
    cfg.fft1_foldcorr = _ptr(keep["foldcorr"])
    cfg.fft_avg1num = setup.avg1num
    cfg.mix1_n = setup.mix1_n
    cfg.mix1_interleave_points = setup.mix1_interleave_points
    cfg.mix1_crossover_points = setup.mix1_crossover_points
    cfg.mix1_fqwin = _ptr(keep["fqwin"])
    cfg.mix1_window = _ptr(keep["mwin"])
    cfg.mix1_cos2win = _ptr(keep["cos2"])
    cfg.mix1_sin2win = _ptr(keep["sin2"])
    cfg.fftx_points_per_hz = setup.fftx_points_per_hz
    cfg.mix1_lowest_fq = setup.mix1_lowest_fq
    cfg.mix1_highest_fq = setup.mix1_highest_fq
    cfg.max_batch = max_batch
    cfg.pg_ch2_c1, cfg.pg_ch2_c2 = float(pg_ch2[0]), float(pg_ch2[1])
    cfg.fft1_inverted_window = _ptr(keep["invwin"])
    return cfg, keep


class Plan:
    """Thin handle around lb200_plan."""

    def __init__(self, setup, device=0, window=None, filtercorr=None, max_batch=0, foldcorr=None, sample_shift=0,
                 pg_ch2=(1.0, 0.0), inverted_window=None):
        self.lib = load_library()
        self.setup = setup
        self.cfg, keep = make_config(setup, device, window, filtercorr, max_batch, foldcorr, sample_shift, pg_ch2,
                                     inverted_window)
        h = C.c_void_p()
        rc = self.lib.lb200_create(C.byref(self.cfg), C.byref(h))
        if rc:
            raise Lb200Error(rc, "lb200_create")
        self.h = h
        del keep

    def close(self):
        if getattr(self, "h", None):
            self.lib.lb200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def stream(self):
        return self.lib.lb200_stream(self.h)

    def synchronize(self):
        rc = self.lib.lb200_synchronize(self.h)
        if rc:
            raise Lb200Error(rc, "synchronize")

    def launches(self):
        return int(self.lib.lb200_launch_count(self.h))

    def h2d_bytes(self):
        return int(self.lib.lb200_h2d_bytes(self.h))

    def d2h_bytes(self):
        return int(self.lib.lb200_d2h_bytes(self.h))

    def _fft1_args(self, timf1_ptr, timf1_bytes, ref, nblocks, fft1_ptr, fft1_floats, fft1_pa,
                   apply_fc, sumsq_ptr, sumsq_floats, sumsq_pa, counter, power_ptr):
        a = Fft1Args()
        a.timf1 = Ring(timf1_ptr, timf1_bytes)
        a.timf1p_ref = ref
        a.nblocks = nblocks
        a.fft1_float = Ring(fft1_ptr, fft1_floats)
        a.fft1_pa = fft1_pa
        a.apply_filtercorr = 1 if apply_fc else 0
        a.fft1_sumsq = Ring(sumsq_ptr, sumsq_floats)
        a.fft1_sumsq_pa = sumsq_pa
        a.fft1_sumsq_counter = counter
        a.power_rows = power_ptr
        return a

    def fft1_dev(self, *, timf1, timf1_bytes, ref, nblocks, fft1, fft1_floats, fft1_pa=0, apply_fc=True,
                 sumsq=None, sumsq_floats=0, sumsq_pa=0, counter=0, power=None, rings=1, ring_stride=0, pa_stride=0):
        """All pointers are raw device addresses (ints)."""
        a = self._fft1_args(timf1, timf1_bytes, ref, nblocks, fft1, fft1_floats, fft1_pa, apply_fc,
                            sumsq, sumsq_floats, sumsq_pa, counter, power)
        a.no_of_rings, a.timf1_ring_stride, a.fft1_pa_stride = rings, ring_stride, pa_stride
        rc = self.lib.lb200_fft1_dev(self.h, C.byref(a))
        if rc:
            raise Lb200Error(rc, "lb200_fft1_dev")

    def fft1_host(self, *, timf1, ref, nblocks, fft1, fft1_pa=0, apply_fc=True, sumsq=None, sumsq_pa=0,
                  counter=0, power=None, keep_on_device=False, corrsum=None, corr_rows=None, xypower_rows=None):
        """numpy arrays standing in for Linrad's host rings (sizes must be powers of two)."""
        a = self._fft1_args(timf1.ctypes.data, timf1.nbytes, ref, nblocks, fft1.ctypes.data, fft1.size, fft1_pa,
                            apply_fc, _ptr(sumsq), 0 if sumsq is None else sumsq.size, sumsq_pa, counter,
                            _ptr(power))
        a.flags = FFT1_SPECTRUM_STAYS_ON_DEVICE if keep_on_device else 0
        if corrsum is not None:
            a.fft1_corrsum = Ring(corrsum.ctypes.data, corrsum.size)
        a.corr_rows = _ptr(corr_rows)
        a.xypower_rows = _ptr(xypower_rows)
        rc = self.lib.lb200_fft1(self.h, C.byref(a))
        if rc:
            raise Lb200Error(rc, "lb200_fft1")

    def _mix1_args(self, fft1_ptr, fft1_floats, fft1_px, nblocks, states, timf3_ptr, timf3_floats, timf3_pa):
        a = Mix1Args()
        a.fft1_float = Ring(fft1_ptr, fft1_floats)
        a.fft1_px = fft1_px
        a.nblocks = nblocks
        a.no_of_channels = len(states)
        a.state = states
        a.timf3_float = Ring(timf3_ptr, timf3_floats)
        a.timf3_pa = timf3_pa
        return a

    def mix1_dev(self, *, fft1, fft1_floats, fft1_px, nblocks, states, timf3, timf3_floats, timf3_pa):
        a = self._mix1_args(fft1, fft1_floats, fft1_px, nblocks, states, timf3, timf3_floats, timf3_pa)
        rc = self.lib.lb200_mix1_dev(self.h, C.byref(a))
        if rc:
            raise Lb200Error(rc, "lb200_mix1_dev")

    def mix1_host(self, *, fft1, fft1_px, nblocks, states, timf3, timf3_floats, timf3_pa):
        a = self._mix1_args(fft1.ctypes.data, fft1.size, fft1_px, nblocks, states, timf3.ctypes.data,
                            timf3_floats, timf3_pa)
        rc = self.lib.lb200_mix1(self.h, C.byref(a))
        if rc:
            raise Lb200Error(rc, "lb200_mix1")


def _wg_args(sumsq_ptr, sumsq_floats, sumsq_pa, nrows, slowsum, wsum, yfac, waterf, waterf_size, state):
    a = WgArgs()
    a.fft1_sumsq = Ring(sumsq_ptr, sumsq_floats)
    a.fft1_sumsq_pa = sumsq_pa
    a.nrows = nrows
    a.fft1_slowsum = slowsum
    a.wg_waterf_sum = wsum
    a.wg_waterf_yfac = yfac
    a.wg_waterf = waterf
    a.wg_waterf_size = waterf_size
    a.state = C.pointer(state)
    return a


def wide_graph_host(plan, wg, state, *, sumsq, sumsq_pa, nrows, slowsum, wsum, yfac, waterf, waterf_size):
    """update_fft1_slowsum for nrows new rows, then fft1_waterfall, on numpy host buffers."""
    a = _wg_args(sumsq.ctypes.data, sumsq.size, sumsq_pa, nrows, slowsum.ctypes.data, wsum.ctypes.data,
                 yfac.ctypes.data, waterf.ctypes.data, waterf_size, state)
    rc = plan.lib.lb200_update_fft1_slowsum(plan.h, C.byref(wg), C.byref(a))
    if rc:
        raise Lb200Error(rc, "lb200_update_fft1_slowsum")
    rc = plan.lib.lb200_fft1_waterfall(plan.h, C.byref(wg), C.byref(a))
    if rc:
        raise Lb200Error(rc, "lb200_fft1_waterfall")


def wide_graph_dev(plan, wg, state, *, sumsq, sumsq_floats, sumsq_pa, nrows, slowsum, wsum, yfac, waterf, waterf_size):
    """same on raw device addresses"""
    a = _wg_args(sumsq, sumsq_floats, sumsq_pa, nrows, slowsum, wsum, yfac, waterf, waterf_size, state)
    rc = plan.lib.lb200_update_fft1_slowsum_dev(plan.h, C.byref(wg), C.byref(a))
    if rc:
        raise Lb200Error(rc, "lb200_update_fft1_slowsum_dev")
    rc = plan.lib.lb200_fft1_waterfall_dev(plan.h, C.byref(wg), C.byref(a))
    if rc:
        raise Lb200Error(rc, "lb200_fft1_waterfall_dev")


def expand_rawdat_host(plan, packed, out_bytes):
    packed = np.ascontiguousarray(packed, np.uint8)
    out = np.zeros(out_bytes // 4, np.int32)
    rc = plan.lib.lb200_expand_rawdat(plan.h, packed.ctypes.data, out.ctypes.data, out_bytes)
    if rc:
        raise Lb200Error(rc, "lb200_expand_rawdat")
    return out


def widen_24bit_host(plan, pcm):
    pcm = np.ascontiguousarray(pcm, np.uint8)
    n = pcm.size // 3
    out = np.zeros(n, np.int32)
    rc = plan.lib.lb200_widen_24bit(plan.h, pcm.ctypes.data, out.ctypes.data, n)
    if rc:
        raise Lb200Error(rc, "lb200_widen_24bit")
    return out


def make_timf2_host(plan, *, fft1, fft1_px, nblocks, liminfo, timf2, timf2_pwr, timf2_pa, att_n=0):
    """make_timf2 (timf2.c:31) on numpy host rings; returns fft1_lowlevel_points"""
    a = Timf2Args()
    a.fft1_float = Ring(fft1.ctypes.data, fft1.size)
    a.fft1_px = fft1_px
    a.nblocks = nblocks
    lim = np.ascontiguousarray(liminfo, np.float32)
    a.liminfo = lim.ctypes.data
    a.timf2_float = Ring(timf2.ctypes.data, timf2.size)
    a.timf2_pwr_float = timf2_pwr.ctypes.data
    a.timf2_pa = timf2_pa
    a.first_bckfft_att_n = att_n
    low = C.c_int(0)
    a.fft1_lowlevel_points = C.pointer(low)
    plan.lib.lb200_make_timf2.argtypes = [C.c_void_p, C.POINTER(Timf2Args)]
    rc = plan.lib.lb200_make_timf2(plan.h, C.byref(a))
    if rc:
        raise Lb200Error(rc, "lb200_make_timf2")
    return low.value


def make_timf2_dev(plan, *, fft1, fft1_floats, fft1_px, nblocks, liminfo, timf2, timf2_floats, timf2_pwr, timf2_pa, att_n=0):
    """make_timf2 on raw device addresses (liminfo on the device too)"""
    a = Timf2Args()
    a.fft1_float = Ring(fft1, fft1_floats)
    a.fft1_px = fft1_px
    a.nblocks = nblocks
    a.liminfo = liminfo
    a.timf2_float = Ring(timf2, timf2_floats)
    a.timf2_pwr_float = timf2_pwr
    a.timf2_pa = timf2_pa
    a.first_bckfft_att_n = att_n
    plan.lib.lb200_make_timf2_dev.argtypes = [C.c_void_p, C.POINTER(Timf2Args)]
    rc = plan.lib.lb200_make_timf2_dev(plan.h, C.byref(a))
    if rc:
        raise Lb200Error(rc, "lb200_make_timf2_dev")


def widen_8bit_host(plan, pcm8):
    """8-bit unsigned PCM -> int16 (rxin.c:1573-1583)"""
    pcm8 = np.ascontiguousarray(pcm8, np.uint8)
    out = np.zeros(pcm8.size, np.int16)
    rc = plan.lib.lb200_widen_8bit(plan.h, C.c_void_p(pcm8.ctypes.data), C.c_void_p(out.ctypes.data), C.c_size_t(pcm8.size))
    if rc:
        raise Lb200Error(rc, "lb200_widen_8bit")
    return out


def float_to_int32_host(plan, z):
    """32-bit float samples -> int32 (rxin.c:1624-1634)"""
    z = np.ascontiguousarray(z, np.float32)
    out = np.zeros(z.size, np.int32)
    rc = plan.lib.lb200_float_to_int32(plan.h, C.c_void_p(z.ctypes.data), C.c_void_p(out.ctypes.data), C.c_size_t(z.size))
    if rc:
        raise Lb200Error(rc, "lb200_float_to_int32")
    return out


class Reducer:
    """lb200_reduce_*: sum of the ranks' averaged power spectra on the root (copy-engine push over
    NVLink + one add kernel).  `exchange` is a callable that all-gathers one 64-byte handle per rank
    (e.g. torch.distributed.all_gather_object) and returns the list ordered by rank."""

    def __init__(self, plan, rank, world, floats, exchange=None, root=0, depth=2):
        self.plan, self.rank, self.world, self.root = plan, rank, world, root
        lib = plan.lib
        lib.lb200_reduce_create.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_int, C.POINTER(C.c_void_p)]
        for name in ("lb200_reduce_export", "lb200_reduce_push", "lb200_reduce_sum"):
            getattr(lib, name).argtypes = [C.c_void_p, C.c_void_p]
        lib.lb200_reduce_connect.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        for name in ("lb200_reduce_rows_released", "lb200_reduce_result_ready", "lb200_reduce_synchronize", "lb200_reduce_destroy"):
            getattr(lib, name).argtypes = [C.c_void_p]
        lib.lb200_reduce_destroy.restype = None
        h = C.c_void_p()
        self._check(lib.lb200_reduce_create(plan.h, rank, world, floats, depth, C.byref(h)), "lb200_reduce_create")
        self.h = h
        if world > 1:
            mine = C.create_string_buffer(64)
            self._check(lib.lb200_reduce_export(self.h, mine), "lb200_reduce_export")
            handles = exchange(bytes(mine.raw))
            assert len(handles) == world and all(len(x) == 64 for x in handles)
            self._check(lib.lb200_reduce_connect(self.h, b"".join(handles), root), "lb200_reduce_connect")

    def _check(self, rc, what):
        if rc:
            raise Lb200Error(rc, what)

    def push(self, rows_ptr):
        self._check(self.plan.lib.lb200_reduce_push(self.h, C.c_void_p(rows_ptr)), "lb200_reduce_push")

    def rows_released(self):
        self._check(self.plan.lib.lb200_reduce_rows_released(self.h), "lb200_reduce_rows_released")

    def sum(self, out_ptr):
        self._check(self.plan.lib.lb200_reduce_sum(self.h, C.c_void_p(out_ptr)), "lb200_reduce_sum")

    def result_ready(self):
        self._check(self.plan.lib.lb200_reduce_result_ready(self.h), "lb200_reduce_result_ready")

    def synchronize(self):
        self._check(self.plan.lib.lb200_reduce_synchronize(self.h), "lb200_reduce_synchronize")

    def close(self):
        if self.h:
            self.plan.lib.lb200_reduce_destroy(self.h)
            self.h = None


def new_states(selfreqs):
    """Mix1 state array as buf.c:1259-1263 / wide_graph.c:174 initialise it."""
    arr = (Mix1State * len(selfreqs))()
    for i, f in enumerate(selfreqs):
        arr[i].mix1_selfreq = f
        arr[i].mix1_point = -1
    return arr


def advance_mix1_states(plan, states, ntransforms):
    """Step the per-selection mixer state over `ntransforms` transforms without computing them
    (set_mix1_phases mix1.c:781-861 + the running phase sum of do_mix1): what a rank that starts in
    the middle of a stream does to arrive at the sequential run's state (shard.block_ranges)."""
    s = plan.setup
    count = s.mix1_new_points if s.mix1_interleave_points else s.mix1_size
    for st in states:
        if st.mix1_selfreq < 0:
            continue
        one = (Mix1State * 1).from_address(C.addressof(st))
        for _ in range(ntransforms):
            rc = plan.lib.lb200_set_mix1_phases(C.byref(plan.cfg), one, C.c_float(st.mix1_selfreq))
            if rc:
                raise Lb200Error(rc, "lb200_set_mix1_phases")
            st.mix1_phase = plan.lib.lb200_phase_advance(st.mix1_phase, st.mix1_phase_rot, count)
    return states
