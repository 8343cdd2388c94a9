"""oracle/refwrap.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes wrapper around oracle/_ref/libref_oracle.so: the reference's own C files
(fft0.c fft1.c fft1_re.c mix1.c + *var.c) compiled unmodified by oracle/Makefile
and driven by oracle/ref_driver.c.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this module.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(_HERE, "_ref", "libref_oracle.so")
SHIM_SO = os.path.join(_HERE, "_ref", "libref_shim.so")     # same harness, hot path through lb200_shim.c (GPU)
CUFFT_SO = os.path.join(_HERE, "_ref", "libref_cufft.so")   # same harness, reference built with -DHAVE_CUFFT=1 (fft_cntrl row 19; GPU)

DWORD_INPUT, TWO_CHANNELS, IQ_DATA = 1, 2, 4


class RefCfg(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "input_mode", "rf_channels", "ad_speed", "fft1_n", "fft1_version", "sinpow",
        "fft1_gain", "mix1_red_n", "avg1num", "avg2num", "waterfall_avgnum", "direction",
        "n_sel", "first_xpoint", "xpoints", "xpoints_per_pixel", "pixels_per_xpoint",
        "wf_lines", "sample_shift", "correlation", "afc", "afc_mix")]


def available():
    return os.path.exists(REF_SO)


def shim_available():
    return os.path.exists(SHIM_SO)


def cufft_available():
    return os.path.exists(CUFFT_SO)


# the reference's own 18-bit codec, assembled from getiq64.s through oracle/nasm2gas.py (oracle/Makefile)
GETIQ_BIN = os.path.join(_HERE, "_ref", "getiq_check")


def getiq_available():
    return os.path.exists(GETIQ_BIN)


def _getiq(mode, nwords, payload):
    import subprocess
    r = subprocess.run([GETIQ_BIN, mode, str(nwords)], input=payload, capture_output=True)
    if r.returncode != 0:
        raise RuntimeError(f"getiq_check {mode} failed rc={r.returncode}")
    return r.stdout


def ref_expand_rawdat(packed):
    """expand_rawdat (getiq64.s:158-220) itself: 9 packed bytes -> 4 int32 words"""
    packed = np.ascontiguousarray(packed, np.uint8)
    nwords = packed.size // 9 * 4
    return np.frombuffer(_getiq("expand", nwords, packed.tobytes()), np.int32).copy()


def ref_compress_rawdat(words, net=False):
    """compress_rawdat_disk / _net (getiq64.s:39-156) themselves: 4 int32 words -> 9 bytes"""
    words = np.ascontiguousarray(words, np.int32)
    return np.frombuffer(_getiq("compress_net" if net else "compress", words.size, words.tobytes()), np.uint8).copy()


class RefOracle:
    """One instance at a time (the reference keeps its state in globals)."""

    def __init__(self, *, input_mode, rf_channels, ad_speed, fft1_n, fft1_version, sinpow=2,
                 fft1_gain=2000, mix1_red_n=4, avg1num=5, avg2num=4, waterfall_avgnum=10,
                 direction=1, n_sel=0, first_xpoint=0, xpoints=None, xpoints_per_pixel=1,
                 pixels_per_xpoint=1, wf_lines=8, sample_shift=0, timf1_bytes=None, max_fft1n=8, through_shim=False,
                 correlation=0, afc=0, afc_mix=0, cufft=False, gpu_batch_n=4):
        self.lib = C.CDLL(CUFFT_SO if cufft else SHIM_SO if through_shim else REF_SO)
        L = self.lib
        L.ref_init.argtypes = [C.POINTER(RefCfg), C.c_int, C.c_int]
        L.ref_process.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_process_timed.argtypes = [C.c_void_p, C.c_int]      # without it the pointer travels as a C int
        L.ref_set_selfreq.argtypes = [C.c_int, C.c_double]
        L.ref_set_foldcorr.argtypes = [C.c_void_p]
        L.ref_set_ch2_phasing.argtypes = [C.c_float, C.c_float]
        L.ref_sel_state.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
        L.ref_make_window.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.ref_fftback.argtypes = [C.c_int, C.c_int, C.c_void_p]
        for f in ("ref_points_per_hz", "ref_filtercorr_start"):
            getattr(L, f).restype = C.c_float
        for f in ("ref_window", "ref_filtercorr", "ref_desired", "ref_sumsq", "ref_slowsum",
                  "ref_waterf", "ref_waterf_yfac", "ref_waterf_sum", "ref_mix1_fqwin",
                  "ref_mix1_window", "ref_mix1_cos2win", "ref_mix1_sin2win", "ref_corrsum", "ref_slowcorr",
                  "ref_slowcorr_tot", "ref_fft1_power", "ref_fft1_xypower"):
            getattr(L, f).restype = C.c_void_p
        L.ref_timf3.restype = C.c_void_p
        L.ref_timf3.argtypes = [C.c_int]
        n = 1 << fft1_n
        if xpoints is None:
            xpoints = n
        self.cfg = RefCfg(input_mode, rf_channels, ad_speed, fft1_n, fft1_version, sinpow,
                          fft1_gain, mix1_red_n, avg1num, avg2num, waterfall_avgnum, direction,
                          n_sel, first_xpoint, xpoints, xpoints_per_pixel, pixels_per_xpoint,
                          wf_lines, sample_shift, correlation, afc, afc_mix)
        frame = (4 if input_mode & IQ_DATA else 2) * rf_channels
        if input_mode & DWORD_INPUT:
            frame *= 2
        if timf1_bytes is None:
            timf1_bytes = 1
            while timf1_bytes < 8 * n * frame:
                timf1_bytes *= 2
        if cufft:
            # the reference's GPU row transforms 2^gpu.fft1_batch_n blocks per fft1_b call (buf.c:248-258)
            L.ref_set_gpu_batch_n(gpu_batch_n)
            timf1_bytes *= 1 << gpu_batch_n
            while max_fft1n < 2 << gpu_batch_n:
                max_fft1n *= 2
        rc = L.ref_init(C.byref(self.cfg), timf1_bytes, max_fft1n)
        if rc != 0:
            raise RuntimeError(f"ref_init failed rc={rc}")
        self.n_sel = n_sel
        self.fft1_size = L.ref_fft1_size()
        self.fft1_block = L.ref_fft1_block()
        self.interleave_points = L.ref_interleave_points()
        self.new_points = L.ref_new_points()
        self.timf1_blockbytes = L.ref_timf1_blockbytes()
        self.mix1_size = L.ref_mix1_size()
        self.mix1_interleave = L.ref_mix1_interleave()
        self.mix1_new_points = L.ref_mix1_new_points()
        self.mix1_crossover = L.ref_mix1_crossover()
        self.timf3_block = L.ref_timf3_block()
        self.timf3_size = L.ref_timf3_size()
        self.wg_xpixels = L.ref_wg_xpixels()
        self.sumsq_bufsize = L.ref_sumsq_bufsize()
        self.points_per_hz = L.ref_points_per_hz()
        self.first_point = L.ref_first_point()
        self.last_point = L.ref_last_point()
        self.muln = L.ref_fft1_muln()              # transforms per fft1_b call (timf1_blockbytes covers all of them)

    def _arr(self, fn, count, dtype=np.float32):
        ptr = getattr(self.lib, fn)()
        buf = (C.c_char * (count * np.dtype(dtype).itemsize)).from_address(ptr)
        return np.frombuffer(buf, dtype=dtype).copy()

    def set_foldcorr(self, table):
        """install fft1_foldcorr (twice_rxchan*fft1_size floats) and raise CALIQ; None = off"""
        if table is None:
            self.lib.ref_set_foldcorr(None)
        else:
            t = np.ascontiguousarray(table, np.float32)
            self.lib.ref_set_foldcorr(t.ctypes.data)

    def set_ch2_phasing(self, c1, c2):
        self.lib.ref_set_ch2_phasing(float(c1), float(c2))

    def set_selfreq(self, ss, hz):
        self.lib.ref_set_selfreq(ss, float(hz))

    def sel_state(self, ss):
        f = np.zeros(4, np.float32)
        i = np.zeros(2, np.int32)
        self.lib.ref_sel_state(ss, f.ctypes.data, i.ctypes.data)
        return dict(phase=f[0], phase_step=f[1], phase_rot=f[2], old_phase=f[3],
                    point=int(i[0]), old_point=int(i[1]))

    def process(self, raw, want_raw=False):
        """raw: contiguous array of whole blocks in timf1 memory format."""
        raw = np.ascontiguousarray(raw)
        nbytes = raw.nbytes
        assert nbytes % self.timf1_blockbytes == 0, (nbytes, self.timf1_blockbytes)
        nb = nbytes // self.timf1_blockbytes
        nt = nb * self.muln
        fft1 = np.zeros((nt, self.fft1_block), np.float32)
        rawout = np.zeros((nt, self.fft1_block), np.float32) if want_raw else None
        t3 = np.zeros((nt, max(self.n_sel, 1), self.timf3_block), np.float32) if self.n_sel else None
        rc = self.lib.ref_process(raw.ctypes.data, nb, fft1.ctypes.data,
                                  rawout.ctypes.data if want_raw else None,
                                  t3.ctypes.data if t3 is not None else None)
        if rc != 0:
            raise RuntimeError(f"reference raised lirerr({rc})")
        return dict(fft1=fft1, raw=rawout, timf3=t3)

    # ---- second FFT front end: the reference's own make_timf2 on the blocks held in fft1_float
    def timf2_setup(self, att_n=0, pow_size=1 << 16):
        self.lib.ref_timf2_float.restype = C.POINTER(C.c_float)
        self.lib.ref_timf2_pwr_float.restype = C.POINTER(C.c_float)
        self.lib.ref_inverted_window.restype = C.POINTER(C.c_float)
        self.lib.ref_lowlevel_fraction.restype = C.c_float
        self.lib.ref_timf2_setup(att_n, pow_size)
        self.timf2_pow_size = pow_size

    def set_fft1_block(self, index, values):
        v = np.ascontiguousarray(values, np.float32)
        self.lib.ref_set_fft1_block(index, v.ctypes.data_as(C.c_void_p))

    def make_timf2(self, liminfo, px, nblocks):
        lim = np.ascontiguousarray(liminfo, np.float32)
        self.lib.ref_set_liminfo(lim.ctypes.data_as(C.c_void_p))
        rc = self.lib.ref_make_timf2(px, nblocks)
        if rc != 0:
            raise RuntimeError(f"reference raised lirerr({rc})")
        n = self.lib.ref_timf2_size()
        t2 = np.ctypeslib.as_array(self.lib.ref_timf2_float(), shape=(n,)).copy()
        pw = np.ctypeslib.as_array(self.lib.ref_timf2_pwr_float(), shape=(self.timf2_pow_size,)).copy()
        return dict(timf2=t2, pwr=pw, timf2_pa=self.lib.ref_timf2_pa(), input_block=self.lib.ref_timf2_input_block(),
                    lowlevel_points=self.lib.ref_lowlevel_points(), lowlevel_fraction=float(self.lib.ref_lowlevel_fraction()))

    def inverted_window(self):
        p = self.lib.ref_inverted_window()
        if not p:
            return None
        return np.ctypeslib.as_array(p, shape=(self.fft1_size // 2 + 1,)).copy()

    # ---- third FFT: the transform half of the reference's own make_fft3_all (fft3.c:215-470)
    def fft3_setup(self, n, sinpow, ring_floats, new_points=0):
        self.lib.ref_set_fft3_new_points(new_points)
        self.lib.ref_fft3_window.restype = C.POINTER(C.c_float)
        self.lib.ref_make_fft3.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        rc = self.lib.ref_fft3_setup(n, sinpow, ring_floats)
        if rc != 0:
            raise RuntimeError(f"ref_fft3_setup failed rc={rc}")
        self.fft3_size = 1 << n
        self.fft3_ring_floats = ring_floats

    def make_fft3(self, ring, px):
        ring = np.ascontiguousarray(ring, np.float32)
        assert ring.size == self.fft3_ring_floats
        out = np.zeros(2 * self.cfg.rf_channels * self.fft3_size, np.float32)
        rc = self.lib.ref_make_fft3(ring.ctypes.data, px, out.ctypes.data)
        if rc != 0:
            raise RuntimeError(f"reference raised lirerr({rc})")
        return out

    def fft3_window(self):
        return np.ctypeslib.as_array(self.lib.ref_fft3_window(), shape=(self.fft3_size,)).copy()

    def process_timed(self, raw, nblocks):
        return self.lib.ref_process_timed(raw.ctypes.data, nblocks)

    # snapshots of reference state
    def window(self, count=None):
        return self._arr("ref_window", count or self.fft1_size)

    def filtercorr(self):
        return self._arr("ref_filtercorr", self.fft1_block)

    def desired(self):
        return self._arr("ref_desired", self.fft1_size)

    def sumsq(self):
        return self._arr("ref_sumsq", self.sumsq_bufsize)

    def sumsq_pa(self):
        return self.lib.ref_sumsq_pa()

    def sumsq_counter(self):
        return self.lib.ref_sumsq_counter()

    def fft1_power(self):
        """fft1_power (fft1afc_flag > 0, one channel): max_fft1n rows of fft1_size floats"""
        return self._arr("ref_fft1_power", self.lib.ref_max_fft1n() * self.lib.ref_fft1_size())

    def fft1_xypower(self):
        """fft1_xypower (two channels): max_fft1n rows of fft1_size TWOCHAN_POWER {x2, y2, im_xy, re_xy}"""
        return self._arr("ref_fft1_xypower", 4 * self.lib.ref_max_fft1n() * self.lib.ref_fft1_size())

    def corrsum(self):
        """fft1_corrsum ring (fft1_correlation_flag == 1): 2*fft1_sumsq_bufsize floats"""
        return self._arr("ref_corrsum", 2 * self.lib.ref_sumsq_bufsize())

    def slowcorr(self):
        return self._arr("ref_slowcorr", 2 * self.lib.ref_fft1_size())

    def slowcorr_tot(self):
        return self._arr("ref_slowcorr_tot", 2 * self.lib.ref_fft1_size(), np.float64)

    def slowsum(self):
        return self._arr("ref_slowsum", self.fft1_size)

    def waterf(self):
        return self._arr("ref_waterf", self.lib.ref_waterf_size(), np.int16)

    def waterf_ptr(self):
        return self.lib.ref_waterf_ptr()

    def wg_state(self):
        """the scalar wide-graph globals (fft1.c:104-223, 4526-4605)"""
        L = self.lib
        return dict(fft1_sumsq_pwg=L.ref_sumsq_pwg(), fft1_sumsq_recalc=L.ref_sumsq_recalc(),
                    wg_waterf_sum_counter=L.ref_waterf_sum_counter(), wg_waterf_ptr=L.ref_waterf_ptr(),
                    latest_wg_spectrum=L.ref_latest_wg_spectrum(), wg_first_point=L.ref_wg_first_point(),
                    wg_last_point=L.ref_wg_last_point(), first_fft_bandwidth=L.ref_first_fft_bandwidth())

    def set_change_fft1_flag(self, v):
        self.lib.ref_set_change_fft1_flag(int(v))

    def waterf_yfac(self):
        return self._arr("ref_waterf_yfac", self.fft1_size)

    def waterf_sum(self):
        return self._arr("ref_waterf_sum", self.fft1_size)

    def mix1_fqwin(self):
        return self._arr("ref_mix1_fqwin", self.mix1_size // 2 + 1)

    def mix1_window(self):
        return self._arr("ref_mix1_window", self.mix1_size)

    def mix1_cos2win(self):
        return self._arr("ref_mix1_cos2win", max(self.mix1_crossover, 1))

    def mix1_sin2win(self):
        return self._arr("ref_mix1_sin2win", max(self.mix1_crossover, 1))

    def timf3_ring(self, ss):
        ptr = self.lib.ref_timf3(ss)
        buf = (C.c_char * (self.timf3_size * 4)).from_address(ptr)
        return np.frombuffer(buf, dtype=np.float32).copy()

    def timf3_pa(self):
        return self.lib.ref_timf3_pa()

    def make_window(self, mo, sz, n, count=None):
        w = np.zeros((count or sz) + 32, np.float32)
        self.lib.ref_make_window(mo, sz, n, w.ctypes.data)
        return w[:count or sz]

    def fftback(self, x):
        x = np.ascontiguousarray(x, np.complex64).copy()
        n = int(np.log2(len(x)))
        self.lib.ref_fftback(len(x), n, x.ctypes.data)
        return x
