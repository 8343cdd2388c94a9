"""oracle/port.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Numpy restatement of the reference's wideband path, function by function, written from the
reference's arithmetic (file:line cited per function).  It is the checker that travels when
oracle/_ref (the compiled reference itself) is not available, and an independent referee in
float64 for the places where the reference's own float32 kernels disagree with each other.

Parity pin: tests/test_oracle_cpu.py checks every function here against oracle/_ref (the
reference's own C files compiled by oracle/Makefile) in this container, and against the golden
fixtures under tests/golden/ (generated from oracle/_ref by tests/golden/make_golden.py) on
machines without /root/reference.  The reference ships no golden vectors of its own
(SURVEY.md 8(c)).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
import ctypes as C
import math
import os

import numpy as np

f32 = np.float32
PI_L = 3.1415926535897932      # globdef.h:93
FFT1_SMALL = 1e-20             # fft1def.h
_HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(_HERE, "liboracle_port.so")
_lib = None


def clib():
    """liboracle_port.so: the byte/integer stages in plain C (oracle_port.c)."""
    global _lib
    if _lib is None:
        _lib = C.CDLL(PORT_SO)
        _lib.port_expand_rawdat.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        _lib.port_compress_rawdat.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        _lib.port_widen_24bit.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        _lib.port_widen_8bit.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        _lib.port_float_to_int32.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        _lib.port_phase_chain.argtypes = [C.c_float, C.c_float, C.c_int, C.c_void_p]
        _lib.port_phase_chain.restype = C.c_float
    return _lib


# ------------------------------------------------------------------------------------------
# integer stages
def expand_rawdat(packed, out_bytes):
    """getiq64.s:158-220 (see oracle_port.c)."""
    packed = np.ascontiguousarray(packed, np.uint8)
    out = np.zeros(out_bytes // 4, np.int32)
    clib().port_expand_rawdat(packed.ctypes.data, out.ctypes.data, out_bytes)
    return out


def expand_rawdat_numpy(packed, out_bytes):
    """Same in numpy, independent of the C restatement."""
    g = out_bytes // 16
    p = np.ascontiguousarray(packed, np.uint8)[: 9 * g].reshape(g, 9).astype(np.uint32)
    out = np.empty((g, 4), np.uint32)
    for i in range(4):
        hi = p[:, 2 * i] | (p[:, 2 * i + 1] << 8)
        out[:, i] = (hi << 16) | (((p[:, 8] >> (2 * i)) & 3) << 14)
    out += np.uint32(0x2000)
    return out.reshape(-1).view(np.int32)


def compress_rawdat(words):
    """getiq64.s:39-96."""
    words = np.ascontiguousarray(words, np.int32)
    out = np.zeros(words.size // 4 * 9, np.uint8)
    clib().port_compress_rawdat(words.ctypes.data, out.ctypes.data, words.nbytes)
    return out


def widen_8bit(b):
    """rxin.c:1573-1583."""
    b = np.ascontiguousarray(b, np.uint8)
    out = np.zeros(b.size, np.int16)
    clib().port_widen_8bit(b.ctypes.data, out.ctypes.data, out.size)
    return out


def float_to_int32(z):
    """rxin.c:1624-1634 (the host's own float -> int conversion, i.e. the reference's on x86-64)."""
    z = np.ascontiguousarray(z, np.float32)
    out = np.zeros(z.size, np.int32)
    clib().port_float_to_int32(z.ctypes.data, out.ctypes.data, out.size)
    return out


def widen_24bit(b):
    """rxin.c:1603-1614."""
    b = np.ascontiguousarray(b, np.uint8)
    out = np.zeros(b.size // 3, np.int32)
    clib().port_widen_24bit(b.ctypes.data, out.ctypes.data, out.size)
    return out


def phase_chain(phase, rot, count, trace=False):
    """mix1.c:146-153: `t1+=t2` count times in float; returns (final, values before each add)."""
    tr = np.zeros(count, np.float32) if trace else None
    r = clib().port_phase_chain(f32(phase), f32(rot), count, tr.ctypes.data if trace else None)
    return f32(r), tr


# ------------------------------------------------------------------------------------------
# fft1_b: fft1win_* + FFT core + output convention
def timf1_span(ring, mask_bytes, ref, setup):
    """The input words of one transform as float64 [samples, words_per_frame]: the span starts
    fft1_interleave_points frames before timf1p_ref (fft1.c:423-428,700; real: fft1_re.c:44)."""
    s = setup
    iq = bool(s.input_mode & 4)
    nfr = s.fft1_size if iq else 2 * s.fft1_size
    pre = s.fft1_interleave_points * s.frame_bytes * (1 if iq else 2)
    idx = (ref - pre + np.arange(nfr * s.frame_bytes)) & mask_bytes
    raw = ring[idx]
    dt = np.int32 if (s.input_mode & 1) else np.int16
    return raw.view(dt).reshape(nfr, -1).astype(np.float64)


def fft1_b(ring, mask_bytes, ref, setup, window=None, sample_shift=0, foldcorr=None, pg_ch2=None):
    """One call of fft1_b (fft1.c:3302): returns fft1_block floats in fft1_float layout
    (mm floats per bin: re1,im1[,re2,im2]), BEFORE fft1_c.
      complex input (versions 6/7, fft1.c:3495-3506,3788-3796; cores fft0.c:1590,161):
          out[k] = conj( sum_n w[n] x[n] exp(-2 pi i n ((k+N/2) mod N)/N) )     (probed)
          ui.sample_shift (one channel only, fft1.c:770-790): < 0 takes Q from |shift| frames
              earlier, > 0 takes I from shift frames earlier
          CALIQ (fft1.c:3607-3657, 3941-4026), per channel, ib = m..N/2-1, ic = N-ib,
              m = max(1, fft1_first_sym_point): z'[ib] = z[ib] - conj(z[ic]) f[ic],
              z'[ic] = z[ic] - conj(z[ib] f[ib])
          fft1_direction<0 (fft1.c:3628-3680): bins ib in [m, N/2) and their mirrors are exchanged
              with re/im swapped, bins 0 and N/2 swap re/im, bins outside [m, N-m] stay
          channel 2 phasing (fft1.c:4064-4080): ch2 *= (c1 - i c2) on [first_sym, N - first_sym)
      real input (version 2, fft1_re.c:32-131): 2N reals, X_k = sum x[n] w[n] exp(-2 pi i n k/2N),
          out[2k]=Im X_k, out[2k+1]=Re X_k for k=1..N-1, bin 0 = (X_N, X_0) (fft1_re.c:100-114)."""
    s = setup
    N, C_ = s.fft1_size, s.rf_channels
    x = timf1_span(ring, mask_bytes, ref, s)
    w = window if window is not None else s.window
    out = np.zeros((N, 2 * C_), np.float64)
    first_sym = min(N - 1 - s.fft1_last_point, s.fft1_first_point)      # fft1.c:4647-4649
    if s.input_mode & 4:
        if sample_shift and C_ == 1:
            # gather I and Q from different frames of the ring
            dt = np.int32 if (s.input_mode & 1) else np.int16
            words = ring.view(dt)
            wmask = (mask_bytes + 1) // words.itemsize - 1
            p0 = (ref - s.fft1_interleave_points * s.frame_bytes) // words.itemsize
            n = np.arange(N)
            di = -sample_shift if sample_shift > 0 else 0
            dq = sample_shift if sample_shift < 0 else 0
            x = np.stack([words[(p0 + 2 * (n + di)) & wmask], words[((p0 + 2 * (n + dq)) & wmask) + 1]], axis=1).astype(np.float64)
        for c in range(C_):
            z = x[:, 2 * c] + 1j * x[:, 2 * c + 1]
            if w is not None:
                z = z * w.astype(np.float64)
            X = np.fft.fft(z)
            y = np.conj(np.roll(X, -N // 2))
            m = max(1, first_sym)
            ib = np.arange(m, N // 2)
            ic = N - ib
            if foldcorr is not None:
                f = np.asarray(foldcorr, np.float64).reshape(N, 2 * C_)
                f = f[:, 2 * c] + 1j * f[:, 2 * c + 1]
                yb, yc = y[ib].copy(), y[ic].copy()
                y[ib] = yb - np.conj(yc) * f[ic]
                y[ic] = yc - np.conj(yb * f[ib])
            if s.direction < 0:
                yb, yc = y[ib].copy(), y[ic].copy()
                y[ib] = yc.imag + 1j * yc.real
                y[ic] = yb.imag + 1j * yb.real
                for k in (0, N // 2):
                    y[k] = y[k].imag + 1j * y[k].real
            if pg_ch2 is not None and c == 1:
                k = np.arange(first_sym, N - first_sym)
                y[k] = y[k] * (pg_ch2[0] - 1j * pg_ch2[1])
            out[:, 2 * c] = y.real
            out[:, 2 * c + 1] = y.imag
    else:
        for c in range(C_):
            z = x[:, c]
            if w is not None:
                z = z * w.astype(np.float64)
            X = np.fft.fft(z)
            lo, hi = s.fft1_first_point, s.fft1_last_point
            if s.direction > 0:                       # fft1_re.c:100-114; bins outside [lo,hi] are not written
                k0 = max(lo, 1)
                out[k0:hi + 1, 2 * c] = X[k0:hi + 1].imag
                out[k0:hi + 1, 2 * c + 1] = X[k0:hi + 1].real
                out[0, 2 * c] = X[N].real
                out[0, 2 * c + 1] = X[0].real
            else:                                     # fft1_re.c:115-130: bin N-ia = (Re X_ia, Im X_ia), ia=k..m
                out[N - 1, 2 * c] = X[0].real
                out[N - 1, 2 * c + 1] = X[N].real
                k0 = N - 1 - hi
                m = 1 + k0 + hi - lo
                k0 = max(k0, 1)
                for ia in range(k0, m + 1):
                    out[N - ia, 2 * c] = X[ia].real
                    out[N - ia, 2 * c + 1] = X[ia].imag if ia < N else X[N].real
    return out.reshape(-1)


# ------------------------------------------------------------------------------------------
def fft1_corr(block, setup):
    """fft1_correlation_flag == 1 (fft1.c:4146-4152): 2 z1 conj(z2) of a filter-corrected
    two-channel block on [first_point,last_point], as [N,2] floats (zero outside)."""
    s = setup
    N = s.fft1_size
    z = np.asarray(block, np.float64).reshape(N, 4)
    lo, hi = s.fft1_first_point, s.fft1_last_point
    c = np.zeros((N, 2))
    x = 2 * (z[lo:hi + 1, 0] + 1j * z[lo:hi + 1, 1]) * np.conj(z[lo:hi + 1, 2] + 1j * z[lo:hi + 1, 3])
    c[lo:hi + 1, 0] = x.real
    c[lo:hi + 1, 1] = x.imag
    return c


def fft1_c(block, filtercorr, setup):
    """fft1.c:4115-4200: z *= filtercorr on [first_point,last_point]; returns (block', power)
    with power[k] = sum over channels |z|^2 (zero outside the range)."""
    s = setup
    N, mm = s.fft1_size, 2 * s.rf_channels
    z = block.reshape(N, mm).astype(np.float64).copy()
    fc = np.asarray(filtercorr, np.float64).reshape(N, mm)
    lo, hi = s.fft1_first_point, s.fft1_last_point
    pw = np.zeros(N)
    for c in range(s.rf_channels):
        a = z[lo:hi + 1, 2 * c] + 1j * z[lo:hi + 1, 2 * c + 1]
        f = fc[lo:hi + 1, 2 * c] + 1j * fc[lo:hi + 1, 2 * c + 1]
        a = a * f
        z[lo:hi + 1, 2 * c] = a.real
        z[lo:hi + 1, 2 * c + 1] = a.imag
        pw[lo:hi + 1] += np.abs(a) ** 2
    return z.reshape(-1), pw


class SumsqState:
    """fft1_sumsq ring bookkeeping of fft1_c (fft1.c:4115,4507-4523)."""

    def __init__(self, setup, rows=16):
        self.N = setup.fft1_size
        self.avg1num = setup.avg1num
        self.ring = np.zeros(rows * self.N, np.float64)
        self.pa = 0
        self.counter = 0
        self.completed = []          # sumsq_pa of every completed row, in order

    def add(self, power, lo, hi):
        row = self.ring[self.pa: self.pa + self.N]
        if self.counter == 0:
            row[lo:hi + 1] = power[lo:hi + 1]
        else:
            row[lo:hi + 1] += power[lo:hi + 1]
        self.counter += 1
        if self.counter >= self.avg1num:
            self.completed.append(self.pa)
            self.pa = (self.pa + self.N) & (self.ring.size - 1)
            self.counter = 0
            return True
        return False


def new_fft1_averages(sumsq_ring, ptr, N, avg2num, ia, ib, slowsum):
    """wide_graph.c:1003-1051: slowsum[ia..ib] = sum of the latest avg2num rows ending at ptr."""
    size = sumsq_ring.size
    src = (ptr - (avg2num - 1) * N + size) & (size - 1)
    slowsum[ia:ib + 1] = sumsq_ring[src + ia: src + ib + 1]
    for _ in range(1, avg2num):
        src = (src + N) & (size - 1)
        slowsum[ia:ib + 1] += sumsq_ring[src + ia: src + ib + 1]
        np.maximum(slowsum[ia:ib + 1], FFT1_SMALL, out=slowsum[ia:ib + 1])


def waterfall_line(wsum, yfac, lo, npix):
    """fft1.c:129-165, 1:1 pixel mapping: short = clamp(1000*log10(sum*yfac)); values below
    -32767 clamp to -32767, above 32767 to 32767 (fft1.c:150-160)."""
    y = wsum[lo:lo + npix].astype(np.float64) * yfac[lo:lo + npix].astype(np.float64)
    with np.errstate(divide="ignore"):
        v = 1000.0 * np.log10(np.maximum(y, 1e-300))
    return v


# ------------------------------------------------------------------------------------------
# mix1
def set_mix1_phases(st, fq, setup):
    """mix1.c:781-861, float branch; st is a dict with the reference's per-selection globals."""
    M = setup.mix1_size
    Mn = setup.mix1_new_points
    t1 = f32(f32(fq) * f32(setup.fftx_points_per_hz))
    pnt = int(np.float64(t1) + 0.5)
    k = pnt % M
    t2 = f32(M * (pnt // M))
    t2 = f32(f32(t1 - t2) - f32(k))
    t2 = f32(t2 - f32(int(t2)))
    st["phase_rot"] = f32(np.float64(t2) * 2 * PI_L / M)
    k = (k * Mn) % M
    st["old_phase"] = st["phase"]
    st["phase"] = f32(st["phase"] + st["phase_step"])
    st["phase_step"] = f32((k * 2) * PI_L / M)
    st["old_point"] = st["point"] if st["point"] != -1 else pnt
    st["point"] = pnt
    if np.float64(st["phase"]) > PI_L:
        st["phase"] = f32(np.float64(st["phase"]) - 2 * PI_L)
    if np.float64(st["phase"]) < PI_L:          # reference quirk kept (mix1.c:860)
        st["phase"] = f32(np.float64(st["phase"]) + 2 * PI_L)
    return st


def new_sel_state():
    return dict(phase=f32(0), phase_step=f32(0), phase_rot=f32(0), old_phase=f32(0), point=-1, old_point=0)


def mix1_gather(block, point, setup):
    """fft1_mix1_fixed, mix1.c:1015-1030: M bins around `point` in fftback order (upper half
    first), zero outside [first_point, last_point) -- the upper clamp excludes last_point."""
    s = setup
    N, C_, M = s.fft1_size, s.rf_channels, s.mix1_size
    z = block.reshape(N, 2 * C_).astype(np.float64)
    out = np.zeros((C_, M), np.complex128)
    for i in range(M):
        b = point + i if i < M // 2 else point - M + i
        ok = (b < s.fft1_last_point) if i < M // 2 else (b >= s.fft1_first_point)
        if ok and 0 <= b < N:
            for c in range(C_):
                out[c, i] = z[b, 2 * c] + 1j * z[b, 2 * c + 1]
    return out


def mix1_taper(M, fqwin, nch):
    """do_mix1 mix1.c:113-135 (one channel) / 455-491 (two channels: the end points of the
    two-channel loop get a second factor)."""
    h = M // 2
    w = np.empty(M, np.float64)
    for i in range(M):
        if i == 0:
            w[i] = fqwin[h - 1]
        elif i <= h:
            w[i] = fqwin[h - i]
        else:
            w[i] = fqwin[i - h]
    if nch == 2:
        w[M - 1] *= fqwin[h - 1]
        w[h] *= fqwin[0]
    return w


class Mix1Port:
    """do_mix1 (mix1.c:55-272, 453-645) for one selection, carrying the timf3 ring."""

    def __init__(self, setup, timf3_size):
        self.s = setup
        self.size = timf3_size
        self.ring = np.zeros(timf3_size, np.float64)
        self.pa = 0
        self.st = new_sel_state()
        self.taper = mix1_taper(setup.mix1_size, setup.mix1_fqwin, setup.rf_channels)

    def step(self, block, selfreq):
        s = self.s
        C_, M, Mi, Mn = s.rf_channels, s.mix1_size, s.mix1_interleave_points, s.mix1_new_points
        mm = 2 * C_
        mask = self.size - 1
        if selfreq < 0:
            idx = (self.pa + np.arange(mm * Mn)) & mask
            self.ring[idx] = 0
            out = self.ring[idx].copy()
            self.pa = (self.pa + mm * Mn) & mask
            return out
        st = set_mix1_phases(self.st, selfreq, s)
        y = mix1_gather(block, st["point"], s) * self.taper[None, :]
        y = np.fft.fft(y, axis=1)                                # fftback: sum_k y_k e^{-2 pi i n k/M}
        _, tph = phase_chain(st["phase"], st["phase_rot"], Mn, trace=True)
        fin, _ = phase_chain(st["phase"], st["phase_rot"], Mn)
        rot = np.exp(1j * tph.astype(np.float64))
        r2 = f32(np.float64(st["phase_rot"]) - 2 * (st["old_point"] - st["point"]) * PI_L / M)

        def put(sample, c, val):
            o = (self.pa + sample * mm + 2 * c) & mask
            self.ring[o] = val.real
            self.ring[o + 1] = val.imag

        def get(sample, c):
            o = (self.pa + sample * mm + 2 * c) & mask
            return self.ring[o] + 1j * self.ring[o + 1]

        if Mi == 0:                                              # mix1.c:141-155
            for c in range(C_):
                for i in range(M):
                    put(i, c, rot[i] * y[c, i])
        elif Mi == Mn:                                           # mix1.c:161-195
            _, rph = phase_chain(st["old_phase"], r2, M // 2, trace=True)
            rrot = np.exp(1j * rph.astype(np.float64))
            for c in range(C_):
                for i in range(M // 2):
                    put(i, c, rrot[i] * get(i, c) + rot[i] * y[c, i])
                for i in range(M // 2, M):
                    put(i, c, y[c, i])
        else:                                                    # mix1.c:196-270
            cross = s.mix1_crossover_points
            k = Mi // 2 - cross // 2
            _, rph = phase_chain(st["old_phase"], r2, cross, trace=True)
            rrot = np.exp(1j * rph.astype(np.float64))
            sb = Mn // 2 + 1 + cross // 2
            for c in range(C_):
                for i in range(cross):
                    put(i, c, rrot[i] * (s.mix1_cos2win[i] * get(i, c)) + rot[i] * y[c, i + k] * s.mix1_sin2win[i])
                for i in range(cross, Mn):
                    j = (k + i) if i < sb else (k + 2 * sb - 2 - i)
                    put(i, c, rot[i] * y[c, i + k] * s.mix1_window[j])
                for i in range(Mn, Mn + cross):
                    put(i, c, y[c, i + k])
        st["phase"] = fin
        idx = (self.pa + np.arange(mm * Mn)) & mask
        out = self.ring[idx].copy()
        self.pa = (self.pa + mm * Mn) & mask
        return out


# ------------------------------------------------------------------------------------------
def run_path(setup, raw, selbins, nblocks, timf1_bytes=None, timf3_size=None, filtercorr=None, window=None,
             sample_shift=0, foldcorr=None, pg_ch2=None, correlation=0):
    """Drive the whole path the way wideband_dsp/narrowband_dsp do (wcw.c:1036-1085,1706-1716)
    over `nblocks` blocks of raw timf1 data; mirrors oracle/refwrap.RefOracle.process."""
    s = setup
    N = s.fft1_size
    rawb = np.ascontiguousarray(raw).view(np.uint8).reshape(-1)
    tb = timf1_bytes or _pow2(8 * N * s.frame_bytes)
    ring = np.zeros(tb, np.uint8)
    pa = px = 0
    fc = filtercorr if filtercorr is not None else s.filtercorr
    t3size = timf3_size or 16 * s.mix1_size * 2 * s.rf_channels
    mixers = [Mix1Port(s, t3size) for _ in selbins]
    hz = s.ad_speed / N / (1 if s.input_mode & 4 else 2)
    sq = SumsqState(s)
    corr_ring = np.zeros(2 * sq.ring.size)
    fft1_out = np.zeros((nblocks, s.fft1_block))
    raw_out = np.zeros((nblocks, s.fft1_block))
    t3_out = np.zeros((nblocks, max(len(selbins), 1), s.timf3_block))
    for b in range(nblocks):
        src = rawb[b * s.timf1_blockbytes:(b + 1) * s.timf1_blockbytes]
        ring[(pa + np.arange(src.size)) & (tb - 1)] = src
        pa = (pa + src.size) & (tb - 1)
        blk = fft1_b(ring, tb - 1, px, s, window=window, sample_shift=sample_shift, foldcorr=foldcorr, pg_ch2=pg_ch2)
        px = (px + s.timf1_blockbytes) & (tb - 1)
        raw_out[b] = blk
        blk, pw = fft1_c(blk, fc, s)
        if correlation == 1 and s.rf_channels == 2:
            cr = fft1_corr(blk, s)
            row = corr_ring[2 * sq.pa: 2 * (sq.pa + N)].reshape(N, 2)
            if sq.counter == 0:
                row[s.fft1_first_point:s.fft1_last_point + 1] = cr[s.fft1_first_point:s.fft1_last_point + 1]
            else:
                row[s.fft1_first_point:s.fft1_last_point + 1] += cr[s.fft1_first_point:s.fft1_last_point + 1]
        sq.add(pw, s.fft1_first_point, s.fft1_last_point)
        fft1_out[b] = blk
        for i, fb in enumerate(selbins):
            t3_out[b, i] = mixers[i].step(blk, fb * hz if fb >= 0 else -1.0)
    return dict(fft1=fft1_out, raw=raw_out, timf3=t3_out, sumsq=sq.ring, corrsum=corr_ring, sumsq_pa=sq.pa, sumsq_counter=sq.counter,
                states=[m.st for m in mixers], timf3_ring=[m.ring for m in mixers])


def _pow2(x):
    p = 1
    while p < x:
        p *= 2
    return p
