"""Host-side setup logic of the path, restated from the reference's init code so that a
configuration can be built without Linrad itself (bench.py, tests on the GPU box):

  get_wideband_sizes   buf.c:139-332   fft1 / mix1 sizes and interleave
  make_interleave_ratio buf.c:113-136
  make_window          fft0.c:812-921  (natural order, i.e. mo=4; mo=5 erfc taper; mo=3 inverse)
  clear_fft1_filtercorr / make_filcorrstart  fft1.c:4653-4724
  prepare_mixer        buf.c:55-111    crossover tables
  set_fft1_endpoints   fft1.c:4607-4651
  make_wg_yfac         wide_graph.c:956-1001 (waterfall zero point)

Everything here is table/scalar setup that Linrad does once per mode change; none of it is on
the per-sample path.  Float32 roundings follow the reference's C types.
"""
import math
from dataclasses import dataclass, field

import numpy as np

DWORD_INPUT, TWO_CHANNELS, IQ_DATA = 1, 2, 4      # globdef.h:277-279
FLOAT_INPUT = 64                                  # globdef.h:283 (float frames: the timf3 ring as input of the third FFT)
PI_L = 3.1415926535897932                          # globdef.h:93
FFT1_WATERFALL_ZERO = 0.14                         # graphcal.h:12
f32 = np.float32


def make_interleave_ratio(sinpow):
    """buf.c:113-136; the result is stored in a float."""
    if sinpow == 0:
        return f32(0.0)
    if sinpow == 9:
        return f32(0.625)
    if sinpow == 8:
        return f32(0.8)
    return f32(2 * math.asin(math.pow(0.5, 1.0 / sinpow)) / PI_L)


def _sequential_sum(start, step, count):
    """x=start; repeat: x+=step  (double), returns the `count` values before each add."""
    steps = np.full(count, step, np.float64)
    steps[0] = start
    return np.add.accumulate(steps)


def make_window(mo, sz, n):
    """fft0.c:812-921.  mo=4: natural-order normalised window of sz points (sin^n, n=1..7;
    8 Gaussian; 9 erfc).  mo=2: first sz+1 points of the 2*sz window.  mo=3: inverse window
    (indices 0..sz/2).  mo=5: the mix1 frequency-domain taper 0.5*erfc(3.2-13 i/sz), i<=sz/2."""
    if mo == 5:
        e1 = _sequential_sum(3.2, -13.0 / sz, sz // 2 + 1)
        return (f32(0.5) * np.array([math.erfc(v) for v in e1]).astype(f32)).astype(f32)
    if n == 0:
        return None
    size = 2 * sz if mo == 2 else sz
    half = size // 2
    if n == 9:
        e2 = 40.0 / size
        if size < 128:
            e2 /= 1.5
        if size < 64:
            e2 /= 1.7
        e1 = _sequential_sum(4.4, -e2, half + 1)
        win = (f32(0.5) * np.array([math.erfc(v) for v in e1]).astype(f32)).astype(f32)
    elif n == 8:
        e1 = _sequential_sum(0.0, 9.8 / size, half + 1)        # filled from i=size/2 downwards
        NATLOG = 2.718281828459045
        win = np.array([math.pow(NATLOG, -v * v) for v in e1]).astype(f32)[::-1].copy()
    else:
        x = _sequential_sum(0.0, PI_L / size, half + 1)
        win = np.array([math.pow(math.sin(v), float(n)) for v in x]).astype(f32)
    # sumsq+=win[i]*win[i] : float product, double accumulation, in index order
    prod = (win * win).astype(f32).astype(np.float64)
    if n == 8:
        sumsq = np.add.accumulate(prod[::-1])[-1]
    else:
        sumsq = np.add.accumulate(prod)[-1]
    if mo == 3:
        inv = np.ones(half + 1, f32)
        with np.errstate(divide="ignore"):
            inv[1:] = (f32(1) / win[1:]).astype(f32)
        return inv
    z = 1 / math.sqrt(2 * sumsq / size)
    win = (win * f32(z)).astype(f32)
    if mo == 2:
        return win                                        # half table, symmetric use
    full = np.empty(size, f32)
    full[:half + 1] = win
    full[half + 1:] = win[1:half][::-1]                   # win[i]=win[size-i]
    return full


def make_filcorrstart(fft1_size, input_mode, gain, permute2=False, real2complex=False):
    """fft1.c:4653-4671 -> fft1_filtercorr_start (float arithmetic)."""
    s = f32(f32(150) * f32(fft1_size)) * f32(math.pow(float(fft1_size), -0.4))
    s = f32(s)
    if input_mode & DWORD_INPUT:
        s = f32(s * f32(4096))
        if permute2:
            s = f32(s * f32(16))
        if input_mode & IQ_DATA:
            s = f32(s * f32(12))
        if real2complex:
            s = f32(s * f32(32))
    elif real2complex:
        s = f32(s * f32(2))
    return f32(f32(gain) / s)


def clear_fft1_filtercorr(fft1_size, rf_channels, input_mode, gain, permute2=False, real2complex=False):
    """fft1.c:4673-4724: uncalibrated filtercorr (mm floats per bin) and fft1_desired."""
    mm = 2 * rf_channels
    start = make_filcorrstart(fft1_size, input_mode, gain, permute2, real2complex)
    fc = np.zeros((fft1_size, mm), f32)
    fc[:, 0::2] = start
    desired = np.ones(fft1_size, f32)
    t1 = f32(f32(0.125) * f32(PI_L))
    t2 = f32(0)
    i, k = 0, fft1_size - 1
    while float(t2) < 0.5 * PI_L:
        t3 = f32(math.sin(float(t2)) * math.sin(float(t2)))
        if input_mode & IQ_DATA:
            desired[i] = t3
            fc[i, 0::2] = f32(t3 * start)
        desired[k] = t3
        fc[k, 0::2] = f32(t3 * start)
        t2 = f32(t2 + t1)
        i += 1
        k -= 1
    return fc.reshape(-1).copy(), desired


@dataclass
class PathSetup:
    """All reference globals the path needs, for one receive mode."""
    input_mode: int
    rf_channels: int
    ad_speed: int
    fft1_n: int
    sinpow: int = 2
    fft1_gain: int = 2000
    mix1_red_n: int = 4
    avg1num: int = 5
    avg2num: int = 4
    waterfall_avgnum: int = 10
    direction: int = 1
    first_xpoint: int = 0
    xpoints: int = -1
    second_fft: bool = False
    # derived
    fft1_size: int = 0
    fft1_interleave_points: int = 0
    fft1_new_points: int = 0
    frame_bytes: int = 0
    timf1_blockbytes: int = 0
    fft1_block: int = 0
    mix1_n: int = 0
    mix1_size: int = 0
    mix1_interleave_points: int = 0
    mix1_new_points: int = 0
    mix1_crossover_points: int = 0
    timf3_block: int = 0
    fft1_first_point: int = 0
    fft1_last_point: int = 0
    fft1_hz_per_point: float = 0.0
    fftx_points_per_hz: float = 0.0
    mix1_lowest_fq: float = 0.0
    mix1_highest_fq: float = 0.0
    window: np.ndarray = field(default=None, repr=False)
    filtercorr: np.ndarray = field(default=None, repr=False)
    desired: np.ndarray = field(default=None, repr=False)
    mix1_fqwin: np.ndarray = field(default=None, repr=False)
    mix1_window: np.ndarray = field(default=None, repr=False)
    mix1_cos2win: np.ndarray = field(default=None, repr=False)
    mix1_sin2win: np.ndarray = field(default=None, repr=False)

    def __post_init__(self):
        iq = bool(self.input_mode & IQ_DATA)
        N = 1 << self.fft1_n
        self.fft1_size = N
        mm = 2 * self.rf_channels
        self.fft1_block = mm * N
        ad_channels = (2 if iq else 1) * self.rf_channels
        self.frame_bytes = 2 * ad_channels * (2 if self.input_mode & DWORD_INPUT else 1)   # buf.c:296-297
        ratio = make_interleave_ratio(self.sinpow)
        self.mix1_n = max(self.fft1_n - self.mix1_red_n, 3)                             # buf.c:315-322
        M = 1 << self.mix1_n
        self.mix1_size = M
        self.mix1_interleave_points = int(f32(ratio * f32(M))) & 0xfffffffe
        self.fft1_interleave_points = self.mix1_interleave_points * (N // M)             # buf.c:327
        self.fft1_new_points = N - self.fft1_interleave_points
        self.mix1_new_points = M - self.mix1_interleave_points
        self.timf1_blockbytes = self.fft1_new_points * self.frame_bytes * (1 if iq else 2)  # buf.c:601-608
        self.timf3_block = mm * self.mix1_new_points                                      # buf.c:657
        self.fft1_hz_per_point = float(f32(f32(self.ad_speed) / f32(N)))
        if not iq:
            self.fft1_hz_per_point = float(f32(f32(self.fft1_hz_per_point) / f32(2)))
        self.fftx_points_per_hz = float(f32(f32(1) / f32(self.fft1_hz_per_point)))
        # set_fft1_endpoints, fft1.c:4607-4651
        if self.xpoints < 0:
            self.xpoints = N
        wg_first = self.first_xpoint
        wg_last = min(self.first_xpoint + self.xpoints, N - 1)
        if self.second_fft:
            self.fft1_first_point, self.fft1_last_point = 0, N - 1
        else:
            self.fft1_first_point, self.fft1_last_point = wg_first, wg_last
        self.mix1_lowest_fq = float(f32(f32(self.fft1_first_point + 1) * f32(self.fft1_hz_per_point)))   # wide_graph.c:1336-1341
        self.mix1_highest_fq = float(f32(f32(self.fft1_last_point - 1) * f32(self.fft1_hz_per_point)))
        # tables
        if not self.sinpow:
            self.window = None
        elif iq:
            self.window = make_window(4, N, self.sinpow)
        else:
            # real input = fft1 version 2 (fft1_re.c): make_window(2,N,..) is the first half of a
            # 2N-point window and sample 2N-1-i shares w[i] with sample i (fft1_re.c:48-57)
            half = make_window(2, N, self.sinpow)
            self.window = np.concatenate([half[:N], half[:N][::-1]]).astype(f32)
        # version 2 is fft_cntrl row {2,2,...}: permute==2, real2complex==0 (fft1var.c:46)
        self.filtercorr, self.desired = clear_fft1_filtercorr(N, self.rf_channels, self.input_mode, self.fft1_gain,
                                                              permute2=not iq, real2complex=False)
        self.mix1_fqwin = make_window(5, M, 4)
        self._prepare_mixer()

    def _prepare_mixer(self):
        """buf.c:55-111"""
        M, Mi, Mn = self.mix1_size, self.mix1_interleave_points, self.mix1_new_points
        self.mix1_crossover_points = 0
        if self.sinpow in (0, 2):
            return
        win = make_window(3, M, self.sinpow)
        self.mix1_window = win
        if self.sinpow == 9:
            cross = M // 8
        elif self.sinpow == 8:
            cross = M // 16
        else:
            i = Mi // 2
            t1 = win[i]
            cross = 0
            while win[i] < f32(30) * t1 and i > 0:
                i -= 1
                cross += 1
            if cross > 0.75 * Mn:
                cross = int(0.75 * Mn)
            if cross > Mi // 2:
                cross = Mi // 2
        self.mix1_crossover_points = cross
        cos2 = np.zeros(max(cross, 1), f32)
        sin2 = np.zeros(max(cross, 1), f32)
        t1 = f32(0.25 * PI_L / cross) if cross else f32(np.inf)      # C float division, buf.c:97; the loops below are empty then
        j = (M - Mn) // 2
        k = j + cross // 2
        j -= cross // 2
        for i in range(cross):
            cos2[i] = f32(float(win[k]) * math.pow(math.cos(float(t1)), 2.0))
            sin2[i] = f32(float(win[j]) * math.pow(math.sin(float(t1)), 2.0))
            k -= 1
            j += 1
            t1 = f32(float(t1) + 0.5 * PI_L / cross)
        self.mix1_cos2win, self.mix1_sin2win = cos2, sin2

    def waterfall_yfac(self, xpoints_per_pixel=1):
        """make_wg_yfac, wide_graph.c:956-1001 (fft1 waterfall branch)."""
        t1 = f32(f32(FFT1_WATERFALL_ZERO) / f32(self.waterfall_avgnum))
        if xpoints_per_pixel > 1:
            t1 = f32(t1 * f32(self.rf_channels))
        with np.errstate(divide="ignore"):
            y = np.where(self.desired > f32(0.3162278),
                         (t1 / np.power(self.desired.astype(np.float64), 2.0).astype(f32)).astype(f32),
                         f32(t1 * f32(10)))
        y = y.astype(f32)
        y[0] = t1
        y[-1] = t1
        return y

    def selfreq_for_bin(self, fbin):
        """Hz from the lower band edge for a (fractional) fft1 bin."""
        return float(fbin) * self.fft1_hz_per_point
