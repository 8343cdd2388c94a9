"""Deterministic synthetic timf1 input for the five BASELINE.json configurations
(SURVEY.md section 8(d)): complex tones + Gaussian noise, rounded and clipped to the
device word size, laid out exactly as a Linrad input thread leaves it in timf1
(rxin.c:1143-1437; frame = [I,Q] / [I1,Q1,I2,Q2] / [x], int16 or left-justified int32).
"""
import numpy as np

DWORD_INPUT, TWO_CHANNELS, IQ_DATA = 1, 2, 4   # globdef.h:277-279


def make_timf1(input_mode, rf_channels, fft1_size, nblocks, new_points, seed=1,
               tones=((3000.37, 8000.0), (1234.0, 800.0), (7000.5, 80.0)), noise=30.0):
    """Returns an integer array of nblocks*new_points frames (2x that many samples for real
    input) in timf1 memory layout.  Tone frequencies are in fft1 bins of the unshifted DFT,
    scaled to fft1_size/8192 so every N gets tones at the same relative places."""
    rng = np.random.default_rng(seed)
    iq = bool(input_mode & IQ_DATA)
    dword = bool(input_mode & DWORD_INPUT)
    nsamp = nblocks * new_points * (1 if iq else 2)
    ntime = fft1_size if iq else 2 * fft1_size
    t = np.arange(nsamp, dtype=np.float64)
    scale = fft1_size / 8192.0
    chans = []
    base = None
    for c in range(rf_channels):
        if c == 0 or base is None:
            z = np.zeros(nsamp, np.complex128)
            for fbin, amp in tones:
                z += amp * np.exp(2j * np.pi * (fbin * scale) * t / ntime)
            base = z
        else:
            z = base * 0.7 * np.exp(0.9j)
        z = z + noise * (rng.standard_normal(nsamp) + 1j * rng.standard_normal(nsamp))
        chans.append(z)
    cols = []
    for z in chans:
        if iq:
            cols += [z.real, z.imag]
        else:
            cols += [z.real]
    x = np.stack(cols, axis=1)                      # frames x words
    x = np.clip(np.rint(x), -32768, 32767)
    if dword:
        # 24-bit left-justified in int32: a 16-bit signal scaled to 24 bits, low byte zero
        x = (np.rint(x * 256 + rng.integers(-128, 128, x.shape))).astype(np.int64) << 8
        x = np.clip(x, -2**31, 2**31 - 256).astype(np.int32)
    else:
        x = x.astype(np.int16)
    return np.ascontiguousarray(x)
