"""Linrad .raw recordings: the header logic lives in the C ABI (lb200_raw_header_parse, restating
open_savefile modesub.c:656-733); this module is the harness side used by tests/ and bench.py --
a writer that lays the bytes out like the reference's writer (modesub.c:1519-1640) for synthetic
recordings, the 18-bit packer (compress_rawdat, getiq64.s:39-96) and a block iterator for
playback.  Payload of a DWORD_INPUT recording is 18-bit packed; the expansion to left-justified
int32 timf1 words is lb200_expand_rawdat (GPU)."""
import ctypes as C
import struct

import numpy as np

from . import api

REMEMBER_UNKNOWN, REMEMBER_NOTHING, REMEMBER_PERSEUS, REMEMBER_SDR14 = -1, -2, -3, -4
DWORD_INPUT, TWO_CHANNELS, IQ_DATA = 1, 2, 4


def pack_18bit(words):
    """int32 timf1 words (multiple of 4) -> 9 bytes per 4 words: the upper 16 bits of each word
    little-endian, then one byte holding bits 15..14 of the four words (compress_rawdat,
    getiq64.s:39-96; inverse of expand_rawdat up to the +0x2000 half-LSB offset it adds)."""
    w = np.ascontiguousarray(words, np.int32).view(np.uint32).reshape(-1, 4)
    out = np.zeros((w.shape[0], 9), np.uint8)
    hi = (w >> 16).astype(np.uint16)
    out[:, 0:8] = hi.view(np.uint8).reshape(-1, 8)
    two = ((w >> 14) & 3).astype(np.uint8)
    out[:, 8] = two[:, 0] | (two[:, 1] << 2) | (two[:, 2] << 4) | (two[:, 3] << 6)
    return out.reshape(-1)


def header_bytes(rx_input_mode, rx_rf_channels, rx_ad_channels, rx_ad_speed, *, old_format=False,
                 remember_tag=REMEMBER_NOTHING, chunk=b"", diskread_time=0.0, passband_center=0.0,
                 passband_direction=1, save_init_flag=0):
    """The header as the reference's writer emits it (modesub.c:1519-1601); the mode word keeps
    only TWO_CHANNELS, DWORD_INPUT, IQ_DATA and DIGITAL_IQ (modesub.c:1575)."""
    mode = rx_input_mode & (TWO_CHANNELS + DWORD_INPUT + IQ_DATA + 32)
    b = b""
    if old_format:
        b += struct.pack("<i", mode)
    else:
        b += struct.pack("<i", remember_tag)
        if remember_tag in (REMEMBER_PERSEUS, REMEMBER_SDR14):
            b += struct.pack("<i", len(chunk)) + chunk
        b += struct.pack("<ddii", diskread_time, passband_center, passband_direction, mode)
    b += struct.pack("<iiiB", rx_rf_channels, rx_ad_channels, rx_ad_speed, save_init_flag)
    return b


def write_raw(path, frames, rx_input_mode, rx_rf_channels, rx_ad_speed, **hdr):
    """frames: the timf1 words of the recording (int16, or left-justified int32 for DWORD_INPUT)."""
    iq = bool(rx_input_mode & IQ_DATA)
    ad_channels = rx_rf_channels * (2 if iq else 1)
    with open(path, "wb") as f:
        f.write(header_bytes(rx_input_mode, rx_rf_channels, ad_channels, rx_ad_speed, **hdr))
        if rx_input_mode & DWORD_INPUT:
            f.write(pack_18bit(np.asarray(frames, np.int32).reshape(-1)).tobytes())
        else:
            f.write(np.ascontiguousarray(frames, np.int16).tobytes())


def parse_header(data):
    """bytes -> api.RawHeader through the C ABI; raises api.Lb200Error on a corrupted header."""
    lib = api.load_library()
    h = api.RawHeader()
    buf = np.frombuffer(data, np.uint8)
    rc = lib.lb200_raw_header_parse(buf.ctypes.data, buf.size, C.byref(h))
    if rc:
        raise api.Lb200Error(rc, "lb200_raw_header_parse")
    return h


def blocks(path, block_bytes, plan=None):
    """Yield consecutive timf1 blocks of `block_bytes` bytes from a recording (uint8 arrays).
    DWORD_INPUT recordings are expanded on the GPU through `plan` (lb200_expand_rawdat)."""
    lib = api.load_library()
    with open(path, "rb") as f:
        head = f.read(4096)
        h = parse_header(head)
        if h.save_init_flag:
            raise NotImplementedError("recordings that carry calibration data (save_init_flag != 0)")
        f.seek(h.payload_offset)
        fb = lib.lb200_raw_block_bytes(C.byref(h), block_bytes)
        while True:
            chunk = f.read(fb)
            if len(chunk) < fb:
                return
            a = np.frombuffer(chunk, np.uint8)
            if h.rx_input_mode & DWORD_INPUT:
                if plan is None:
                    raise ValueError("an 18-bit recording needs a plan for lb200_expand_rawdat")
                a = api.expand_rawdat_host(plan, a, block_bytes).view(np.uint8)
            yield a
