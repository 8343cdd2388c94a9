"""Linrad .raw recordings: header logic of open_savefile (modesub.c:656-733) in the C ABI, the
harness-side writer / 18-bit packer, and playback of a recording into the GPU path."""
import struct

import numpy as np
import pytest

from linrad_b200 import api, rawfile, sizing
from linrad_b200.synth import make_timf1
from oracle import port

IQ, DW, TWO = sizing.IQ_DATA, sizing.DWORD_INPUT, sizing.TWO_CHANNELS


def test_header_new_format_round_trip():
    b = rawfile.header_bytes(IQ, 1, 2, 96000, diskread_time=4711.0, passband_center=144.3, passband_direction=-1)
    h = rawfile.parse_header(b + b"\x01\x02")
    assert (h.remember_tag, h.chunk_size) == (rawfile.REMEMBER_NOTHING, 0)
    assert (h.diskread_time, h.passband_center, h.passband_direction) == (4711.0, 144.3, -1)
    assert (h.rx_input_mode, h.rx_rf_channels, h.rx_ad_channels, h.rx_ad_speed) == (IQ, 1, 2, 96000)
    assert h.save_init_flag == 0 and h.payload_offset == len(b) == 4 + 8 + 8 + 4 + 4 + 4 + 4 + 4 + 1


def test_header_old_format_and_two_channel_bit():
    # old files start with rx_input_mode itself; TWO_CHANNELS comes from rx_rf_channels (modesub.c:719-721)
    b = struct.pack("<iiiiB", IQ | DW, 2, 4, 192000, 0)
    h = rawfile.parse_header(b)
    assert h.remember_tag == rawfile.REMEMBER_NOTHING and h.passband_direction == 1 and h.diskread_time == 0
    assert h.rx_input_mode == (IQ | DW | TWO) and h.rx_ad_channels == 4 and h.payload_offset == 17
    lib = api.load_library()
    import ctypes as C
    assert lib.lb200_raw_block_bytes(C.byref(h), 65536) == 18 * 65536 // 32        # buf.c:599
    h2 = rawfile.parse_header(rawfile.header_bytes(IQ, 1, 2, 96000))
    assert lib.lb200_raw_block_bytes(C.byref(h2), 16384) == 16384


def test_header_proprietary_chunk():
    chunk = bytes(range(40))
    b = rawfile.header_bytes(IQ | DW, 1, 2, 2000000, remember_tag=rawfile.REMEMBER_PERSEUS, chunk=chunk)
    h = rawfile.parse_header(b)
    assert h.remember_tag == rawfile.REMEMBER_PERSEUS and h.chunk_size == 40 and h.chunk_offset == 8
    assert b[h.chunk_offset: h.chunk_offset + h.chunk_size] == chunk
    assert h.rx_ad_speed == 2000000 and h.payload_offset == len(b)


@pytest.mark.parametrize("bad", ["short", "tag", "direction", "mode", "adch", "adch_mismatch", "chunk"])
def test_header_corrupted(bad):
    good = rawfile.header_bytes(IQ, 1, 2, 96000)
    if bad == "short":
        b = good[:-1]
    elif bad == "tag":
        b = struct.pack("<i", -9) + good[4:]
    elif bad == "direction":
        b = good[:20] + struct.pack("<i", 0) + good[24:]
    elif bad == "mode":
        b = good[:24] + struct.pack("<i", 256) + good[28:]
    elif bad == "adch":
        b = good[:32] + struct.pack("<i", 5) + good[36:]
    elif bad == "adch_mismatch":
        b = good[:28] + struct.pack("<ii", 1, 3) + good[36:]
    else:
        b = struct.pack("<ii", rawfile.REMEMBER_SDR14, 1000) + b"\x00" * 20
    with pytest.raises(api.Lb200Error) as e:
        rawfile.parse_header(b)
    assert e.value.code == 3104


def test_pack_18bit_is_the_reference_packer():
    rng = np.random.default_rng(3)
    words = (rng.integers(-2 ** 23, 2 ** 23, 4096, dtype=np.int64) << 8).astype(np.int32)
    assert np.array_equal(rawfile.pack_18bit(words), port.compress_rawdat(words))


def test_int16_recording_blocks(tmp_path):
    s = sizing.PathSetup(input_mode=IQ, rf_channels=1, ad_speed=96000, fft1_n=10, mix1_red_n=3)
    raw = make_timf1(s.input_mode, 1, s.fft1_size, 6, s.fft1_new_points, seed=2)
    path = str(tmp_path / "a.raw")
    rawfile.write_raw(path, raw, s.input_mode, 1, s.ad_speed, passband_direction=-1)
    got = list(rawfile.blocks(path, s.timf1_blockbytes))
    want = np.ascontiguousarray(raw).view(np.uint8).reshape(-1)
    assert len(got) == 6 and np.array_equal(np.concatenate(got), want[: 6 * s.timf1_blockbytes])


@pytest.mark.gpu
def test_18bit_recording_playback_to_spectrum(tmp_path):
    """BASELINE configs[0]-style playback: .raw file -> header -> 18-bit expansion on the GPU ->
    fft1, against the oracle fed the words expand_rawdat must produce."""
    from oracle import refwrap
    from tests.helpers import CudaStream, rel_rms, run_reference
    s = sizing.PathSetup(input_mode=IQ | DW, rf_channels=1, ad_speed=96000, fft1_n=11, mix1_red_n=3)
    nblocks = 6
    raw = make_timf1(s.input_mode, 1, s.fft1_size, nblocks, s.fft1_new_points, seed=4)
    path = str(tmp_path / "b.raw")
    rawfile.write_raw(path, raw, s.input_mode, 1, s.ad_speed)
    cs = CudaStream(s, [])
    try:
        h = rawfile.parse_header(open(path, "rb").read(64))
        assert h.rx_input_mode == (IQ | DW) and h.rx_ad_speed == 96000
        blocks = list(rawfile.blocks(path, s.timf1_blockbytes, plan=cs.plan))
        assert len(blocks) == nblocks
        words = np.concatenate(blocks).view(np.int32)
        want = port.expand_rawdat(port.compress_rawdat(np.asarray(raw, np.int32).reshape(-1)[: words.size]), words.nbytes)
        assert np.array_equal(words, want)                       # bit-exact codec
        got = cs.process(words, nblocks, chunk=3, mix=False)
        if refwrap.available():
            ref = run_reference(dict(input_mode=IQ | DW, rf_channels=1, ad_speed=96000, fft1_n=11, mix1_red_n=3, version=6),
                                words, [], nblocks)
            assert rel_rms(got["fft1"], ref["fft1"]) <= 1e-5
    finally:
        cs.close()
