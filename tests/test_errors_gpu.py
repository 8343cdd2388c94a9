"""Error behaviour of the C ABI: everything the kernels do not cover fails loudly with the code
the header documents (there is no silent fallback of any kind)."""
import numpy as np
import pytest

from linrad_b200 import api, sizing
from tests.helpers import IQ_DATA, DWORD_INPUT, TWO_CHANNELS

pytestmark = pytest.mark.gpu

UNSUPPORTED, BAD_ARG, BAD_CONFIG = 3103, 3104, 3102


def _setup(**kw):
    base = dict(input_mode=IQ_DATA, rf_channels=1, ad_speed=96000, fft1_n=11, mix1_red_n=3)
    base.update(kw)
    return sizing.PathSetup(**base)


def _code(fn):
    with pytest.raises(api.Lb200Error) as e:
        fn()
    return e.value.code


def test_create_rejects_what_is_not_built():
    # mix1.size above what is built (documented limit: 16384 one channel, 8192 two; the reference allows 32768)
    assert _code(lambda: api.Plan(_setup(fft1_n=17, mix1_red_n=2))) == UNSUPPORTED
    assert _code(lambda: api.Plan(_setup(input_mode=IQ_DATA | TWO_CHANNELS, rf_channels=2, fft1_n=16, mix1_red_n=2))) == UNSUPPORTED
    # IQ-only options on real input / one channel
    fold = np.zeros(2 * 2048, np.float32)
    assert _code(lambda: api.Plan(_setup(input_mode=0), foldcorr=fold)) == UNSUPPORTED
    assert _code(lambda: api.Plan(_setup(), pg_ch2=(0.9, 0.1))) == UNSUPPORTED
    # sizes outside 2^7 .. 2^20
    assert _code(lambda: api.Plan(_setup(fft1_n=6, mix1_red_n=2))) in (UNSUPPORTED, BAD_CONFIG)
    # ui.sample_shift is ignored for two channels and real input, exactly like the reference
    api.Plan(_setup(input_mode=IQ_DATA | TWO_CHANNELS, rf_channels=2), sample_shift=3).close()
    api.Plan(_setup(input_mode=0), sample_shift=-2).close()


def test_calls_reject_bad_arguments():
    s = _setup()
    plan = api.Plan(s)
    try:
        N = s.fft1_size
        timf1 = np.zeros(8 * N * 4, np.uint8)
        fft1 = np.zeros(8 * s.fft1_block, np.float32)
        sumsq = np.zeros(16 * N, np.float32)
        ok = dict(timf1=timf1, ref=0, nblocks=2, fft1=fft1, sumsq=sumsq)
        plan.fft1_host(**ok)
        assert _code(lambda: plan.fft1_host(**dict(ok, nblocks=9))) == BAD_ARG                  # more transforms than the ring holds
        assert _code(lambda: plan.fft1_host(**dict(ok, fft1_pa=5))) == BAD_ARG                  # not on a transform boundary
        assert _code(lambda: plan.fft1_host(**dict(ok, fft1=np.zeros(3 * s.fft1_block, np.float32)))) == BAD_ARG   # ring not a power of two
        assert _code(lambda: plan.fft1_host(**dict(ok, counter=s.avg1num))) == BAD_ARG          # fft1_sumsq_counter out of range
        assert _code(lambda: plan.fft1_host(**dict(ok, corrsum=np.zeros(2 * sumsq.size, np.float32)))) == UNSUPPORTED   # cross spectrum needs two channels
        plan.fft1_host(**dict(ok, nblocks=0))                                                   # nothing to do is not an error
    finally:
        plan.close()


def test_strerror_covers_the_codes():
    lib = api.load_library()
    for code in (0, 3100, 3101, 3102, 3103, 3104, 1211, 1212):
        assert len(lib.lb200_strerror(code)) > 1


def test_mix1_refuses_a_spectrum_that_exists_nowhere():
    """LB200_FFT1_SPECTRUM_STAYS_ON_DEVICE leaves the host ring unwritten.  If the mirror copy is not the final
    spectrum either (raw fft1_b output: apply_filtercorr = 0), lb200_mix1 must not mix the stale host ring."""
    s = _setup()
    plan = api.Plan(s)
    try:
        N = s.fft1_size
        timf1 = np.zeros(8 * N * 4, np.uint8)
        fft1 = np.zeros(8 * s.fft1_block, np.float32)
        sumsq = np.zeros(16 * N, np.float32)
        timf3_size = 16 * s.mix1_size * 2
        timf3 = np.zeros(2 * timf3_size, np.float32)
        states = api.new_states([s.selfreq_for_bin(300.4)])
        mix = dict(fft1=fft1, fft1_px=0, nblocks=2, states=states, timf3=timf3, timf3_floats=timf3_size, timf3_pa=0)
        # kept on the device, filter-corrected: fine
        plan.fft1_host(timf1=timf1, ref=0, nblocks=2, fft1=fft1, sumsq=sumsq, keep_on_device=True)
        plan.mix1_host(**mix)
        # kept on the device but raw: the blocks are in neither place
        plan.fft1_host(timf1=timf1, ref=0, nblocks=2, fft1=fft1, apply_fc=False, keep_on_device=True)
        assert _code(lambda: plan.mix1_host(**mix)) == BAD_ARG
        # written to the host ring again: fine again
        plan.fft1_host(timf1=timf1, ref=0, nblocks=2, fft1=fft1, sumsq=sumsq)
        plan.mix1_host(**mix)
    finally:
        plan.close()
