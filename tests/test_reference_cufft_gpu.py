"""The reference's OWN GPU path (fft_cntrl row 19 "CUDA", fft1.c:3531-3553: fft1win_gpu on the CPU,
cudaMemcpy, cufftExecC2C of 2^gpu.fft1_batch_n transforms, cudaMemcpy back, half swap + re/im swap
fft1.c:3586-3597), compiled from the reference's files with its -DHAVE_CUFFT=1 (oracle/Makefile,
_ref/libref_cufft.so) and run on the same B200: a second, independent realisation of the path by the
reference's author.  Its spectrum is the CPU rows' spectrum times the constant i: cuFFT's forward transform
with re/im swapped is i*conj(X), the CPU rows deliver conj(X) (fft1_direction > 0) -- a global phase that
nothing downstream of fft1 sees (powers, mix1 -> timf3 carries the same constant).  Up to that constant it has
to agree with the reference's CPU row 6 and with this library."""
import numpy as np
import pytest

from linrad_b200 import sizing
from linrad_b200.synth import make_timf1
from oracle import refwrap
from tests.helpers import CONFIGS, CudaStream, rel_rms, parity_record, power_plain_figures

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(not (refwrap.available() and refwrap.cufft_available()), reason="oracle/_ref not built")
@pytest.mark.parametrize("fft1_n,batch_n", [(13, 4), (16, 2)])
def test_reference_cuda_row_agrees_with_cpu_row_and_with_us(fft1_n, batch_n):
    kw = dict(CONFIGS["cfg1"], fft1_n=fft1_n)
    kw.pop("version")
    s = sizing.PathSetup(**kw)
    batch = 1 << batch_n
    nblocks = 2 * batch
    raw = make_timf1(s.input_mode, 1, s.fft1_size, nblocks, s.fft1_new_points, seed=77)
    rawb = np.ascontiguousarray(raw).view(np.uint8).reshape(-1)[: nblocks * s.timf1_blockbytes]
    hz = s.ad_speed / s.fft1_size
    sel = 0.31 * s.fft1_size + 0.37

    g = refwrap.RefOracle(fft1_version=19, n_sel=1, cufft=True, gpu_batch_n=batch_n, **kw)
    assert g.muln == batch and g.timf1_blockbytes == batch * s.timf1_blockbytes
    g.set_selfreq(0, sel * hz)
    out_g = g.process(rawb)
    sumsq_g = g.sumsq()

    c = refwrap.RefOracle(fft1_version=6, n_sel=1, max_fft1n=8, **kw)
    c.set_selfreq(0, sel * hz)
    out_c = c.process(rawb)
    sumsq_c = c.sumsq()

    cs = CudaStream(s, [sel])
    got = cs.process(raw, nblocks, chunk=4)
    cs.close()

    def times_i(a):                      # interleaved re, im -> i * (re + i im)
        z = np.asarray(a).reshape(a.shape[:-1] + (-1, 2))
        return np.stack([-z[..., 1], z[..., 0]], axis=-1).reshape(a.shape)

    assert rel_rms(out_g["fft1"], out_c["fft1"]) > 1.0          # the constant is really there
    out_g["fft1"] = times_i(times_i(times_i(out_g["fft1"])))     # divide by i
    out_g["timf3"] = times_i(times_i(times_i(out_g["timf3"])))
    e_gc = rel_rms(out_g["fft1"], out_c["fft1"])
    e_ug = rel_rms(got["fft1"], out_g["fft1"])
    e_uc = rel_rms(got["fft1"], out_c["fft1"])
    t_gc = rel_rms(out_g["timf3"][:, 0], out_c["timf3"][:, 0])
    t_ug = rel_rms(got["timf3"][:, 0], out_g["timf3"][:, 0])
    rows = nblocks // s.avg1num
    n = rows * s.fft1_size
    w_gc, f_gc = power_plain_figures(sumsq_g[:n], sumsq_c[:n])
    w_ug, f_ug = power_plain_figures(cs.sumsq[:n], sumsq_g[:n])
    parity_record(case=f"reference cuda row 19 N=2^{fft1_n} batch {batch}", fft1_ref_gpu_vs_ref_cpu=e_gc, fft1_ours_vs_ref_gpu=e_ug,
                  fft1_ours_vs_ref_cpu=e_uc, timf3_ref_gpu_vs_ref_cpu=t_gc, timf3_ours_vs_ref_gpu=t_ug,
                  sumsq_worst_plain_ref_gpu_vs_ref_cpu=w_gc, sumsq_frac_over_ref_gpu_vs_ref_cpu=f_gc,
                  sumsq_worst_plain_ours_vs_ref_gpu=w_ug, sumsq_frac_over_ours_vs_ref_gpu=f_ug)
    # the reference's two realisations agree to float rounding, and so do we with either of them
    assert e_gc <= 2e-6, e_gc
    assert e_ug <= 2e-6 and e_uc <= 2e-6, (e_ug, e_uc)
    assert t_gc <= 5e-5 and t_ug <= 5e-5, (t_gc, t_ug)   # the reference's own two rows differ by 3.3e-5 here
