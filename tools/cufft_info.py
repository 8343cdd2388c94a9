"""Informational only: what cuFFT does on the same shapes, device resident.
Not part of the product path and not a parity reference: a plain batched C2C transform of
already-converted complex64 data (no unpack, no window, no filtercorr, no power, no mix1), i.e.
LESS work than fft1_fused_kernel does per transform.  Reports time per batch and the GB/s of the
transform's own compulsory traffic (8N in + 8N out per channel-transform).
usage: python tools/cufft_info.py"""
import json
import torch

def run(n_log2, batch, channels, iters=20):
    N = 1 << n_log2
    x = torch.randn(batch * channels, N, dtype=torch.complex64, device="cuda")
    y = torch.empty_like(x)
    for _ in range(5):
        torch.fft.fft(x, out=y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        torch.fft.fft(x, out=y)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    traffic = 16.0 * N * batch * channels
    return {"fft1_n": n_log2, "transforms": batch, "channels": channels, "ms": ms,
            "GBps_of_16N_bytes": traffic / ms / 1e6, "points_per_us": N * batch * channels / ms / 1e3}

if __name__ == "__main__":
    out = [run(13, 5920, 1), run(14, 2960, 2), run(18, 60, 1)]
    print(json.dumps({"impl": "cufft-informational", "torch": torch.__version__, "results": out}))
