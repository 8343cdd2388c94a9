"""Quick device probes used to calibrate design choices (not part of the product):
L2-resident vs HBM copy bandwidth with torch copies."""
import torch, time
dev = torch.device("cuda")
def bw(nbytes, iters):
    a = torch.empty(nbytes // 4, dtype=torch.float32, device=dev); b = torch.empty_like(a)
    for _ in range(3): b.copy_(a)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): b.copy_(a)
    e1.record(); torch.cuda.synchronize()
    return 2 * nbytes * iters / (e0.elapsed_time(e1) * 1e-3) / 1e9
for mb in (4, 8, 16, 32, 48, 64, 128, 256, 1024):
    print(f"copy {mb} MiB src (+{mb} dst): {bw(mb << 20, 50):.0f} GB/s (read+write)")
p = torch.cuda.get_device_properties(0)
print(p.name, p.multi_processor_count, "SMs, L2", p.L2_cache_size >> 20, "MiB")
