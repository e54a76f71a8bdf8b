"""One-directional HBM bandwidth of this B200: pure read (reductions over 4 GiB), pure write (memset), copy.  The HBM roofline of
MEASURED_PEAKS.json is a copy figure (bytes read + bytes written); kernels that only read (the 1D-CNN, the EER histograms) or only
write (conv1) are bound by the one-directional figures printed here."""
import torch

def rate(name, fn, nbytes, reps=8):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    print(f"{name:48s} {reps * nbytes / (e0.elapsed_time(e1) * 1e-3) / 1e12:6.2f} TB/s", flush=True)

n = 1 << 30                                   # 4 GiB of fp32
x = torch.ones(n, dtype=torch.float32, device="cuda")
y = torch.empty_like(x)
rate("read : torch.sum(fp32), 4 GiB", lambda: x.sum(), 4 * n)
rate("read : torch.max(fp32), 4 GiB", lambda: x.max(), 4 * n)
rate("read : sum of a bf16 view, 4 GiB", lambda: x.view(torch.bfloat16).sum(), 4 * n)
rate("read : torch.count_nonzero(int32 view), 4 GiB", lambda: torch.count_nonzero(x.view(torch.int32)), 4 * n)
rate("write: memset (zero_), 4 GiB", lambda: y.zero_(), 4 * n)
rate("write: fill_(1.0), 4 GiB", lambda: y.fill_(1.0), 4 * n)
rate("copy : y.copy_(x), 4 + 4 GiB touched", lambda: y.copy_(x), 8 * n)
rate("r+w  : x.mul_(1.0001) in place, 4 + 4 GiB", lambda: x.mul_(1.0001), 8 * n)
