import torch, time
n = 2 * 1024**3 // 4
h = torch.empty(n, dtype=torch.float32, pin_memory=True); h.fill_(1.0)
d = torch.empty(n, dtype=torch.float32, device="cuda")
for size_mb in (96, 512, 2048):
    m = size_mb * 1024**2 // 4
    for _ in range(2):
        d[:m].copy_(h[:m], non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = max(2, 4096 // size_mb)
    for _ in range(reps):
        d[:m].copy_(h[:m], non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"H2D {size_mb} MB x{reps}: {reps * m * 4 / dt / 1e9:.1f} GB/s")
# two streams concurrently
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
m = 512 * 1024**2 // 4
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(4):
    with torch.cuda.stream(s1): d[:m].copy_(h[:m], non_blocking=True)
    with torch.cuda.stream(s2): d[m:2*m].copy_(h[m:2*m], non_blocking=True)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"H2D 2 streams: {8 * m * 4 / dt / 1e9:.1f} GB/s")
