"""Pinned host -> device copy rate of this box with 1/2/4 copy streams (context for the e2e number: 243 MB per batch)."""
import torch, time
x = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
d = torch.empty_like(x, device="cuda")
for n in (1, 2, 4):
    streams = [torch.cuda.Stream() for _ in range(n)]
    chunk = x.numel() // n
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        for i, s in enumerate(streams):
            with torch.cuda.stream(s):
                d[i * chunk:(i + 1) * chunk].copy_(x[i * chunk:(i + 1) * chunk], non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 10
    print(f"{n} stream(s): {x.numel() / dt / 1e9:.1f} GB/s")
# with a device-to-host stream running at the same time (the e2e runner returns 2 GB of crops per mosaic)
y = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
e = torch.empty_like(y, device="cuda")
s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    with torch.cuda.stream(s_in):
        d.copy_(x, non_blocking=True)
    with torch.cuda.stream(s_out):
        y.copy_(e, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 10
print(f"H2D with concurrent D2H: {x.numel() / dt / 1e9:.1f} GB/s each way")
