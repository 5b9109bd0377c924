import torch, time
n = 335544320
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for chunk in (n, n // 4, n // 16):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for r in range(5):
        for o in range(0, n, chunk):
            d[o:o + chunk].copy_(h[o:o + chunk], non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    print("H2D pinned 1-D, chunk %d MB: %.1f GB/s" % (chunk >> 20, n / dt / 1e9))
# with a concurrent D2H of 10 % of the bytes on another stream
s2 = torch.cuda.Stream()
h2 = torch.empty(n // 10, dtype=torch.uint8).pin_memory()
d2 = torch.empty(n // 10, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()
t0 = time.perf_counter()
for r in range(5):
    with torch.cuda.stream(s2):
        h2.copy_(d2, non_blocking=True)
    d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 5
print("H2D with concurrent D2H: %.1f GB/s" % (n / dt / 1e9))
