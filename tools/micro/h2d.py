import torch, time
n = 398131200
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for chunk in (n, n // 16, n // 64):
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        t0 = time.perf_counter()
        for off in range(0, n, chunk):
            d[off:off + chunk].copy_(h[off:off + chunk], non_blocking=True)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    print("H2D pinned, %d copies of %.1f MB: %.2f GB/s" % (n // chunk, chunk / 1e6, n / best / 1e9))
# concurrent D2H of a small buffer does not matter; test bidirectional anyway
h2 = torch.empty(4 << 20, dtype=torch.uint8).pin_memory()
s2 = torch.cuda.Stream()
torch.cuda.synchronize()
t0 = time.perf_counter()
for off in range(0, n, n // 16):
    d[off:off + n // 16].copy_(h[off:off + n // 16], non_blocking=True)
    with torch.cuda.stream(s2):
        h2.copy_(d[:4 << 20], non_blocking=True)
torch.cuda.synchronize()
print("with 16 concurrent 4 MB D2H: %.2f GB/s" % (n / (time.perf_counter() - t0) / 1e9))
