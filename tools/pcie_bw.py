import torch, time
for mb in (8, 64, 256):
    n = mb * (1 << 20)
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    for direction in ("d2h", "h2d"):
        for _ in range(3):
            (h.copy_(d, non_blocking=True) if direction == "d2h" else d.copy_(h, non_blocking=True))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        K = 20
        for _ in range(K):
            (h.copy_(d, non_blocking=True) if direction == "d2h" else d.copy_(h, non_blocking=True))
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / K
        print("%s %4d MB: %.1f GB/s" % (direction, mb, n / dt / 1e9))
# both directions at once
n = 64 << 20
d1 = torch.empty(n, dtype=torch.uint8, device="cuda"); h1 = torch.empty(n, dtype=torch.uint8).pin_memory()
d2 = torch.empty(n // 4, dtype=torch.uint8, device="cuda"); h2 = torch.empty(n // 4, dtype=torch.uint8).pin_memory()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    with torch.cuda.stream(s1): h1.copy_(d1, non_blocking=True)
    with torch.cuda.stream(s2): d2.copy_(h2, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 20
print("duplex: d2h 64 MB + h2d 16 MB in %.1f us -> d2h %.1f GB/s" % (dt * 1e6, n / dt / 1e9))
