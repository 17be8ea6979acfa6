import torch, time
for mb in (64, 400):
    n = mb * 1024 * 1024
    h = torch.empty(n, dtype=torch.uint8).pin_memory(); d = torch.empty(n, dtype=torch.uint8, device="cuda")
    for name, (a, b) in {"H2D": (d, h), "D2H": (h, d)}.items():
        for _ in range(2): a.copy_(b, non_blocking=True)
        torch.cuda.synchronize(); t = time.perf_counter()
        for _ in range(5): a.copy_(b, non_blocking=True)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 5
        print(mb, "MB", name, "%.1f GB/s" % (n / dt / 1e9))
# bidirectional
h1 = torch.empty(400<<20, dtype=torch.uint8).pin_memory(); h2 = torch.empty(400<<20, dtype=torch.uint8).pin_memory()
d1 = torch.empty(400<<20, dtype=torch.uint8, device="cuda"); d2 = torch.empty(400<<20, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(5):
    with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 5
print("bidirectional 400 MB each way: %.1f GB/s per direction" % ((400<<20) / dt / 1e9))
