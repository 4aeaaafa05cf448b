"""D2H bandwidth from the GPU to pinned host memory: one copy stream vs two concurrent ones, and chunk sizes."""
import time, torch
dev = torch.device("cuda", 0)
n = 100 << 20
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
def run(nstreams, parts):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        step = n // parts
        for p in range(parts):
            st = (s1, s2)[p % nstreams]
            with torch.cuda.stream(st):
                h[p * step:(p + 1) * step].copy_(d[p * step:(p + 1) * step], non_blocking=True)
    torch.cuda.synchronize()
    return 10 * n / (time.perf_counter() - t0) / 1e9
for ns, parts in ((1, 1), (1, 4), (2, 2), (2, 4), (2, 8)):
    run(ns, parts)
    print(f"streams={ns} parts={parts}: {run(ns, parts):.1f} GB/s")
h2 = torch.empty(30 << 20, dtype=torch.uint8).pin_memory(); d2 = torch.empty(30 << 20, dtype=torch.uint8, device=dev)
s3 = torch.cuda.Stream(dev)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10):
    with torch.cuda.stream(s1): h.copy_(d, non_blocking=True)
    with torch.cuda.stream(s3): d2.copy_(h2, non_blocking=True)
torch.cuda.synchronize()
print(f"d2h 100 MB with concurrent h2d 30 MB: {10 * n / (time.perf_counter() - t0) / 1e9:.1f} GB/s d2h")
