"""HBM bandwidth by access mix (context for the roofline of write-heavy kernels)."""
import torch
dev = torch.device("cuda", 0)
n = 1 << 30
a = torch.empty(n, dtype=torch.uint8, device=dev)
b = torch.empty(n, dtype=torch.uint8, device=dev)
def t(fn, reps=10):
    for _ in range(3): fn()
    best = 1e9
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e))
    return best
af = a.view(torch.float32)
print("copy  (r+w bytes) GB/s", 2 * n / t(lambda: b.copy_(a)) / 1e6)
print("write only        GB/s", n / t(lambda: a.fill_(1)) / 1e6)
print("read only (sum)   GB/s", n / t(lambda: af.sum()) / 1e6)
