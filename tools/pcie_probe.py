"""PCIe copy bandwidth from / to pinned host memory, with and without binding the process to the GPU's CPUs."""
import os, time, torch
dev = torch.device("cuda", 0)
def bw(tag):
    n = 256 << 20
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    for direction in ("d2h", "h2d"):
        for _ in range(2):
            (h.copy_(d, non_blocking=True) if direction == "d2h" else d.copy_(h, non_blocking=True)); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(8):
            (h.copy_(d, non_blocking=True) if direction == "d2h" else d.copy_(h, non_blocking=True))
        torch.cuda.synchronize()
        print(tag, direction, f"{8 * n / (time.perf_counter() - t0) / 1e9:.1f} GB/s")
print("affinity before:", len(os.sched_getaffinity(0)), "cpus")
bw("default ")
try:
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    pynvml.nvmlDeviceSetCpuAffinity(h)
    print("affinity after nvmlDeviceSetCpuAffinity:", len(os.sched_getaffinity(0)), "cpus", sorted(os.sched_getaffinity(0))[:4], "...")
    bw("gpu-numa")
except Exception as e:
    print("nvml affinity failed:", e)
os.system("nvidia-smi topo -m 2>/dev/null | head -8; lscpu | grep -E 'NUMA|Model name|^CPU\\(s\\)' | head -8")
