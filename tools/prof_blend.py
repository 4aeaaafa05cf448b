"""Times the atlas post-processing kernels (csrc/blend.cu) with the library's per-stage CUDA events.
usage: python tools/prof_blend.py [--size 1024] [--iters 1000]"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import worldrenderer_b200 as wr  # noqa: E402
from worldrenderer_b200 import _native  # noqa: E402
from worldrenderer_b200.uv import uv_padding  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=1024)
ap.add_argument("--iters", type=int, default=1000)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--mask", default="islands", choices=["islands", "charts"])
a = ap.parse_args()
dev = torch.device("cuda", 0)
H = W = a.size
g = torch.Generator().manual_seed(0)
src = torch.rand((H, W, 3), generator=g).to(dev)
tgt = torch.rand((H, W, 3), generator=g).to(dev)
yy, xx = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
if a.mask == "islands":
    mask = (((yy // 37 + xx // 53) % 3) != 0).to(dev)  # ~2/3 of the atlas in the solve region, many islands
else:
    mask = (((yy % 256) >= 12) & ((xx % 340) >= 12)).to(dev)  # atlas-like: large charts separated by 12-texel gutters
solver = wr.PoissonBlendingSolver("torch-cuda", str(dev))
ctx = solver._ctx
ctx.profile(True)
for name, fn in (("poisson", lambda: solver(src, mask, tgt, a.iters, inplace=False)),):
    ts = []
    for _ in range(a.reps):
        fn()
        torch.cuda.synchronize()
        ts.append(dict(ctx.profile_read()))
    best = min(ts, key=lambda d: sum(d.values()))
    tot = sum(best.values())
    n = H * W * 3 * float(mask.float().mean())
    print(f"{name} {H}x{W}x3, mask={a.mask} ({float(mask.float().mean()):.2f} of the atlas), {a.iters} sweeps: {tot:.3f} ms total; stages {best}")
    print(f"  {n * a.iters / (best['k_pb_jacobi'] * 1e-3) / 1e12:.2f} T point-sweeps/s over the solve region; "
          f"{a.iters / 8:.0f} launches of {best['k_pb_jacobi'] / (a.iters / 8) * 1e3:.1f} us")
dctx = _native.default_context(dev)
dctx.profile(True)
ts = []
for _ in range(a.reps):
    uv_padding(src, mask, 3)
    torch.cuda.synchronize()
    ts.append(dict(dctx.profile_read()))
best = min(ts, key=lambda d: sum(d.values()))
print(f"uv_padding {H}x{W}x3 radius 3: {sum(best.values()):.3f} ms; stages {best}")
