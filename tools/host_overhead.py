"""Host-side cost of one render() call (tiny mesh, so the GPU is never the bottleneck) and config D style
batches (many meshes, one render() per mesh)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench, worldrenderer_b200 as wr
from worldrenderer_b200 import synth
dev = torch.device("cuda", 0)
cam = wr.get_orthogonal_camera(device="cuda:0", **synth.CANONICAL_RIG)
ctx = wr.NVDiffRastContextWrapper("cuda:0", "cuda")
def mk(v, f):
    m = wr.TexturedMesh(v_pos=torch.tensor(v, dtype=torch.float32, device=dev), t_pos_idx=torch.tensor(f, device=dev)); m.set_stitched_mesh(m.v_pos, m.t_pos_idx); m.v_nrm; return m
tiny = mk(*synth.icosphere(2, 0.5))
for res in (64, 768):
    for _ in range(20): wr.render(ctx, tiny, cam, res, res, render_attr=False)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    n = 300
    for _ in range(n): wr.render(ctx, tiny, cam, res, res, render_attr=False)
    t_issue = time.perf_counter() - t0
    torch.cuda.synchronize(); t_all = time.perf_counter() - t0
    print(f"tiny mesh {res}^2: host issue {t_issue / n * 1e6:.1f} us/call, wall {t_all / n * 1e6:.1f} us/call")
meshes = [mk(*synth.icosphere(50, 0.5)) for _ in range(8)]
for _ in range(3):
    for m in meshes: wr.render(ctx, m, cam, 768, 768, render_attr=False)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5):
    for m in meshes: wr.render(ctx, m, cam, 768, 768, render_attr=False)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"config D style, 8 x 50k-face meshes x 6 views: {5 * 8 * 6 / dt:.0f} views/s, {dt / 40 * 1e6:.0f} us per mesh")
ctx.ctx.profile(True)
acc = {}
for k in range(10):
    wr.render(ctx, tiny, cam, 768, 768, render_attr=False)
    for n, ms in ctx.ctx.profile_read(): acc.setdefault(n, []).append(ms * 1e3)
ctx.ctx.profile(False)
print("tiny mesh 768^2 stages (us):", {n: round(float(np.mean(x[2:])), 1) for n, x in acc.items()})
