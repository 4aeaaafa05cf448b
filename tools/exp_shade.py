"""Marginal-cost experiments for the shading pass (config B shapes)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench, worldrenderer_b200 as wr
from worldrenderer_b200 import synth
from worldrenderer_b200.render import render_geometry_raw
dev = torch.device("cuda", 0)
cam_cpu = wr.get_orthogonal_camera(**synth.CANONICAL_RIG)
cam = wr.Camera(c2w=cam_cpu.c2w.to(dev), w2c=cam_cpu.w2c.to(dev), proj_mtx=cam_cpu.proj_mtx.to(dev), mvp_mtx=cam_cpu.mvp_mtx.to(dev), cam_pos=None)
v, f = bench.terrain_arrays(0)
ctx = wr.NVDiffRastContextWrapper("cuda:0", "cuda")
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
def mk(v, f):
    m = wr.TexturedMesh(v_pos=torch.from_numpy(v).to(dev), t_pos_idx=torch.from_numpy(f).to(dev)); m.set_stitched_mesh(m.v_pos, m.t_pos_idx); m.v_nrm; return m
full = mk(v, f)
empty = mk(v, f[:1] * 0)   # one degenerate face: nothing covered
def run(name, mesh, **kw):
    acc = {}
    ctx.ctx.profile(True)
    for k in range(12):
        flush.fill_(k)
        render_geometry_raw(ctx, mesh, cam, 768, 768, **kw)
        for n, ms in ctx.ctx.profile_read(): acc.setdefault(n, []).append(ms * 1e3)
    ctx.ctx.profile(False)
    print(f"{name:34s}", {n: round(float(np.mean(x[2:])), 1) for n, x in acc.items() if n in ("k_shade", "k_depth_finalize", "k_setup_triangles")})
S = wr.SimpleNormalization()
C = wr.DepthControlNetNormalization()
run("full all outputs controlnet", full, depth_normalization_strategy=C)
run("full all outputs simple", full, depth_normalization_strategy=S)
run("full no normal", full, want_normal=False, depth_normalization_strategy=S)
run("full no normal no depth", full, want_normal=False, want_depth=False)
run("full mask only", full, want_normal=False, want_depth=False, want_pos=False)
run("empty all outputs simple", empty, depth_normalization_strategy=S)
run("empty all outputs controlnet", empty, depth_normalization_strategy=C)
run("empty mask only", empty, want_normal=False, want_depth=False, want_pos=False)
