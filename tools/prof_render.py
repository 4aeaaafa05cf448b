"""Small driver for ncu / stage timing: config B render steps (and optionally config C bake steps)."""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import worldrenderer_b200 as wr  # noqa: E402
from worldrenderer_b200 import synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--per-view", action="store_true")
ap.add_argument("--quick", action="store_true", help="with --per-view: only the all-views table")
ap.add_argument("--bake", action="store_true")
ap.add_argument("--mesh", default="terrain")
ap.add_argument("--depth", default="controlnet")
args = ap.parse_args()

dev = torch.device("cuda", 0)
cam_cpu = wr.get_orthogonal_camera(**synth.CANONICAL_RIG)
cam = wr.Camera(c2w=cam_cpu.c2w.to(dev), w2c=cam_cpu.w2c.to(dev), proj_mtx=cam_cpu.proj_mtx.to(dev),
                mvp_mtx=cam_cpu.mvp_mtx.to(dev), cam_pos=cam_cpu.cam_pos.to(dev))
if args.mesh == "terrain":
    v, f = bench.terrain_arrays(0)
else:
    v, f = synth.icosphere(50, 0.5)
    v, f = v.astype(np.float32), f.astype(np.int64)
mesh = wr.TexturedMesh(v_pos=torch.from_numpy(v).to(dev), t_pos_idx=torch.from_numpy(f).to(dev))
mesh.set_stitched_mesh(mesh.v_pos, mesh.t_pos_idx)
mesh.v_nrm
ctx = wr.NVDiffRastContextWrapper("cuda:0", "cuda")
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)


STRAT = {"controlnet": wr.DepthControlNetNormalization(), "simple": wr.SimpleNormalization(),
         "none": None}[args.depth]


def stage_table(c, reps):
    acc = {}
    ctx.ctx.profile(True)
    for k in range(reps):
        flush.fill_(k)
        wr.render(ctx, mesh, c, 768, 768, render_attr=False, depth_normalization_strategy=STRAT)
        for n, ms in ctx.ctx.profile_read():
            acc.setdefault(n, []).append(ms * 1e3)
    ctx.ctx.profile(False)
    return {n: float(np.mean(x)) for n, x in acc.items()}


for _ in range(args.steps):
    flush.fill_(1)
    out = wr.render(ctx, mesh, cam, 768, 768, render_attr=False, depth_normalization_strategy=STRAT)
torch.cuda.synchronize()
if args.per_view:
    print("all views (us):", {k: round(x, 1) for k, x in stage_table(cam, 30).items()})
    for b in range(0 if args.quick else 6):
        t = stage_table(cam[b], 10)
        print(f"view {b} covered={int(out.mask[b].sum())} (us):", {k: round(x, 1) for k, x in t.items()})
if args.bake:
    with torch.no_grad():
        b = bench.bench_bake(ctx, dev, flush)
    print(b)
    from worldrenderer_b200.uv import fused_view_maps, fused_unproject
    import contextlib, io
    v, f = synth.icosphere(50, 0.5)
    vt, ft = synth.cell_atlas_uv(f.shape[0])
    m = wr.TexturedMesh(v_pos=torch.tensor(v, dtype=torch.float32), t_pos_idx=torch.tensor(f), v_tex=torch.tensor(vt, dtype=torch.float32),
                        t_tex_idx=torch.tensor(ft), texture=torch.zeros((1024, 1024, 3)))
    m.set_stitched_mesh(m.v_pos, m.t_pos_idx); m.to(dev); m.v_nrm
    images = torch.from_numpy(synth.view_images(6, 768, 768, seed=1)).to(dev)
    pre = wr.uv_precompute(ctx, m, 1024, 1024)
    ctx.ctx.profile(True)
    acc = {}
    for k in range(8):
        flush.fill_(k)
        wr.uv_precompute(ctx, m, 1024, 1024)
        for n, ms in ctx.ctx.profile_read(): acc.setdefault("pre:" + n, []).append(ms * 1e3)
        _, geo, att = fused_view_maps(ctx, m, cam, images, 768, 768, 5)
        for n, ms in ctx.ctx.profile_read(): acc.setdefault("prep:" + n, []).append(ms * 1e3)
        fused_unproject(ctx, pre, cam, 768, 768, geo, att, aoi_cos_thresh=0.2, depth_grad_thresh=0.1, alpha=3.0)
        for n, ms in ctx.ctx.profile_read(): acc.setdefault("unproj:" + n, []).append(ms * 1e3)
    ctx.ctx.profile(False)
    print("bake stages (us):", {n: round(float(np.mean(x[2:])), 1) for n, x in acc.items()})
