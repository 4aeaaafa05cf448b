"""How close is 'within 1e-5' really?  Counts bit-level mismatches between the CUDA path and the oracle."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases, worldrenderer_b200 as wr
from oracle import render_oracle
from oracle.render_oracle import DepthSpec
from worldrenderer_b200.render import render_geometry_raw
from test_gpu_render_parity import make_mesh
ctx = wr.NVDiffRastContextWrapper("cuda:0", "cuda")
v, f = cases.terrain_mesh(128, 64)
mesh = make_mesh(v, f, ctx.device)
for name, cam in [("ortho", cases.canonical_cameras(device=ctx.device)), ("persp", cases.perspective_cameras(device=ctx.device)), ("inside", cases.inside_cameras(device=ctx.device))]:
    raw = render_geometry_raw(ctx, mesh, cam, 200, 264, want_tri_id=True, want_rast=True, depth_normalization_strategy=wr.DepthControlNetNormalization())
    ref = render_oracle.render(v, f, cam.mvp_mtx.cpu().numpy(), cam.w2c.cpu().numpy(), 200, 264, v_nrm=mesh.v_nrm.cpu().numpy(), depth=DepthSpec("controlnet"))
    for k in ("rast", "pos", "normal", "depth"):
        a, b = raw[k].cpu().numpy(), ref[k]
        neq = (a.view(np.uint32) != b.view(np.uint32)) & ~((a == 0) & (b == 0))
        print(f"{name:7s} {k:7s} bit-mismatches {int(neq.sum()):8d} of {a.size:9d}   max abs diff {float(np.abs(a - b).max()):.3e}")
