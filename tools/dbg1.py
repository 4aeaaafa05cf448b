import sys; sys.path.insert(0,"tests"); sys.path.insert(0,".")
import numpy as np, torch
import cases, worldrenderer_b200 as wr
from worldrenderer_b200 import synth
from worldrenderer_b200.render import render_geometry_raw
from test_gpu_render_parity import make_mesh
ctx = wr.NVDiffRastContextWrapper("cuda:0","cuda")
v, f = synth.terrain(1000, 500, 0)
v = v / np.abs(v).max() * 0.5
v = np.stack([v[:, 0], -v[:, 2], v[:, 1]], -1).astype(np.float32)
mesh = make_mesh(v, f.astype(np.int32), ctx.device)
cam = cases.canonical_cameras(device=ctx.device)
raw = render_geometry_raw(ctx, mesh, cam, 768, 768, want_tri_id=True)
for b in range(6):
    m = raw["mask"][b]
    rows = m.any(1).nonzero().flatten(); cols = m.any(0).nonzero().flatten()
    r0,r1,c0,c1 = int(rows.min()),int(rows.max()),int(cols.min()),int(cols.max())
    sub = m[r0:r1+1,c0:c1+1]
    holes = (~sub).nonzero()
    print(b, "bbox", r0,r1,c0,c1, "sum", int(m.sum()), "holes", holes.shape[0], holes[:8].tolist())
print(v.min(0), v.max(0))
