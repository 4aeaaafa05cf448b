import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import worldrenderer_b200 as wr
from worldrenderer_b200 import synth
dev = torch.device("cuda", 0)
cam = wr.get_orthogonal_camera(device="cuda:0", **synth.CANONICAL_RIG)
ctx = wr.NVDiffRastContextWrapper("cuda:0", "cuda")
def mk(v, f):
    m = wr.TexturedMesh(v_pos=torch.tensor(v, dtype=torch.float32, device=dev), t_pos_idx=torch.tensor(f, device=dev)); m.set_stitched_mesh(m.v_pos, m.t_pos_idx); m.v_nrm; return m
def stages(mesh, c, res):
    ctx.ctx.profile(True); acc = {}
    for k in range(10):
        wr.render(ctx, mesh, c, res, res, render_attr=False)
        for n, ms in ctx.ctx.profile_read(): acc.setdefault(n, []).append(ms * 1e3)
    ctx.ctx.profile(False)
    return {n: round(float(np.mean(x[2:])), 1) for n, x in acc.items() if n in ("k_setup_triangles", "k_raster_queues")}
for freq in (1, 2, 4, 8, 16):
    m = mk(*synth.icosphere(freq, 0.5))
    print(f"icosphere f={freq:2d} faces={20*freq*freq:5d}", "6v 768:", stages(m, cam, 768), "1v 768:", stages(m, cam[0], 768), "6v 256:", stages(m, cam, 256))
