"""Config E on one GPU: where a rank's time goes when it holds 32 / 16 / 8 / 4 of the 32 views (per-kernel stage
tables; no exchange)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import worldrenderer_b200 as wr
from worldrenderer_b200 import parallel, synth
from worldrenderer_b200.uv import fused_view_maps, fused_unproject, uv_finalize
dev = torch.device("cuda", 0)
NV, RES, UV = 32, 2048, 4096
v, f = synth.terrain(2500, 1000, 0)
v = v / np.abs(v).max() * 0.5
v = np.ascontiguousarray(np.stack([v[:, 0], -v[:, 2], v[:, 1]], -1), np.float32)
vt = synth.terrain_uv(2500, 1000).astype(np.float32)
mesh = wr.TexturedMesh(v_pos=torch.from_numpy(v).to(dev), t_pos_idx=torch.from_numpy(f).to(dev), v_tex=torch.from_numpy(vt).to(dev),
                       t_tex_idx=torch.from_numpy(f).to(dev), texture=torch.zeros((UV, UV, 3), device=dev))
mesh.set_stitched_mesh(mesh.v_pos, mesh.t_pos_idx); mesh.v_nrm
cam = wr.get_orthogonal_camera(elevation_deg=[20.0] * NV, distance=[1.0] * NV, left=-0.55, right=0.55, bottom=-0.55, top=0.55,
                               azimuth_deg=list(np.linspace(0, 360, NV + 1)[:-1]), device=str(dev))
ctx = wr.NVDiffRastContextWrapper("cuda:0", "cuda")
pre = wr.uv_precompute(ctx, mesh, UV, UV)
kw = dict(aoi_cos_thresh=0.2, depth_grad_thresh=0.1, alpha=3.0)
def ev():
    return torch.cuda.Event(enable_timing=True)
for n in (32, 16, 8, 4):
    c = cam[:n]
    img = torch.rand((n, RES, RES, 3), device=dev)
    accum = torch.empty((UV, UV, 5), device=dev)
    def run():
        e = [ev() for _ in range(4)]
        e[0].record()
        _, geo, att = fused_view_maps(ctx, mesh, c, img, RES, RES, 5)
        e[1].record()
        fused_unproject(ctx, pre, c, RES, RES, geo, att, accumulate_only=True, accum=accum, add_to_accum=False, **kw)
        e[2].record()
        out = uv_finalize(ctx, accum, pre.uv_attr)
        e[3].record()
        return e
    for _ in range(3): run()
    torch.cuda.synchronize()
    ts = []
    for _ in range(6):
        e = run(); torch.cuda.synchronize()
        ts.append([e[i].elapsed_time(e[i + 1]) for i in range(3)])
    t = np.mean(ts, 0)
    ctx.ctx.profile(True)
    fused_view_maps(ctx, mesh, c, img, RES, RES, 5)
    st = dict(ctx.ctx.profile_read())
    ctx.ctx.profile(False)
    print(f"{n:2d} views: view maps {t[0]:.3f} ms, unproject {t[1]:.3f} ms, finalize {t[2]:.3f} ms, total {t.sum():.3f} ms | last native call stages (us):",
          {k: round(x * 1e3, 1) for k, x in st.items()})
    del img, accum
