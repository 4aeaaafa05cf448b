"""One GPU: config E baked at once vs as the sum of 8 interleaved view shards -- where do the colours differ?"""
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
img = torch.rand((NV, RES, RES, 3), device=dev)
kw = dict(aoi_cos_thresh=0.2, depth_grad_thresh=0.1, alpha=3.0)
def accumulate(sl):
    c, im = cam[sl], img[sl].contiguous()
    _, geo, att = fused_view_maps(ctx, mesh, c, im, RES, RES, 5)
    _, _, acc, _, _ = fused_unproject(ctx, pre, c, RES, RES, geo, att, accumulate_only=True, **kw)
    return acc
full = accumulate(slice(0, NV))
parts = [accumulate(slice(r, NV, 8)) for r in range(8)]
summed = torch.stack(parts).sum(0)
a1, m1 = uv_finalize(ctx, full, pre.uv_attr)
a8, m8 = uv_finalize(ctx, summed, pre.uv_attr)
print("masks equal", bool(torch.equal(m1, m8)), "max abs colour diff", float((a1 - a8).abs().max()))
d = (a1 - a8).abs().max(-1).values
idx = torch.nonzero(d > 1e-4)
print("texels above 1e-4:", idx.shape[0], "of", int(m1.sum()))
for y, x in idx[:8].tolist():
    print((y, x), "diff %.4f" % float(d[y, x]), "full acc", [round(t, 7) for t in full[y, x].tolist()], "sum acc", [round(t, 7) for t in summed[y, x].tolist()],
          "per-shard sum_w", [round(float(p[y, x, 3]), 7) for p in parts])
den = full[..., 3]
print("sum_w percentiles over covered texels:", np.percentile(den[m1].cpu().numpy(), [0, 0.01, 0.1, 1, 50]).tolist())
