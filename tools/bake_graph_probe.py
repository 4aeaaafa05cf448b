"""Config C bake: eager CameraProjection call vs its CUDA-graph replay (wr.BakeGraph), with and without return_dict."""
import contextlib, io, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import worldrenderer_b200 as wr
from worldrenderer_b200 import synth
dev = torch.device('cuda', 0)
H = W = 768; uv = 1024; NV = 6
v, f = synth.icosphere(50, 0.5)
vt, ft = synth.cell_atlas_uv(f.shape[0])
mesh = wr.TexturedMesh(v_pos=torch.tensor(v, dtype=torch.float32), t_pos_idx=torch.tensor(f, dtype=torch.int64),
                       v_tex=torch.tensor(vt, dtype=torch.float32), t_tex_idx=torch.tensor(ft, dtype=torch.int64),
                       texture=torch.zeros((uv, uv, 3), dtype=torch.float32))
mesh.set_stitched_mesh(mesh.v_pos, mesh.t_pos_idx); mesh.to(dev); mesh.v_nrm
cam = wr.get_orthogonal_camera(device=str(dev), **synth.CANONICAL_RIG)
images = torch.from_numpy(synth.view_images(NV, H, W, seed=1)).to(dev)
proj = wr.CameraProjection("torch-cuda", None, str(dev), "cuda")
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
def timed(fn, reps=20):
    for _ in range(3): fn()
    ms = []
    for k in range(reps):
        flush.fill_(k & 0xFF)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    return float(np.median(ms))
base = dict(uv_size=uv, poisson_blending=False, uv_padding=False, depth_grad_dilation=5, uv_exp_blend_alpha=3,
            uv_exp_blend_view_weight=torch.ones(NV), aoi_cos_valid_threshold=0.2, depth_grad_threshold=0.1,
            iou_rejection_threshold=None)
with contextlib.redirect_stdout(io.StringIO()):
    rows = []
    for name, kw in (("return_dict", dict(base, return_dict=True)), ("atlas only", dict(base)),
                     ("return_dict + padding", dict(base, return_dict=True, uv_padding=True)),
                     ("return_dict + padding + poisson 1000", dict(base, return_dict=True, uv_padding=True, poisson_blending=True, pb_num_iters=1000))):
        eager = timed(lambda: proj(images, mesh, cam, **kw), reps=10)
        g = wr.BakeGraph(proj, images, mesh, cam, **kw)
        graph = timed(g.replay, reps=10)
        rows.append((name, eager, graph))
for name, e, g in rows:
    print(f"{name}: eager {e*1e3:.1f} us, graph replay {g*1e3:.1f} us")
