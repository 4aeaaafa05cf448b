"""One config-C bake (CameraProjection, return_dict=True) and one config-A render inside a cudaProfiler range:
    ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/prof_r02_bake \
        python tools/prof_bake.py"""
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
proj = wr.CameraProjection(None, None, str(dev), "cuda")
kw = dict(uv_size=uv, poisson_blending=False, uv_padding=False, depth_grad_dilation=5, uv_exp_blend_alpha=3,
          uv_exp_blend_view_weight=torch.ones(NV), aoi_cos_valid_threshold=0.2, depth_grad_threshold=0.1,
          iou_rejection_threshold=None, return_dict=True)
ctx = wr.NVDiffRastContextWrapper(str(dev), "cuda")
with contextlib.redirect_stdout(io.StringIO()):
    for _ in range(3):
        proj(images, mesh, cam, **kw)
        wr.render(ctx, mesh, cam, H, W, render_attr=False)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    proj(images, mesh, cam, **kw)
    wr.render(ctx, mesh, cam, H, W, render_attr=False)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("done")
