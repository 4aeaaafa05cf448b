"""Bit-level comparison of the bake (step-by-step API and fused CameraProjection) with the oracle."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases, worldrenderer_b200 as wr
from oracle import render_oracle
from test_gpu_render_parity import make_mesh
from worldrenderer_b200 import synth
ctx = wr.NVDiffRastContextWrapper("cuda:0", "cuda")
dev = ctx.device

def cmp(name, a, b, sel=None):
    a = np.asarray(a, np.float32); b = np.asarray(b, np.float32)
    if sel is not None: a, b = a[sel], b[sel]
    neq = (a.view(np.uint32) != b.view(np.uint32)) & ~((a == 0) & (b == 0))
    d = np.abs(a - b); rel = d / np.maximum(np.abs(b), 1e-30)
    big = (d > 1e-5 * np.abs(b) + 1e-6).sum()
    print(f"  {name:16s} bit-mismatch {int(neq.sum()):9d} / {a.size:9d}  max abs {d.max():.3e}  max rel(|b|>1e-3) {rel[np.abs(b) > 1e-3].max() if (np.abs(b)>1e-3).any() else 0:.3e}  outside 1e-5/1e-6: {int(big)}")

for freq, views, uvs, alpha, vw in [(8, 96, 128, 3.0, None), (8, 96, 128, 3.0, [1.0, 0.5, 1.0, 2.0, 1.0, 1.0]), (8, 96, 128, 6.0, None), (50, 768, 1024, 3.0, None)]:
    print(f"== icosphere {freq}, views {views}^2, atlas {uvs}^2, alpha {alpha}, view weights {vw}")
    v, f = cases.icosphere_mesh(freq)
    mesh = make_mesh(v, f, dev, with_uv=True, tex_size=uvs, seed=1)
    cam = cases.canonical_cameras(device=dev)
    images = synth.view_images(6, views, views, seed=1)
    vnp, fnp = mesh.v_pos.cpu().numpy(), mesh.t_pos_idx.cpu().numpy().astype(np.int32)
    pre = wr.uv_precompute(ctx, mesh, uvs, uvs)
    geo = wr.uv_render_geometry(ctx, mesh, cam, views, views, pre, compute_depth_grad=True, depth_grad_dilation=5)
    attr = wr.uv_render_attr(torch.from_numpy(images), geo)
    rpre = render_oracle.uv_precompute(vnp, fnp, mesh.v_tex.cpu().numpy(), mesh.t_tex_idx.cpu().numpy(), uvs, uvs)
    rgeo = render_oracle.uv_render_geometry(vnp, fnp, mesh.v_nrm.cpu().numpy(), fnp, cam.mvp_mtx.cpu().numpy(), cam.w2c.cpu().numpy(), views, views, rpre, True, 5)
    rattr = render_oracle.uv_render_attr(images, rgeo)
    inside = rpre["uv_mask"]
    cmp("uv_pos", pre.uv_pos.cpu().numpy(), rpre["uv_pos"])
    for n in ["view_position", "view_normal", "view_depth", "view_aoi_cos"]:
        cmp(n, getattr(geo, n).cpu().numpy(), rgeo[n])
    cmp("view_depth_grad", geo.view_depth_grad[:, 0].cpu().numpy(), rgeo["view_depth_grad"])
    for n in ["uv_pos_ndc", "uv_pos_proj", "uv_pos_error", "uv_aoi_cos", "uv_depth_grad"]:
        cmp(n, getattr(geo, n).cpu().numpy(), rgeo[n], (slice(None), inside))
    cmp("uv_attr_proj", attr.uv_attr_proj.cpu().numpy(), rattr["uv_attr_proj"], (slice(None), inside))
    vwt = None if vw is None else torch.tensor(vw)
    proj = wr.CameraProjection(None, None, str(dev), "cuda")
    out = proj(torch.from_numpy(images), mesh, cam, uv_size=uvs, poisson_blending=False, uv_padding=False, depth_grad_dilation=5,
               uv_exp_blend_alpha=alpha, uv_exp_blend_view_weight=vwt, aoi_cos_valid_threshold=0.2, depth_grad_threshold=0.1,
               iou_rejection_threshold=None, return_dict=True)
    ref = render_oracle.camera_projection(images, vnp, fnp, mesh.v_nrm.cpu().numpy(), fnp, mesh.v_tex.cpu().numpy(),
                                          mesh.t_tex_idx.cpu().numpy().astype(np.int32), mesh.texture.cpu().numpy(), cam.mvp_mtx.cpu().numpy(),
                                          cam.w2c.cpu().numpy(), uvs, aoi_cos_valid_threshold=0.2, depth_grad_threshold=0.1,
                                          uv_exp_blend_alpha=alpha, uv_exp_blend_view_weight=None if vw is None else np.asarray(vw, np.float32), depth_grad_dilation=5)
    same_mask = out.uv_proj_mask.cpu().numpy() == ref["uv_proj_mask"]
    vsame = (ref["blend"]["uv_valid_mask"] == ref["blend"]["uv_valid_mask"])
    print(f"  uv_proj_mask differs at {int((~same_mask).sum())} texels")
    cmp("uv_proj", out.uv_proj.cpu().numpy(), ref["uv_proj"], same_mask)
    if hasattr(out, "uv_blend_weight") and out.uv_blend_weight is not None:
        cmp("uv_blend_weight", out.uv_blend_weight.cpu().numpy(), ref["blend"]["uv_blend_weight"])
