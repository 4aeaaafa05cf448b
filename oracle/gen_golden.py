"""Generates tests/golden/*.npz by running the UNMODIFIED reference Python on CPU.  TEST INFRASTRUCTURE.

Run in the build container only (needs /root/reference):   python -m oracle.gen_golden

The reference's render.py / uv.py / projection.py / camera.py / mesh.py execute as they are; only the
absent `nvdiffrast.torch` operators are served by the C oracle (oracle/ref_shim.py).  The fixtures
pin (a) the host-side API (cameras, load_mesh) of worldrenderer_b200 and (b) the NumPy restatement
oracle/render_oracle.py, which is what the GPU box compares the CUDA kernels with.

Inputs are stored next to the outputs so the tests never need the reference tree.
"""
from __future__ import annotations

import json
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from worldrenderer_b200 import synth  # noqa: E402


def _np(t):
    return None if t is None else t.detach().cpu().numpy()


def cameras(mu):
    rig = synth.CANONICAL_RIG
    o = mu.get_orthogonal_camera(**rig)
    p = mu.get_camera(elevation_deg=[10.0, -20.0, 35.0, 60.0], distance=[1.8] * 4, fovy_deg=[40.0] * 4,
                      azimuth_deg=[0.0, 75.0, 160.0, 250.0], aspect_wh=4 / 3, near=0.05, far=20.0)
    q = mu.get_camera(elevation_deg=[0.0] * 3, distance=[2.0] * 3, fovy_deg=[50.0] * 3, azimuth_deg=None, num_views=3)
    frames = json.load(open(os.path.join(ref_shim.REFERENCE_ROOT, "mvadapter", "test", "camera_path.json")))
    sel = [frames[i] for i in (0, 33, 66, 99)]
    c2w = torch.tensor([f["matrix_world"] for f in sel], dtype=torch.float32)  # uniform scale 0.6 in the rotation
    fov = torch.tensor([f["fov_deg"] for f in sel], dtype=torch.float32)
    j = mu.get_camera(c2w=c2w, fovy_deg=fov, aspect_wh=720 / 480)
    out = {"json_c2w": _np(c2w), "json_fov": _np(fov)}
    for name, cam in (("ortho", o), ("persp", p), ("ring", q), ("json", j)):
        for fld in ("c2w", "w2c", "proj_mtx", "mvp_mtx", "cam_pos"):
            out[f"{name}_{fld}"] = _np(getattr(cam, fld))
    w = torch.linalg.inv(p.c2w)
    k = mu.get_camera(w2c=w, proj_mtx=p.proj_mtx)
    out["w2conly_mvp_mtx"] = _np(k.mvp_mtx)
    np.savez_compressed(os.path.join(OUT, "cameras.npz"), **out)


def load_mesh_cases(mu):
    rng = np.random.default_rng(3)
    v, f = synth.icosphere(2, 1.0)
    v = v * np.array([1.5, 0.7, 1.1]) + np.array([0.3, -0.2, 0.1]) + rng.normal(0, 0.01, v.shape)
    out = {"vertices": v, "faces": f}
    combos = {
        "default": {},
        "rescale": dict(rescale=True),
        "center_rescale": dict(rescale=True, move_to_center=True, scale=0.45),
        "zup": dict(shape_init_mesh_up="+z", shape_init_mesh_front="-y", rescale=True),
        "x2y": dict(front_x_to_y=True, rescale=True),
    }
    with tempfile.TemporaryDirectory() as d:
        path = synth.save_npz(os.path.join(d, "m.npz"), v, f)
        for name, kw in combos.items():
            m, off, sc = mu.load_mesh(path, return_transform=True, **kw)
            out[f"{name}_v_pos"] = _np(m.v_pos)
            out[f"{name}_t_pos_idx"] = _np(m.t_pos_idx)
            out[f"{name}_offset"] = np.zeros(0) if off is None else np.asarray(off)
            out[f"{name}_scale"] = np.zeros(0) if sc is None else np.asarray(sc)
            out[f"{name}_v_nrm"] = _np(m.v_nrm)
    np.savez_compressed(os.path.join(OUT, "load_mesh.npz"), **out)


def _sphere_mesh(mu, freq, tex_size, seed=0):
    v, f = synth.icosphere(freq, 0.5)
    vt, ft = synth.cell_atlas_uv(f.shape[0])
    rng = np.random.default_rng(seed)
    tex = rng.uniform(0, 1, (tex_size, tex_size, 3)).astype(np.float32)
    m = mu.TexturedMesh(v_pos=torch.tensor(v, dtype=torch.float32), t_pos_idx=torch.tensor(f, dtype=torch.int64),
                        v_tex=torch.tensor(vt, dtype=torch.float32), t_tex_idx=torch.tensor(ft, dtype=torch.int64),
                        texture=torch.tensor(tex))
    m.set_stitched_mesh(m.v_pos, m.t_pos_idx)
    return m


def _terrain_mesh(mu, nx, ny):
    v, f = synth.terrain(nx, ny, 0)
    v = v / np.abs(v).max() * 0.5
    v = np.stack([v[:, 0], -v[:, 2], v[:, 1]], -1)
    m = mu.TexturedMesh(v_pos=torch.tensor(v, dtype=torch.float32), t_pos_idx=torch.tensor(f, dtype=torch.int64))
    m.set_stitched_mesh(m.v_pos, m.t_pos_idx)
    return m


def _mesh_inputs(m):
    d = {"v_pos": _np(m.v_pos), "t_pos_idx": _np(m.t_pos_idx).astype(np.int32)}
    if m.v_tex is not None:
        d.update(v_tex=_np(m.v_tex), t_tex_idx=_np(m.t_tex_idx).astype(np.int32), texture=_np(m.texture))
    return d


def render_cases(mu):
    ctx = mu.NVDiffRastContextWrapper("cpu", "cuda")
    m = _sphere_mesh(mu, 6, 16)
    cam = mu.get_orthogonal_camera(**synth.CANONICAL_RIG)
    out = _mesh_inputs(m)
    out.update(mvp=_np(cam.mvp_mtx), w2c=_np(cam.w2c), v_nrm=_np(m.v_nrm))
    H = W = 64
    r = mu.render(ctx, m, cam, H, W, render_attr=True, attr_background=0.25)
    out.update(mask=_np(r.mask), pos=_np(r.pos), normal=_np(r.normal), depth_controlnet=_np(r.depth), attr_linear=_np(r.attr))
    r = mu.render(ctx, m, cam, H, W, render_attr=True, texture_filter_mode="nearest",
                  depth_normalization_strategy=mu.Zero123PlusPlusNormalization())
    out.update(depth_zero123pp=_np(r.depth), attr_nearest=_np(r.attr))
    r = mu.render(ctx, m, cam, H, W, render_attr=False, depth_normalization_strategy=mu.SimpleNormalization())
    out.update(depth_simple=_np(r.depth))
    r = mu.render(ctx, m, cam, H, W, render_attr=False, depth_normalization_strategy=None, normal_background=0.5)
    out.update(depth_none=_np(r.depth), normal_bg05=_np(r.normal))
    np.savez_compressed(os.path.join(OUT, "render_sphere.npz"), **out)

    m = _terrain_mesh(mu, 32, 16)
    cams = {
        "persp": mu.get_camera(elevation_deg=[10.0, -20.0, 35.0, 60.0], distance=[1.8] * 4, fovy_deg=[40.0] * 4,
                               azimuth_deg=[0.0, 75.0, 160.0, 250.0], aspect_wh=64 / 48),
        "inside": mu.get_camera(elevation_deg=[5.0, 40.0], distance=[0.3, 0.45], fovy_deg=[70.0, 90.0],
                                azimuth_deg=[20.0, 200.0], near=0.05, far=10.0, aspect_wh=64 / 48),
    }
    out = _mesh_inputs(m)
    out["v_nrm"] = _np(m.v_nrm)
    for name, cam in cams.items():
        r = mu.render(ctx, m, cam, 48, 64, render_attr=False)
        out.update({f"{name}_mvp": _np(cam.mvp_mtx), f"{name}_w2c": _np(cam.w2c), f"{name}_mask": _np(r.mask),
                    f"{name}_pos": _np(r.pos), f"{name}_normal": _np(r.normal), f"{name}_depth": _np(r.depth)})
    np.savez_compressed(os.path.join(OUT, "render_terrain.npz"), **out)


def bake_cases(mu):
    m = _sphere_mesh(mu, 6, 64, seed=2)
    cam = mu.get_orthogonal_camera(**synth.CANONICAL_RIG)
    images = synth.view_images(6, 48, 48, seed=1)
    proj = mu.CameraProjection(pb_backend="torch-native", bg_remover=None, device="cpu", context_type="cuda")
    out = _mesh_inputs(m)
    out.update(mvp=_np(cam.mvp_mtx), w2c=_np(cam.w2c), v_nrm=_np(m.v_nrm), images=images)
    vw = torch.tensor([1.0, 0.5, 1.0, 2.0, 1.0, 1.0])
    out["view_weight"] = _np(vw)
    variants = {
        "a": dict(aoi_cos_valid_threshold=0.2, depth_grad_threshold=0.1, uv_exp_blend_alpha=3.0,
                  uv_exp_blend_view_weight=vw, depth_grad_dilation=5),
        "b": dict(aoi_cos_valid_threshold=-1.0, depth_grad_threshold=None, uv_exp_blend_alpha=3.0,
                  uv_exp_blend_view_weight=torch.ones(6), depth_grad_dilation=5),
        "c": dict(aoi_cos_valid_threshold=0.3, depth_grad_threshold=0.1, uv_exp_blend_alpha=6.0, depth_grad_dilation=3),
    }
    for name, kw in variants.items():
        r = proj(torch.from_numpy(images), m, cam, uv_size=64, poisson_blending=False, uv_padding=False,
                 iou_rejection_threshold=None, return_dict=True, **kw)
        out.update({f"{name}_uv_proj": _np(r.uv_proj), f"{name}_uv_proj_mask": _np(r.uv_proj_mask),
                    f"{name}_uv_depth_grad": _np(r.uv_depth_grad), f"{name}_uv_aoi_cos": _np(r.uv_aoi_cos)})
    # with view masks (the rendered masks themselves) -> exercises uv_mask_proj and the IoU branch
    ctx = proj.ctx
    masks = mu.render(ctx, m, cam, 48, 48, render_attr=False).mask.float()
    r = proj(torch.from_numpy(images), m, cam, masks=masks, uv_size=64, poisson_blending=False, uv_padding=False,
             return_dict=True)
    out.update(masks=_np(masks), m_uv_proj=_np(r.uv_proj), m_uv_proj_mask=_np(r.uv_proj_mask))
    # step-by-step intermediates for one configuration
    pre = mu.uv.uv_precompute(ctx, m, 64, 64)
    geo = mu.uv.uv_render_geometry(ctx, m, cam, 48, 48, pre, compute_depth_grad=True, depth_grad_dilation=5)
    out.update(uv_mask=_np(pre.uv_mask), uv_pos=_np(pre.uv_pos), uv_pos_ndc=_np(geo.uv_pos_ndc),
               uv_pos_error=_np(geo.uv_pos_error), view_aoi_cos=_np(geo.view_aoi_cos),
               view_depth_grad=_np(geo.view_depth_grad)[:, 0], view_depth=_np(geo.view_depth))
    np.savez_compressed(os.path.join(OUT, "bake_sphere.npz"), **out)


def operator_cases(mu):
    """The dr.* operator boundary itself: the clip-space positions the REFERENCE hands to dr.rasterize (its own
    torch.matmul, utils.py:127-129, and its own UV -> clip construction, uv.py:28-38) are recorded together with
    what came back, so that the GPU operators can be checked bit for bit on exactly those inputs -- no restated
    clip transform in between."""
    dr = sys.modules["nvdiffrast.torch"]
    rec = {"rasterize": [], "interpolate": []}
    orig_r, orig_i = dr.rasterize, dr.interpolate

    def rasterize(ctx, pos, tri, resolution, *a, **kw):
        out = orig_r(ctx, pos, tri, resolution, *a, **kw)
        rec["rasterize"].append((_np(pos).copy(), _np(tri).copy(), tuple(int(x) for x in resolution), _np(out[0]).copy()))
        return out

    def interpolate(attr, rast, tri, *a, **kw):
        out = orig_i(attr, rast, tri, *a, **kw)
        rec["interpolate"].append((_np(attr).copy(), _np(rast).copy(), _np(tri).copy(), _np(out[0]).copy()))
        return out

    dr.rasterize, dr.interpolate = rasterize, interpolate
    try:
        ctx = mu.NVDiffRastContextWrapper("cpu", "cuda")
        m = _sphere_mesh(mu, 6, 16)
        mu.render(ctx, m, mu.get_orthogonal_camera(**synth.CANONICAL_RIG), 64, 64, render_attr=True)
        t = _terrain_mesh(mu, 32, 16)
        mu.render(ctx, t, mu.get_camera(elevation_deg=[10.0, -20.0, 35.0, 60.0], distance=[1.8] * 4, fovy_deg=[40.0] * 4,
                                        azimuth_deg=[0.0, 75.0, 160.0, 250.0], aspect_wh=64 / 48), 48, 64, render_attr=False)
        mu.render(ctx, t, mu.get_camera(elevation_deg=[5.0, 40.0], distance=[0.3, 0.45], fovy_deg=[70.0, 90.0],
                                        azimuth_deg=[20.0, 200.0], near=0.05, far=10.0, aspect_wh=64 / 48), 48, 64,
                  render_attr=False)
        mu.uv.uv_precompute(ctx, m, 64, 64)
    finally:
        dr.rasterize, dr.interpolate = orig_r, orig_i
    out = {"n_rasterize": len(rec["rasterize"]), "n_interpolate": len(rec["interpolate"])}
    for k, (pos, tri, res, rast) in enumerate(rec["rasterize"]):
        out.update({f"r{k}_pos": pos, f"r{k}_tri": tri.astype(np.int32), f"r{k}_res": np.asarray(res), f"r{k}_rast": rast})
    for k, (attr, rast, tri, val) in enumerate(rec["interpolate"]):
        out.update({f"i{k}_attr": attr, f"i{k}_rast": rast, f"i{k}_tri": tri.astype(np.int32), f"i{k}_out": val})
    np.savez_compressed(os.path.join(OUT, "operators.npz"), **out)
    print("operators:", out["n_rasterize"], "rasterize calls,", out["n_interpolate"], "interpolate calls")


def _reference_block(path, first, last):
    """Source lines of the reference from the line containing `first` to the one containing `last` (inclusive),
    dedented -- executed, never stored: the tangent-space bake of pipeline_texture.py is inline code of a method."""
    import textwrap
    lines = open(path).read().splitlines()
    a = next(i for i, l in enumerate(lines) if first in l)
    b = next(i for i, l in enumerate(lines) if last in l and i > a)
    return textwrap.dedent("\n".join(lines[a:b + 1]))


def tangent_cases(mu):
    """TexturedMesh.v_tang (mesh.py:121-167), render(render_tangent=True) (render.py:280-284) and the tangent-space
    rotation of the normal modality (mvadapter/test/utils/pipeline_texture.py:358-398), all run by the reference."""
    import types
    import torch.nn.functional as F
    ctx = mu.NVDiffRastContextWrapper("cpu", "cuda")
    m = _sphere_mesh(mu, 6, 16)
    cam = mu.get_orthogonal_camera(**synth.CANONICAL_RIG)
    out = _mesh_inputs(m)
    out.update(mvp=_np(cam.mvp_mtx), w2c=_np(cam.w2c), v_nrm=_np(m.v_nrm), v_tang=_np(m.v_tang))
    render_out = mu.render(ctx, m, cam, 64, 64, render_attr=False, render_depth=False, render_normal=True,
                           render_tangent=True)
    out.update(normal=_np(render_out.normal), tangent=_np(render_out.tangent), mask=_np(render_out.mask))
    rng = np.random.default_rng(11)
    nimg = rng.normal(0, 1, (6, 64, 64, 3))
    nimg[..., 2] = np.abs(nimg[..., 2]) + 0.5
    nimg = (nimg / np.linalg.norm(nimg, axis=-1, keepdims=True) * 0.5 + 0.5).astype(np.float32)
    out["normal_images"] = nimg
    src = _reference_block(os.path.join(ref_shim.REFERENCE_ROOT, "mvadapter", "test", "utils", "pipeline_texture.py"),
                           "# compute UV tangent space", "mod_tensor = (mod_tensor * 0.5 + 0.5).clamp(0, 1)")
    ns = {"torch": torch, "F": F, "render_out": render_out, "mod_tensor": torch.from_numpy(nimg),
          "self": types.SimpleNamespace(device="cpu")}
    exec(src, ns)
    out["tangent_space"] = _np(ns["mod_tensor"])
    np.savez_compressed(os.path.join(OUT, "tangent.npz"), **out)


def poisson_cases(mu):
    """The reference's own PoissonBlendingSolver (blend.py:186-324) on CPU, "torch-native" backend, unmodified;
    then its uv_blend / CameraProjection with Poisson blending + padding, where only cvcuda.inpaint is served by the
    oracle's fill (ref_shim._make_cvcuda) -- that pins the control flow of uv.py:426-461, not the fill."""
    import importlib
    blend = importlib.import_module("mvadapter.utils.mesh_utils.blend")
    solver = blend.PoissonBlendingSolver("torch-native", "cpu")
    rng = np.random.default_rng(7)
    H, W = 56, 72
    yy, xx = np.mgrid[0:H, 0:W]
    src = (0.5 + 0.4 * np.sin(0.3 * xx + 0.2 * yy)[..., None] * np.array([1, 0.8, 0.6]) + 0.05 * rng.random((H, W, 3)))
    tgt = (0.4 + 0.3 * np.cos(0.15 * xx - 0.25 * yy)[..., None] * np.array([0.7, 1, 0.5]) + 0.05 * rng.random((H, W, 3)))
    src, tgt = src.astype(np.float32), tgt.astype(np.float32)
    mask = np.zeros((H, W), np.float32)
    mask[5:30, 3:60] = 1
    mask[10:12, 10:20] = 0
    mask[0:3, :] = 1          # touches the border: removed by blend.py:233-236
    mask[40:56, 60:72] = 1
    mask[35, 5] = 1           # isolated pixel
    out = {"src": src, "tgt": tgt, "mask": mask}
    for gm in ("src", "max", "avg"):
        for it in (0, 1, 9, 250):
            r = solver(torch.from_numpy(src), torch.from_numpy(mask), torch.from_numpy(tgt).clone(), it,
                       inplace=False, grad_mode=gm)
            out[f"{gm}_{it}"] = _np(r)
    r = solver(torch.from_numpy(src), torch.from_numpy(np.repeat(mask[..., None], 3, -1)), torch.from_numpy(tgt).clone(),
               16, inplace=False)
    out["mask3_16"] = _np(r)

    # bake with the post-processing tail
    m = _sphere_mesh(mu, 6, 64, seed=2)
    cam = mu.get_orthogonal_camera(**synth.CANONICAL_RIG)
    images = synth.view_images(6, 48, 48, seed=1)
    proj = mu.CameraProjection(pb_backend="torch-native", bg_remover=None, device="cpu", context_type="cuda")
    kw = dict(uv_size=64, iou_rejection_threshold=None, aoi_cos_valid_threshold=0.2, depth_grad_threshold=0.1,
              uv_exp_blend_alpha=3.0, depth_grad_dilation=5)
    out["bake_pad"] = _np(proj(torch.from_numpy(images), m, cam, poisson_blending=False, uv_padding=True, **kw))
    out["bake_pad_scratch"] = _np(proj(torch.from_numpy(images), m, cam, poisson_blending=False, uv_padding=True,
                                       from_scratch=True, **kw))
    out["bake_pb"] = _np(proj(torch.from_numpy(images), m, cam, poisson_blending=True, pb_num_iters=40,
                              uv_padding=True, **kw))
    out["bake_pb_noborder"] = _np(proj(torch.from_numpy(images), m, cam, poisson_blending=True, pb_num_iters=40,
                                       pb_keep_original_border=False, uv_padding=True, **kw))
    np.savez_compressed(os.path.join(OUT, "poisson.npz"), **out)


def smart_paint_inpaint(img: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """Deterministic stand-in for the inpainting network: img [1,3,H,W], mask [1,1,H,W] -> [1,3,H,W]."""
    H, W = img.shape[-2:]
    yy, xx = torch.meshgrid(torch.arange(H, device=img.device, dtype=torch.float32),
                            torch.arange(W, device=img.device, dtype=torch.float32), indexing="ij")
    pat = torch.stack([0.5 + 0.4 * torch.sin(0.05 * xx), 0.5 + 0.4 * torch.cos(0.04 * yy),
                       0.5 + 0.4 * torch.sin(0.03 * (xx + yy))])[None]
    return img * (1 - mask) + pat * mask


def smart_paint_mesh():
    """Asymmetric blob (no two candidate views score alike) with a cell atlas."""
    v, f = synth.icosphere(5, 0.45)
    v = v * np.array([1.0, 0.75, 0.6]) + 0.08 * np.sin(4.0 * v[:, [1, 2, 0]]) + np.array([0.03, -0.02, 0.01])
    vt, ft = synth.cell_atlas_uv(f.shape[0])
    return v.astype(np.float32), f.astype(np.int64), vt.astype(np.float32), ft.astype(np.int64)


def smart_paint_case(mu):
    """The reference's SmartPainter (smart_paint.py:37-335) on CPU, unmodified.  Test infrastructure swaps in:
    the C oracle for nvdiffrast, the oracle's fill for cvcuda.inpaint (ref_shim), the solver backend class
    (load_inline needs a GPU; Poisson blending is off in this loop anyway), and a recording proxy for the module's
    `np` so that the per-round view scores can be stored."""
    import importlib
    blend = importlib.import_module("mvadapter.utils.mesh_utils.blend")
    blend.PBTorchCUDAKernelBackend = blend.PBTorchNativeBackend
    sp = importlib.import_module("mvadapter.utils.mesh_utils.smart_paint")
    rec = {"scores": []}

    class _NP:
        def __getattr__(self, name):
            return getattr(np, name)

        def max(self, x):
            rec["scores"].append(np.asarray(x, dtype=np.float64))
            return np.max(x)

    sp.np = _NP()
    v, f, vt, ft = smart_paint_mesh()
    uv = 96
    rng = np.random.default_rng(5)
    tex = rng.uniform(0.2, 0.8, (uv, uv, 3)).astype(np.float32)
    m = mu.TexturedMesh(v_pos=torch.from_numpy(v), t_pos_idx=torch.from_numpy(f), v_tex=torch.from_numpy(vt),
                        t_tex_idx=torch.from_numpy(ft), texture=torch.from_numpy(tex))
    m.set_stitched_mesh(m.v_pos, m.t_pos_idx)
    inpaint_mask = np.zeros((uv, uv), bool)
    inpaint_mask[:, : uv // 2] = True      # the left half of the atlas is unpainted
    inpaint_mask[10:30, 60:90] = True
    painter = sp.SmartPainter("cpu", context_type="cuda")
    torch.manual_seed(0)
    tex_out, valid_out = painter("case", m, smart_paint_inpaint, torch.from_numpy(tex), torch.from_numpy(inpaint_mask),
                                 min_rounds=2, max_rounds=3)
    out = {"v_pos": v, "t_pos_idx": f.astype(np.int32), "v_tex": vt, "t_tex_idx": ft.astype(np.int32), "texture": tex,
           "inpaint_mask": inpaint_mask, "texture_out": _np(tex_out), "valid_out": _np(valid_out),
           "view_scores": np.stack(rec["scores"])}
    np.savez_compressed(os.path.join(OUT, "smart_paint.npz"), **out)
    print("smart_paint rounds", len(rec["scores"]), "best", [int(np.argmax(s)) for s in rec["scores"]],
          "top-2 margins", [float(np.sort(s)[-1] - np.sort(s)[-2]) for s in rec["scores"]])


def strategy_cases(mu):
    """The validity / blend-weight strategy classes (uv.py:248-370) on synthetic per-view tensors: every option the
    reference's callers can reach -- depth-gradient threshold with and without a gradient map, view masks present /
    absent, first_view_dominate, per-view weights, linear and softmax normalisation."""
    import contextlib
    import io
    rng = np.random.default_rng(11)
    nv, h, w = 5, 24, 20
    geo_in = dict(uv_pos_error=(rng.random((nv, h, w)) ** 4 * 4e-3).astype(np.float32),
                  uv_aoi_cos=(rng.random((nv, h, w)) * 1.2 - 0.2).astype(np.float32),
                  uv_depth_grad=(rng.random((nv, h, w)) * 0.25).astype(np.float32))
    uv_mask = rng.random((h, w)) < 0.8
    uv_mask_proj = (rng.random((nv, h, w)) ** 0.1).astype(np.float32)   # mostly foreground
    view_weight = np.array([1.0, 0.5, 2.0, 1.0, 3.0], np.float32)
    out = dict(geo_in, uv_mask=uv_mask, uv_mask_proj=uv_mask_proj, view_weight=view_weight)

    def geo(with_grad=True):
        z = torch.zeros(1)
        return mu.uv.UVRenderGeometryOutput(
            uv_pos_proj=z, uv_pos_error=torch.from_numpy(geo_in["uv_pos_error"]),
            uv_aoi_cos=torch.from_numpy(geo_in["uv_aoi_cos"]), uv_pos_ndc=z, view_mask=z, view_normal=z,
            view_aoi_cos=z, view_position=z, view_depth=z,
            uv_depth_grad=torch.from_numpy(geo_in["uv_depth_grad"]) if with_grad else None)

    pre = mu.uv.UVPrecomputeOutput(height=h, width=w, uv_attr=torch.zeros(h, w, 3), uv_mask=torch.from_numpy(uv_mask),
                                   uv_pos=torch.zeros(h, w, 3))
    attr = mu.uv.UVRenderAttrOutput(uv_attr_proj=torch.zeros(1), uv_mask_proj=torch.from_numpy(uv_mask_proj))
    attr_nomask = mu.uv.UVRenderAttrOutput(uv_attr_proj=torch.zeros(1), uv_mask_proj=None)
    validity = {   # name -> (constructor kwargs, gradient map present, view masks present)
        "v_default": (dict(), True, True),
        "v_thresholds": (dict(pos_error_eps=5e-4, aoi_cos_thresh=0.3, mask_thresh=0.5, depth_grad_thresh=0.1), True, True),
        "v_grad_missing": (dict(depth_grad_thresh=0.1), False, True),
        "v_no_view_mask": (dict(aoi_cos_thresh=0.2, depth_grad_thresh=0.15), True, False),
        "v_first_view": (dict(aoi_cos_thresh=0.2, first_view_dominate=True), True, True),
    }
    with contextlib.redirect_stdout(io.StringIO()):   # the reference prints a warning in two of the cases
        for name, (kw, with_grad, with_mask) in validity.items():
            v = mu.uv.SimpleUVValidityStrategy(**kw)(pre, geo(with_grad), attr if with_mask else attr_nomask)
            out[name] = _np(v)
    valid = torch.from_numpy(out["v_default"])
    blends = {
        "w_linear_a1": dict(alpha=1.0),
        "w_linear_a3": dict(alpha=3.0),
        "w_linear_a6_vw": dict(alpha=6.0, view_weight=torch.from_numpy(view_weight)),
        "w_softmax_a2": dict(alpha=2.0, normalization="softmax"),
        "w_softmax_a3_vw": dict(alpha=3.0, normalization="softmax", view_weight=torch.from_numpy(view_weight)),
    }
    for name, kw in blends.items():
        out[name] = _np(mu.uv.ExponentialBlend(**kw)(pre, geo(), attr, valid.clone()))
    np.savez_compressed(os.path.join(OUT, "strategies.npz"), **out)


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    mu = ref_shim.load_reference()
    import importlib
    mu.uv = importlib.import_module("mvadapter.utils.mesh_utils.uv")
    cameras(mu)
    load_mesh_cases(mu)
    render_cases(mu)
    bake_cases(mu)
    poisson_cases(mu)
    smart_paint_case(mu)
    operator_cases(mu)
    tangent_cases(mu)
    strategy_cases(mu)
    for n in sorted(os.listdir(OUT)):
        print(n, os.path.getsize(os.path.join(OUT, n)))


if __name__ == "__main__":
    main()
