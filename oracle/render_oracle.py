"""NumPy restatement of the reference's render / UV-bake Python layers.  TEST INFRASTRUCTURE ONLY.

The GPU box has no /root/reference, so the parity tests there cannot import the reference's
Python; this module restates it over the C oracle (oracle/shim.py).  It is pinned against the
unmodified reference (run through oracle/ref_shim.py in the build container) by
tests/test_oracle_golden.py using the fixtures in tests/golden/.

Each function cites the reference lines it follows (paths relative to
/root/reference/mvadapter/utils/mesh_utils/).  All arithmetic is float32 NumPy with one ufunc
per operation; where the reference leaves the summation order to a library (torch.matmul,
.sum(-1)) the order written here is the normative one of DESIGN.md section 3/4, and the CUDA
kernels reproduce it.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import os

import numpy as np

from . import shim

f32 = np.float32


def _a(x) -> np.ndarray:
    return np.ascontiguousarray(x, dtype=f32)


# ---------------------------------------------------------------------------------------------
# mesh.py:85-119  TexturedMesh._compute_vertex_normal
# ---------------------------------------------------------------------------------------------

def vertex_normals(v_pos, tri) -> np.ndarray:
    v = _a(v_pos)
    t = np.asarray(tri, np.int64).reshape(-1, 3)
    e1 = v[t[:, 1]] - v[t[:, 0]]
    e2 = v[t[:, 2]] - v[t[:, 0]]
    fn = np.stack([e1[:, 1] * e2[:, 2] - e1[:, 2] * e2[:, 1],
                   e1[:, 2] * e2[:, 0] - e1[:, 0] * e2[:, 2],
                   e1[:, 0] * e2[:, 1] - e1[:, 1] * e2[:, 0]], -1).astype(f32)
    acc = np.zeros_like(v)
    for k in range(3):  # mesh.py:106-108: three scatter_add_ passes (sum order is unspecified there)
        np.add.at(acc, t[:, k], fn)
    sq = (acc * acc).sum(-1, keepdims=True)
    acc = np.where(sq > f32(1e-20), acc, np.array([0, 0, 1], f32))  # mesh.py:111-113
    n = np.sqrt((acc * acc).sum(-1, keepdims=True))
    return (acc / np.maximum(n, f32(1e-12))).astype(f32)  # F.normalize, eps 1e-12


def _normalize_rows(x: np.ndarray) -> np.ndarray:
    """F.normalize(x, dim=-1): x / max(|x|, 1e-12), the squares summed left to right."""
    n = np.sqrt((x[..., 0] * x[..., 0] + x[..., 1] * x[..., 1]) + x[..., 2] * x[..., 2]).astype(f32)
    return (x / np.maximum(n, f32(1e-12))[..., None]).astype(f32)


def vertex_tangents(v_pos, tri, v_tex, tri_tex, v_nrm) -> np.ndarray:
    """mesh.py:121-167: per-face tangent from the UV gradients, mean over the faces of a vertex, normalised, made
    perpendicular to the vertex normal, normalised again.  A vertex without a face is 0 / 0 = NaN (as there)."""
    v, vt, n = _a(v_pos), _a(v_tex), _a(v_nrm)
    t = np.asarray(tri, np.int64).reshape(-1, 3)
    tt = np.asarray(tri_tex, np.int64).reshape(-1, 3)
    uve1, uve2 = vt[tt[:, 1]] - vt[tt[:, 0]], vt[tt[:, 2]] - vt[tt[:, 0]]                     # mesh.py:135-136
    pe1, pe2 = v[t[:, 1]] - v[t[:, 0]], v[t[:, 2]] - v[t[:, 0]]                               # mesh.py:137-138
    nom = (pe1 * uve2[:, 1:2] - pe2 * uve1[:, 1:2]).astype(f32)                               # mesh.py:140
    den = (uve1[:, 0:1] * uve2[:, 1:2] - uve1[:, 1:2] * uve2[:, 0:1]).astype(f32)             # mesh.py:141
    with np.errstate(all="ignore"):
        den = np.where(den > 0, np.maximum(den, f32(1e-6)), np.minimum(den, f32(-1e-6))).astype(f32)  # mesh.py:144-146
        tang = (nom / den).astype(f32)
        acc = np.zeros_like(n)
        cnt = np.zeros_like(n)
        for k in range(3):                                                                     # mesh.py:149-155
            np.add.at(acc, t[:, k], tang)
            np.add.at(cnt, t[:, k], f32(1))
        out = _normalize_rows((acc / cnt).astype(f32))                                        # mesh.py:156-159
        d = ((out[:, 0] * n[:, 0] + out[:, 1] * n[:, 1]) + out[:, 2] * n[:, 2]).astype(f32)
        return _normalize_rows((out - d[:, None] * n).astype(f32))                            # mesh.py:160-162


def _cross(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    return np.stack([a[..., 1] * b[..., 2] - a[..., 2] * b[..., 1], a[..., 2] * b[..., 0] - a[..., 0] * b[..., 2],
                     a[..., 0] * b[..., 1] - a[..., 1] * b[..., 0]], -1).astype(f32)


def tangent_space_normals(normal, tangent, image, view_axis) -> np.ndarray:
    """mvadapter/test/utils/pipeline_texture.py:358-396: view normal images [B,H,W,3] in [0,1] (geometry tangent frame
    of each view) -> colours of the same normals in the UV tangent frame (rendered tangent, bitangent, normal)."""
    vN, vT, img = _a(normal), _a(tangent), _a(image)
    g = np.broadcast_to(_a(view_axis)[:, None, None, :], vN.shape)
    T, B_, N = _normalize_rows(vT), _normalize_rows(_cross(vN, vT)), _normalize_rows(vN)      # :359-362
    gb = _cross(vN, g)                                                                         # :365-382
    gt = _cross(gb, vN)
    GT, GB = _normalize_rows(gt), _normalize_rows(gb)                                          # :383-385
    m = (img * f32(2) - f32(1)).astype(f32)                                                    # :388
    world = _normalize_rows(((m[..., 0:1] * GT + m[..., 1:2] * GB) + m[..., 2:3] * N).astype(f32))  # :389-395
    rows = [((world[..., 0] * F_[..., 0] + world[..., 1] * F_[..., 1]) + world[..., 2] * F_[..., 2]).astype(f32)
            for F_ in (T, B_, N)]
    out = _normalize_rows(np.stack(rows, -1))                                                  # :398-400
    return np.clip(out * f32(0.5) + f32(0.5), f32(0), f32(1)).astype(f32)                     # :401


# ---------------------------------------------------------------------------------------------
# render.py:164-217  depth normalisation strategies (specs are plain tuples here)
# ---------------------------------------------------------------------------------------------

@dataclass
class DepthSpec:
    """kind: 'controlnet' (render.py:164-183), 'zero123pp' (:186-196), 'simple' (:199-217), 'none'."""
    kind: str = "controlnet"
    far_clip: float = 0.25
    near_clip: float = 1.0
    scale: float = 1.0
    offset: float = -1.0
    clamp: bool = True
    bg_value: Optional[float] = None

    def background(self) -> float:
        if self.bg_value is not None:
            return self.bg_value
        return {"controlnet": 0.0, "zero123pp": 0.8, "simple": 1.0}.get(self.kind, 0.0)


def normalize_depth(depth: np.ndarray, mask: np.ndarray, spec: DepthSpec) -> np.ndarray:
    d = depth.astype(f32)
    if spec.kind == "none":
        return d
    if spec.kind in ("controlnet", "zero123pp"):
        B = d.shape[0]
        lo = d.reshape(B, -1).min(-1)[:, None, None]
        hi = d.reshape(B, -1).max(-1)[:, None, None]
        n = np.clip((d - lo) / ((hi - lo) + f32(1e-5)), f32(0), f32(1))
        if spec.kind == "controlnet":
            n = f32(1.0) - n
            n = n * f32(spec.near_clip - spec.far_clip) + f32(spec.far_clip)
        d = n.astype(f32)
    elif spec.kind == "simple":
        d = d * f32(spec.scale) + f32(spec.offset)
        if spec.clamp:
            d = np.clip(d, f32(0), f32(1))
    else:
        raise ValueError(spec.kind)
    d = d.copy()
    d[~mask] = f32(spec.background())
    return d


# ---------------------------------------------------------------------------------------------
# render.py:220-286  render()   (geometry outputs; attr path = render_attr below)
# ---------------------------------------------------------------------------------------------

def render(v_pos, tri, mvp, w2c, height: int, width: int, v_nrm=None, tri_nrm=None,
           depth: Optional[DepthSpec] = DepthSpec(), normal_background: float = 0.0,
           v_tex=None, tri_tex=None, texture=None, attr_background: float = 0.5,
           texture_filter_mode: str = "linear", nthreads: int = 0, elementwise: str = "numpy") -> dict:
    """elementwise="by_view": the element-wise NumPy tail (view depth, normalisers, normal normalisation) of each
    view runs on its own host thread -- same arithmetic, bit for bit (tests/test_oracle_golden.py).  bench.py uses
    it for the CPU arm so that the arm scales with the host cores, as the reference's torch code would on CPU."""
    if elementwise == "by_view" and texture is None:
        return _render_by_view(v_pos, tri, mvp, w2c, height, width, v_nrm, tri_nrm, depth, normal_background, nthreads)
    v = _a(v_pos)
    tri = np.asarray(tri, np.int32).reshape(-1, 3)
    mvp = _a(mvp).reshape(-1, 4, 4)
    w2c = _a(w2c).reshape(-1, 4, 4)
    clip = shim.clip_positions(v, mvp, nthreads)                      # utils.py:127-129
    rast, ids = shim.rasterize(clip, tri, (height, width), nthreads)  # render.py:241
    mask = rast[..., 3] > 0                                           # render.py:242
    pos = shim.interpolate(v[None], rast, tri, nthreads)              # render.py:244
    out = {"rast": rast, "tri_id": ids, "mask": mask, "pos": pos}
    if depth is not None:
        # render.py:248-249 + utils.py:132-139: only -z of w2c * (p,1) is used
        m = w2c[:, None, None, 2, :]
        zv = ((m[..., 0] * pos[..., 0] + m[..., 1] * pos[..., 1]) + m[..., 2] * pos[..., 2]) + m[..., 3]
        d = (-zv).astype(f32)
        out["depth_view"] = d.copy()
        B = d.shape[0]
        lo = d.reshape(B, -1).min(-1)[:, None, None]                  # render.py:251-255 (min incl. background)
        d = np.where(mask, d, lo).astype(f32)
        out["depth"] = normalize_depth(d, mask, depth)
    if v_nrm is not None:
        tn = tri if tri_nrm is None else np.asarray(tri_nrm, np.int32).reshape(-1, 3)
        n = shim.interpolate(_a(v_nrm)[None], rast, tn, nthreads)     # render.py:275
        ln = np.sqrt(((n[..., 0] * n[..., 0] + n[..., 1] * n[..., 1]) + n[..., 2] * n[..., 2]))
        n = (n / np.maximum(ln, f32(1e-12))[..., None]).astype(f32)   # render.py:276
        n[~mask] = f32(normal_background)                             # render.py:277
        out["normal"] = n
    if texture is not None:
        tt = np.asarray(tri_tex, np.int32).reshape(-1, 3)
        tex_c = shim.interpolate(_a(v_tex)[None], rast, tt, nthreads)          # render.py:261
        fg = shim.texture(_a(texture)[None], tex_c, texture_filter_mode, "wrap", nthreads)  # render.py:267
        out["attr"] = np.where(mask[..., None], fg, f32(attr_background)).astype(f32)      # render.py:268-269
    return out


def _render_by_view(v_pos, tri, mvp, w2c, height, width, v_nrm, tri_nrm, depth, normal_background, nthreads):
    """render() with the element-wise NumPy tail of every view on its own host thread (NumPy releases the GIL inside
    its loops).  Per-view minima / maxima make the views independent, so this is the same arithmetic as render(),
    bit for bit; only the C operators see all views at once."""
    from concurrent.futures import ThreadPoolExecutor
    v = _a(v_pos)
    tri = np.asarray(tri, np.int32).reshape(-1, 3)
    mvp = _a(mvp).reshape(-1, 4, 4)
    w2c = _a(w2c).reshape(-1, 4, 4)
    B = mvp.shape[0]
    clip = shim.clip_positions(v, mvp, nthreads)
    rast, ids = shim.rasterize(clip, tri, (height, width), nthreads)
    pos = shim.interpolate(v[None], rast, tri, nthreads)
    nrm = None
    if v_nrm is not None:
        tn = tri if tri_nrm is None else np.asarray(tri_nrm, np.int32).reshape(-1, 3)
        nrm = shim.interpolate(_a(v_nrm)[None], rast, tn, nthreads)
    out = {"rast": rast, "tri_id": ids, "mask": np.empty(rast.shape[:3], bool), "pos": pos}
    if depth is not None:
        out["depth_view"] = np.empty(rast.shape[:3], f32)
        out["depth"] = np.empty(rast.shape[:3], f32)
    if nrm is not None:
        out["normal"] = np.empty(nrm.shape, f32)

    def tail(b):
        mask = rast[b:b + 1, ..., 3] > 0
        out["mask"][b:b + 1] = mask
        p = pos[b:b + 1]
        if depth is not None:
            m = w2c[b:b + 1, None, None, 2, :]
            zv = ((m[..., 0] * p[..., 0] + m[..., 1] * p[..., 1]) + m[..., 2] * p[..., 2]) + m[..., 3]
            d = (-zv).astype(f32)
            out["depth_view"][b:b + 1] = d
            lo = d.reshape(1, -1).min(-1)[:, None, None]
            d = np.where(mask, d, lo).astype(f32)
            out["depth"][b:b + 1] = normalize_depth(d, mask, depth)
        if nrm is not None:
            n = nrm[b:b + 1]
            ln = np.sqrt(((n[..., 0] * n[..., 0] + n[..., 1] * n[..., 1]) + n[..., 2] * n[..., 2]))
            n = (n / np.maximum(ln, f32(1e-12))[..., None]).astype(f32)
            n[~mask] = f32(normal_background)
            out["normal"][b:b + 1] = n

    workers = max(1, min(B, nthreads if nthreads > 0 else (os.cpu_count() or 1)))
    if workers == 1:
        for b in range(B):
            tail(b)
    else:
        with ThreadPoolExecutor(workers) as ex:
            list(ex.map(tail, range(B)))
    return out


# ---------------------------------------------------------------------------------------------
# F.grid_sample(mode='bilinear', padding_mode='zeros', align_corners=False) as used at
# uv.py:143,155,164,200,213.  img [H,W,C], ndc [...,2] -> [...,C]
# ---------------------------------------------------------------------------------------------

def grid_sample(img: np.ndarray, ndc: np.ndarray) -> np.ndarray:
    img = _a(img)
    if img.ndim == 2:
        img = img[..., None]
    H, W, C = img.shape
    gx, gy = ndc[..., 0].astype(f32), ndc[..., 1].astype(f32)
    with np.errstate(all="ignore"):
        ix = ((gx + f32(1)) * f32(W) - f32(1)) / f32(2)
        iy = ((gy + f32(1)) * f32(H) - f32(1)) / f32(2)
        x0f, y0f = np.floor(ix), np.floor(iy)
        tx, ty = ix - x0f, iy - y0f
    out = np.zeros(gx.shape + (C,), f32)
    bad = ~(np.isfinite(ix) & np.isfinite(iy))
    big = f32(1e9)
    x0 = np.clip(np.where(bad, -big, x0f), -big, big).astype(np.int64)
    y0 = np.clip(np.where(bad, -big, y0f), -big, big).astype(np.int64)
    w00 = (f32(1) - tx) * (f32(1) - ty)
    w10 = tx * (f32(1) - ty)
    w01 = (f32(1) - tx) * ty
    w11 = tx * ty
    for dx, dy, wgt in ((0, 0, w00), (1, 0, w10), (0, 1, w01), (1, 1, w11)):
        xi, yi = x0 + dx, y0 + dy
        ok = (xi >= 0) & (xi < W) & (yi >= 0) & (yi < H) & ~bad
        v = img[np.clip(yi, 0, H - 1), np.clip(xi, 0, W - 1)]
        out = out + np.where(ok[..., None], v * wgt[..., None], f32(0))
    return out.astype(f32)


# ---------------------------------------------------------------------------------------------
# uv.py:24-53  uv_precompute
# ---------------------------------------------------------------------------------------------

def uv_precompute(v_pos, tri, v_tex, tri_tex, height: int, width: int, nthreads: int = 0) -> dict:
    vt = _a(v_tex)
    uvc = vt * f32(2.0) - f32(1.0)                                     # uv.py:31
    clip = np.concatenate([uvc, np.zeros_like(uvc[:, :1]), np.ones_like(uvc[:, :1])], -1)
    rast, ids = shim.rasterize(clip[None], np.asarray(tri_tex, np.int32), (height, width), nthreads)
    uv_mask = rast[0, ..., 3] > 0                                      # uv.py:41
    uv_pos = shim.interpolate(_a(v_pos)[None], rast, np.asarray(tri, np.int32), nthreads)[0]  # uv.py:43
    return {"uv_mask": uv_mask, "uv_pos": uv_pos, "tri_id": ids[0]}


# ---------------------------------------------------------------------------------------------
# uv.py:72-184  uv_render_geometry ; uv.py:193-222  uv_render_attr
# ---------------------------------------------------------------------------------------------

def sobel_dilate(depth: np.ndarray, dilation: int) -> np.ndarray:
    """uv.py:122-141: zero-padded Sobel cross-correlation, magnitude, max_pool(k=d, s=1, p=d//2)."""
    B, H, W = depth.shape
    p = np.zeros((B, H + 2, W + 2), f32)
    p[:, 1:-1, 1:-1] = depth
    def s(dy, dx):
        return p[:, dy:dy + H, dx:dx + W]
    # kernels written out row by row in the tap order of a 3x3 cross-correlation
    gx = (((((s(0, 0) - s(0, 2)) + f32(2) * s(1, 0)) - f32(2) * s(1, 2)) + s(2, 0)) - s(2, 2)).astype(f32)
    gy = (((((s(0, 0) + f32(2) * s(0, 1)) + s(0, 2)) - s(2, 0)) - f32(2) * s(2, 1)) - s(2, 2)).astype(f32)
    g = np.sqrt(gx * gx + gy * gy).astype(f32)
    d = int(dilation)
    pad = d // 2
    q = np.full((B, H + 2 * pad, W + 2 * pad), -np.inf, f32)
    q[:, pad:pad + H, pad:pad + W] = g
    Ho, Wo = H + 2 * pad - d + 1, W + 2 * pad - d + 1
    out = np.full((B, Ho, Wo), -np.inf, f32)
    for dy in range(d):
        for dx in range(d):
            out = np.maximum(out, q[:, dy:dy + Ho, dx:dx + Wo])
    return out


def uv_render_geometry(v_pos, tri, v_nrm, tri_nrm, mvp, w2c, view_h: int, view_w: int, pre: dict,
                       compute_depth_grad: bool = True, depth_grad_dilation: int = 1, nthreads: int = 0) -> dict:
    mvp = _a(mvp).reshape(-1, 4, 4)
    w2c = _a(w2c).reshape(-1, 4, 4)
    B = mvp.shape[0]
    uv_pos = pre["uv_pos"]
    Hu, Wu, _ = uv_pos.shape
    clip = shim.clip_positions(uv_pos.reshape(-1, 3), mvp, nthreads).reshape(B, Hu, Wu, 4)  # uv.py:87-89
    with np.errstate(all="ignore"):
        ndc = (clip[..., :2] / clip[..., 3:4]).astype(f32)                                   # uv.py:90
    r = render(v_pos, tri, mvp, w2c, view_h, view_w, v_nrm=v_nrm, tri_nrm=tri_nrm,
               depth=DepthSpec("simple", scale=1.0, offset=0.0, clamp=False, bg_value=1e2),
               nthreads=nthreads)                                                             # uv.py:92-104
    mask, nrm = r["mask"], r["normal"]
    R = w2c[:, None, None, :3, :3]
    ncs = np.stack([(R[..., i, 0] * nrm[..., 0] + R[..., i, 1] * nrm[..., 1]) + R[..., i, 2] * nrm[..., 2]
                    for i in range(3)], -1).astype(f32)                                       # uv.py:108-110
    ln = np.sqrt((ncs[..., 0] * ncs[..., 0] + ncs[..., 1] * ncs[..., 1]) + ncs[..., 2] * ncs[..., 2])
    ncs = (ncs / np.maximum(ln, f32(1e-12))[..., None]).astype(f32)                          # uv.py:111
    ncs[~mask] = nrm[~mask]                                                                   # uv.py:112
    aoi = np.clip(ncs[..., 2], f32(0), f32(1)).astype(f32)                                    # uv.py:113-119
    out = {"uv_pos_ndc": ndc, "view_mask": mask, "view_position": r["pos"], "view_normal": nrm,
           "view_aoi_cos": aoi, "view_depth": r["depth"], "view_tri_id": r["tri_id"]}
    if compute_depth_grad:
        dg = sobel_dilate(r["depth"], depth_grad_dilation)                                    # uv.py:122-141
        out["view_depth_grad"] = dg
        out["uv_depth_grad"] = np.stack([grid_sample(dg[b], ndc[b])[..., 0] for b in range(B)])  # uv.py:143-145
    else:
        out["view_depth_grad"] = None
        out["uv_depth_grad"] = None
    proj = np.stack([grid_sample(r["pos"][b], ndc[b]) for b in range(B)])                     # uv.py:155-160
    diff = proj - uv_pos[None]
    out["uv_pos_proj"] = proj
    out["uv_pos_error"] = np.sqrt((diff[..., 0] * diff[..., 0] + diff[..., 1] * diff[..., 1])
                                  + diff[..., 2] * diff[..., 2]).astype(f32)                  # uv.py:162
    out["uv_aoi_cos"] = np.stack([grid_sample(aoi[b], ndc[b])[..., 0] for b in range(B)])     # uv.py:164-169
    return out


def uv_render_attr(images, geo: dict, masks=None) -> dict:
    images = _a(images)
    ndc = geo["uv_pos_ndc"]
    B = images.shape[0]
    out = {"uv_attr_proj": np.stack([grid_sample(images[b], ndc[b]) for b in range(B)])}      # uv.py:200-205
    if masks is not None:
        m = _a(masks)
        if m.ndim == 4:
            m = m.mean(-1).astype(f32)                                                        # uv.py:211-212
        out["uv_mask_proj"] = np.stack([grid_sample(m[b], ndc[b])[..., 0] for b in range(B)]) # uv.py:213-218
    else:
        out["uv_mask_proj"] = None
    return out


# ---------------------------------------------------------------------------------------------
# uv.py:248-298 SimpleUVValidityStrategy ; uv.py:317-348 ExponentialBlend ; uv.py:385-468 uv_blend
# (non-Poisson branch, do_uv_padding=False)
# ---------------------------------------------------------------------------------------------

def uv_validity(pre: dict, geo: dict, attr: Optional[dict], pos_error_eps=1e-3, aoi_cos_thresh=0.1,
                mask_thresh=0.9, depth_grad_thresh=None, first_view_dominate=False) -> np.ndarray:
    valid = (geo["uv_pos_error"] < f32(pos_error_eps)) & (geo["uv_aoi_cos"] > f32(aoi_cos_thresh))
    if depth_grad_thresh is not None and geo.get("uv_depth_grad") is not None:
        valid &= geo["uv_depth_grad"] < f32(depth_grad_thresh)
    valid &= pre["uv_mask"][None]
    if attr is not None and attr.get("uv_mask_proj") is not None:
        valid &= attr["uv_mask_proj"] > f32(mask_thresh)
    if first_view_dominate:
        valid[1:] &= ~valid[0:1]
    return valid


def exponential_blend(geo: dict, valid: np.ndarray, alpha=1.0, view_weight=None, normalization="linear") -> np.ndarray:
    w = (geo["uv_aoi_cos"] * valid.astype(f32)).astype(f32)
    with np.errstate(all="ignore"):
        if view_weight is not None:
            expo = (f32(alpha) / _a(view_weight))[:, None, None]
            w = np.power(w, expo).astype(f32)
        else:
            w = np.power(w, f32(alpha)).astype(f32)
        if normalization == "linear":
            s = np.maximum(w.sum(0, keepdims=True, dtype=f32), f32(1e-5))
            return np.clip(w / s, f32(0), f32(1)).astype(f32)
    if normalization == "softmax":
        w = np.where(valid, w, f32(-1e5))
        e = np.exp(w - w.max(0, keepdims=True))
        return (e / e.sum(0, keepdims=True)).astype(f32)
    raise ValueError(normalization)


def uv_blend(pre: dict, geo: dict, attr: Optional[dict], uv_attr_old=None, **kw) -> dict:
    vkw = {k: kw[k] for k in ("pos_error_eps", "aoi_cos_thresh", "mask_thresh", "depth_grad_thresh",
                              "first_view_dominate") if k in kw}
    bkw = {k: kw[k] for k in ("alpha", "view_weight", "normalization") if k in kw}
    valid = uv_validity(pre, geo, attr, **vkw)
    weight = exponential_blend(geo, valid, **bkw)
    valid_any = valid.any(0)                                                                   # uv.py:411
    out = {"uv_valid_mask": valid, "uv_blend_weight": weight, "uv_valid_mask_blend": valid_any,
           "uv_attr_blend": None}
    if attr is None:
        return out
    blend = (attr["uv_attr_proj"] * weight[..., None]).sum(0, dtype=f32)                      # uv.py:421-423
    old = np.zeros_like(blend) if uv_attr_old is None else _a(uv_attr_old)
    va = valid_any[..., None].astype(f32)
    stitched = (blend * va + old * (f32(1) - va)).astype(f32)                                  # uv.py:452-455
    pkw = {k: kw[k] for k in ("do_uv_padding", "uv_padding_radius", "pad_unseen_area", "poisson_blending",
                              "pb_num_iters", "pb_keep_original_border", "pb_grad_mode") if k in kw}
    out["uv_attr_blend"] = atlas_postprocess(blend, stitched, valid_any, pre["uv_mask"], old, **pkw)
    return out


def atlas_postprocess(blend, stitched, valid_any, uv_mask, old, do_uv_padding=False, uv_padding_radius=3,
                      pad_unseen_area=False, poisson_blending=False, pb_num_iters=1000,
                      pb_keep_original_border=True, pb_grad_mode="src", nthreads: int = 0):
    """uv.py:426-461: optional Poisson blend (blend.py, exact sweep count) and seam padding (the oracle's
    statement of the fill that stands in for cvcuda.inpaint, wr_oracle_blend.c)."""
    if poisson_blending:
        assert do_uv_padding                                                                   # uv.py:427
        padded = shim.uv_padding(blend, valid_any, uv_padding_radius, nthreads)                # uv.py:429-431
        if pb_keep_original_border:
            tgt = old                                                                          # uv.py:432-433
        else:
            tgt = shim.uv_padding(stitched, uv_mask, uv_padding_radius, nthreads)              # uv.py:435-443
        res = shim.poisson_blend(padded, valid_any, tgt, pb_num_iters, pb_grad_mode, nthreads)  # uv.py:445-452
    else:
        res = stitched
    if do_uv_padding:                                                                          # uv.py:457-461
        res = shim.uv_padding(res, valid_any if pad_unseen_area else uv_mask, uv_padding_radius, nthreads)
    return res


# ---------------------------------------------------------------------------------------------
# projection.py:54-204  CameraProjection.__call__ (warp_images=False; the IoU rejection of :125-138 is
# evaluated and reported)
# ---------------------------------------------------------------------------------------------

def camera_projection(images, v_pos, tri, v_nrm, tri_nrm, v_tex, tri_tex, texture, mvp, w2c, uv_size: int,
                      masks=None, iou_rejection_threshold=0.8, aoi_cos_valid_threshold=0.3,
                      depth_grad_dilation=5, depth_grad_threshold=0.1, uv_exp_blend_alpha=6.0,
                      uv_exp_blend_view_weight=None, poisson_blending=False, pb_num_iters=1000,
                      pb_keep_original_border=True, from_scratch=False, uv_padding=False,
                      nthreads: int = 0) -> Optional[dict]:
    images = _a(images)
    Nv, H, W, _ = images.shape
    m = None
    if masks is not None:
        m = _a(masks)
        if m.ndim == 4:
            m = m.mean(-1).astype(f32)                                                        # projection.py:97-98
    pre = uv_precompute(v_pos, tri, v_tex, tri_tex, uv_size, uv_size, nthreads)                # projection.py:111
    geo = uv_render_geometry(v_pos, tri, v_nrm, tri_nrm, mvp, w2c, H, W, pre, True,
                             depth_grad_dilation, nthreads)                                   # projection.py:114-123
    if m is not None and iou_rejection_threshold is not None:                                 # projection.py:125-138
        g = (m > f32(0.5)).astype(f32)
        rm = geo["view_mask"].astype(f32)
        inter = g * rm
        union = g + rm - inter
        iou = inter.sum((1, 2)) / union.sum((1, 2))
        if iou.min() < iou_rejection_threshold:
            return None
    attr = uv_render_attr(images, geo, m)                                                     # projection.py:165-169
    bl = uv_blend(pre, geo, attr, uv_attr_old=texture, aoi_cos_thresh=aoi_cos_valid_threshold,
                  depth_grad_thresh=depth_grad_threshold, alpha=uv_exp_blend_alpha,
                  view_weight=uv_exp_blend_view_weight, do_uv_padding=uv_padding, pad_unseen_area=from_scratch,
                  poisson_blending=poisson_blending, pb_num_iters=pb_num_iters,
                  pb_keep_original_border=pb_keep_original_border)                            # projection.py:170-188
    return {"uv_proj": bl["uv_attr_blend"], "uv_proj_mask": bl["uv_valid_mask_blend"],
            "uv_depth_grad": geo["uv_depth_grad"], "uv_aoi_cos": geo["uv_aoi_cos"],
            "pre": pre, "geo": geo, "attr": attr, "blend": bl}
