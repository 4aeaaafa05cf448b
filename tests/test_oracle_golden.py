"""Pins oracle/render_oracle.py (the NumPy restatement the GPU box compares the kernels with) against
outputs of the UNMODIFIED reference Python (render.py, uv.py, projection.py, mesh.py) recorded by
oracle/gen_golden.py in tests/golden/.  Runs on CPU, no reference tree needed.

The reference computes clip positions with torch.matmul (utils.py:129), whose summation order is the BLAS
library's; the oracle uses the contract's fixed order.  Both feed the same C rasterizer, so a vertex that
lands within one rounding of a snap boundary can move a triangle edge by 1/16 pixel: a handful of pixels
may differ in coverage.  The tests therefore allow a tiny, explicitly counted, mismatch budget on masks
and compare the float maps on the pixels where both agree."""
import os

import numpy as np
import pytest

from oracle import render_oracle
from oracle.render_oracle import DepthSpec

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MASK_BUDGET = 2e-4   # fraction of pixels allowed to differ in coverage (see module docstring)
RTOL, ATOL = 2e-5, 2e-6


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name)))


def agree(mask_a, mask_b):
    diff = mask_a != mask_b
    assert diff.mean() <= MASK_BUDGET, f"{diff.sum()} of {diff.size} pixels differ in coverage"
    return ~diff


def test_vertex_normals_match_reference():
    g = load("render_sphere.npz")
    n = render_oracle.vertex_normals(g["v_pos"], g["t_pos_idx"])
    np.testing.assert_allclose(n, g["v_nrm"], rtol=1e-5, atol=1e-6)
    g = load("render_terrain.npz")
    n = render_oracle.vertex_normals(g["v_pos"], g["t_pos_idx"])
    np.testing.assert_allclose(n, g["v_nrm"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("key,spec", [
    ("depth_controlnet", DepthSpec("controlnet")),
    ("depth_zero123pp", DepthSpec("zero123pp")),
    ("depth_simple", DepthSpec("simple", scale=1.0, offset=-1.0, clamp=True)),
    ("depth_none", DepthSpec("none")),
])
def test_render_sphere_matches_reference(key, spec):
    g = load("render_sphere.npz")
    r = render_oracle.render(g["v_pos"], g["t_pos_idx"], g["mvp"], g["w2c"], 64, 64, v_nrm=g["v_nrm"], depth=spec,
                             v_tex=g["v_tex"], tri_tex=g["t_tex_idx"], texture=g["texture"], attr_background=0.25)
    ok = agree(r["mask"], g["mask"])
    np.testing.assert_allclose(r["pos"][ok], g["pos"][ok], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(r["normal"][ok], g["normal"][ok], rtol=RTOL, atol=ATOL)
    # min / max normalisation propagates a one-pixel coverage difference to every pixel only through lo / hi,
    # which moves by far less than the tolerance here
    np.testing.assert_allclose(r["depth"][ok], g[key][ok], rtol=1e-4, atol=1e-5)
    if key == "depth_controlnet":
        np.testing.assert_allclose(r["attr"][ok], g["attr_linear"][ok], rtol=1e-4, atol=1e-5)


def test_render_sphere_nearest_texture_and_normal_background():
    g = load("render_sphere.npz")
    r = render_oracle.render(g["v_pos"], g["t_pos_idx"], g["mvp"], g["w2c"], 64, 64, v_nrm=g["v_nrm"],
                             depth=DepthSpec("none"), normal_background=0.5, v_tex=g["v_tex"], tri_tex=g["t_tex_idx"],
                             texture=g["texture"], attr_background=0.5, texture_filter_mode="nearest")
    ok = agree(r["mask"], g["mask"])
    np.testing.assert_allclose(r["normal"][ok], g["normal_bg05"][ok], rtol=RTOL, atol=ATOL)
    # nearest filtering is discontinuous: a texel boundary within rounding of the sample flips the texel
    close = np.isclose(r["attr"], g["attr_nearest"], rtol=1e-4, atol=1e-5).all(-1)
    assert (close | ~ok).mean() > 0.999


@pytest.mark.parametrize("cam", ["persp", "inside"])
def test_render_terrain_perspective_matches_reference(cam):
    g = load("render_terrain.npz")
    r = render_oracle.render(g["v_pos"], g["t_pos_idx"], g[f"{cam}_mvp"], g[f"{cam}_w2c"], 48, 64, v_nrm=g["v_nrm"],
                             depth=DepthSpec("controlnet"))
    ok = agree(r["mask"], g[f"{cam}_mask"])
    assert g[f"{cam}_mask"].sum() > 500
    np.testing.assert_allclose(r["pos"][ok], g[f"{cam}_pos"][ok], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(r["normal"][ok], g[f"{cam}_normal"][ok], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(r["depth"][ok], g[f"{cam}_depth"][ok], rtol=1e-4, atol=1e-5)


def _bake(g, **kw):
    return render_oracle.camera_projection(g["images"], g["v_pos"], g["t_pos_idx"], g["v_nrm"], g["t_pos_idx"],
                                           g["v_tex"], g["t_tex_idx"], g["texture"], g["mvp"], g["w2c"], 64, **kw)


def _stable(ref, aoi_thr, dg_thr, eps=1e-3):
    geo = ref["geo"]
    near = np.abs(geo["uv_pos_error"] - eps) < 5e-5 * eps + 1e-7
    near |= np.abs(geo["uv_aoi_cos"] - aoi_thr) < 5e-5
    if dg_thr is not None:
        near |= np.abs(geo["uv_depth_grad"] - dg_thr) < 5e-5 * max(1.0, dg_thr)
    return ~near.any(0)


@pytest.mark.parametrize("name,kw", [
    ("a", dict(aoi_cos_valid_threshold=0.2, depth_grad_threshold=0.1, uv_exp_blend_alpha=3.0, depth_grad_dilation=5)),
    ("b", dict(aoi_cos_valid_threshold=-1.0, depth_grad_threshold=None, uv_exp_blend_alpha=3.0, depth_grad_dilation=5)),
    ("c", dict(aoi_cos_valid_threshold=0.3, depth_grad_threshold=0.1, uv_exp_blend_alpha=6.0, depth_grad_dilation=3)),
])
def test_bake_matches_reference(name, kw):
    g = load("bake_sphere.npz")
    vw = {"a": g["view_weight"], "b": np.ones(6, np.float32), "c": None}[name]
    ref = _bake(g, iou_rejection_threshold=None, uv_exp_blend_view_weight=vw, **kw)
    np.testing.assert_array_equal(ref["pre"]["uv_mask"], g["uv_mask"])
    inside = g["uv_mask"]
    np.testing.assert_allclose(ref["uv_aoi_cos"][:, inside], g[f"{name}_uv_aoi_cos"][:, inside], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(ref["uv_depth_grad"][:, inside], g[f"{name}_uv_depth_grad"][:, inside], rtol=1e-4, atol=2e-3)  # Sobel of values ~1e2: cancellation
    stable = _stable(ref, kw["aoi_cos_valid_threshold"], kw["depth_grad_threshold"])
    assert stable.mean() > 0.98
    same = ref["uv_proj_mask"] == g[f"{name}_uv_proj_mask"]
    assert same[stable].mean() > 0.9995
    sel = stable & same
    np.testing.assert_allclose(ref["uv_proj"][sel], g[f"{name}_uv_proj"][sel], rtol=2e-4, atol=1e-5)
    assert g[f"{name}_uv_proj_mask"].sum() > 300  # the comparison above is not vacuous


def test_bake_with_view_masks_matches_reference():
    g = load("bake_sphere.npz")
    ref = _bake(g, masks=g["masks"])
    assert ref is not None
    stable = _stable(ref, 0.3, 0.1) & ~(np.abs(ref["attr"]["uv_mask_proj"] - 0.9) < 1e-4).any(0)
    same = ref["uv_proj_mask"] == g["m_uv_proj_mask"]
    assert same[stable].mean() > 0.9995
    sel = stable & same
    np.testing.assert_allclose(ref["uv_proj"][sel], g["m_uv_proj"][sel], rtol=2e-4, atol=1e-5)
    # IoU rejection branch (projection.py:125-138)
    bad = np.zeros_like(g["masks"]); bad[:, :5, :5] = 1
    assert _bake(g, masks=bad) is None


def test_bake_intermediates_match_reference():
    g = load("bake_sphere.npz")
    pre = render_oracle.uv_precompute(g["v_pos"], g["t_pos_idx"], g["v_tex"], g["t_tex_idx"], 64, 64)
    np.testing.assert_array_equal(pre["uv_mask"], g["uv_mask"])
    np.testing.assert_allclose(pre["uv_pos"], g["uv_pos"], rtol=RTOL, atol=ATOL)
    geo = render_oracle.uv_render_geometry(g["v_pos"], g["t_pos_idx"], g["v_nrm"], g["t_pos_idx"], g["mvp"], g["w2c"],
                                           48, 48, pre, True, 5)
    inside = g["uv_mask"]
    np.testing.assert_allclose(geo["uv_pos_ndc"][:, inside], g["uv_pos_ndc"][:, inside], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(geo["view_aoi_cos"], g["view_aoi_cos"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(geo["view_depth"], g["view_depth"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(geo["view_depth_grad"], g["view_depth_grad"], rtol=1e-4, atol=1e-3)
    err_ok = np.isclose(geo["uv_pos_error"][:, inside], g["uv_pos_error"][:, inside], rtol=1e-3, atol=1e-5)
    assert err_ok.mean() > 0.999


def test_validity_and_blend_restatements_match_reference_strategies():
    """oracle uv_validity / exponential_blend against recordings of the reference's SimpleUVValidityStrategy and
    ExponentialBlend (uv.py:248-348) with every option (gen_golden.py strategy_cases)."""
    g = dict(np.load(os.path.join(GOLDEN, "strategies.npz")))
    pre = {"uv_mask": g["uv_mask"]}
    geo = {k: g[k] for k in ("uv_pos_error", "uv_aoi_cos", "uv_depth_grad")}
    geo_nograd = dict(geo, uv_depth_grad=None)
    attr = {"uv_mask_proj": g["uv_mask_proj"]}
    cases_ = {
        "v_default": (dict(), geo, attr),
        "v_thresholds": (dict(pos_error_eps=5e-4, aoi_cos_thresh=0.3, mask_thresh=0.5, depth_grad_thresh=0.1), geo, attr),
        "v_grad_missing": (dict(depth_grad_thresh=0.1), geo_nograd, attr),
        "v_no_view_mask": (dict(aoi_cos_thresh=0.2, depth_grad_thresh=0.15), geo, {"uv_mask_proj": None}),
        "v_first_view": (dict(aoi_cos_thresh=0.2, first_view_dominate=True), geo, attr),
    }
    for name, (kw, ge, at) in cases_.items():
        np.testing.assert_array_equal(render_oracle.uv_validity(pre, ge, at, **kw), g[name], err_msg=name)
    valid = g["v_default"]
    blends = {
        "w_linear_a1": dict(alpha=1.0),
        "w_linear_a3": dict(alpha=3.0),
        "w_linear_a6_vw": dict(alpha=6.0, view_weight=g["view_weight"]),
        "w_softmax_a2": dict(alpha=2.0, normalization="softmax"),
        "w_softmax_a3_vw": dict(alpha=3.0, normalization="softmax", view_weight=g["view_weight"]),
    }
    for name, kw in blends.items():
        got = render_oracle.exponential_blend(geo, valid.copy(), **kw)
        np.testing.assert_allclose(got, g[name], rtol=2e-6, atol=1e-7, err_msg=name)
