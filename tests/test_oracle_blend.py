"""CPU tests of the oracle for the atlas post-processing tail (oracle/wr_oracle_blend.c):

* Poisson blending is PINNED: tests/golden/poisson.npz holds outputs of the reference's own
  PoissonBlendingSolver (blend.py:186-324, "torch-native" backend, run unmodified on CPU by oracle/gen_golden.py).
* The seam fill is the oracle's own statement (cvcuda.inpaint is absent): known-answer tests derived by hand, plus
  the reference's uv_blend / CameraProjection control flow (uv.py:426-461) recorded with that fill plugged in."""
import os

import numpy as np
import pytest

from oracle import render_oracle, shim

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name)))


# ---------------------------------------------------------------------------------------------- Poisson

@pytest.mark.parametrize("mode", ["src", "max", "avg"])
@pytest.mark.parametrize("iters", [0, 1, 9, 250])
def test_poisson_matches_reference(mode, iters):
    g = load("poisson.npz")
    out = shim.poisson_blend(g["src"], g["mask"] > 0.5, g["tgt"], iters, mode)
    # conv2d / sum(-1) of the reference have library summation order; values are in [0, 1]
    np.testing.assert_allclose(out, g[f"{mode}_{iters}"], rtol=0, atol=2e-6)
    outside = ~(g["mask"] > 0.5)
    np.testing.assert_array_equal(out[outside], g["tgt"][outside])


def test_poisson_three_channel_mask_matches_reference():
    g = load("poisson.npz")
    m3 = np.repeat(g["mask"][..., None], 3, -1)
    out = shim.poisson_blend(g["src"], m3.mean(-1) > 0.5, g["tgt"], 16, "src")  # blend.py:229-230
    np.testing.assert_allclose(out, g["mask3_16"], rtol=0, atol=2e-6)


def test_poisson_constant_images_are_a_fixed_point():
    src = np.full((20, 24, 3), 0.25, np.float32)
    tgt = np.full((20, 24, 3), 0.75, np.float32)
    mask = np.zeros((20, 24), bool)
    mask[4:15, 3:20] = True
    out = shim.poisson_blend(src, mask, tgt, 37, "src")
    # zero laplacian, constant boundary: (k_in * c + k_out * c) / 4 == c exactly, away from the image border of src
    np.testing.assert_array_equal(out, tgt)


def test_poisson_border_is_never_solved_and_result_is_clamped():
    rng = np.random.default_rng(0)
    src = (rng.random((12, 14, 1)) * 8).astype(np.float32)  # large gradients: unclamped iterate leaves [0, 1]
    tgt = rng.random((12, 14, 1)).astype(np.float32)
    mask = np.ones((12, 14), bool)
    out = shim.poisson_blend(src, mask, tgt, 30, "src")
    np.testing.assert_array_equal(out[0], tgt[0])
    np.testing.assert_array_equal(out[-1], tgt[-1])
    np.testing.assert_array_equal(out[:, 0], tgt[:, 0])
    np.testing.assert_array_equal(out[:, -1], tgt[:, -1])
    assert out.min() >= 0.0 and out.max() <= 1.0
    assert (out[1:-1, 1:-1] == 0.0).any() or (out[1:-1, 1:-1] == 1.0).any()


def test_poisson_converges_to_the_discrete_poisson_equation():
    rng = np.random.default_rng(1)
    H, W = 18, 18
    yy, xx = np.mgrid[0:H, 0:W]
    src = (0.5 + 0.02 * np.sin(0.7 * xx) * np.cos(0.5 * yy))[..., None].astype(np.float32)
    tgt = (0.5 + 0.05 * rng.random((H, W, 1))).astype(np.float32)
    mask = np.zeros((H, W), bool)
    mask[3:15, 3:15] = True
    x = shim.poisson_blend(src, mask, tgt, 4000, "src")[..., 0].astype(np.float64)
    s = src[..., 0].astype(np.float64)
    lap = lambda a: 4 * a[1:-1, 1:-1] - a[:-2, 1:-1] - a[2:, 1:-1] - a[1:-1, :-2] - a[1:-1, 2:]
    res = (lap(x) - lap(s))[mask[1:-1, 1:-1]]
    assert np.abs(res).max() < 5e-6


# ---------------------------------------------------------------------------------------------- seam fill

def test_inpaint_single_known_pixel_floods_the_image():
    img = np.zeros((9, 13, 3), np.uint8)
    img[4, 7] = (10, 200, 77)
    mask = np.ones((9, 13), bool)
    mask[4, 7] = False
    out = shim.inpaint_u8(img, mask, 3)
    assert (out == np.array([10, 200, 77], np.uint8)).all()


def test_inpaint_keeps_known_pixels_and_degenerate_masks():
    rng = np.random.default_rng(2)
    img = rng.integers(0, 256, (17, 11, 3), dtype=np.uint8)
    mask = rng.random((17, 11)) < 0.6
    out = shim.inpaint_u8(img, mask, 2)
    np.testing.assert_array_equal(out[~mask], img[~mask])
    np.testing.assert_array_equal(shim.inpaint_u8(img, np.ones_like(mask), 3), img)   # nothing known: unchanged
    np.testing.assert_array_equal(shim.inpaint_u8(img, np.zeros_like(mask), 3), img)  # nothing to fill


def test_inpaint_radius_zero_copies_the_nearest_known_pixel():
    img = np.zeros((6, 20, 1), np.uint8)
    img[:, :4] = 50
    img[:, 16:] = 200
    mask = np.ones((6, 20), bool)
    mask[:, :4] = False
    mask[:, 16:] = False
    out = shim.inpaint_u8(img, mask, 0)[..., 0]
    assert (out[:, 4:10] == 50).all() and (out[:, 10:16] == 200).all()
    # equidistant would be broken towards the smaller seed index; columns 9 | 10 are 6 away from 3 | 16: no tie here
    assert out[0, 9] == 50 and out[0, 10] == 200


def test_inpaint_weights_by_hand():
    # one row: known pixels at columns 0 (value 0) and 1 (value 90); pixel 2 is filled with radius 1
    img = np.array([[[0], [90], [0], [0]]], np.uint8)
    mask = np.array([[0, 0, 1, 1]], bool)
    out = shim.inpaint_u8(img, mask, 1)[0, :, 0]
    # p = 2: nearest known q = 1, D = 1; window around q (radius 1): t = 0 (d^2 = 4), t = 1 (d^2 = 1)
    w0, w1 = np.float32(2) / np.float32(5), np.float32(2) / np.float32(2)
    want2 = np.rint((w0 * np.float32(0) + w1 * np.float32(90)) / (w0 + w1))
    # p = 3: q = 1, D = 4; t = 0 (d^2 = 9), t = 1 (d^2 = 4)
    v0, v1 = np.float32(5) / np.float32(10), np.float32(5) / np.float32(5)
    want3 = np.rint((v0 * np.float32(0) + v1 * np.float32(90)) / (v0 + v1))
    assert out[2] == want2 == 64 and out[3] == want3 == 60


def test_inpaint_values_stay_within_the_known_range_and_seeds_are_near():
    rng = np.random.default_rng(3)
    img = rng.integers(40, 200, (40, 56, 3), dtype=np.uint8)
    mask = np.ones((40, 56), bool)
    for _ in range(12):
        r, c = rng.integers(0, 36), rng.integers(0, 52)
        mask[r:r + 4, c:c + 4] = False
    out = shim.inpaint_u8(img, mask, 3)
    assert out[mask].min() >= img[~mask].min() and out[mask].max() <= img[~mask].max()


def test_uv_padding_quantises_like_the_reference():
    rng = np.random.default_rng(4)
    attr = rng.random((16, 16, 3)).astype(np.float32) * 1.2 - 0.1
    inside = np.zeros((16, 16), bool)
    inside[4:12, 4:12] = True
    out = shim.uv_padding(attr, inside, 3)
    want = (np.clip(attr, 0, 1) * np.float32(255)).astype(np.uint8).astype(np.float32) / np.float32(255)  # cv_ops.py:23-35
    np.testing.assert_array_equal(out[inside], want[inside])
    assert np.isin(np.rint(out * 255), np.arange(256)).all()


# ---------------------------------------------------------------------------------------------- bake tail

def _bake(g, **kw):
    return render_oracle.camera_projection(g["images"], g["v_pos"], g["t_pos_idx"], g["v_nrm"], g["t_pos_idx"],
                                           g["v_tex"], g["t_tex_idx"], g["texture"], g["mvp"], g["w2c"], 64,
                                           iou_rejection_threshold=None, aoi_cos_valid_threshold=0.2,
                                           depth_grad_threshold=0.1, uv_exp_blend_alpha=3.0, depth_grad_dilation=5, **kw)


@pytest.mark.parametrize("name,kw", [
    ("bake_pad", dict(uv_padding=True)),
    ("bake_pad_scratch", dict(uv_padding=True, from_scratch=True)),
    ("bake_pb", dict(uv_padding=True, poisson_blending=True, pb_num_iters=40)),
    ("bake_pb_noborder", dict(uv_padding=True, poisson_blending=True, pb_num_iters=40, pb_keep_original_border=False)),
])
def test_bake_tail_matches_reference_control_flow(name, kw):
    """uv.py:426-461 run by the reference itself (with the oracle's fill behind cvcuda.inpaint) against the
    oracle's restatement.  A validity flip at a threshold changes one texel and, through the fill,
    its neighbourhood, so the comparison is a fraction of texels within one quantisation step."""
    g = load("bake_sphere.npz")
    want = load("poisson.npz")[name]
    got = _bake(g, **kw)["uv_proj"]
    assert got.shape == want.shape
    close = np.abs(got - want).max(-1) <= 1.5 / 255
    assert close.mean() > 0.995, close.mean()
    assert np.abs(got - want).mean() < 2e-4
