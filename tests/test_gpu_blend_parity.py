"""GPU parity of the atlas post-processing tail (csrc/blend.cu) against oracle/wr_oracle_blend.c.

Both sides evaluate the same individually rounded fp32 operations in the same order, so every comparison
with the oracle is BIT-EXACT (assert_array_equal); the committed outputs of the reference's own solver
(tests/golden/poisson.npz) are matched within the summation-order tolerance of its conv2d / sum(-1)."""
import os

import numpy as np
import pytest
import torch

import cases
import worldrenderer_b200 as wr
from oracle import render_oracle, shim
from test_gpu_render_parity import make_mesh
from worldrenderer_b200 import cv_ops, synth
from worldrenderer_b200.uv import uv_padding

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _images(H, W, C, seed):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W]
    src = 0.5 + 0.4 * np.sin(0.3 * xx + 0.2 * yy)[..., None] * np.linspace(1, 0.6, C) + 0.05 * rng.random((H, W, C))
    tgt = 0.4 + 0.3 * np.cos(0.15 * xx - 0.25 * yy)[..., None] * np.linspace(0.7, 1, C) + 0.05 * rng.random((H, W, C))
    return src.astype(np.float32), tgt.astype(np.float32)


def _blob_mask(H, W, seed, density=0.55):
    rng = np.random.default_rng(seed)
    m = np.zeros((H, W), bool)
    for _ in range(max(3, H * W // 900)):
        r, c = rng.integers(0, H), rng.integers(0, W)
        h, w = rng.integers(1, max(2, H // 3)), rng.integers(1, max(2, W // 3))
        m[r:r + h, c:c + w] = True
    m &= rng.random((H, W)) < (0.5 + density / 2)
    return m


@pytest.mark.parametrize("mode", ["src", "max", "avg"])
@pytest.mark.parametrize("iters", [0, 1, 9, 250])
def test_poisson_golden_inputs(cuda_device, mode, iters):
    g = dict(np.load(os.path.join(GOLDEN, "poisson.npz")))
    solver = wr.PoissonBlendingSolver("torch-native", str(cuda_device))
    tgt = torch.from_numpy(g["tgt"]).to(cuda_device)
    out = solver(torch.from_numpy(g["src"]), torch.from_numpy(g["mask"]), tgt, iters, inplace=False, grad_mode=mode)
    assert out.data_ptr() != tgt.data_ptr()
    np.testing.assert_array_equal(tgt.cpu().numpy(), g["tgt"])  # inplace=False leaves the target alone
    want = shim.poisson_blend(g["src"], g["mask"] > 0.5, g["tgt"], iters, mode)
    np.testing.assert_array_equal(out.cpu().numpy(), want)
    np.testing.assert_allclose(out.cpu().numpy(), g[f"{mode}_{iters}"], rtol=0, atol=2e-6)  # the reference itself


@pytest.mark.parametrize("H,W,C,iters", [(1, 1, 3, 4), (3, 3, 1, 5), (50, 113, 3, 8), (130, 250, 4, 21), (49, 225, 2, 16),
                                         (97, 96, 3, 100)])
def test_poisson_shapes_and_tile_edges(cuda_device, H, W, C, iters):
    src, tgt = _images(H, W, C, seed=H * W)
    mask = _blob_mask(H, W, seed=W) | (np.arange(W)[None, :] % 112 < 3) & (np.arange(H)[:, None] >= 0)  # across tile seams
    solver = wr.PoissonBlendingSolver("torch-native", str(cuda_device))
    for mode in ("src", "max"):
        out = solver(torch.from_numpy(src), torch.from_numpy(mask), torch.from_numpy(tgt), iters, inplace=False,
                     grad_mode=mode)
        np.testing.assert_array_equal(out.cpu().numpy(), shim.poisson_blend(src, mask, tgt, iters, mode))


def test_poisson_full_atlas_1000_sweeps(cuda_device):
    H = W = 1024
    src, tgt = _images(H, W, 3, seed=5)
    mask = _blob_mask(H, W, seed=6)
    assert mask.mean() > 0.2
    solver = wr.PoissonBlendingSolver("torch-cuda", str(cuda_device))
    out = solver(torch.from_numpy(src), torch.from_numpy(mask), torch.from_numpy(tgt), 1000, inplace=False)
    np.testing.assert_array_equal(out.cpu().numpy(), shim.poisson_blend(src, mask, tgt, 1000, "src"))


def test_poisson_backend_sweep_count_and_inplace(cuda_device):
    src, tgt = _images(40, 44, 3, seed=9)
    mask = _blob_mask(40, 44, seed=10)
    s, t = torch.from_numpy(src).to(cuda_device), torch.from_numpy(tgt).to(cuda_device)
    m = torch.from_numpy(mask).to(cuda_device)
    # pointer-swapping backends return sweep num_iters - 1 for odd counts (blend.py:88-100, 166-169)
    for backend, eff in (("torch-native", 9), ("torch-cuda", 8), ("triton", 8)):
        out = wr.PoissonBlendingSolver(backend, str(cuda_device))(s, m, t, 9, inplace=False)
        np.testing.assert_array_equal(out.cpu().numpy(), shim.poisson_blend(src, mask, tgt, eff, "src"))
    with pytest.raises(ValueError):
        wr.PoissonBlendingSolver("cupy", str(cuda_device))
    t2 = t.clone()
    out = wr.PoissonBlendingSolver("torch-native", str(cuda_device))(s, m.float(), t2, 12)  # inplace=True is the default
    assert out.data_ptr() == t2.data_ptr()
    np.testing.assert_array_equal(t2.cpu().numpy(), shim.poisson_blend(src, mask, tgt, 12, "src"))


@pytest.mark.parametrize("H,W,C,radius,known", [(1, 1, 3, 3, 1.0), (7, 300, 3, 3, 0.05), (64, 64, 1, 0, 0.02),
                                                (200, 131, 4, 5, 0.3), (512, 512, 3, 3, 0.001), (33, 17, 3, 3, 0.0)])
def test_inpaint_u8_bit_exact(cuda_device, H, W, C, radius, known):
    rng = np.random.default_rng(H + W)
    img = rng.integers(0, 256, (H, W, C), dtype=np.uint8)
    mask = rng.random((H, W)) >= known
    out = cv_ops.inpaint_cvc(torch.from_numpy(img).to(cuda_device), torch.from_numpy(mask).to(cuda_device), radius)
    assert out.dtype == torch.uint8
    np.testing.assert_array_equal(out.cpu().numpy(), shim.inpaint_u8(img, mask, radius))


def test_inpaint_float_contract_and_uv_padding(cuda_device):
    rng = np.random.default_rng(11)
    H, W = 256, 320
    attr = (rng.random((H, W, 3)) * 1.2 - 0.1).astype(np.float32)
    inside = _blob_mask(H, W, seed=12)
    out = uv_padding(torch.from_numpy(attr).to(cuda_device), torch.from_numpy(inside).to(cuda_device), 3)
    np.testing.assert_array_equal(out.cpu().numpy(), shim.uv_padding(attr, inside, 3))
    # inpaint_cvc on a float image in [0, 1]: same quantisation, float result (cv_ops.py:23-35)
    a01 = np.clip(attr, 0, 1)
    got = cv_ops.inpaint_cvc(torch.from_numpy(a01).to(cuda_device), torch.from_numpy(~inside).to(cuda_device), 3)
    assert got.dtype == torch.float32
    # the final "/ 255.0" is torch's here (a multiplication by the rounded reciprocal on CUDA): one ulp
    np.testing.assert_allclose(got.cpu().numpy(), shim.uv_padding(a01, inside, 3), rtol=2e-7, atol=0)
    np.testing.assert_array_equal(np.rint(got.cpu().numpy() * 255), np.rint(shim.uv_padding(a01, inside, 3) * 255))
    both = cv_ops.batch_inpaint_cvc(torch.from_numpy(np.stack([a01, a01])).to(cuda_device),
                                    torch.from_numpy(np.stack([~inside, ~inside])).to(cuda_device), 3)
    np.testing.assert_array_equal(both[1].cpu().numpy(), got.cpu().numpy())


def test_full_size_padding_properties(cuda_device):
    """4096^2 atlas (config E size): known texels come back quantised, every texel is filled, idempotent."""
    H = W = 4096
    g = torch.Generator(device="cpu").manual_seed(0)
    attr = torch.rand((H, W, 3), generator=g).to(cuda_device)
    inside = torch.zeros((H, W), dtype=torch.bool, device=cuda_device)
    inside[100:900, 50:3000] = True
    inside[2000:2100, 2000:4090] = True
    out = uv_padding(attr, inside, 3)
    q8 = (attr.clamp(0, 1) * 255).to(torch.uint8)                 # cv_ops.py:23-24
    out8 = torch.round(out * 255).to(torch.uint8)
    assert torch.equal(out8[inside], q8[inside])
    assert torch.equal(out, out8.float().cpu().div(255.0).to(out.device))  # exactly u8 / 255 (true division)
    again = uv_padding(out, inside, 3)
    assert torch.equal(again, out)                                # k/255 survives the (x * 255) truncation for every k
    q = q8.float() / 255.0
    far = out[3500, 100]   # far from every chart: equals an average of known texels, hence inside their range
    assert (far >= q[inside].min()).all() and (far <= q[inside].max()).all()


def _bake_setup(device):
    v, f = cases.icosphere_mesh(8)
    mesh = make_mesh(v, f, device, with_uv=True, tex_size=128, seed=1)
    cam = cases.canonical_cameras(device=device)
    images = synth.view_images(6, 96, 96, seed=1)
    return mesh, cam, images


@pytest.mark.parametrize("kw", [dict(poisson_blending=False, uv_padding=True),
                                dict(poisson_blending=False, uv_padding=True, from_scratch=True),
                                dict(poisson_blending=True, uv_padding=True, pb_num_iters=64),
                                dict(poisson_blending=True, uv_padding=True, pb_num_iters=64, pb_keep_original_border=False)])
def test_camera_projection_with_tail(wr_ctx, kw):
    mesh, cam, images = _bake_setup(wr_ctx.device)
    proj = wr.CameraProjection("torch-cuda", None, str(wr_ctx.device), "cuda")
    common = dict(iou_rejection_threshold=None, aoi_cos_valid_threshold=0.2, depth_grad_threshold=0.1,
                  uv_exp_blend_alpha=3.0)
    out, valid = proj(torch.from_numpy(images), mesh, cam, uv_size=128, return_uv_projection_mask=True, **common, **kw)
    n32 = lambda t: t.cpu().numpy().astype(np.int32)
    ref = render_oracle.camera_projection(
        images, mesh.v_pos.cpu().numpy(), n32(mesh.t_pos_idx), mesh.v_nrm.cpu().numpy(), n32(mesh.t_pos_idx),
        mesh.v_tex.cpu().numpy(), n32(mesh.t_tex_idx), mesh.texture.cpu().numpy(), cam.mvp_mtx.cpu().numpy(),
        cam.w2c.cpu().numpy(), 128, **common, **kw)
    same_valid = valid.cpu().numpy() == ref["uv_proj_mask"]
    assert same_valid.mean() > 0.999
    got, want = out.cpu().numpy(), ref["uv_proj"]
    # a validity flip at a threshold moves a texel and, through the fill, its neighbours: count, do not forbid
    close = np.abs(got - want).max(-1) <= 1.5 / 255
    assert close.mean() > 0.995, close.mean()
    if same_valid.all():
        np.testing.assert_allclose(got, want, rtol=0, atol=1.01 / 255)


def test_tail_given_identical_inputs_is_bit_exact(wr_ctx):
    """The same tail fed with the oracle's own blend / validity: no threshold effects left, exact equality."""
    from worldrenderer_b200.uv import UVPrecomputeOutput, atlas_postprocess
    rng = np.random.default_rng(21)
    Hu = Wu = 160
    uv_mask = _blob_mask(Hu, Wu, seed=22)
    valid_any = uv_mask & (rng.random((Hu, Wu)) < 0.7)
    blend = rng.random((Hu, Wu, 3)).astype(np.float32) * valid_any[..., None]
    old = rng.random((Hu, Wu, 3)).astype(np.float32)
    stitched = np.where(valid_any[..., None], blend, old)
    dev = wr_ctx.device
    pre = UVPrecomputeOutput(height=Hu, width=Wu, uv_attr=torch.from_numpy(old).to(dev),
                             uv_mask=torch.from_numpy(uv_mask).to(dev), uv_pos=None)
    solver = wr.PoissonBlendingSolver("torch-native", str(dev))
    for kw in (dict(do_uv_padding=True), dict(do_uv_padding=True, pad_unseen_area=True),
               dict(do_uv_padding=True, poisson_blending=True, pb_num_iters=33),
               dict(do_uv_padding=True, poisson_blending=True, pb_num_iters=33, pb_keep_original_border=False,
                    pb_grad_mode="max")):
        want = render_oracle.atlas_postprocess(blend, stitched, valid_any, uv_mask, old, **kw)
        for b in (torch.from_numpy(blend).to(dev), None):  # the fused bake passes blend=None
            got = atlas_postprocess(b, torch.from_numpy(stitched).to(dev), torch.from_numpy(valid_any).to(dev), pre,
                                    pb_solver=solver, **kw)
            np.testing.assert_array_equal(got.cpu().numpy(), want)


def test_render_between_blend_calls_is_unaffected(wr_ctx):
    """The blend scratch lives behind the raster's self-cleaning prefix of the same context."""
    from worldrenderer_b200 import _native
    v, f = cases.icosphere_mesh(6)
    mesh = make_mesh(v, f, wr_ctx.device)
    cam = cases.canonical_cameras(device=wr_ctx.device)
    first = wr.render(wr_ctx, mesh, cam, 96, 96, render_attr=False)
    src, tgt = _images(300, 300, 3, seed=1)
    mask = _blob_mask(300, 300, seed=2)
    c = wr_ctx.ctx
    s, t = torch.from_numpy(src).to(wr_ctx.device), torch.from_numpy(tgt).to(wr_ctx.device)
    m = torch.from_numpy(mask).to(wr_ctx.device).view(torch.uint8)
    out = torch.empty_like(t)
    c.check(_native.lib().wr_poisson_blend(c.handle, s.data_ptr(), m.data_ptr(), t.data_ptr(), 300, 300, 3, 24, 0,
                                           out.data_ptr(), c.stream()), "wr_poisson_blend")
    np.testing.assert_array_equal(out.cpu().numpy(), shim.poisson_blend(src, mask, tgt, 24, "src"))
    second = wr.render(wr_ctx, mesh, cam, 96, 96, render_attr=False)
    assert torch.equal(first.mask, second.mask) and torch.equal(first.pos, second.pos)
    assert torch.equal(first.depth, second.depth)
