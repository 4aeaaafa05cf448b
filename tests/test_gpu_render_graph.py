"""RenderGraph (worldrenderer_b200/graph.py): a CUDA-graph replay of render() calls gives the same tensors as the
eager calls, for one job and for a config-D style batch of meshes, and follows in-place updates of its inputs."""
import numpy as np
import pytest
import torch

import cases
import worldrenderer_b200 as wr

pytestmark = pytest.mark.gpu


def _mesh(v, f, dev):
    m = wr.TexturedMesh(v_pos=torch.tensor(v, dtype=torch.float32), t_pos_idx=torch.tensor(f, dtype=torch.int64))
    m.set_stitched_mesh(m.v_pos, m.t_pos_idx)
    m.to(dev)
    return m


def _same(a, b):
    for name in ("mask", "pos", "depth", "normal"):
        assert torch.equal(getattr(a, name), getattr(b, name)), name


@pytest.mark.parametrize("lanes", [1, 2, 3])
def test_graph_replay_equals_eager(wr_ctx, lanes):
    dev = wr_ctx.device
    cam = cases.canonical_cameras(device=dev)
    meshes = [_mesh(*cases.terrain_mesh(96, 64, seed=s), dev) for s in range(3)] + [_mesh(*cases.icosphere_mesh(8), dev)]
    g = wr.RenderGraph(wr_ctx, [(m, cam) for m in meshes], 64, 64, lanes=lanes, render_attr=False)
    for _ in range(3):   # replays are idempotent (the packed buffer cleans itself)
        outs = g.replay()
    torch.cuda.synchronize()
    for m, o in zip(meshes, outs):
        _same(o, wr.render(wr_ctx, m, cam, 64, 64, render_attr=False))
    # in-place update of a captured input is picked up by the next replay (the cached normals stay as they are,
    # for the eager call and the replay alike)
    with torch.no_grad():
        meshes[0].v_pos.mul_(0.8)
    want = wr.render(wr_ctx, meshes[0], cam, 64, 64, render_attr=False)
    outs = g.replay()
    torch.cuda.synchronize()
    _same(outs[0], want)


@pytest.mark.parametrize("stagger", [False, True])
@pytest.mark.parametrize("view_lanes", [2, 3, 4, 6])
def test_view_lanes_equal_eager(wr_ctx, view_lanes, stagger):
    """One job, its views split over concurrent lanes that render into slices of the job's outputs; staggered: group
    k + 1 waits for the raster-done event of group k (wr_render_args.raster_done_event)."""
    dev = wr_ctx.device
    cam = cases.canonical_cameras(device=dev)
    meshes = [_mesh(*cases.terrain_mesh(96, 64, seed=1), dev), _mesh(*cases.icosphere_mesh(8), dev)]
    g = wr.RenderGraph(wr_ctx, [(m, cam) for m in meshes], 64, 64, view_lanes=view_lanes, stagger=stagger,
                       render_attr=False)
    for _ in range(3):
        outs = g.replay()
    torch.cuda.synchronize()
    for m, o in zip(meshes, outs):
        assert o.mask.shape == (6, 64, 64)
        _same(o, wr.render(wr_ctx, m, cam, 64, 64, render_attr=False))
    # a normaliser that allocates its own result goes through the copy path
    g2 = wr.RenderGraph(wr_ctx, [(meshes[0], cam)], 64, 64, view_lanes=2, render_attr=False,
                        depth_normalization_strategy=wr.Zero123PlusPlusNormalization())
    o = g2.replay()[0]
    torch.cuda.synchronize()
    _same(o, wr.render(wr_ctx, meshes[0], cam, 64, 64, render_attr=False,
                       depth_normalization_strategy=wr.Zero123PlusPlusNormalization()))


@pytest.mark.parametrize("tail", [False, True])
def test_bake_graph_replay_equals_eager(wr_ctx, tail):
    """BakeGraph: one CameraProjection call captured as a CUDA graph returns what the eager call returns, also
    after the images were updated in place; with the padding + Poisson tail as well."""
    from test_gpu_bake_parity import _setup
    from worldrenderer_b200 import synth
    dev = wr_ctx.device
    mesh, cam, images = _setup(dev)
    img = torch.from_numpy(images).to(dev)
    proj = wr.CameraProjection("torch-cuda" if tail else None, None, str(dev), "cuda")
    kw = dict(uv_size=128, poisson_blending=tail, uv_padding=tail, pb_num_iters=40, depth_grad_dilation=5,
              uv_exp_blend_alpha=3.0, uv_exp_blend_view_weight=torch.tensor([1.0, 0.5, 1.0, 2.0, 1.0, 1.0]),
              aoi_cos_valid_threshold=0.2, depth_grad_threshold=0.1, iou_rejection_threshold=None, return_dict=True)
    g = wr.BakeGraph(proj, img, mesh, cam, **kw)
    for step in range(2):
        if step == 1:
            img.copy_(torch.from_numpy(synth.view_images(6, 96, 96, seed=5)).to(dev))
        got = g.replay()
        torch.cuda.synchronize()
        want = proj(img, mesh, cam, **kw)
        assert torch.equal(got.uv_proj, want.uv_proj) and torch.equal(got.uv_proj_mask, want.uv_proj_mask)
        assert torch.equal(got.uv_depth_grad, want.uv_depth_grad) and torch.equal(got.uv_aoi_cos, want.uv_aoi_cos)
    with pytest.raises(ValueError):
        wr.BakeGraph(proj, img.cpu(), mesh, cam, **kw)
    with pytest.raises(ValueError):
        wr.BakeGraph(proj, img, mesh, cam, masks=torch.ones((6, 96, 96), device=dev), **dict(kw, iou_rejection_threshold=0.8))
