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


@pytest.mark.parametrize("view_lanes", [2, 3, 4])
def test_view_lanes_equal_eager(wr_ctx, view_lanes):
    """One job, its views split over concurrent lanes that render into slices of the job's outputs."""
    dev = wr_ctx.device
    cam = cases.canonical_cameras(device=dev)
    meshes = [_mesh(*cases.terrain_mesh(96, 64, seed=1), dev), _mesh(*cases.icosphere_mesh(8), dev)]
    g = wr.RenderGraph(wr_ctx, [(m, cam) for m in meshes], 64, 64, view_lanes=view_lanes, render_attr=False)
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
