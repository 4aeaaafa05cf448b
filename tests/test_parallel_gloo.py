"""Multi-rank host logic on CPU: world_size 2 over gloo (127.0.0.1).  The per-rank bake accumulators are
produced here by the ORACLE (this is a test of the decomposition and the collective plumbing, not of the
kernels): shard the views, accumulate (sum w rgb, sum w, sum valid) per rank, all-reduce, finalise, and
compare with the oracle's single-process bake."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from worldrenderer_b200 import parallel

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_chunk_bounds_partition_the_atlas_on_block_boundaries():
    """Texel ranges of the chunked multi-GPU bake (sharded_bake(chunks=...)): a partition of the atlas, inner
    boundaries on the exchange kernels' 1024-texel blocks, never an empty range -- every rank derives the same list
    from (ntex, chunks, shape) alone, which is what makes the per-chunk device barriers pair up."""
    for ntex in [1 << 16, 1000 * 1000, 1024 * 1024, 4096 * 4096, 1500 * 1500 + 4, 5000]:
        for n in [1, 2, 3, 4, 8, 12, 64]:
            for shape in ("equal", "falling"):
                b = parallel.chunk_bounds(ntex, n, shape)
                assert 1 <= len(b) <= n and b[0][0] == 0 and b[-1][1] == ntex
                assert all(lo < hi for lo, hi in b)
                assert all(b[i][1] == b[i + 1][0] and b[i][1] % 1024 == 0 for i in range(len(b) - 1))
                sizes = [hi - lo for lo, hi in b]
                if shape == "equal" and len(b) == n:
                    assert max(sizes[:-1], default=0) - min(sizes[:-1], default=0) <= 1024
                if shape == "falling":
                    assert all(sizes[i] + 1024 >= sizes[i + 1] for i in range(len(sizes) - 2))
    assert parallel.chunk_bounds(4096 * 4096, 8) == [(k << 21, (k + 1) << 21) for k in range(8)]
    assert parallel.chunk_bounds(10 * 1024, 4, "falling") == [(0, 4096), (4096, 7168), (7168, 9216), (9216, 10240)]
    with pytest.raises(ValueError):
        parallel.chunk_bounds(1 << 20, 4, "rising")


def test_shard_bounds_partition_everything_once():
    for n in [0, 1, 5, 6, 7, 32, 64]:
        for world in [1, 2, 3, 4, 8]:
            b = parallel.shard_bounds(n, world)
            assert len(b) == world and b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
    assert parallel.my_shard(10, 1, 4) == (3, 6)
    assert parallel.my_shard(10) == (0, 10)  # no process group: one shard
    # both kinds of shard_slice partition range(n); interleaved shards differ in size by at most one as well
    for n in [0, 1, 7, 32]:
        for world in [1, 3, 8]:
            for il in (False, True):
                parts = [list(range(n))[parallel.shard_slice(n, r, world, interleave=il)] for r in range(world)]
                assert sorted(sum(parts, [])) == list(range(n))
                assert max(map(len, parts)) - min(map(len, parts)) <= 1
    assert list(range(32))[parallel.shard_slice(32, 3, 8, interleave=True)] == [3, 11, 19, 27]
    import worldrenderer_b200 as wr
    from worldrenderer_b200 import synth
    cam = wr.get_orthogonal_camera(**synth.CANONICAL_RIG)
    assert torch.equal(parallel.shard_camera(cam, 1, 2, interleave=True).mvp_mtx, cam.mvp_mtx[1::2])
    assert torch.equal(parallel.shard_camera(cam, 1, 2).mvp_mtx, cam.mvp_mtx[3:6])


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _local_accumulators(g, lo, hi):
    """Oracle accumulators of views [lo, hi) -- same algebra as wr_uv_unproject's accumulate mode."""
    from oracle import render_oracle as ro
    pre = ro.uv_precompute(g["v_pos"], g["t_pos_idx"], g["v_tex"], g["t_tex_idx"], 64, 64)
    acc = np.zeros((64, 64, 5), np.float32)
    if hi > lo:
        geo = ro.uv_render_geometry(g["v_pos"], g["t_pos_idx"], g["v_nrm"], g["t_pos_idx"], g["mvp"][lo:hi],
                                    g["w2c"][lo:hi], 48, 48, pre, True, 5)
        attr = ro.uv_render_attr(g["images"][lo:hi], geo)
        valid = ro.uv_validity(pre, geo, attr, aoi_cos_thresh=0.2, depth_grad_thresh=0.1)
        w = np.power(geo["uv_aoi_cos"] * valid.astype(np.float32), np.float32(3.0)).astype(np.float32)
        acc[..., :3] = (attr["uv_attr_proj"] * w[..., None]).sum(0)
        acc[..., 3] = w.sum(0)
        acc[..., 4] = valid.sum(0)
    return acc, pre


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = dict(np.load(os.path.join(GOLDEN, "bake_sphere.npz")))
        lo, hi = parallel.my_shard(6)
        assert (lo, hi) == parallel.shard_bounds(6, world)[rank]
        acc, pre = _local_accumulators(g, lo, hi)
        accum = torch.from_numpy(acc)
        parallel.all_reduce_accumulators(accum)
        atlas, any_ = parallel.finalize_accumulators_reference(accum, torch.from_numpy(g["texture"]))
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), atlas=atlas.numpy(), any=any_.numpy())
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_bake_decomposition_over_gloo(tmp_path, world):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    from oracle import render_oracle as ro
    g = dict(np.load(os.path.join(GOLDEN, "bake_sphere.npz")))
    ref = ro.camera_projection(g["images"], g["v_pos"], g["t_pos_idx"], g["v_nrm"], g["t_pos_idx"], g["v_tex"],
                               g["t_tex_idx"], g["texture"], g["mvp"], g["w2c"], 64, iou_rejection_threshold=None,
                               aoi_cos_valid_threshold=0.2, depth_grad_threshold=0.1, uv_exp_blend_alpha=3.0,
                               depth_grad_dilation=5)
    outs = [dict(np.load(os.path.join(str(tmp_path), f"rank{r}.npz"))) for r in range(world)]
    for o in outs:
        np.testing.assert_array_equal(o["any"], ref["uv_proj_mask"])       # integer-exact across ranks
        np.testing.assert_allclose(o["atlas"], ref["uv_proj"], rtol=1e-5, atol=1e-6)
        np.testing.assert_array_equal(o["atlas"], outs[0]["atlas"])         # every rank holds the same atlas
    assert ref["uv_proj_mask"].sum() > 300


def test_all_reduce_is_a_noop_without_a_group():
    t = torch.arange(10, dtype=torch.float32).reshape(1, 2, 5).clone()
    assert torch.equal(parallel.all_reduce_accumulators(t.clone()), t)
