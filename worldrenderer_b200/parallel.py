"""Multi-GPU partitioning of the geometry path (one process per GPU, torch.distributed).

The reference is single-GPU; this module is the new functionality
BASELINE.json asks for:

* Rendering shards with NO data-path collective: units (mesh, view) are independent.  `shard_bounds`
  gives rank r a contiguous, balanced slice of the meshes (config D) or of the views (config E).
* The bake has ONE exchange step.  Because
      sum_v attr_v * (w_v / max(sum_v w_v, 1e-5))  ==  (sum_v attr_v * w_v) / max(sum_v w_v, 1e-5)
  (uv.py:342-344, 421-423; the outer clamp(0, 1) is a no-op for w >= 0), each rank accumulates
  [Huv, Wuv, 5] = (sum w r, sum w g, sum w b, sum w, sum valid) over its local views with
  wr_uv_unproject, the ranks all-reduce that tensor (NCCL over NVLink / NVSwitch; gloo in the CPU
  tests) and every rank finalises with wr_uv_finalize.  The valid count travels as a float sum of small
  integers, which is exact, so `uv_valid_mask_blend` is identical on every rank and to the 1-GPU result;
  the colour sums differ from the 1-GPU result only by fp32 summation order (1e-5 relative).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from .camera import Camera


def shard_bounds(n: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous balanced partition of range(n) into `world` slices (first n % world slices get one more)."""
    if world <= 0:
        raise ValueError("world size must be positive")
    base, extra = divmod(n, world)
    out, start = [], 0
    for r in range(world):
        size = base + (1 if r < extra else 0)
        out.append((start, start + size))
        start += size
    return out


def my_shard(n: int, rank: Optional[int] = None, world: Optional[int] = None) -> Tuple[int, int]:
    if rank is None or world is None:
        if dist.is_available() and dist.is_initialized():
            rank, world = dist.get_rank(), dist.get_world_size()
        else:
            rank, world = 0, 1
    return shard_bounds(n, world)[rank]


def shard_camera(cam: Camera, rank: Optional[int] = None, world: Optional[int] = None) -> Camera:
    """The views of `cam` owned by this rank (possibly an empty camera batch)."""
    n = cam.mvp_mtx.shape[0]
    lo, hi = my_shard(n, rank, world)
    return cam[lo:hi]


def all_reduce_accumulators(accum: torch.Tensor, group=None) -> torch.Tensor:
    """SUM all-reduce of the packed bake accumulators, in place.  No-op outside a process group."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(accum, op=dist.ReduceOp.SUM, group=group)
    return accum


def finalize_accumulators_reference(accum: torch.Tensor, old_attr: Optional[torch.Tensor]):
    """Plain torch statement of wr_uv_finalize (uv.py:452-455), used by the CPU tests of the decomposition."""
    den = accum[..., 3:4].clamp(min=1e-5)
    any_ = accum[..., 4] > 0.5
    va = any_[..., None].to(accum.dtype)
    old = torch.zeros_like(accum[..., :3]) if old_attr is None else old_attr
    return (accum[..., :3] / den) * va + old * (1.0 - va), any_


def render_mesh_shard(ctx, meshes: Sequence, cam: Camera, height: int, width: int, rank: Optional[int] = None,
                      world: Optional[int] = None, **render_kwargs):
    """Config D: every rank renders its contiguous slice of `meshes` (all views each).  Returns
    (first mesh index, [RenderOutput, ...]).  No communication."""
    from .render import render
    lo, hi = my_shard(len(meshes), rank, world)
    return lo, [render(ctx, meshes[i], cam, height, width, **render_kwargs) for i in range(lo, hi)]


def sharded_bake(ctx, mesh, cam_local: Camera, images_local: torch.Tensor, uv_size: int, *, view_masks_local=None,
                 aoi_cos_valid_threshold: float = 0.3, depth_grad_dilation: int = 5,
                 depth_grad_threshold: Optional[float] = 0.1, uv_exp_blend_alpha: float = 6.0,
                 uv_exp_blend_view_weight_local=None, group=None):
    """Config E: this rank holds `cam_local` / `images_local` (its share of the views, possibly none);
    the mesh is replicated.  Returns (atlas [uv,uv,3], valid_any [uv,uv] bool), identical on all ranks."""
    from .uv import fused_unproject, fused_view_maps, uv_finalize, uv_precompute
    pre = uv_precompute(ctx, mesh, uv_size, uv_size)
    n_local = cam_local.mvp_mtx.shape[0]
    if n_local > 0:
        H, W = int(images_local.shape[1]), int(images_local.shape[2])
        _, geo, att = fused_view_maps(ctx, mesh, cam_local, images_local, H, W, int(depth_grad_dilation))
        _, _, accum, _, _ = fused_unproject(
            ctx, pre, cam_local, H, W, geo, att, view_masks=view_masks_local, aoi_cos_thresh=aoi_cos_valid_threshold,
            depth_grad_thresh=depth_grad_threshold, alpha=uv_exp_blend_alpha,
            view_weight=uv_exp_blend_view_weight_local, accumulate_only=True)
    else:
        accum = torch.zeros((uv_size, uv_size, 5), dtype=torch.float32, device=ctx.device)
    all_reduce_accumulators(accum, group)
    return uv_finalize(ctx, accum, pre.uv_attr)
