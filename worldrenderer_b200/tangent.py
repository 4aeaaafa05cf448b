"""Tangent-space normal bake helper.

The reference's texture pipeline bakes the "normal" modality into the UV TANGENT space of the mesh: inline in
`mvadapter/test/utils/pipeline_texture.py:344-398` it renders normal + tangent maps (`render(..., render_tangent=True)`),
reads every view's normal image in the geometry tangent frame of its camera, and re-expresses the resulting world
normal in the rendered (tangent, bitangent, normal) frame before handing the images to `CameraProjection`.  Here that
block is one CUDA kernel (`k_tangent_space_normals`, csrc/shade.cu) behind `wr_tangent_space_normals`.
"""
from __future__ import annotations

from typing import Optional, Sequence, Union

import torch

from . import _native
from .render import RenderOutput

# geometry tangent axis of the six canonical views (pipeline_texture.py:370-381)
CANONICAL_VIEW_TANGENTS = ((1.0, 0.0, 0.0), (0.0, 1.0, 0.0), (-1.0, 0.0, 0.0), (0.0, -1.0, 0.0), (-1.0, 0.0, 0.0),
                           (-1.0, 0.0, 0.0))


def view_normals_to_tangent_space(normal_images: torch.Tensor, render_out: RenderOutput,
                                  view_tangents: Optional[Union[torch.Tensor, Sequence[Sequence[float]]]] = None
                                  ) -> torch.Tensor:
    """[B,H,W,3] normal images in [0,1] (geometry tangent space of each view) -> [B,H,W,3] colours of the same normals
    in the mesh's UV tangent space.  `render_out` holds the `normal` and `tangent` maps of the same views
    (`render(ctx, mesh, cameras, H, W, render_normal=True, render_tangent=True)`); `view_tangents` [B,3] defaults to
    the canonical six-view table of the reference."""
    if render_out.normal is None or render_out.tangent is None:
        raise ValueError("view_normals_to_tangent_space needs render(..., render_normal=True, render_tangent=True)")
    vN, vT = render_out.normal, render_out.tangent
    if vN.device.type != "cuda":
        raise RuntimeError("view_normals_to_tangent_space runs on CUDA tensors only; there is no CPU path")
    B, H, W, _ = vN.shape
    if view_tangents is None:
        if B != len(CANONICAL_VIEW_TANGENTS):
            raise ValueError("view_tangents must be given for anything but the canonical six views")
        view_tangents = CANONICAL_VIEW_TANGENTS
    axis = torch.as_tensor(view_tangents, dtype=torch.float32, device=vN.device).reshape(B, 3).contiguous()
    img = normal_images.to(device=vN.device, dtype=torch.float32).contiguous()
    if tuple(img.shape) != (B, H, W, 3) or tuple(vT.shape) != (B, H, W, 3):
        raise ValueError(f"normal images {tuple(img.shape)} / tangent map {tuple(vT.shape)} do not match the normal map "
                         f"{tuple(vN.shape)}")
    n, t = vN.to(torch.float32).contiguous(), vT.to(torch.float32).contiguous()
    out = torch.empty_like(n)
    ctx = _native.default_context(vN.device)
    ctx.check(_native.lib().wr_tangent_space_normals(ctx.handle, _native.ptr(n), _native.ptr(t), _native.ptr(img),
                                                     _native.ptr(axis), B, H, W, _native.ptr(out), ctx.stream()),
              "wr_tangent_space_normals")
    return out
