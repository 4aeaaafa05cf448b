"""Small tensor / image helpers of the geometry path.

Mirrors mvadapter/utils/mesh_utils/utils.py of the reference: `tensor_to_image` (:22-44),
`image_to_tensor` (:47-63), `make_image_grid` (:91-120), `get_clip_space_position` (:127-129),
`transform_points_homo` (:132-139).  The two transforms are kept as plain torch ops for callers
that use them directly; `render()` and the bake do not call them -- the same arithmetic is fused
into the CUDA kernels (csrc/raster.cu k_snap_vertices, csrc/shade.cu k_shade).
"""
from __future__ import annotations

import math
from datetime import datetime
from typing import List, Optional, Union

import numpy as np
import torch
from PIL import Image

LIST_TYPE = Union[list, np.ndarray, torch.Tensor]
IMAGE_TYPE = Union[Image.Image, List[Image.Image], np.ndarray, torch.Tensor]
SINGLE_IMAGE_TYPE = Union[Image.Image, np.ndarray, torch.Tensor]


def _as_uint8(arr: np.ndarray) -> np.ndarray:
    """float images are scaled by 255 and truncated, masks become 0 / 255, uint8 passes through."""
    if arr.dtype in (np.float32, np.float16):
        return (arr * 255).astype(np.uint8)
    if arr.dtype == np.bool_:
        return arr.astype(np.uint8) * 255
    assert arr.dtype == np.uint8
    return arr


def tensor_to_image(data, batched: bool = False, format: str = "HWC"):
    """float [0,1] / bool / uint8 array or tensor -> PIL image, or a list of them when `batched`
    (utils.py:22-44).  `format="CHW"` moves the channel axis last first."""
    if isinstance(data, Image.Image):
        return data
    arr = _as_uint8(data.detach().cpu().numpy() if isinstance(data, torch.Tensor) else data)
    if format == "CHW" and arr.ndim == (4 if batched else 3):
        arr = np.moveaxis(arr, -3, -1)
    return [Image.fromarray(a) for a in arr] if batched else Image.fromarray(arr)


def image_to_tensor(image: IMAGE_TYPE, return_type: str = "pt", device: Optional[str] = None):
    """PIL image(s) are scaled by 1/255; arrays and tensors are taken as they are, cast to float32
    (utils.py:47-63).  A single PIL image comes back without the batch dimension."""
    assert return_type in ["np", "pt"]
    single = isinstance(image, Image.Image)
    frames = [image] if single else image
    if isinstance(frames, list):
        frames = np.stack([np.array(f) for f in frames], axis=0).astype(np.float32) / 255.0
    if return_type == "pt" and isinstance(frames, np.ndarray):
        frames = torch.tensor(frames, device=device)
    if isinstance(frames, torch.Tensor):
        frames = frames.to(dtype=torch.float32, device=device)
    return frames[0] if single else frames


def largest_factor_near_sqrt(n: int) -> int:
    """Largest divisor of n that does not exceed sqrt(n) (1 for n < 1)."""
    return max((d for d in range(1, math.isqrt(max(n, 1)) + 1) if n % d == 0), default=1)


def make_image_grid(images: List[Image.Image], rows: Optional[int] = None, cols: Optional[int] = None,
                    resize: Optional[int] = None) -> Image.Image:
    """Pastes equally sized images row by row into one RGB sheet (utils.py:91-120).  A missing dimension is derived
    from the other; with neither the sheet is made as square as the count allows."""
    count = len(images)
    if rows is None and cols is None:
        rows = largest_factor_near_sqrt(count)
    if rows is None:
        assert count % cols == 0
        rows = count // cols
    elif cols is None:
        assert count % rows == 0
        cols = count // rows
    assert count == rows * cols
    tiles = images if resize is None else [t.resize((resize, resize)) for t in images]
    tw, th = tiles[0].size
    sheet = Image.new("RGB", size=(cols * tw, rows * th))
    for k, tile in enumerate(tiles):
        r, c = divmod(k, cols)
        sheet.paste(tile, box=(c * tw, r * th))
    return sheet


def get_current_timestamp(fmt: str = "%Y%m%d%H%M%S") -> str:
    return datetime.now().strftime(fmt)


def get_clip_space_position(pos: torch.Tensor, mvp_mtx: torch.Tensor) -> torch.Tensor:
    """[V,3] world positions x [B,4,4] mvp -> [B,V,4] clip positions."""
    ones = torch.ones([pos.shape[0], 1]).to(pos)
    return torch.matmul(torch.cat([pos, ones], dim=-1), mvp_mtx.permute(0, 2, 1))


def transform_points_homo(pos: torch.Tensor, mtx: torch.Tensor) -> torch.Tensor:
    """Applies [B,4,4] to [B,...,3] points, returns xyz (no perspective divide)."""
    batch = pos.shape[0]
    lead = pos.shape[1:-1]
    flat = pos.reshape(batch, -1, 3)
    rot = torch.einsum("bij,bnj->bni", mtx[:, :3, :3], flat)
    out = rot + mtx[:, None, :3, 3]
    return out.reshape(batch, *lead, 3)
