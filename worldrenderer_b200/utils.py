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


def tensor_to_image(data, batched: bool = False, format: str = "HWC"):
    """float [0,1] / bool / uint8 array -> PIL image(s)."""
    if isinstance(data, Image.Image):
        return data
    if isinstance(data, torch.Tensor):
        data = data.detach().cpu().numpy()
    if data.dtype in (np.float32, np.float16):
        data = (data * 255).astype(np.uint8)
    elif data.dtype == np.bool_:
        data = data.astype(np.uint8) * 255
    assert data.dtype == np.uint8
    if format == "CHW":
        if batched and data.ndim == 4:
            data = data.transpose((0, 2, 3, 1))
        elif not batched and data.ndim == 3:
            data = data.transpose((1, 2, 0))
    if batched:
        return [Image.fromarray(d) for d in data]
    return Image.fromarray(data)


def image_to_tensor(image: IMAGE_TYPE, return_type: str = "pt", device: Optional[str] = None):
    """PIL image(s) are scaled by 1/255; arrays and tensors are taken as they are (cast to float32)."""
    assert return_type in ["np", "pt"]
    single = isinstance(image, Image.Image)
    if single:
        image = [image]
    if isinstance(image, list):
        image = np.stack([np.array(im) for im in image], axis=0).astype(np.float32) / 255.0
    if isinstance(image, np.ndarray) and return_type == "pt":
        image = torch.tensor(image, device=device)
    if isinstance(image, torch.Tensor):
        image = image.to(dtype=torch.float32, device=device)
    return image[0] if single else image


def largest_factor_near_sqrt(n: int) -> int:
    root = int(math.sqrt(n))
    if root * root == n:
        return root
    for i in range(root, 0, -1):
        if n % i == 0:
            return i
    return 1


def make_image_grid(images: List[Image.Image], rows: Optional[int] = None, cols: Optional[int] = None,
                    resize: Optional[int] = None) -> Image.Image:
    if rows is None and cols is not None:
        assert len(images) % cols == 0
        rows = len(images) // cols
    elif cols is None and rows is not None:
        assert len(images) % rows == 0
        cols = len(images) // rows
    elif rows is None and cols is None:
        rows = largest_factor_near_sqrt(len(images))
        cols = len(images) // rows
    assert len(images) == rows * cols
    if resize is not None:
        images = [im.resize((resize, resize)) for im in images]
    w, h = images[0].size
    grid = Image.new("RGB", size=(cols * w, rows * h))
    for i, im in enumerate(images):
        grid.paste(im, box=(i % cols * w, i // cols * h))
    return grid


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
