"""Image operators of the bake that the reference takes from CV-CUDA.

Mirrors mvadapter/utils/mesh_utils/cv_ops.py: `inpaint_cvc` (:11-35) and `batch_inpaint_cvc` (:38-51) keep
their signatures, dtype handling and quantisation (float images become uint8 by truncation of x * 255, the
result is uint8 / 255).  The fill itself is NOT cvcuda's: that operator is third-party, absent from the
reference tree and from this image, so its exact output could not be pinned.  `wr_inpaint_u8`
(csrc/blend.cu) fills every masked pixel from its nearest known pixel's neighbourhood.
`batch_erode` / `batch_dilate` (:54-93, cvcuda.morphology) have no caller anywhere in the reference and are not
provided.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _native


def inpaint_cvc(image: torch.Tensor, mask: torch.Tensor, padding_size: int,
                return_dtype: Optional[torch.dtype] = None):
    input_dtype = image.dtype
    image = image.detach()
    mask = mask.detach()
    if image.dtype != torch.uint8:
        image = (image * 255).to(torch.uint8)  # cv_ops.py:23-24
    if mask.dtype != torch.uint8:
        mask = (mask * 255).to(torch.uint8)    # cv_ops.py:25-26
    if image.ndim != 3 or mask.shape != image.shape[:2]:
        raise ValueError(f"inpaint_cvc: image [H,W,C] and mask [H,W] expected, got {tuple(image.shape)}, "
                         f"{tuple(mask.shape)}")
    image = image.contiguous()
    mask = mask.contiguous()
    H, W, C = image.shape
    out = torch.empty_like(image)
    c = _native.default_context(image.device)
    c.check(_native.lib().wr_inpaint_u8(c.handle, _native.ptr(image), _native.ptr(mask), H, W, C, int(padding_size),
                                        _native.ptr(out), c.stream()), "wr_inpaint_u8")
    if return_dtype == torch.uint8 or input_dtype == torch.uint8:
        return out
    return out.to(dtype=input_dtype) / 255.0


def batch_inpaint_cvc(images: torch.Tensor, masks: torch.Tensor, padding_size: int,
                      return_dtype: Optional[torch.dtype] = None):
    return torch.stack([inpaint_cvc(i, m, padding_size, return_dtype) for i, m in zip(images, masks)], dim=0)


def batch_erode(masks, kernel_size, return_dtype=None):
    raise NotImplementedError("batch_erode (cvcuda.morphology; no caller in the reference) is outside the bake path")


def batch_dilate(masks, kernel_size, return_dtype=None):
    raise NotImplementedError("batch_dilate (cvcuda.morphology; no caller in the reference) is outside the bake path")
