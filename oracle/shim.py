"""ctypes binding of the C oracle (oracle/wr_oracle.c).  TEST INFRASTRUCTURE ONLY.

Exposes numpy-level `rasterize / interpolate / texture / clip_positions` with the argument
meaning of the `nvdiffrast.torch` calls the reference makes (render.py:55,79,111;
utils.py:127-129).  Nothing under worldrenderer_b200/ may import this module.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional, Tuple

import numpy as np

from . import build as _build

_LIB: Optional[ctypes.CDLL] = None

_f32p = ctypes.POINTER(ctypes.c_float)
_i32p = ctypes.POINTER(ctypes.c_int32)
_u8p = ctypes.POINTER(ctypes.c_uint8)


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        path = _build.build()
        L = ctypes.CDLL(path)
        L.wro_rasterize.restype = ctypes.c_int
        L.wro_rasterize.argtypes = [_f32p, ctypes.c_int, ctypes.c_int, ctypes.c_int, _i32p, ctypes.c_int,
                                    ctypes.c_int, ctypes.c_int, _f32p, _i32p, ctypes.c_int]
        L.wro_interpolate.restype = ctypes.c_int
        L.wro_interpolate.argtypes = [_f32p, ctypes.c_int, ctypes.c_int, ctypes.c_int, _f32p, ctypes.c_int,
                                      ctypes.c_int, ctypes.c_int, _i32p, ctypes.c_int, _f32p, ctypes.c_int]
        L.wro_texture.restype = ctypes.c_int
        L.wro_texture.argtypes = [_f32p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, _f32p,
                                  ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                  _f32p, ctypes.c_int]
        L.wro_clip_positions.restype = ctypes.c_int
        L.wro_clip_positions.argtypes = [_f32p, ctypes.c_int, _f32p, ctypes.c_int, _f32p, ctypes.c_int]
        L.wro_num_threads.restype = ctypes.c_int
        L.wro_poisson_blend.restype = ctypes.c_int
        L.wro_poisson_blend.argtypes = [_f32p, _u8p, _f32p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                        ctypes.c_int, _f32p, ctypes.c_int]
        L.wro_inpaint.restype = ctypes.c_int
        L.wro_inpaint.argtypes = [_u8p, _u8p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, _u8p,
                                  ctypes.c_int]
        _LIB = L
    return _LIB


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.int32)


def _fp(a: np.ndarray):
    return a.ctypes.data_as(_f32p)


def _ip(a: np.ndarray):
    return a.ctypes.data_as(_i32p)


def num_threads() -> int:
    return int(lib().wro_num_threads())


def rasterize(pos, tri, resolution: Tuple[int, int], nthreads: int = 0):
    """pos [B,V,4] (instanced) or [V,4] (one view); tri [F,3]; -> rast [B,H,W,4] f32, tri_id [B,H,W] i32."""
    pos = _f32(pos)
    tri = _i32(tri).reshape(-1, 3)
    H, W = int(resolution[0]), int(resolution[1])
    if pos.ndim == 2:
        pos = pos[None]
    assert pos.ndim == 3 and pos.shape[-1] == 4
    B, V = pos.shape[0], pos.shape[1]
    rast = np.empty((B, H, W, 4), np.float32)
    ids = np.empty((B, H, W), np.int32)
    rc = lib().wro_rasterize(_fp(pos), B, V, 1, _ip(tri), tri.shape[0], H, W, _fp(rast), _ip(ids), nthreads)
    if rc != 0:
        raise RuntimeError(f"wro_rasterize failed with status {rc}")
    return rast, ids


def interpolate(attr, rast, tri, nthreads: int = 0):
    """attr [1|B,V,A] or [V,A]; rast [B,H,W,4]; tri [F,3] -> [B,H,W,A]."""
    attr = _f32(attr)
    rast = _f32(rast)
    tri = _i32(tri).reshape(-1, 3)
    if attr.ndim == 2:
        attr = attr[None]
    B, H, W, _ = rast.shape
    out = np.empty((B, H, W, attr.shape[-1]), np.float32)
    rc = lib().wro_interpolate(_fp(attr), attr.shape[0], attr.shape[1], attr.shape[2], _fp(rast), B, H, W,
                               _ip(tri), tri.shape[0], _fp(out), nthreads)
    if rc != 0:
        raise RuntimeError(f"wro_interpolate failed with status {rc}")
    return out


_FILTER = {"nearest": 0, "linear": 1, "auto": 1}
_BOUNDARY = {"wrap": 0, "clamp": 1, "zero": 2}


def texture(tex, uv, filter_mode: str = "auto", boundary_mode: str = "wrap", nthreads: int = 0):
    """tex [1|B,TH,TW,C]; uv [B,H,W,2] -> [B,H,W,C]."""
    tex = _f32(tex)
    uv = _f32(uv)
    if filter_mode not in _FILTER or boundary_mode not in _BOUNDARY:
        raise NotImplementedError(f"oracle texture: {filter_mode}/{boundary_mode}")
    B, H, W, _ = uv.shape
    out = np.empty((B, H, W, tex.shape[-1]), np.float32)
    rc = lib().wro_texture(_fp(tex), tex.shape[0], tex.shape[1], tex.shape[2], tex.shape[3], _fp(uv), B, H, W,
                           _FILTER[filter_mode], _BOUNDARY[boundary_mode], _fp(out), nthreads)
    if rc != 0:
        raise RuntimeError(f"wro_texture failed with status {rc}")
    return out


def clip_positions(v_pos, mvp, nthreads: int = 0):
    """v_pos [V,3]; mvp [B,4,4] -> [B,V,4] in the contract's fixed operation order."""
    v_pos = _f32(v_pos)
    mvp = _f32(mvp).reshape(-1, 4, 4)
    out = np.empty((mvp.shape[0], v_pos.shape[0], 4), np.float32)
    rc = lib().wro_clip_positions(_fp(v_pos), v_pos.shape[0], _fp(mvp), mvp.shape[0], _fp(out), nthreads)
    if rc != 0:
        raise RuntimeError(f"wro_clip_positions failed with status {rc}")
    return out


_GRAD_MODE = {"src": 0, "max": 1, "avg": 2}


def poisson_blend(src, mask, tgt, num_iters: int, grad_mode: str = "src", nthreads: int = 0):
    """PoissonBlendingSolver.__call__ (blend.py:214-324) with exactly `num_iters` Jacobi sweeps.
    src, tgt [H,W,C] f32; mask [H,W] bool (already thresholded) -> [H,W,C] f32."""
    src = _f32(src)
    tgt = _f32(tgt)
    m = np.ascontiguousarray(np.asarray(mask) != 0, dtype=np.uint8)
    H, W, C = tgt.shape
    assert src.shape == tgt.shape and m.shape == (H, W)
    out = np.empty_like(tgt)
    rc = lib().wro_poisson_blend(_fp(src), m.ctypes.data_as(_u8p), _fp(tgt), H, W, C, int(num_iters),
                                 _GRAD_MODE[grad_mode], _fp(out), nthreads)
    if rc != 0:
        raise RuntimeError(f"wro_poisson_blend failed with status {rc}")
    return out


def inpaint_u8(img, mask, radius: int, nthreads: int = 0):
    """The seam fill that stands in for cvcuda.inpaint (cv_ops.py:32): img [H,W,C] u8, mask [H,W] (non-zero = fill)."""
    img = np.ascontiguousarray(img, dtype=np.uint8)
    m = np.ascontiguousarray(np.asarray(mask) != 0, dtype=np.uint8)
    H, W, C = img.shape
    assert m.shape == (H, W)
    out = np.empty_like(img)
    rc = lib().wro_inpaint(img.ctypes.data_as(_u8p), m.ctypes.data_as(_u8p), H, W, C, int(radius),
                           out.ctypes.data_as(_u8p), nthreads)
    if rc != 0:
        raise RuntimeError(f"wro_inpaint failed with status {rc}")
    return out


def uv_padding(attr, inside_mask, radius: int, nthreads: int = 0):
    """uv_padding (uv.py:373-382) over inpaint_cvc's quantisation (cv_ops.py:23-35): float in, float out."""
    a = np.clip(_f32(attr), 0.0, 1.0)
    q = (a * np.float32(255.0)).astype(np.uint8)  # truncation, like tensor.to(torch.uint8)
    out = inpaint_u8(q, ~(np.asarray(inside_mask) != 0), radius, nthreads)
    return out.astype(np.float32) / np.float32(255.0)
