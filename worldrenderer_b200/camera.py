"""Camera construction for the geometry path (host side, tiny tensors, plain torch).

Same API and conventions as mvadapter/utils/mesh_utils/camera.py of the reference:
z-up world, cameras looking at the origin (`get_c2w` :23-65), GL-style projection matrices with the
y axis negated so that image row 0 is the top (`get_projection_matrix` :68-87,
`get_orthogonal_projection_matrix` :90-110), `Camera` (:113-149), `get_camera` (:152-191),
`get_orthogonal_camera` (:194-223).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch
import torch.nn.functional as F

from .utils import LIST_TYPE


def list_to_pt(x: LIST_TYPE, dtype: Optional[torch.dtype] = None, device: Optional[str] = None) -> torch.Tensor:
    """Lists / arrays become tensors on `device`; tensors only change dtype (they stay where they are)."""
    return x.to(dtype=dtype) if isinstance(x, torch.Tensor) else torch.tensor(x, dtype=dtype, device=device)


def _radians(deg: LIST_TYPE, device) -> torch.Tensor:
    return list_to_pt(deg, dtype=torch.float32, device=device) * math.pi / 180


def _look_at_origin(eye: torch.Tensor) -> torch.Tensor:
    """[B,3] camera positions -> [B,4,4] camera-to-world, columns (right, up, -forward, eye), world up = +z."""
    n = eye.shape[0]
    z_up = eye.new_tensor([0, 0, 1])[None, :].repeat(n, 1)
    fwd = F.normalize(torch.zeros_like(eye) - eye, dim=-1)
    side = F.normalize(torch.cross(fwd, z_up, dim=-1), dim=-1)
    upv = F.normalize(torch.cross(side, fwd, dim=-1), dim=-1)
    m = eye.new_zeros(n, 4, 4)
    m[:, :3, 0], m[:, :3, 1], m[:, :3, 2], m[:, :3, 3] = side, upv, -fwd, eye
    m[:, 3, 3] = 1.0
    return m


def get_c2w(elevation_deg: LIST_TYPE, distance: LIST_TYPE, azimuth_deg: Optional[LIST_TYPE],
            num_views: Optional[int] = 1, device: Optional[str] = None) -> torch.Tensor:
    """Camera-to-world matrices [B,4,4] of cameras on a sphere around the origin (camera.py:23-65).  Without
    azimuths, `num_views` cameras are spread evenly over the full turn."""
    if azimuth_deg is None:
        assert num_views is not None, "num_views must be provided if azimuth_deg is None."
        azimuth_deg = torch.linspace(0, 360, num_views + 1, dtype=torch.float32, device=device)[:-1]
    az, el = _radians(azimuth_deg, device), _radians(elevation_deg, device)
    rad = list_to_pt(distance, dtype=torch.float32, device=device)
    ring = rad * torch.cos(el)
    return _look_at_origin(torch.stack([ring * torch.cos(az), ring * torch.sin(az), rad * torch.sin(el)], dim=-1))


def _diag_projection(sx, sy, sz, tx, ty, tz, perspective: bool, n: int, device) -> torch.Tensor:
    """[n,4,4] matrices with the given diagonal scales and last-column offsets; the bottom row is (0,0,-1,0) for a
    perspective and (0,0,0,1) for an orthographic projection."""
    m = torch.zeros(n, 4, 4, dtype=torch.float32, device=device)
    for i, (scale, shift) in enumerate(((sx, tx), (sy, ty), (sz, tz))):
        m[:, i, i] = scale
        m[:, i, 3] = shift
    m[:, 3, 2 if perspective else 3] = -1 if perspective else 1
    return m


def get_projection_matrix(fovy_deg: LIST_TYPE, aspect_wh: float = 1.0, near: float = 0.1, far: float = 100.0,
                          device: Optional[str] = None) -> torch.Tensor:
    """GL perspective projection with y negated (image row 0 is the top), camera.py:68-87."""
    half = torch.tan(_radians(fovy_deg, device) / 2)
    depth = far - near
    return _diag_projection(1 / (aspect_wh * half), -1 / half, -(far + near) / depth, 0.0, 0.0, -2 * far * near / depth,
                            True, half.shape[0], device)


def get_orthogonal_projection_matrix(batch_size: int, left: float, right: float, bottom: float, top: float,
                                     near: float = 0.1, far: float = 100.0,
                                     device: Optional[str] = None) -> torch.Tensor:
    """GL orthographic projection with y negated, camera.py:90-110."""
    w, h, d = right - left, top - bottom, far - near
    return _diag_projection(2 / w, -2 / h, -2 / d, -(right + left) / w, -(top + bottom) / h, -(far + near) / d, False,
                            batch_size, device)


_FIELDS = ("c2w", "w2c", "proj_mtx", "mvp_mtx", "cam_pos")


@dataclass
class Camera:
    c2w: Optional[torch.Tensor]
    w2c: torch.Tensor
    proj_mtx: torch.Tensor
    mvp_mtx: torch.Tensor
    cam_pos: Optional[torch.Tensor]

    def __getitem__(self, index):
        """A sub-batch of views.  An int keeps the batch dimension: cam[i] is a 1-view camera (camera.py:121-137)."""
        if isinstance(index, int):
            index = slice(index, index + 1)
        elif not isinstance(index, (slice, list)):
            raise NotImplementedError
        picked = {name: getattr(self, name) for name in _FIELDS}
        # (the matrices of a strided sub-batch are made contiguous once, here, not by every call that reads them)
        return Camera(**{name: None if t is None else (t[index] if name == "cam_pos" else t[index].contiguous())
                         for name, t in picked.items()})

    def to(self, device: Optional[str] = None):
        for name in _FIELDS:
            t = getattr(self, name)
            if t is not None:
                setattr(self, name, t.to(device))

    def __len__(self):
        return self.c2w.shape[0]


def _assemble(c2w: Optional[torch.Tensor], w2c: Optional[torch.Tensor], proj: torch.Tensor) -> Camera:
    if w2c is None:
        # torch.linalg.inv hands back column-major matrices (strides (16, 1, 4)): the same values made row-major once
        # here, instead of a copy kernel in front of every render() / bake that reads them
        w2c = torch.linalg.inv(c2w).contiguous()
    return Camera(c2w=c2w, w2c=w2c, proj_mtx=proj, mvp_mtx=proj @ w2c, cam_pos=None if c2w is None else c2w[:, :3, 3])


def get_camera(elevation_deg: Optional[LIST_TYPE] = None, distance: Optional[LIST_TYPE] = None,
               fovy_deg: Optional[LIST_TYPE] = None, azimuth_deg: Optional[LIST_TYPE] = None,
               num_views: Optional[int] = 1, c2w: Optional[torch.Tensor] = None, w2c: Optional[torch.Tensor] = None,
               proj_mtx: Optional[torch.Tensor] = None, aspect_wh: float = 1.0, near: float = 0.1,
               far: float = 100.0, perturb_camera_position: Optional[float] = None,
               device: Optional[str] = None) -> Camera:
    """Perspective camera batch from spherical parameters, or from a given c2w / w2c / projection
    (camera.py:152-191).  With a given w2c there is no c2w and no camera position."""
    if w2c is not None:
        c2w = None
    elif c2w is None:
        c2w = get_c2w(elevation_deg, distance, azimuth_deg, num_views, device)
        if perturb_camera_position is not None:
            # The reference (:170-178) draws the noise and then never uses the perturbed position;
            # only the RNG side effect is observable, so only that is kept.
            torch.randn_like(c2w[:, :3, 3])
    if proj_mtx is None:
        proj_mtx = get_projection_matrix(fovy_deg, aspect_wh=aspect_wh, near=near, far=far, device=device)
    return _assemble(c2w, w2c, proj_mtx)


def get_orthogonal_camera(elevation_deg: LIST_TYPE, distance: LIST_TYPE, left: float, right: float, bottom: float,
                          top: float, azimuth_deg: Optional[LIST_TYPE] = None, num_views: Optional[int] = 1,
                          near: float = 0.1, far: float = 100.0, device: Optional[str] = None) -> Camera:
    """Orthographic rig on the same sphere (camera.py:194-223)."""
    c2w = get_c2w(elevation_deg, distance, azimuth_deg, num_views, device)
    proj = get_orthogonal_projection_matrix(c2w.shape[0], left, right, bottom, top, near=near, far=far, device=device)
    return _assemble(c2w, None, proj)
