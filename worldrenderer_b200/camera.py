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
    if isinstance(x, (list, np.ndarray)):
        return torch.tensor(x, dtype=dtype, device=device)
    return x.to(dtype=dtype)


def get_c2w(elevation_deg: LIST_TYPE, distance: LIST_TYPE, azimuth_deg: Optional[LIST_TYPE],
            num_views: Optional[int] = 1, device: Optional[str] = None) -> torch.Tensor:
    """Camera-to-world matrices [B,4,4] of cameras on a sphere around the origin, world up = +z."""
    if azimuth_deg is None:
        assert num_views is not None, "num_views must be provided if azimuth_deg is None."
        azimuth_deg = torch.linspace(0, 360, num_views + 1, dtype=torch.float32, device=device)[:-1]
    else:
        num_views = len(azimuth_deg)
    azimuth = list_to_pt(azimuth_deg, dtype=torch.float32, device=device) * math.pi / 180
    elevation = list_to_pt(elevation_deg, dtype=torch.float32, device=device) * math.pi / 180
    dist = list_to_pt(distance, dtype=torch.float32, device=device)
    eye = torch.stack([dist * torch.cos(elevation) * torch.cos(azimuth),
                       dist * torch.cos(elevation) * torch.sin(azimuth),
                       dist * torch.sin(elevation)], dim=-1)
    world_up = torch.tensor([0, 0, 1], dtype=torch.float32, device=device)[None, :].repeat(num_views, 1)
    forward = F.normalize(torch.zeros_like(eye) - eye, dim=-1)
    right = F.normalize(torch.cross(forward, world_up, dim=-1), dim=-1)
    up = F.normalize(torch.cross(right, forward, dim=-1), dim=-1)
    top = torch.cat([torch.stack([right, up, -forward], dim=-1), eye[:, :, None]], dim=-1)  # [B,3,4]
    c2w = torch.cat([top, torch.zeros_like(top[:, :1])], dim=1)
    c2w[:, 3, 3] = 1.0
    return c2w


def get_projection_matrix(fovy_deg: LIST_TYPE, aspect_wh: float = 1.0, near: float = 0.1, far: float = 100.0,
                          device: Optional[str] = None) -> torch.Tensor:
    fovy = list_to_pt(fovy_deg, dtype=torch.float32, device=device) * math.pi / 180
    t = torch.tan(fovy / 2)
    proj = torch.zeros(fovy.shape[0], 4, 4, dtype=torch.float32, device=device)
    proj[:, 0, 0] = 1 / (aspect_wh * t)
    proj[:, 1, 1] = -1 / t  # y flipped: row 0 of the image is the top
    proj[:, 2, 2] = -(far + near) / (far - near)
    proj[:, 2, 3] = -2 * far * near / (far - near)
    proj[:, 3, 2] = -1
    return proj


def get_orthogonal_projection_matrix(batch_size: int, left: float, right: float, bottom: float, top: float,
                                     near: float = 0.1, far: float = 100.0,
                                     device: Optional[str] = None) -> torch.Tensor:
    proj = torch.zeros(batch_size, 4, 4, dtype=torch.float32, device=device)
    proj[:, 0, 0] = 2 / (right - left)
    proj[:, 1, 1] = -2 / (top - bottom)  # y flipped
    proj[:, 2, 2] = -2 / (far - near)
    proj[:, 0, 3] = -(right + left) / (right - left)
    proj[:, 1, 3] = -(top + bottom) / (top - bottom)
    proj[:, 2, 3] = -(far + near) / (far - near)
    proj[:, 3, 3] = 1
    return proj


@dataclass
class Camera:
    c2w: Optional[torch.Tensor]
    w2c: torch.Tensor
    proj_mtx: torch.Tensor
    mvp_mtx: torch.Tensor
    cam_pos: Optional[torch.Tensor]

    def __getitem__(self, index):
        # an int keeps the batch dimension (cam[i] is a 1-view camera), like the reference (:121-137)
        if isinstance(index, int):
            sel = slice(index, index + 1)
        elif isinstance(index, (slice, list)):
            sel = index
        else:
            raise NotImplementedError
        return Camera(
            c2w=None if self.c2w is None else self.c2w[sel],
            w2c=self.w2c[sel],
            proj_mtx=self.proj_mtx[sel],
            mvp_mtx=self.mvp_mtx[sel],
            cam_pos=None if self.cam_pos is None else self.cam_pos[sel],
        )

    def to(self, device: Optional[str] = None):
        if self.c2w is not None:
            self.c2w = self.c2w.to(device)
        self.w2c = self.w2c.to(device)
        self.proj_mtx = self.proj_mtx.to(device)
        self.mvp_mtx = self.mvp_mtx.to(device)
        if self.cam_pos is not None:
            self.cam_pos = self.cam_pos.to(device)

    def __len__(self):
        return self.c2w.shape[0]


def get_camera(elevation_deg: Optional[LIST_TYPE] = None, distance: Optional[LIST_TYPE] = None,
               fovy_deg: Optional[LIST_TYPE] = None, azimuth_deg: Optional[LIST_TYPE] = None,
               num_views: Optional[int] = 1, c2w: Optional[torch.Tensor] = None, w2c: Optional[torch.Tensor] = None,
               proj_mtx: Optional[torch.Tensor] = None, aspect_wh: float = 1.0, near: float = 0.1,
               far: float = 100.0, perturb_camera_position: Optional[float] = None,
               device: Optional[str] = None) -> Camera:
    """Perspective camera batch from spherical parameters, or from given c2w / w2c / projection."""
    if w2c is None:
        if c2w is None:
            c2w = get_c2w(elevation_deg, distance, azimuth_deg, num_views, device)
            if perturb_camera_position is not None:
                # The reference (:170-178) draws the noise and then never uses the perturbed position;
                # only the RNG side effect is observable, so only that is kept.
                torch.randn_like(c2w[:, :3, 3])
        cam_pos = c2w[:, :3, 3]
        w2c = torch.linalg.inv(c2w)
    else:
        cam_pos = None
        c2w = None
    if proj_mtx is None:
        proj_mtx = get_projection_matrix(fovy_deg, aspect_wh=aspect_wh, near=near, far=far, device=device)
    return Camera(c2w=c2w, w2c=w2c, proj_mtx=proj_mtx, mvp_mtx=proj_mtx @ w2c, cam_pos=cam_pos)


def get_orthogonal_camera(elevation_deg: LIST_TYPE, distance: LIST_TYPE, left: float, right: float, bottom: float,
                          top: float, azimuth_deg: Optional[LIST_TYPE] = None, num_views: Optional[int] = 1,
                          near: float = 0.1, far: float = 100.0, device: Optional[str] = None) -> Camera:
    c2w = get_c2w(elevation_deg, distance, azimuth_deg, num_views, device)
    w2c = torch.linalg.inv(c2w)
    proj_mtx = get_orthogonal_projection_matrix(batch_size=c2w.shape[0], left=left, right=right, bottom=bottom,
                                                top=top, near=near, far=far, device=device)
    return Camera(c2w=c2w, w2c=w2c, proj_mtx=proj_mtx, mvp_mtx=proj_mtx @ w2c, cam_pos=c2w[:, :3, 3])
