"""SmartPainter: iterative view selection + inpainting of the unseen parts of a texture.

Drop-in for mvadapter/utils/mesh_utils/smart_paint.py of the reference (`SmartPainter.__init__` :38-46,
`__call__` :48-335).  It is a CALLER of the geometry path -- every round renders 108 candidate views, scores
them, renders the best one at 1024^2, hands the image to the user's `inpaint_func` and bakes the result back
with `CameraProjection` -- and that is how it is built here: the candidate views are one fused `wr_render`
(score-map texture + angle-of-incidence map in a single pass, no normal / position maps), the 2 x 108
`.sum().item()` round trips of the reference's scoring loop (:148-158) are one `wr_view_scores` call and a
single read-back, and the bake is `CameraProjection` with seam padding.  The mask morphology of the single
best view (:163-229) is a handful of torch image ops, as in the reference.

`self.last_trace` (not in the reference) keeps, per round, the view scores, the chosen view and the size of the
inpaint mask, for tests and debugging.
"""
from __future__ import annotations

import os
from itertools import product
from typing import Callable, List, Optional

import numpy as np
import torch
import torch.nn.functional as F

from . import _native
from .camera import Camera, get_camera
from .mesh import TexturedMesh, mesh_use_texture
from .projection import CameraProjection
from .render import DepthControlNetNormalization, NVDiffRastContextWrapper, render, render_geometry_raw
from .utils import make_image_grid, tensor_to_image
from .uv import uv_padding

# smart_paint.py:64-76 -- the candidate rig
_ELEVATIONS = range(-60, 61, 15)
_AZIMUTHS = range(0, 360, 30)
_DISTANCE, _FOVY = 1.2, 40
_SCORE_RES, _INPAINT_RES = 256, 1024          # :109, :163
_ATTR_EPS, _AOI_MIN, _MARGIN = 1e-3, 0.1, 0.3  # :150-155


def candidate_cameras(device) -> Camera:
    elev, azim, dist, fovy = zip(*product(_ELEVATIONS, _AZIMUTHS, [_DISTANCE], [_FOVY]))
    return get_camera(elevation_deg=list(elev), azimuth_deg=list(azim), distance=list(dist), fovy_deg=list(fovy),
                      perturb_camera_position=0.1, device=device)


def score_views(ctx: NVDiffRastContextWrapper, mesh: TexturedMesh, cams: Camera, score_map: torch.Tensor,
                size: int = _SCORE_RES) -> np.ndarray:
    """Per-view score of smart_paint.py:148-158 as float64 [B]: share of pixels that show unpainted surface
    (score < 1e-3) at a usable angle, plus the surplus of the viewing angle over the stored score elsewhere."""
    tex = score_map.to(torch.float32)[..., None].contiguous()  # one channel is enough: the reference reads attr[..., 0]
    raw = render_geometry_raw(ctx, mesh, cams, size, size, want_pos=False, want_depth=False, want_normal=False,
                              want_attr=True, want_geo=True, attr_background=1.0, texture_override=tex,
                              texture_filter_mode="nearest")
    attr, geo = raw["attr"], raw["geo"]
    B = geo.shape[0]
    count = torch.empty((B,), dtype=torch.int32, device=geo.device)
    fsum = torch.empty((B,), dtype=torch.float32, device=geo.device)
    c = ctx.ctx
    c.check(_native.lib().wr_view_scores(c.handle, _native.ptr(attr), attr.shape[-1], _native.ptr(geo), B, size, size,
                                         _ATTR_EPS, _AOI_MIN, _MARGIN, _native.ptr(count), _native.ptr(fsum),
                                         c.stream()), "wr_view_scores")
    both = torch.stack([count.to(torch.float64), fsum.to(torch.float64)]).cpu().numpy()  # the round's one read-back
    return (both[0] + both[1]) / float(size * size)


def _erode(mask: torch.Tensor, radius: int) -> torch.Tensor:   # :165-177
    k = 2 * radius + 1
    return (-F.max_pool2d(-(mask[None, None].float()), kernel_size=k, stride=1, padding=radius)).squeeze().bool()


def _dilate(mask: torch.Tensor, radius: int) -> torch.Tensor:  # :179-189
    k = 2 * radius + 1
    return F.max_pool2d(mask[None, None].float(), kernel_size=k, stride=1, padding=radius).squeeze().bool()


def _occlusion_boundary(depth: torch.Tensor, dilation: int, thresh: float) -> torch.Tensor:  # :206-231
    gx = torch.tensor([[1, 0, -1], [2, 0, -2], [1, 0, -1]], device=depth.device, dtype=torch.float32).view(1, 1, 3, 3)
    gy = torch.tensor([[1, 2, 1], [0, 0, 0], [-1, -2, -1]], device=depth.device, dtype=torch.float32).view(1, 1, 3, 3)
    d = depth[None, None]
    grad = (F.conv2d(d, gx, padding=1) ** 2 + F.conv2d(d, gy, padding=1) ** 2).sqrt()
    edge = (grad > thresh)[0][0]
    return _dilate(edge, dilation) if dilation > 0 else edge


class SmartPainter:
    def __init__(self, device: str, context_type: str = "gl"):
        self.device = device
        self.cam_proj = CameraProjection(pb_backend="torch-cuda", bg_remover=None, device=device,
                                         context_type=context_type)
        self.ctx = NVDiffRastContextWrapper(device=self.device, context_type=context_type)
        self.last_trace: List[dict] = []

    def _best_view_inputs(self, mesh: TexturedMesh, cam: Camera, score_map: torch.Tensor, texture: torch.Tensor):
        """The three things the inpainting step needs from the chosen view (:233-271): where to inpaint, the
        current colours, and nothing else -- score + angle + depth come from one fused render, colours from a
        second one with linear filtering."""
        size = _INPAINT_RES
        tex = score_map.to(torch.float32)[..., None].contiguous()
        raw = render_geometry_raw(self.ctx, mesh, cam, size, size, want_pos=False, want_depth=True, want_normal=False,
                                  want_attr=True, want_geo=True, attr_background=1.0, texture_override=tex,
                                  texture_filter_mode="nearest",
                                  depth_normalization_strategy=DepthControlNetNormalization())  # render()'s default (:248)
        score = raw["attr"][0, :, :, 0]
        aoi = raw["geo"][0, :, :, 3]
        mask = (score < _ATTR_EPS) | (aoi - score > _MARGIN)                 # :245-247
        occ = _occlusion_boundary(raw["depth"][0], dilation=0, thresh=0.1)   # :248-250
        mask = _dilate(_erode(mask, 3), 5) & ~occ                            # :251-254
        with mesh_use_texture(mesh, texture):
            image = render(self.ctx, mesh, cam, height=size, width=size, texture_filter_mode="linear").attr[0]
        return mask, image, occ

    def __call__(self, mod_name: str, mesh: TexturedMesh, inpaint_func: Callable, uv_texture: torch.Tensor,
                 uv_inpaint_mask: torch.Tensor, max_view_score_thresh: float = 0.02, min_rounds: int = 3,
                 max_rounds: int = 8, uv_padding_end: bool = True, debug_dir: Optional[str] = None,
                 debug_visualize_details: bool = False):
        cams = candidate_cameras(self.device)
        texture = uv_texture.clone()
        painted = ~uv_inpaint_mask.clone()
        score_map = torch.zeros_like(painted, dtype=torch.float32)
        score_map[painted] = 1.0
        self.last_trace = []

        best_score, rnd = 1.0, 0
        while rnd < min_rounds or (best_score > max_view_score_thresh and rnd < max_rounds):   # :96-98
            scores = score_views(self.ctx, mesh, cams, score_map)
            best_score = float(np.max(scores))
            best = int(np.argmax(scores))
            cam = cams[best:best + 1]

            mask, image, occ = self._best_view_inputs(mesh, cam, score_map, texture)
            result = inpaint_func(image.permute(2, 0, 1)[None], mask.float()[None, None])[0].permute(1, 2, 0)  # :277-281
            if debug_dir is not None:
                if debug_visualize_details:
                    tensor_to_image(occ).save(os.path.join(debug_dir, f"{mod_name}_occ_boundary_{rnd:02d}.jpg"))
                make_image_grid([tensor_to_image(image), tensor_to_image(mask), tensor_to_image(result)], rows=1).save(
                    os.path.join(debug_dir, f"{mod_name}_inpaint_result_{rnd:02d}.jpg"))

            with mesh_use_texture(mesh, texture):                                                       # :293-309
                proj = self.cam_proj(result[None], mesh, cam, masks=mask[None].float(), from_scratch=False,
                                     poisson_blending=False, depth_grad_dilation=3, uv_exp_blend_alpha=3,
                                     aoi_cos_valid_threshold=0.1, uv_size=mesh.uv_size, uv_padding=True,
                                     iou_rejection_threshold=None, return_dict=True)
            texture = proj.uv_proj
            new_valid = proj.uv_proj_mask
            painted = new_valid | painted
            gained = torch.where(new_valid, proj.uv_aoi_cos[0], torch.zeros_like(proj.uv_aoi_cos[0]))   # :322-327
            if debug_dir is not None and debug_visualize_details:
                tensor_to_image(new_valid).save(os.path.join(debug_dir, f"{mod_name}_uv_inpaint_mask_{rnd:02d}.jpg"))
                make_image_grid([tensor_to_image(score_map), tensor_to_image(gained),
                                 tensor_to_image(torch.max(score_map, gained))]).save(
                    os.path.join(debug_dir, f"{mod_name}_score_map_{rnd:02d}.jpg"))
            score_map = torch.max(score_map, gained)
            self.last_trace.append({"view_score": scores, "best_view": best, "inpaint_pixels": int(mask.sum()),
                                    "new_texels": int(new_valid.sum())})
            rnd += 1

        if uv_padding_end:
            texture = uv_padding(texture, painted, 3)
        return texture, painted

