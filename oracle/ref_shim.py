"""Import the UNMODIFIED reference package on CPU, with `nvdiffrast.torch` served by the C oracle.

TEST INFRASTRUCTURE ONLY, and only usable where /root/reference exists (the build container;
the GPU box has no copy).  It is how tests/golden/ is generated (oracle/gen_golden.py) and how
the NumPy restatement in oracle/render_oracle.py is pinned against the reference's own Python
(render.py, uv.py, projection.py, camera.py, mesh.py, utils.py run as they are).

The reference imports several third-party modules that are absent from this image
(nvdiffrast, trimesh, cvcuda, imageio, matplotlib, omegaconf, ...).  None of them is touched
on the hot path, so they are replaced by empty stub modules; `nvdiffrast.torch` gets a real
CPU implementation of the four names the path calls, and `cvcuda` the two names cv_ops.py's
inpaint_cvc calls (served by the oracle's own seam fill, see _make_cvcuda).
"""
from __future__ import annotations

import importlib
import os
import sys
import types

import numpy as np
import torch

from . import shim

REFERENCE_ROOT = os.environ.get("WR_REFERENCE_ROOT", "/root/reference")


class _AnyAttr(types.ModuleType):
    """Module whose every missing attribute is a harmless placeholder class."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        cls = type(name, (), {"__init__": lambda self, *a, **k: None})
        setattr(self, name, cls)
        return cls


def _make_dr() -> types.ModuleType:
    dr = types.ModuleType("nvdiffrast.torch")

    class _Ctx:
        def __init__(self, device=None, **_):
            self.device = device

    def rasterize(ctx, pos, tri, resolution, ranges=None, grad_db=True):
        assert ranges is None, "range mode is not used by the reference"
        rast, _ = shim.rasterize(pos.detach().cpu().numpy(), tri.detach().cpu().numpy(), tuple(resolution))
        rast = torch.from_numpy(rast)
        return rast, torch.zeros_like(rast)

    def interpolate(attr, rast, tri, rast_db=None, diff_attrs=None):
        out = shim.interpolate(attr.detach().cpu().numpy(), rast.detach().cpu().numpy(), tri.detach().cpu().numpy())
        out = torch.from_numpy(out)
        return out, out.new_zeros(*out.shape[:-1], 0)

    def texture(tex, uv, uv_da=None, mip_level_bias=None, mip=None, filter_mode="auto",
                boundary_mode="wrap", max_mip_level=None):
        out = shim.texture(tex.detach().cpu().numpy(), uv.detach().cpu().numpy(), filter_mode, boundary_mode)
        return torch.from_numpy(out)

    def antialias(*a, **k):
        raise NotImplementedError("dr.antialias is outside the hot path (render.py:271, off by default)")

    dr.RasterizeCudaContext = _Ctx
    dr.RasterizeGLContext = _Ctx
    dr.rasterize = rasterize
    dr.interpolate = interpolate
    dr.texture = texture
    dr.antialias = antialias
    return dr


def _make_cvcuda() -> types.ModuleType:
    """`cvcuda` as far as the reference's cv_ops.py:1-35 touches it.  `inpaint` is served by the oracle's
    statement of the seam fill (wr_oracle_blend.c) -- NOT by CV-CUDA's algorithm, which is absent here; it lets
    the reference's own uv_padding / uv_blend control flow (uv.py:373-382, 426-461) run unmodified on CPU."""
    cv = types.ModuleType("cvcuda")

    class _T:
        def __init__(self, t):
            self.t = t

        def cuda(self):
            return self.t

    def as_tensor(x, layout=None):
        return _T(x)

    def inpaint(image, mask, radius):
        out = shim.inpaint_u8(image.t.detach().cpu().numpy(), mask.t.detach().cpu().numpy(), int(radius))
        return _T(torch.from_numpy(out))

    cv.as_tensor = as_tensor
    cv.inpaint = inpaint
    return cv


_STUBS = ["cv2", "trimesh", "imageio", "matplotlib", "matplotlib.pyplot", "matplotlib.cm",
          "matplotlib.colors", "omegaconf", "pytorch_lightning", "jaxtyping", "typeguard",
          "spandrel", "gltflib", "pymeshlab", "open3d"]


def load_reference():
    """Returns the reference's `mvadapter.utils.mesh_utils` package, running on CPU."""
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "mvadapter")):
        raise FileNotFoundError(f"reference tree not found at {REFERENCE_ROOT}")
    if "nvdiffrast.torch" not in sys.modules or not hasattr(sys.modules["nvdiffrast.torch"], "rasterize"):
        pkg = types.ModuleType("nvdiffrast")
        pkg.__path__ = []
        dr = _make_dr()
        pkg.torch = dr
        sys.modules["nvdiffrast"] = pkg
        sys.modules["nvdiffrast.torch"] = dr
    if "cvcuda" not in sys.modules or not hasattr(sys.modules["cvcuda"], "inpaint"):
        sys.modules["cvcuda"] = _make_cvcuda()
    for name in _STUBS:
        try:
            importlib.import_module(name)
        except Exception:
            mod = _AnyAttr(name)
            mod.__path__ = []
            sys.modules[name] = mod
            if "." in name:
                parent, child = name.rsplit(".", 1)
                setattr(sys.modules[parent], child, mod)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    prev = sys.dont_write_bytecode
    sys.dont_write_bytecode = True  # the reference tree is read-only
    try:
        return importlib.import_module("mvadapter.utils.mesh_utils")
    finally:
        sys.dont_write_bytecode = prev
