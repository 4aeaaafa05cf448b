"""Triangle mesh container and loader of the geometry path.

Mirrors mvadapter/utils/mesh_utils/mesh.py of the reference: `TexturedMesh` (:24-185) with lazily
computed vertex normals (`_compute_vertex_normal` :85-119 -> CUDA kernels k_face_normals_scatter /
k_normalize_vertex_normals through wr_vertex_normals) and tangents (:121-167), `mesh_use_texture`
(:188-195) and `load_mesh` (:198-345: .npz fast path :212-222, centring :238-241, rescale :245-248,
axis remap :250-274, front_x_to_y :277-286, UV flip :296-298).

Differences that are deliberate: the mesh keeps cached int32 copies of its index tensors (the
reference converts int64 -> int32 on every rasterize / interpolate call, render.py:57-58,80), and
normals are computed on the GPU only -- a mesh on the CPU raises instead of silently taking a
PyTorch path.  GLB export (:348-526) is outside the hot path and not provided.
"""
from __future__ import annotations

from contextlib import contextmanager
from dataclasses import dataclass, field
from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from . import _native


def dot(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    return torch.sum(x * y, -1, keepdim=True)


@dataclass
class TexturedMesh:
    v_pos: torch.Tensor
    t_pos_idx: torch.Tensor

    # texture coordinates
    v_tex: Optional[torch.Tensor] = None
    t_tex_idx: Optional[torch.Tensor] = None

    # texture map
    texture: Optional[torch.Tensor] = None

    # vertices, faces after vertex merging
    _stitched_v_pos: Optional[torch.Tensor] = None
    _stitched_t_pos_idx: Optional[torch.Tensor] = None

    _v_nrm: Optional[torch.Tensor] = None
    _v_tang: Optional[torch.Tensor] = None

    # (name) -> (key, int32 tensor): contiguous int32 copies of the index tensors
    _i32_cache: Dict[str, tuple] = field(default_factory=dict, repr=False, compare=False)

    # ------------------------------------------------------------------ lazily derived attributes
    @property
    def v_nrm(self) -> torch.Tensor:
        if self._v_nrm is None:
            self._v_nrm = self._compute_vertex_normal()
        return self._v_nrm

    @property
    def v_tang(self) -> torch.Tensor:
        if self._v_tang is None:
            self._v_tang = self._compute_tangent()
        return self._v_tang

    def set_vertex_normal(self, v_nrm: torch.Tensor) -> None:
        assert v_nrm.shape == self.v_pos.shape
        self._v_nrm = v_nrm.to(self.v_pos)

    def set_stitched_mesh(self, v_pos: torch.Tensor, t_pos_idx: torch.Tensor) -> None:
        self._stitched_v_pos = v_pos
        self._stitched_t_pos_idx = t_pos_idx

    @property
    def stitched_v_pos(self) -> torch.Tensor:
        if self._stitched_v_pos is None:
            print("Warning: Stitched vertices not available, using original vertices!")
            return self.v_pos
        return self._stitched_v_pos

    @property
    def stitched_t_pos_idx(self) -> torch.Tensor:
        if self._stitched_t_pos_idx is None:
            print("Warning: Stitched faces not available, using original faces!")
            return self.t_pos_idx
        return self._stitched_t_pos_idx

    @property
    def uv_size(self) -> Optional[int]:
        return None if self.texture is None else self.texture.shape[0]

    # ------------------------------------------------------------------ int32 index cache
    def index_i32(self, name: str) -> torch.Tensor:
        """Contiguous int32 copy of `t_pos_idx` / `t_tex_idx` / `stitched_t_pos_idx`, cached until the
        source tensor is replaced, modified in place or moved."""
        if name == "stitched_t_pos_idx":
            src = self._stitched_t_pos_idx if self._stitched_t_pos_idx is not None else self.t_pos_idx
        else:
            src = getattr(self, name)
        key = (src.data_ptr(), src._version, tuple(src.shape), src.dtype, src.device)
        hit = self._i32_cache.get(name)
        # the entry keeps the SOURCE tensor alive and is matched by identity: a replaced index tensor whose storage
        # the caching allocator hands out again at the same address can never alias a stale copy
        if hit is not None and hit[1] is src and hit[0] == key:
            return hit[2]
        out = src.to(torch.int32).contiguous()
        self._i32_cache[name] = (key, src, out)
        return out

    # ------------------------------------------------------------------ normals / tangents
    def _compute_vertex_normal(self) -> torch.Tensor:
        """Area-weighted vertex normals of the stitched mesh (mesh.py:85-119), on the GPU."""
        if self._stitched_v_pos is None or self._stitched_t_pos_idx is None:
            print("Warning: Stitched vertices and faces not available, computing vertex normals on "
                  "original mesh, which can be erroneous!")
            v_pos, name = self.v_pos, "t_pos_idx"
        else:
            v_pos, name = self._stitched_v_pos, "stitched_t_pos_idx"
        if v_pos.device.type != "cuda":
            raise RuntimeError("TexturedMesh.v_nrm: vertex normals are computed by a CUDA kernel; move the mesh to "
                               "a CUDA device first (mesh.to('cuda')). There is no CPU path.")
        tri = self.index_i32(name)
        v = v_pos.to(torch.float32).contiguous()
        out = torch.empty_like(v)
        ctx = _shared_context(v.device)
        ctx.check(_native.lib().wr_vertex_normals(ctx.handle, _native.ptr(v), v.shape[0], _native.ptr(tri),
                                                  tri.shape[0], _native.ptr(out), ctx.stream()),
                  "wr_vertex_normals")
        if torch.is_anomaly_enabled():
            assert torch.all(torch.isfinite(out))
        return out

    def _compute_tangent(self) -> torch.Tensor:
        """Per-vertex tangents from the UV gradients of the faces (mesh.py:121-167), on the GPU
        (k_face_tangents_scatter / k_finish_vertex_tangents through wr_vertex_tangents)."""
        nrm = self.v_nrm
        if nrm.device.type != "cuda":
            raise RuntimeError("TexturedMesh.v_tang: tangents are computed by a CUDA kernel; move the mesh to a CUDA "
                               "device first (mesh.to('cuda')). There is no CPU path.")
        if self.v_tex is None or self.t_tex_idx is None:
            raise ValueError("TexturedMesh.v_tang needs UV coordinates (v_tex, t_tex_idx)")
        v = self.v_pos.to(torch.float32).contiguous()
        vt = self.v_tex.to(torch.float32).contiguous()
        tri, tri_t = self.index_i32("t_pos_idx"), self.index_i32("t_tex_idx")
        if nrm.shape != v.shape:
            raise ValueError("v_tang: v_nrm must have one row per vertex of v_pos (mesh.py:131 accumulates into "
                             "zeros_like(v_nrm) by position index)")
        if tri.shape != tri_t.shape:
            raise ValueError("t_pos_idx and t_tex_idx must list the same faces")
        nrm = nrm.to(torch.float32).contiguous()
        out = torch.empty_like(v)
        ctx = _shared_context(v.device)
        ctx.check(_native.lib().wr_vertex_tangents(ctx.handle, _native.ptr(v), v.shape[0], _native.ptr(tri),
                                                   _native.ptr(vt), vt.shape[0], _native.ptr(tri_t), tri.shape[0],
                                                   _native.ptr(nrm), _native.ptr(out), ctx.stream()),
                  "wr_vertex_tangents")
        if torch.is_anomaly_enabled():
            assert torch.all(torch.isfinite(out))
        return out

    def to(self, device: Optional[str] = None):
        moved = {}   # tensors that alias each other (an unstitched mesh shares v_pos / t_pos_idx) keep doing so
        for name in ("v_pos", "t_pos_idx", "v_tex", "t_tex_idx", "texture", "_stitched_v_pos",
                     "_stitched_t_pos_idx", "_v_nrm", "_v_tang"):
            val = getattr(self, name)
            if val is not None:
                if id(val) not in moved:
                    moved[id(val)] = val.to(device)
                setattr(self, name, moved[id(val)])
        self._i32_cache.clear()


_CONTEXTS: Dict[int, "_native.NativeContext"] = {}


def _shared_context(device: torch.device) -> "_native.NativeContext":
    """Process-wide native context per device for mesh-level kernels (normals)."""
    index = device.index if device.index is not None else torch.cuda.current_device()
    ctx = _CONTEXTS.get(index)
    if ctx is None:
        ctx = _native.NativeContext(torch.device("cuda", index))
        _CONTEXTS[index] = ctx
    return ctx


@contextmanager
def mesh_use_texture(mesh: TexturedMesh, texture: torch.Tensor):
    saved = mesh.texture
    mesh.texture = texture
    try:
        yield
    finally:
        mesh.texture = saved


_AXES = {"+x": (1, 0, 0), "+y": (0, 1, 0), "+z": (0, 0, 1), "-x": (-1, 0, 0), "-y": (0, -1, 0), "-z": (0, 0, -1)}


class _ArrayMesh:
    """Minimal stand-in for a trimesh.Trimesh when loading .npz files."""
    vertices = None
    faces = None


def load_mesh(mesh_path: str, rescale: bool = False, move_to_center: bool = False, scale: float = 0.5,
              flip_uv: bool = True, merge_vertices: bool = True, default_uv_size: Optional[int] = None,
              shape_init_mesh_up: str = "+y", shape_init_mesh_front: str = "+x", front_x_to_y: bool = False,
              device: Optional[str] = None, return_transform: bool = False):
    """Loads a mesh into the z-up frame used by the camera rig.  `.npz` files (keys `vertices`, `faces`)
    need no third-party package; any other format goes through trimesh, which must be installed."""
    if mesh_path.endswith(".npz"):
        data = np.load(mesh_path)
        mesh = _ArrayMesh()
        mesh.vertices = data["vertices"]
        mesh.faces = data["faces"]
        merge_vertices = False
    else:
        try:
            import trimesh
        except ImportError as e:  # pragma: no cover - trimesh is absent from the build image
            raise ImportError("load_mesh: formats other than .npz need the `trimesh` package") from e
        scene = trimesh.load(mesh_path, force="mesh", process=False)
        if isinstance(scene, trimesh.Trimesh):
            mesh = scene
        elif isinstance(scene, trimesh.scene.Scene):
            mesh = trimesh.Trimesh()
            for obj in scene.geometry.values():
                mesh = trimesh.util.concatenate([mesh, obj])
        else:
            raise ValueError(f"Unknown mesh type at {mesh_path}.")

    vertex_normals = getattr(mesh, "vertex_normals", None)

    transform_offset = None
    if move_to_center:
        transform_offset = mesh.vertices.mean(0)
        mesh.vertices = mesh.vertices - transform_offset

    transform_scale = None
    if rescale:
        extent = np.abs(mesh.vertices).max()
        mesh.vertices = mesh.vertices / extent * scale
        transform_scale = extent / scale

    if shape_init_mesh_up not in _AXES or shape_init_mesh_front not in _AXES:
        raise ValueError(f"shape_init_mesh_up and shape_init_mesh_front must be one of {list(_AXES)}.")
    if shape_init_mesh_up[1] == shape_init_mesh_front[1]:
        raise ValueError("shape_init_mesh_up and shape_init_mesh_front must be orthogonal.")
    z_axis = np.array(_AXES[shape_init_mesh_up])
    x_axis = np.array(_AXES[shape_init_mesh_front])
    y_axis = np.cross(z_axis, x_axis)
    mesh_to_std = np.linalg.inv(np.stack([x_axis, y_axis, z_axis], axis=0).T)
    mesh.vertices = np.dot(mesh_to_std, mesh.vertices.T).T
    if vertex_normals is not None:
        vertex_normals = np.dot(mesh_to_std, vertex_normals.T).T
    if front_x_to_y:
        old_x = mesh.vertices[:, 0].copy()
        mesh.vertices[:, 0] = mesh.vertices[:, 1]
        mesh.vertices[:, 1] = -old_x
        if vertex_normals is not None:
            old_nx = vertex_normals[:, 0].copy()
            vertex_normals[:, 0] = vertex_normals[:, 1]
            vertex_normals[:, 1] = -old_nx

    v_pos = torch.tensor(mesh.vertices, dtype=torch.float32)
    t_pos_idx = torch.tensor(mesh.faces, dtype=torch.int64)

    visual = getattr(mesh, "visual", None)
    if visual is not None and getattr(visual, "uv", None) is not None:
        v_tex = torch.tensor(visual.uv, dtype=torch.float32)
        if flip_uv:
            v_tex[:, 1] = 1.0 - v_tex[:, 1]
        t_tex_idx = t_pos_idx.clone()
        if default_uv_size is not None or getattr(visual.material, "baseColorTexture", None) is None:
            assert default_uv_size is not None
            texture = torch.zeros((default_uv_size, default_uv_size, 3), dtype=torch.float32)
        else:
            texture = torch.tensor(np.array(visual.material.baseColorTexture) / 255.0, dtype=torch.float32)[..., :3]
    else:
        v_tex = t_tex_idx = texture = None

    out = TexturedMesh(v_pos=v_pos, t_pos_idx=t_pos_idx, v_tex=v_tex, t_tex_idx=t_tex_idx, texture=texture)

    if vertex_normals is not None:
        out.set_vertex_normal(F.normalize(torch.tensor(vertex_normals, dtype=torch.float32), dim=-1))

    if merge_vertices and vertex_normals is None:
        mesh.merge_vertices(merge_tex=True)  # only when the file carries no normals
        out.set_stitched_mesh(torch.tensor(mesh.vertices, dtype=torch.float32),
                              torch.tensor(mesh.faces, dtype=torch.int64))
    else:
        out.set_stitched_mesh(out.v_pos, out.t_pos_idx)

    out.to(device)

    if return_transform:
        return out, transform_offset, transform_scale
    return out


def replace_mesh_texture_and_save(*args, **kwargs):
    """GLB export (reference mesh.py:348-526) is file I/O through trimesh / gltflib, outside the
    geometry path this package replaces."""
    raise NotImplementedError("replace_mesh_texture_and_save: GLB export is outside the scope of worldrenderer_b200")
