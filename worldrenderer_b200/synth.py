"""Deterministic synthetic inputs for the benchmark configurations.

Meshes are produced as NumPy arrays (float64 vertices like a file loader would give,
int64 faces) and can be written as the `.npz` files `load_mesh` accepts
(reference: mvadapter/utils/mesh_utils/mesh.py:212-222, keys `vertices`, `faces`).
"""
from __future__ import annotations

from typing import Tuple

import numpy as np

# ---------------------------------------------------------------------------------------------
# geodesic icosphere: 20 * f^2 faces, 10 * f^2 + 2 vertices (f = 50 -> 50 000 / 25 002)
# ---------------------------------------------------------------------------------------------

def _icosahedron() -> Tuple[np.ndarray, np.ndarray]:
    g = (1.0 + 5.0 ** 0.5) / 2.0
    v = np.array(
        [[-1, g, 0], [1, g, 0], [-1, -g, 0], [1, -g, 0],
         [0, -1, g], [0, 1, g], [0, -1, -g], [0, 1, -g],
         [g, 0, -1], [g, 0, 1], [-g, 0, -1], [-g, 0, 1]], dtype=np.float64)
    f = np.array(
        [[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11],
         [1, 5, 9], [5, 11, 4], [11, 10, 2], [10, 7, 6], [7, 1, 8],
         [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9],
         [4, 9, 5], [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]], dtype=np.int64)
    return v / np.linalg.norm(v[0]), f


def icosphere(frequency: int = 50, radius: float = 0.5) -> Tuple[np.ndarray, np.ndarray]:
    """Class-I geodesic sphere. Returns (vertices [V,3] f64, faces [F,3] i64), outward winding."""
    f = int(frequency)
    base_v, base_f = _icosahedron()
    # lattice points (i, j) with i + j <= f inside one face, row-major by i
    ii, jj = np.meshgrid(np.arange(f + 1), np.arange(f + 1), indexing="ij")
    keep = (ii + jj) <= f
    ii, jj = ii[keep], jj[keep]
    kk = f - ii - jj
    lut = -np.ones((f + 1, f + 1), np.int64)
    lut[ii, jj] = np.arange(ii.size)
    # small triangles of the lattice: "up" (i,j),(i+1,j),(i,j+1) and "down" (i+1,j),(i+1,j+1),(i,j+1)
    ui, uj = np.nonzero(((np.add.outer(np.arange(f), np.arange(f))) < f))
    up = np.stack([lut[ui, uj], lut[ui + 1, uj], lut[ui, uj + 1]], -1)
    di, dj = np.nonzero(((np.add.outer(np.arange(f), np.arange(f))) < f - 1))
    down = np.stack([lut[di + 1, dj], lut[di + 1, dj + 1], lut[di, dj + 1]], -1)
    local = np.concatenate([up, down], 0)

    pts, faces = [], []
    for n, (a, b, c) in enumerate(base_f):
        p = (kk[:, None] * base_v[a] + ii[:, None] * base_v[b] + jj[:, None] * base_v[c]) / f
        pts.append(p / np.linalg.norm(p, axis=1, keepdims=True))
        faces.append(local + n * ii.size)
    pts = np.concatenate(pts, 0)
    faces = np.concatenate(faces, 0)
    # weld the duplicated edge / corner points (neighbours are >= ~1/f apart)
    key = np.round(pts * 1e6).astype(np.int64)
    _, first, inverse = np.unique(key, axis=0, return_index=True, return_inverse=True)
    inverse = inverse.reshape(-1)
    verts = pts[first] * radius
    faces = inverse[faces]
    # make every face wind counter-clockwise seen from outside
    n = np.cross(verts[faces[:, 1]] - verts[faces[:, 0]], verts[faces[:, 2]] - verts[faces[:, 0]])
    flip = (n * verts[faces].mean(1)).sum(-1) < 0
    faces[flip] = faces[flip][:, [0, 2, 1]]
    return verts, faces.astype(np.int64)


# ---------------------------------------------------------------------------------------------
# procedural terrain: (nx x ny) quads -> 2 nx ny faces (1000 x 500 -> 1 000 000 / 501 501)
# ---------------------------------------------------------------------------------------------

def terrain(nx: int = 1000, ny: int = 500, seed: int = 0) -> Tuple[np.ndarray, np.ndarray]:
    """Height field z = sum_k 2^-k 0.15 sin(2^k (a_k x + b_k y) + phi_k), x in [-1,1], y in [-.5,.5]."""
    rng = np.random.default_rng(seed)
    a = rng.uniform(1.0, 4.0, 6) * rng.choice([-1.0, 1.0], 6)
    b = rng.uniform(1.0, 4.0, 6) * rng.choice([-1.0, 1.0], 6)
    phi = rng.uniform(0.0, 2.0 * np.pi, 6)
    x = np.linspace(-1.0, 1.0, nx + 1)
    y = np.linspace(-0.5, 0.5, ny + 1)
    X, Y = np.meshgrid(x, y, indexing="xy")  # [ny+1, nx+1]
    Z = np.zeros_like(X)
    for k in range(6):
        Z += 2.0 ** -k * 0.15 * np.sin(2.0 ** k * (a[k] * X + b[k] * Y) + phi[k])
    verts = np.stack([X, Y, Z], -1).reshape(-1, 3)
    j, i = np.meshgrid(np.arange(ny), np.arange(nx), indexing="ij")
    v00 = (j * (nx + 1) + i).reshape(-1)
    v10, v01, v11 = v00 + 1, v00 + nx + 1, v00 + nx + 2
    faces = np.stack([np.stack([v00, v10, v11], -1), np.stack([v00, v11, v01], -1)], 1).reshape(-1, 3)
    return verts, faces.astype(np.int64)


def terrain_uv(nx: int = 1000, ny: int = 500, margin: float = 0.02) -> np.ndarray:
    """Planar per-vertex UVs for `terrain` (same vertex order): one chart, no overlap."""
    u = np.linspace(margin, 1.0 - margin, nx + 1)
    v = np.linspace(margin, 1.0 - margin, ny + 1)
    U, Vv = np.meshgrid(u, v, indexing="xy")
    return np.stack([U, Vv], -1).reshape(-1, 2)


def cell_atlas_uv(num_faces: int, pad: float = 0.08) -> Tuple[np.ndarray, np.ndarray]:
    """One triangle pair per square atlas cell; unique UV vertices per face.

    Returns (v_tex [3F,2] f64, t_tex_idx [F,3] i64).  With F = 50 000 the grid is 159 x 158
    cells ("one-triangle-pair-per-cell atlas").
    """
    pairs = (num_faces + 1) // 2
    n = int(np.ceil(np.sqrt(pairs)))
    cell = 1.0 / n
    fidx = np.arange(num_faces)
    p = fidx // 2
    cx, cy = (p % n) * cell, (p // n) * cell
    lo, hi = pad * cell, (1.0 - pad) * cell
    gap = 0.5 * pad * cell
    lower = np.array([[lo, lo + gap], [lo, hi], [hi - gap, hi]])  # below-diagonal triangle
    upper = np.array([[lo + gap, lo], [hi, hi - gap], [hi, lo]])  # above-diagonal triangle
    corners = np.where((fidx % 2 == 0)[:, None, None], lower[None], upper[None])
    v_tex = corners + np.stack([cx, cy], -1)[:, None, :]
    return v_tex.reshape(-1, 2), np.arange(3 * num_faces, dtype=np.int64).reshape(-1, 3)


def view_images(num_views: int, height: int, width: int, seed: int = 1) -> np.ndarray:
    """Smooth colour fields img[v,y,x,c] = .5 + .5 sin(w_c . (x,y) + phi_vc), float32 [Nv,H,W,3]."""
    rng = np.random.default_rng(seed)
    omega = rng.uniform(0.01, 0.06, (3, 2))
    phi = rng.uniform(0.0, 2.0 * np.pi, (num_views, 3))
    y, x = np.meshgrid(np.arange(height, dtype=np.float64), np.arange(width, dtype=np.float64), indexing="ij")
    arg = omega[None, :, 0, None, None] * x + omega[None, :, 1, None, None] * y + phi[:, :, None, None]
    # channels-last IN MEMORY, like a decoded image (a transposed view would make every consumer copy it)
    return np.ascontiguousarray((0.5 + 0.5 * np.sin(arg)).transpose(0, 2, 3, 1), dtype=np.float32)


def save_npz(path: str, vertices: np.ndarray, faces: np.ndarray) -> str:
    np.savez(path, vertices=vertices, faces=faces)
    return path


# canonical 6-view orthographic rig of the texture pipeline
# (reference: mvadapter/test/utils/pipeline_texture.py:226-230, 277-286)
CANONICAL_RIG = dict(
    elevation_deg=[0, 0, 0, 0, 89.99, -89.99],
    distance=[1.0] * 6,
    left=-0.55, right=0.55, bottom=-0.55, top=0.55,
    azimuth_deg=[x - 90 for x in [0, 90, 180, 270, 180, 180]],
)
