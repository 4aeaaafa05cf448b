"""Slow NumPy twin of the raster contract.  TEST INFRASTRUCTURE ONLY.

An independent second statement of `dr.rasterize` / `dr.interpolate` (reference call sites
render.py:55,79; uv.py:40-43) used to cross-check oracle/wr_oracle.c on small cases: one
Python iteration per triangle, exact Python/NumPy int64 edge functions, every fp32 operation
issued as its own NumPy ufunc (so no FMA contraction can occur).
"""
from __future__ import annotations

import numpy as np

f32 = np.float32
COORD_LIMIT = f32(4194304.0)
GUARD = f32(16.0)
EMPTY = np.uint64(0xFFFFFFFFFFFFFFFF)


def _depth_key(zw: np.ndarray) -> np.ndarray:
    u = zw.astype(f32).view(np.uint32)
    neg = (u & np.uint32(0x80000000)) != 0
    return np.where(neg, ~u, u | np.uint32(0x80000000)).astype(np.uint32)


def _project_snap(p, W, H):
    """p: (4,) f32 -> (X, Y, zw) or None."""
    x, y, z, w = (f32(v) for v in p)
    if not (w > 0):
        return None
    with np.errstate(all="ignore"):
        rw = f32(1.0) / w
        fx = (x * f32(8 * W)) * rw
        fy = (y * f32(8 * H)) * rw
        if not (abs(fx) <= COORD_LIMIT) or not (abs(fy) <= COORD_LIMIT):
            return None
        return int(np.rint(fx)), int(np.rint(fy)), z * rw


def _plane_dist(k, p):
    x, y, z, w = (f32(v) for v in p)
    if k == 0:
        return z + w
    if k == 1:
        return w - z
    if k == 2:
        return x + GUARD * w
    if k == 3:
        return GUARD * w - x
    if k == 4:
        return y + GUARD * w
    return GUARD * w - y


def _clip(poly):
    for k in range(6):
        d = [_plane_dist(k, p) for p in poly]
        out = []
        n = len(poly)
        for i in range(n):
            j = (i + 1) % n
            in_i, in_j = d[i] >= 0, d[j] >= 0
            if in_i:
                out.append(poly[i])
            if in_i != in_j:
                with np.errstate(all="ignore"):
                    t = d[i] / (d[i] - d[j])
                    out.append(((poly[j] - poly[i]) * t + poly[i]).astype(f32))
        poly = out
        if len(poly) < 3:
            return []
    return poly


def _top_left(ax, ay, bx, by):
    dx, dy = bx - ax, by - ay
    return dy < 0 or (dy == 0 and dx > 0)


def _raster_snapped(X, Y, ZW, tid, W, H, buf):
    X = [int(v) for v in X]
    Y = [int(v) for v in Y]
    ZW = [f32(v) for v in ZW]
    area2 = (X[1] - X[0]) * (Y[2] - Y[0]) - (Y[1] - Y[0]) * (X[2] - X[0])
    if area2 == 0:
        return
    if area2 < 0:
        X[1], X[2] = X[2], X[1]
        Y[1], Y[2] = Y[2], Y[1]
        ZW[1], ZW[2] = ZW[2], ZW[1]
        area2 = -area2
    ox, oy = 8 - 8 * W, 8 - 8 * H
    c0 = max(-((-(min(X) - ox)) // 16), 0)
    c1 = min((max(X) - ox) // 16, W - 1)
    r0 = max(-((-(min(Y) - oy)) // 16), 0)
    r1 = min((max(Y) - oy) // 16, H - 1)
    if c0 > c1 or r0 > r1:
        return
    px = (16 * np.arange(c0, c1 + 1, dtype=np.int64) + ox)[None, :]
    py = (16 * np.arange(r0, r1 + 1, dtype=np.int64) + oy)[:, None]

    def edge(a, b):
        return (X[b] - X[a]) * (py - Y[a]) - (Y[b] - Y[a]) * (px - X[a])

    e0, e1, e2 = edge(1, 2), edge(2, 0), edge(0, 1)
    b0_ = 0 if _top_left(X[1], Y[1], X[2], Y[2]) else 1
    b1_ = 0 if _top_left(X[2], Y[2], X[0], Y[0]) else 1
    b2_ = 0 if _top_left(X[0], Y[0], X[1], Y[1]) else 1
    cover = (e0 >= b0_) & (e1 >= b1_) & (e2 >= b2_)
    if not cover.any():
        return
    with np.errstate(all="ignore"):
        inv = f32(1.0) / f32(np.int64(area2))
        b0 = e0.astype(f32) * inv
        b1 = e1.astype(f32) * inv
        b2 = (f32(1.0) - b0) - b1
        zw = ((ZW[0] * b0) + (ZW[1] * b1)) + (ZW[2] * b2)
        zw = (zw + f32(0.0)).astype(f32)
        ok = cover & (zw >= -1) & (zw <= 1)
    packed = (_depth_key(zw).astype(np.uint64) << np.uint64(32)) | np.uint64(tid)
    view = buf[r0:r1 + 1, c0:c1 + 1]
    upd = ok & (packed < view)
    view[upd] = packed[upd]


def rasterize_ids(pos: np.ndarray, tri: np.ndarray, H: int, W: int) -> np.ndarray:
    """pos [V,4] f32 clip space, tri [F,3] -> tri_id [H,W] int32 (-1 = background)."""
    pos = np.asarray(pos, f32)
    tri = np.asarray(tri, np.int64).reshape(-1, 3)
    V = pos.shape[0]
    buf = np.full((H, W), EMPTY, np.uint64)
    for t, (i0, i1, i2) in enumerate(tri):
        if min(i0, i1, i2) < 0 or max(i0, i1, i2) >= V:
            continue
        p = pos[[i0, i1, i2]]
        if not np.isfinite(p).all():
            continue
        x, y, z, w = p[:, 0], p[:, 1], p[:, 2], p[:, 3]
        if (x < -w).all() or (x > w).all() or (y < -w).all() or (y > w).all() or (z < -w).all() or (z > w).all():
            continue
        snapped = [_project_snap(q, W, H) for q in p]
        if all(s is not None for s in snapped):
            _raster_snapped([s[0] for s in snapped], [s[1] for s in snapped], [s[2] for s in snapped], t, W, H, buf)
            continue
        poly = _clip([q.copy() for q in p])
        if len(poly) < 3:
            continue
        sn = [_project_snap(q, W, H) for q in poly]
        if any(s is None for s in sn):
            continue
        for i in range(1, len(sn) - 1):
            a, b, c = sn[0], sn[i], sn[i + 1]
            _raster_snapped([a[0], b[0], c[0]], [a[1], b[1], c[1]], [a[2], b[2], c[2]], t, W, H, buf)
    ids = (buf & np.uint64(0xFFFFFFFF)).astype(np.int64)
    ids[buf == EMPTY] = -1
    return ids.astype(np.int32)


def shade(pos: np.ndarray, tri: np.ndarray, ids: np.ndarray) -> np.ndarray:
    """(u, v, z/w, id+1) per pixel from the unsnapped vertices."""
    pos = np.asarray(pos, f32)
    tri = np.asarray(tri, np.int64).reshape(-1, 3)
    H, W = ids.shape
    rast = np.zeros((H, W, 4), f32)
    fg = ids >= 0
    if not fg.any():
        return rast
    r, c = np.nonzero(fg)
    t = ids[r, c].astype(np.int64)
    p0, p1, p2 = pos[tri[t, 0]], pos[tri[t, 1]], pos[tri[t, 2]]
    with np.errstate(all="ignore"):
        fx = (2 * c + 1 - W).astype(f32) / f32(W)
        fy = (2 * r + 1 - H).astype(f32) / f32(H)
        p0x, p0y = p0[:, 0] - fx * p0[:, 3], p0[:, 1] - fy * p0[:, 3]
        p1x, p1y = p1[:, 0] - fx * p1[:, 3], p1[:, 1] - fy * p1[:, 3]
        p2x, p2y = p2[:, 0] - fx * p2[:, 3], p2[:, 1] - fy * p2[:, 3]
        a0 = p1x * p2y - p1y * p2x
        a1 = p2x * p0y - p2y * p0x
        a2 = p0x * p1y - p0y * p1x
        iw = f32(1.0) / ((a0 + a1) + a2)
        b0, b1 = a0 * iw, a1 * iw
        z = ((p0[:, 2] * a0) + (p1[:, 2] * a1)) + (p2[:, 2] * a2)
        w = ((p0[:, 3] * a0) + (p1[:, 3] * a1)) + (p2[:, 3] * a2)
        zw = z / w
        u = np.where(b0 >= 0, np.minimum(b0, f32(1)), f32(0))
        v = np.where(b1 >= 0, np.minimum(b1, f32(1)), f32(0))
        d = np.where(zw >= -1, np.minimum(zw, f32(1)), f32(-1))
    rast[r, c, 0], rast[r, c, 1], rast[r, c, 2] = u, v, d
    rast[r, c, 3] = (t + 1).astype(f32)
    return rast


def rasterize(pos, tri, resolution):
    pos = np.asarray(pos, f32)
    if pos.ndim == 2:
        pos = pos[None]
    H, W = resolution
    ids = np.stack([rasterize_ids(p, tri, H, W) for p in pos])
    rast = np.stack([shade(p, tri, i) for p, i in zip(pos, ids)])
    return rast, ids


def interpolate(attr, rast, tri):
    attr = np.asarray(attr, f32)
    if attr.ndim == 2:
        attr = attr[None]
    tri = np.asarray(tri, np.int64).reshape(-1, 3)
    B, H, W, _ = rast.shape
    V, A = attr.shape[1], attr.shape[2]
    out = np.zeros((B, H, W, A), f32)
    for b in range(B):
        ab = attr[0 if attr.shape[0] == 1 else b]
        ids = rast[b, ..., 3].astype(np.int64) - 1
        ok = (ids >= 0) & (ids < tri.shape[0])
        r, c = np.nonzero(ok)
        t = tri[ids[r, c]]
        good = ((t >= 0) & (t < V)).all(-1)
        r, c, t = r[good], c[good], t[good]
        u, v = rast[b, r, c, 0:1], rast[b, r, c, 1:2]
        w = (f32(1.0) - u) - v
        out[b, r, c] = ((ab[t[:, 0]] * u) + (ab[t[:, 1]] * v)) + (ab[t[:, 2]] * w)
    return out
