"""Independent NumPy / pure-Python twin of oracle/wr_oracle_blend.c (TEST INFRASTRUCTURE ONLY, small inputs).

Written from the statements in the C file's comments, not from its code, so that the two can be compared bit for
bit: `poisson_blend` (vectorised NumPy, the same individually rounded float32 operations in the same order) and
`inpaint_u8` (plain loops: jump flooding, then the inverse-square-distance average around the nearest known pixel).
"""
from __future__ import annotations

import numpy as np

f32 = np.float32


def _shift(a: np.ndarray, dr: int, dc: int) -> np.ndarray:
    """out[r, c] = a[r + dr, c + dc], zero outside the image."""
    out = np.zeros_like(a)
    H, W = a.shape[:2]
    rs, re = max(0, -dr), min(H, H - dr)
    cs, ce = max(0, -dc), min(W, W - dc)
    if rs < re and cs < ce:
        out[rs:re, cs:ce] = a[rs + dr:re + dr, cs + dc:ce + dc]
    return out


_NBR = ((-1, 0), (1, 0), (0, -1), (0, 1))  # up, down, left, right (blend.py:289-297)


def poisson_blend(src, mask, tgt, num_iters: int, grad_mode: str = "src") -> np.ndarray:
    src = np.asarray(src, f32)
    tgt = np.asarray(tgt, f32)
    m = np.asarray(mask) != 0
    H, W, _ = tgt.shape
    m = m.copy()
    m[0, :] = m[-1, :] = False
    m[:, 0] = m[:, -1] = False
    if grad_mode == "src":
        lap = f32(4) * src
        for d in _NBR:
            lap = lap - _shift(src, *d)
    else:
        lap = None
        for d in _NBR:
            ds = src - _shift(src, *d)
            dt = tgt - _shift(tgt, *d)
            pick = np.where(np.abs(ds) > np.abs(dt), ds, dt) if grad_mode == "max" else (ds + dt) * f32(0.5)
            lap = pick if lap is None else lap + pick
    outside = np.where(m[..., None], f32(0), tgt).astype(f32)
    fq = None
    for d in _NBR:
        v = _shift(outside, *d)
        fq = v if fq is None else fq + v
    b = (lap + fq).astype(f32)
    x = np.where(m[..., None], tgt, f32(0)).astype(f32)
    for _ in range(int(num_iters)):
        s = ((_shift(x, -1, 0) + _shift(x, 1, 0)) + _shift(x, 0, -1)) + _shift(x, 0, 1)
        x = np.where(m[..., None], (s + b) * f32(0.25), f32(0)).astype(f32)
    return np.where(m[..., None], np.clip(x, f32(0), f32(1)), tgt).astype(f32)


def inpaint_u8(img, mask, radius: int) -> np.ndarray:
    img = np.asarray(img, np.uint8)
    fill = np.asarray(mask) != 0
    H, W, C = img.shape
    out = img.copy()
    if fill.all() or not fill.any():
        return out
    NONE = -1
    seed = np.where(fill, NONE, np.arange(H * W).reshape(H, W)).astype(np.int64)
    top = 1
    while top < max(H, W):
        top <<= 1
    steps = []
    s = top >> 1
    while s >= 1:
        steps.append(s)
        s >>= 1
    steps.append(1)

    def d2(r, c, sd):
        return (r - sd // W) ** 2 + (c - sd % W) ** 2

    for step in steps:
        nxt = seed.copy()
        for r in range(H):
            for c in range(W):
                best = seed[r, c]
                for j in (-step, 0, step):
                    for k in (-step, 0, step):
                        rr, cc = r + j, c + k
                        if not (0 <= rr < H and 0 <= cc < W):
                            continue
                        sd = seed[rr, cc]
                        if sd == NONE:
                            continue
                        if best == NONE or d2(r, c, sd) < d2(r, c, best) or (d2(r, c, sd) == d2(r, c, best) and sd < best):
                            best = sd
                nxt[r, c] = best
        seed = nxt
    for r in range(H):
        for c in range(W):
            if not fill[r, c]:
                continue
            q = int(seed[r, c])
            qr, qc = q // W, q % W
            num = f32(1 + d2(r, c, q))
            acc = [f32(0)] * C
            ws = f32(0)
            for tr in range(qr - radius, qr + radius + 1):
                for tc in range(qc - radius, qc + radius + 1):
                    if not (0 <= tr < H and 0 <= tc < W) or (tr - qr) ** 2 + (tc - qc) ** 2 > radius * radius or fill[tr, tc]:
                        continue
                    w = num / f32(1 + (r - tr) ** 2 + (c - tc) ** 2)
                    for ch in range(C):
                        acc[ch] = acc[ch] + w * f32(img[tr, tc, ch])
                    ws = ws + w
            for ch in range(C):
                out[r, c, ch] = np.uint8(min(f32(255), np.rint(acc[ch] / ws)))
    return out
