// dr.texture sampling core (2-D, no mip maps), shared by k_texture and the fused render pass.
// Operation order: DESIGN.md 3.6 (same as oracle/wr_oracle.c wro_texture).
#pragma once
#include "common.cuh"

__device__ __forceinline__ int wrap_i(int i, int n) { int m = i % n; return m < 0 ? m + n : m; }
__device__ __forceinline__ float clampf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }

// Samples nch (<= MAXC) consecutive channels of tex [TH,TW,C] (tb already points at the first one) at (u, v).
template <int MAXC>
__device__ __forceinline__ void sample_texture(const float *tb, int TH, int TW, int C, int nch, float u, float v,
                                               int filter, int boundary, float *acc)
{
    if (boundary == 0) { u = u - floorf(u); v = v - floorf(v); }
    float x = u * (float)TW, y = v * (float)TH;
    if (boundary == 1) { x = clampf(x, 0.0f, (float)TW); y = clampf(y, 0.0f, (float)TH); }
    for (int a = 0; a < MAXC; ++a) acc[a] = 0.0f;
    if (!isfinite(x) || !isfinite(y)) return;
    int ix[2], iy[2];
    float wx[2], wy[2];
    int taps;
    if (filter == 0) {
        ix[0] = (int)floorf(x); iy[0] = (int)floorf(y);
        ix[1] = ix[0]; iy[1] = iy[0];
        wx[0] = wy[0] = 1.0f; wx[1] = wy[1] = 0.0f;
        taps = 1;
    } else {
        const float xs = x - 0.5f, ys = y - 0.5f;
        const float x0 = floorf(xs), y0 = floorf(ys);
        ix[0] = (int)x0; ix[1] = ix[0] + 1;
        iy[0] = (int)y0; iy[1] = iy[0] + 1;
        wx[1] = xs - x0; wx[0] = 1.0f - wx[1];
        wy[1] = ys - y0; wy[0] = 1.0f - wy[1];
        taps = 2;
    }
    for (int j = 0; j < taps; ++j) {
        for (int i = 0; i < taps; ++i) {
            int tx = ix[i], ty = iy[j];
            if (boundary == 0) { tx = wrap_i(tx, TW); ty = wrap_i(ty, TH); }
            else if (boundary == 1) {
                tx = tx < 0 ? 0 : (tx > TW - 1 ? TW - 1 : tx);
                ty = ty < 0 ? 0 : (ty > TH - 1 ? TH - 1 : ty);
            } else if (tx < 0 || tx >= TW || ty < 0 || ty >= TH) continue;
            const float wgt = wx[i] * wy[j];
            const float *s = tb + ((size_t)ty * TW + tx) * C;
            for (int a = 0; a < MAXC; ++a)
                if (a < nch) acc[a] = acc[a] + __ldg(s + a) * wgt;
        }
    }
}
