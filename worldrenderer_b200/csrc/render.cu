// Fused render() shading pass (reference render.py:220-286 + utils.py:127-139).
//
// Input: the packed (depth_key << 32 | triangle id) buffer left by the raster stages.  Output: mask,
// world position, view depth (raw or SimpleNormalization), normal, optionally triangle id, the
// nvdiffrast-layout rast tensor and the textured attribute map -- everything render() returns, written
// once.  Operation order of every expression: DESIGN.md 3.4-3.6 / 4 (same as oracle/).
//
//   k_shade           one pixel per thread, no shared memory and no block barrier: a warp whose 32 pixels
//                     are background (78% of the warps of config B) is one 8-byte load and its stores.
//                     Measured alternatives that LOST on config B (profiles/README.md): transposing the
//                     3-channel stores through shared memory (+7%), forcing 6 blocks/SM (+3%, spills),
//                     four pixels per thread with 16-byte loads/stores (+22%: the covered path serialises).
//   k_depth_finalize  second depth pass for normalisers that need the per-view min / max.
//
// The shading kernel resets every packed entry it consumes to WR_EMPTY_PIXEL, which leaves the
// buffer clean for the next call (no clear pass).
#include "common.cuh"
#include "texture.cuh"

namespace {

struct ShadeParams {
    wr_render_args a;
    unsigned long long *packed;  // [B,H,W]
    uint8_t *mask;               // [B,H,W] coverage: the caller's out_mask or scratch
    uint32_t *range;             // [B,4] zero-initialised: [0] = max of ~ordered(d) over all pixels (i.e. the
                                 // minimum), [1] = max of ordered(d) over covered pixels (0 = none)
};

struct PixelGeo {
    float px, py, pz;   // world position (zeros on background)
    float nx, ny, nz;   // normalised normal (background value on background)
    float tx, ty, tz;   // normalised tangent (generic instantiation only)
    float u, v, w;      // clamped barycentrics, w = (1 - u) - v
    float zw;           // clamped z/w (only when requested)
};

__device__ __forceinline__ float warp_min(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xFFFFFFFFu, v, o));
    return v;
}
__device__ __forceinline__ float warp_max(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xFFFFFFFFu, v, o));
    return v;
}

__device__ __forceinline__ float apply_simple(float d, float scale, float offset, int clamp)
{
    d = d * scale + offset;
    if (clamp) d = fminf(fmaxf(d, 0.0f), 1.0f);
    return d;
}

// Everything render() derives for a covered pixel (c, r) won by triangle `id`.  m = mvp of the view.
__device__ __forceinline__ void shade_covered(const wr_render_args &A, const float *m, int id, int c, int r,
                                              bool want_normal, bool want_zw, bool want_tangent, PixelGeo &g)
{
    const int W = A.W, H = A.H;
    const int i0 = __ldg(A.tri + 3 * (size_t)id), i1 = __ldg(A.tri + 3 * (size_t)id + 1),
              i2 = __ldg(A.tri + 3 * (size_t)id + 2);
    const float *q0 = A.v_pos + 3 * (size_t)i0, *q1 = A.v_pos + 3 * (size_t)i1, *q2 = A.v_pos + 3 * (size_t)i2;
    const float x0 = __ldg(q0), y0 = __ldg(q0 + 1), z0 = __ldg(q0 + 2);
    const float x1 = __ldg(q1), y1 = __ldg(q1 + 1), z1 = __ldg(q1 + 2);
    const float x2 = __ldg(q2), y2 = __ldg(q2 + 1), z2 = __ldg(q2 + 2);
    // the normal gathers depend on the indices only: issue them with the position gathers so that both
    // round trips overlap (and overlap the barycentric math below)
    float n0x = 0.f, n0y = 0.f, n0z = 0.f, n1x = 0.f, n1y = 0.f, n1z = 0.f, n2x = 0.f, n2y = 0.f, n2z = 0.f;
    if (want_normal) {
        int j0 = i0, j1 = i1, j2 = i2;
        if (A.tri_nrm) {
            j0 = __ldg(A.tri_nrm + 3 * (size_t)id); j1 = __ldg(A.tri_nrm + 3 * (size_t)id + 1);
            j2 = __ldg(A.tri_nrm + 3 * (size_t)id + 2);
        }
        if ((unsigned)j0 < (unsigned)A.Vn && (unsigned)j1 < (unsigned)A.Vn && (unsigned)j2 < (unsigned)A.Vn) {
            const float *n0 = A.v_nrm + 3 * (size_t)j0, *n1 = A.v_nrm + 3 * (size_t)j1, *n2 = A.v_nrm + 3 * (size_t)j2;
            n0x = __ldg(n0); n0y = __ldg(n0 + 1); n0z = __ldg(n0 + 2);
            n1x = __ldg(n1); n1y = __ldg(n1 + 1); n1z = __ldg(n1 + 2);
            n2x = __ldg(n2); n2y = __ldg(n2 + 1); n2z = __ldg(n2 + 2);
        }
    }
    // clip-space vertices, utils.py:127-129 in the contract's operation order
    const float c0x = ((m[0] * x0 + m[1] * y0) + m[2] * z0) + m[3];
    const float c0y = ((m[4] * x0 + m[5] * y0) + m[6] * z0) + m[7];
    const float c0w = ((m[12] * x0 + m[13] * y0) + m[14] * z0) + m[15];
    const float c1x = ((m[0] * x1 + m[1] * y1) + m[2] * z1) + m[3];
    const float c1y = ((m[4] * x1 + m[5] * y1) + m[6] * z1) + m[7];
    const float c1w = ((m[12] * x1 + m[13] * y1) + m[14] * z1) + m[15];
    const float c2x = ((m[0] * x2 + m[1] * y2) + m[2] * z2) + m[3];
    const float c2y = ((m[4] * x2 + m[5] * y2) + m[6] * z2) + m[7];
    const float c2w = ((m[12] * x2 + m[13] * y2) + m[14] * z2) + m[15];
    const float fx = (float)(2 * c + 1 - W) / (float)W;
    const float fy = (float)(2 * r + 1 - H) / (float)H;
    const float p0x = c0x - fx * c0w, p0y = c0y - fy * c0w;
    const float p1x = c1x - fx * c1w, p1y = c1y - fy * c1w;
    const float p2x = c2x - fx * c2w, p2y = c2y - fy * c2w;
    const float a0 = p1x * p2y - p1y * p2x;
    const float a1 = p2x * p0y - p2y * p0x;
    const float a2 = p0x * p1y - p0y * p1x;
    const float iw = 1.0f / ((a0 + a1) + a2);
    const float b0 = a0 * iw, b1 = a1 * iw;
    const float u = (b0 >= 0.0f) ? (b0 > 1.0f ? 1.0f : b0) : 0.0f;
    const float v = (b1 >= 0.0f) ? (b1 > 1.0f ? 1.0f : b1) : 0.0f;
    const float w = (1.0f - u) - v;
    g.u = u; g.v = v; g.w = w;
    if (want_zw) {
        const float c0z = ((m[8] * x0 + m[9] * y0) + m[10] * z0) + m[11];
        const float c1z = ((m[8] * x1 + m[9] * y1) + m[10] * z1) + m[11];
        const float c2z = ((m[8] * x2 + m[9] * y2) + m[10] * z2) + m[11];
        const float zc = ((c0z * a0) + (c1z * a1)) + (c2z * a2);
        const float wc = ((c0w * a0) + (c1w * a1)) + (c2w * a2);
        const float zw = zc / wc;
        g.zw = (zw >= -1.0f) ? (zw > 1.0f ? 1.0f : zw) : -1.0f;
    }
    g.px = ((x0 * u) + (x1 * v)) + (x2 * w);
    g.py = ((y0 * u) + (y1 * v)) + (y2 * w);
    g.pz = ((z0 * u) + (z1 * v)) + (z2 * w);
    if (want_normal) {
        const float ix = ((n0x * u) + (n1x * v)) + (n2x * w);
        const float iy = ((n0y * u) + (n1y * v)) + (n2y * w);
        const float iz = ((n0z * u) + (n1z * v)) + (n2z * w);
        const float ln = sqrtf((ix * ix + iy * iy) + iz * iz);
        const float dn = fmaxf(ln, 1e-12f);
        g.nx = ix / dn; g.ny = iy / dn; g.nz = iz / dn;
    }
    if (want_tangent) {  // render.py:280-284: same faces as the normals (stitched_t_pos_idx)
        int j0 = i0, j1 = i1, j2 = i2;
        if (A.tri_nrm) {
            j0 = __ldg(A.tri_nrm + 3 * (size_t)id); j1 = __ldg(A.tri_nrm + 3 * (size_t)id + 1);
            j2 = __ldg(A.tri_nrm + 3 * (size_t)id + 2);
        }
        float ix = 0.f, iy = 0.f, iz = 0.f;
        if ((unsigned)j0 < (unsigned)A.Vn && (unsigned)j1 < (unsigned)A.Vn && (unsigned)j2 < (unsigned)A.Vn) {
            const float *t0 = A.v_tang + 3 * (size_t)j0, *t1 = A.v_tang + 3 * (size_t)j1, *t2 = A.v_tang + 3 * (size_t)j2;
            ix = ((__ldg(t0) * u) + (__ldg(t1) * v)) + (__ldg(t2) * w);
            iy = ((__ldg(t0 + 1) * u) + (__ldg(t1 + 1) * v)) + (__ldg(t2 + 1) * w);
            iz = ((__ldg(t0 + 2) * u) + (__ldg(t1 + 2) * v)) + (__ldg(t2 + 2) * w);
        }
        const float ln = sqrtf((ix * ix + iy * iy) + iz * iz);
        const float dn = fmaxf(ln, 1e-12f);
        g.tx = ix / dn; g.ty = iy / dn; g.tz = iz / dn;
    }
}

constexpr int kShadeRows = 4;  // rows per thread: four independent 8-byte loads in flight before any use

// Output sets with their own instantiation (no per-pixel pointer tests, ~30% fewer issued instructions on
// config B); every other combination runs the generic instantiation (OUTS < 0, runtime tests).
constexpr int kOutNormal = 1, kOutDepth = 2, kOutTwoPass = 4, kOutGeo = 8;
constexpr int kOutsRenderDefault = kOutNormal | kOutDepth | kOutTwoPass;  // mask + pos + normal + min/max depth
constexpr int kOutsSimpleDepth = kOutNormal | kOutDepth;                  // mask + pos + normal + simple depth
constexpr int kOutsBakeView = kOutGeo | kOutDepth;                        // mask + (pos, aoi_cos) + simple depth

// One column strip of kShadeRows pixels per thread.  grid = (ceil(W/128), ceil(H/kShadeRows), B), 128 threads.
template <int OUTS>
__global__ void __launch_bounds__(128) k_shade(ShadeParams P)
{
    const wr_render_args &A = P.a;
    constexpr bool kGeneric = OUTS < 0;
    const bool has_mask = kGeneric ? (P.mask != nullptr) : true;
    const bool has_geo = kGeneric ? (A.out_geo != nullptr) : ((OUTS & kOutGeo) != 0);
    const bool has_pos = kGeneric ? (A.out_pos != nullptr) : !has_geo;
    const bool has_normal = kGeneric ? (A.out_normal != nullptr) : ((OUTS & kOutNormal) != 0);
    const bool need_normal = has_normal || has_geo;
    const bool has_depth = kGeneric ? (A.out_depth != nullptr) : ((OUTS & kOutDepth) != 0);
    const bool two_pass = kGeneric ? (has_depth && A.depth_mode != WR_DEPTH_SIMPLE) : ((OUTS & kOutTwoPass) != 0);
    const bool has_id = kGeneric && A.out_tri_id != nullptr;
    const bool has_rast = kGeneric && A.out_rast != nullptr;
    const bool has_attr = kGeneric && A.out_attr != nullptr;
    const bool has_tangent = kGeneric && A.out_tangent != nullptr;

    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int r0 = blockIdx.y * kShadeRows;
    const int b = blockIdx.z;
    const int W = A.W, H = A.H;
    const bool live = c < W;

    // Per-block depth range in shared memory.  The only block barrier sits at kernel entry, where no warp
    // waits on memory yet; afterwards warps retire independently and the last one to finish publishes --
    // and only if the block improves on the range it saw at entry (after the first wave almost none does).
    __shared__ uint32_t s_lo, s_hi, s_done, s_seen_lo, s_seen_hi;
    if (two_pass) {
        if (threadIdx.x == 0) {
            s_lo = 0u; s_hi = 0u; s_done = 0u;
            s_seen_lo = *reinterpret_cast<volatile uint32_t *>(P.range + 4 * b);
            s_seen_hi = *reinterpret_cast<volatile uint32_t *>(P.range + 4 * b + 1);
        }
        __syncthreads();
    }

    const size_t o0 = ((size_t)b * H + r0) * W + (live ? c : 0);
    const int nrows = live ? min(kShadeRows, H - r0) : 0;
    unsigned long long pk[kShadeRows];
#pragma unroll
    for (int k = 0; k < kShadeRows; ++k) pk[k] = (k < nrows) ? P.packed[o0 + (size_t)k * W] : WR_EMPTY_PIXEL;
    float m[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) m[j] = __ldg(A.mvp + 16 * b + j);  // independent of the packed ids: in flight with them
    float wz0 = 0.f, wz1 = 0.f, wz2 = 0.f, wz3 = 0.f;
    if (has_depth) {
        const float *m2 = A.w2c + 16 * b + 8;
        wz0 = __ldg(m2); wz1 = __ldg(m2 + 1); wz2 = __ldg(m2 + 2); wz3 = __ldg(m2 + 3);
    }
    const float nbx = A.normal_bg[0], nby = A.normal_bg[1], nbz = A.normal_bg[2];
    float rot[9];
    if (has_geo) {
#pragma unroll
        for (int j = 0; j < 9; ++j) rot[j] = __ldg(A.w2c + 16 * b + 4 * (j / 3) + (j % 3));  // R = w2c[:3,:3]
    }
    float lo = INFINITY, hi = -INFINITY;

#pragma unroll 1
    for (int k = 0; k < nrows; ++k) {
        const int r = r0 + k;
        const size_t o = o0 + (size_t)k * W;
        const bool covered = pk[k] != WR_EMPTY_PIXEL;
        int id = -1;
        PixelGeo g;
        g.px = g.py = g.pz = 0.f;
        g.nx = nbx; g.ny = nby; g.nz = nbz;
        g.tx = A.tangent_bg[0]; g.ty = A.tangent_bg[1]; g.tz = A.tangent_bg[2];
        g.u = g.v = g.w = 0.f; g.zw = 0.f;
        if (covered) {
            P.packed[o] = WR_EMPTY_PIXEL;  // self-cleaning
            id = (int)(uint32_t)(pk[k] & 0xFFFFFFFFull);
            shade_covered(A, m, id, c, r, need_normal, has_rast, has_tangent, g);
        }
        if (has_mask) P.mask[o] = covered ? 1 : 0;
        if (has_geo) {
            // uv.py:108-119: camera-space normal, re-normalised, z clamped to [0,1]; background keeps the
            // (un-rotated) background normal
            float aoi = g.nz;
            if (covered) {
                const float x = (rot[0] * g.nx + rot[1] * g.ny) + rot[2] * g.nz;
                const float y = (rot[3] * g.nx + rot[4] * g.ny) + rot[5] * g.nz;
                const float z = (rot[6] * g.nx + rot[7] * g.ny) + rot[8] * g.nz;
                const float ln = sqrtf((x * x + y * y) + z * z);
                aoi = z / fmaxf(ln, 1e-12f);
            }
            aoi = fminf(fmaxf(aoi, 0.0f), 1.0f);
            reinterpret_cast<float4 *>(A.out_geo)[o] = make_float4(g.px, g.py, g.pz, aoi);
        }
        if (has_pos) { float *d = A.out_pos + 3 * o; d[0] = g.px; d[1] = g.py; d[2] = g.pz; }
        if (has_normal) { float *d = A.out_normal + 3 * o; d[0] = g.nx; d[1] = g.ny; d[2] = g.nz; }
        if (has_tangent) { float *d = A.out_tangent + 3 * o; d[0] = g.tx; d[1] = g.ty; d[2] = g.tz; }
        if (has_id) A.out_tri_id[o] = id;
        if (has_rast)
            reinterpret_cast<float4 *>(A.out_rast)[o] =
                covered ? make_float4(g.u, g.v, g.zw, (float)(id + 1)) : make_float4(0.f, 0.f, 0.f, 0.f);
        if (has_attr) {
            float *d = A.out_attr + (size_t)A.TC * o;
            if (covered) {
                const int t0 = __ldg(A.tri_tex + 3 * (size_t)id), t1 = __ldg(A.tri_tex + 3 * (size_t)id + 1),
                          t2 = __ldg(A.tri_tex + 3 * (size_t)id + 2);
                float tu = 0.f, tv = 0.f;
                if ((unsigned)t0 < (unsigned)A.Vt && (unsigned)t1 < (unsigned)A.Vt && (unsigned)t2 < (unsigned)A.Vt) {
                    const float *e0 = A.v_tex + 2 * (size_t)t0, *e1 = A.v_tex + 2 * (size_t)t1, *e2 = A.v_tex + 2 * (size_t)t2;
                    tu = ((__ldg(e0) * g.u) + (__ldg(e1) * g.v)) + (__ldg(e2) * g.w);
                    tv = ((__ldg(e0 + 1) * g.u) + (__ldg(e1 + 1) * g.v)) + (__ldg(e2 + 1) * g.w);
                }
                for (int c0 = 0; c0 < A.TC; c0 += 4) {
                    float acc[4];
                    sample_texture<4>(A.texture + c0, A.TH, A.TW, A.TC, min(4, A.TC - c0), tu, tv, A.tex_filter, 0, acc);
                    for (int j = 0; j < 4 && c0 + j < A.TC; ++j) d[c0 + j] = acc[j];
                }
            } else {
                for (int j = 0; j < A.TC; ++j) d[j] = A.attr_bg;
            }
        }
        if (has_depth) {
            // view depth = -(w2c * (p,1)).z (render.py:248-249, utils.py:132-139); background uses p = 0
            const float zv = ((wz0 * g.px + wz1 * g.py) + wz2 * g.pz) + wz3;
            const float d = -zv;
            if (!two_pass) {
                A.out_depth[o] = covered ? apply_simple(d, A.depth_p0, A.depth_p1, A.depth_clamp) : A.depth_bg;
            } else {
                A.out_depth[o] = d;
                lo = fminf(lo, d);
                if (covered) hi = fmaxf(hi, d);
            }
        }
    }

    if (two_pass) {
        // per-view (min over all pixels, max over covered pixels): warp shuffle -> shared atomics -> the
        // last warp of the block issues at most one global atomic pair
        lo = warp_min(lo);
        hi = warp_max(hi);
        if ((threadIdx.x & 31) == 0) {
            if (lo < INFINITY) atomicMax(&s_lo, ~wr_float_ordered(lo));
            if (hi > -INFINITY) atomicMax(&s_hi, wr_float_ordered(hi));
            __threadfence_block();
            const uint32_t nwarps = (blockDim.x + 31) >> 5;
            if (atomicAdd(&s_done, 1u) == nwarps - 1) {
                __threadfence_block();
                const uint32_t klo = *reinterpret_cast<volatile uint32_t *>(&s_lo);
                const uint32_t khi = *reinterpret_cast<volatile uint32_t *>(&s_hi);
                if (klo > s_seen_lo) atomicMax(P.range + 4 * b, klo);
                if (khi > s_seen_hi) atomicMax(P.range + 4 * b + 1, khi);
            }
        }
    }
}

// Second depth pass (render.py:250-257): background <- per-view min, then the normaliser.
// kFinGroups x 4 pixels per thread; the groups of a thread are a block-stride apart so every load
// instruction stays coalesced and all of a thread's loads are in flight together.
constexpr int kFinGroups = 1;

__device__ __forceinline__ float finalize_one(float d, bool covered, float lo, float den, int mode, float p0, float p1,
                                              float bg)
{
    float x = covered ? d : lo;
    if (mode == WR_DEPTH_CONTROLNET || mode == WR_DEPTH_ZERO123PP) {
        float n = (x - lo) / den;
        n = fminf(fmaxf(n, 0.0f), 1.0f);
        if (mode == WR_DEPTH_CONTROLNET) {
            n = 1.0f - n;
            n = n * p1 + p0;
        }
        x = covered ? n : bg;
    }
    return x;
}

__global__ void __launch_bounds__(256) k_depth_finalize(float *depth, const uint8_t *mask, const uint32_t *range,
                                                        long long npix_view, int mode, float p0, float p1, float bg,
                                                        int vec)
{
    const int b = blockIdx.y;
    const float lo = wr_ordered_float(~range[4 * b]);
    const uint32_t hik = range[4 * b + 1];
    const float hi = hik == 0u ? lo : wr_ordered_float(hik);  // no covered pixel: filled image is constant lo
    const float den = (hi - lo) + 1e-5f;
    float *dv = depth + (size_t)b * npix_view;
    const uint8_t *mv = mask + (size_t)b * npix_view;
    const long long base = (long long)blockIdx.x * (256 * kFinGroups) + threadIdx.x;  // group index
    if (vec) {
        const long long ngroups = npix_view >> 2;
        float4 d[kFinGroups];
        uchar4 m[kFinGroups];
#pragma unroll
        for (int k = 0; k < kFinGroups; ++k) {
            const long long gi = base + 256 * k;
            if (gi < ngroups) {
                d[k] = reinterpret_cast<const float4 *>(dv)[gi];
                m[k] = reinterpret_cast<const uchar4 *>(mv)[gi];
            }
        }
#pragma unroll
        for (int k = 0; k < kFinGroups; ++k) {
            const long long gi = base + 256 * k;
            if (gi < ngroups) {
                float4 o;
                o.x = finalize_one(d[k].x, m[k].x, lo, den, mode, p0, p1, bg);
                o.y = finalize_one(d[k].y, m[k].y, lo, den, mode, p0, p1, bg);
                o.z = finalize_one(d[k].z, m[k].z, lo, den, mode, p0, p1, bg);
                o.w = finalize_one(d[k].w, m[k].w, lo, den, mode, p0, p1, bg);
                reinterpret_cast<float4 *>(dv)[gi] = o;
            }
        }
    } else {
        for (int k = 0; k < kFinGroups; ++k) {
            const long long i0 = (base + 256 * k) * 4;
            for (int j = 0; j < 4; ++j)
                if (i0 + j < npix_view) dv[i0 + j] = finalize_one(dv[i0 + j], mv[i0 + j] != 0, lo, den, mode, p0, p1, bg);
        }
    }
}

}  // namespace

extern "C" int wr_render(wr_ctx *ctx, const wr_render_args *args, void *stream_)
{
    if (!ctx || !args) return WR_ERR_INVALID_ARGUMENT;
    const wr_render_args &A = *args;
    if (A.B < 0 || A.V < 0 || A.F < 0 || A.H <= 0 || A.W <= 0 || A.H > 8192 || A.W > 8192) return WR_ERR_INVALID_ARGUMENT;
    if (A.F >= (1 << 30)) return WR_ERR_UNSUPPORTED;
    if (A.B == 0) return WR_OK;
    if (!A.mvp || (A.V > 0 && !A.v_pos) || (A.F > 0 && !A.tri)) return WR_ERR_INVALID_ARGUMENT;
    if (reinterpret_cast<uintptr_t>(A.mvp) & 15u) return WR_ERR_INVALID_ARGUMENT;  // read as float4 rows
    if (A.out_depth && !A.w2c) return WR_ERR_INVALID_ARGUMENT;
    if (A.out_normal && !A.v_nrm) return WR_ERR_INVALID_ARGUMENT;
    if (A.out_geo && (!A.v_nrm || !A.w2c)) return WR_ERR_INVALID_ARGUMENT;
    if (A.out_tangent && !A.v_tang) return WR_ERR_INVALID_ARGUMENT;
    if (A.out_attr && (!A.v_tex || !A.tri_tex || !A.texture || A.TH <= 0 || A.TW <= 0 || A.TC <= 0)) return WR_ERR_INVALID_ARGUMENT;
    if (A.depth_mode < WR_DEPTH_NONE || A.depth_mode > WR_DEPTH_SIMPLE) return WR_ERR_INVALID_ARGUMENT;
    if (A.tex_filter < 0 || A.tex_filter > 1) return WR_ERR_UNSUPPORTED;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) return wr_set_cuda_error(ctx, e, "cudaSetDevice");

    VtxSrc src;
    src.pos = A.v_pos; src.mvp = A.mvp; src.V = A.V; src.batched = 0;
    RasterResult res;
    const bool two_pass = A.out_depth && A.depth_mode != WR_DEPTH_SIMPLE;
    const size_t npix = (size_t)A.B * A.H * A.W;
    void *extra = nullptr;
    wr_stage_begin(ctx);
    int rc = wr_run_raster(ctx, src, A.B, A.tri, A.F, nullptr, A.H, A.W, (two_pass && !A.out_mask) ? npix : 0, &res,
                           &extra, stream);
    if (rc != WR_OK) return rc;

    ShadeParams P;
    P.a = A;
    P.packed = res.packed;
    P.mask = A.out_mask ? A.out_mask : (two_pass ? static_cast<uint8_t *>(extra) : nullptr);
    P.range = reinterpret_cast<uint32_t *>(res.view_stats);
    wr_stage(ctx, stream, "k_shade");
    {
        const dim3 grid(wr_div_up(A.W, 128), wr_div_up(A.H, kShadeRows), A.B);
        const bool extras = A.out_tri_id || A.out_rast || A.out_attr || A.out_tangent;
        const bool plain = P.mask && A.out_pos && A.out_normal && A.out_depth && !A.out_geo && !extras;
        const bool bake = P.mask && A.out_geo && A.out_depth && !two_pass && !A.out_pos && !A.out_normal && !extras;
        if (plain && two_pass) k_shade<kOutsRenderDefault><<<grid, 128, 0, stream>>>(P);
        else if (plain) k_shade<kOutsSimpleDepth><<<grid, 128, 0, stream>>>(P);
        else if (bake) k_shade<kOutsBakeView><<<grid, 128, 0, stream>>>(P);
        else k_shade<-1><<<grid, 128, 0, stream>>>(P);
    }
    WR_CHECK_LAUNCH(ctx, "k_shade");
    wr_raster_consumed(ctx, &res);
    if (two_pass) {
        const long long npv = (long long)A.H * A.W;
        const int vec = (npv & 3) == 0 && (reinterpret_cast<uintptr_t>(A.out_depth) & 15u) == 0 &&
                        (reinterpret_cast<uintptr_t>(P.mask) & 3u) == 0;
        wr_stage(ctx, stream, "k_depth_finalize");
        k_depth_finalize<<<dim3(wr_div_up(wr_div_up(npv, 4), 256 * kFinGroups), A.B), 256, 0, stream>>>(
            A.out_depth, P.mask, P.range, npv, A.depth_mode, A.depth_p0, A.depth_p1, A.depth_bg, vec);
        WR_CHECK_LAUNCH(ctx, "k_depth_finalize");
    }
    wr_stage(ctx, stream, "end");
    return WR_OK;
}
