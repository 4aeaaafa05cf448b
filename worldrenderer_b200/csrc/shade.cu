// Per-pixel stages: dr.interpolate, dr.texture, vertex normals and the fused render() shading pass.
// Reference: render.py:64-120 (operators), render.py:220-286 (render), mesh.py:85-119 (normals).
// Operation order of every expression follows DESIGN.md section 3.4-3.6 / 4 (same as oracle/).
#include "common.cuh"

namespace {

// ------------------------------------------------------------------------------------------
// dr.interpolate
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_interpolate(const float *attr, int attr_B, int V, int A, const float *rast,
                                                     long long npix_total, long long npix_view, const int32_t *tri,
                                                     int F, float *out)
{
    const long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= npix_total) return;
    const int b = (int)(o / npix_view);
    const float4 r4 = __ldg(reinterpret_cast<const float4 *>(rast) + o);
    float *dst = out + o * A;
    const long long id = (long long)r4.w - 1;
    bool valid = id >= 0 && id < F;
    int i0 = 0, i1 = 0, i2 = 0;
    if (valid) {
        i0 = __ldg(tri + 3 * id); i1 = __ldg(tri + 3 * id + 1); i2 = __ldg(tri + 3 * id + 2);
        valid = (unsigned)i0 < (unsigned)V && (unsigned)i1 < (unsigned)V && (unsigned)i2 < (unsigned)V;
    }
    if (!valid) {
        for (int a = 0; a < A; ++a) dst[a] = 0.0f;
        return;
    }
    const float *ab = attr + (attr_B == 1 ? 0 : (size_t)b * V * A);
    const float u = r4.x, v = r4.y, w = (1.0f - u) - v;
    const float *a0 = ab + (size_t)i0 * A, *a1 = ab + (size_t)i1 * A, *a2 = ab + (size_t)i2 * A;
    for (int a = 0; a < A; ++a) dst[a] = ((__ldg(a0 + a) * u) + (__ldg(a1 + a) * v)) + (__ldg(a2 + a) * w);
}

// ------------------------------------------------------------------------------------------
// dr.texture (2-D, no mip maps)
// ------------------------------------------------------------------------------------------
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

__global__ void __launch_bounds__(256) k_texture(const float *tex, int tex_B, int TH, int TW, int C, const float *uv,
                                                 long long npix_total, long long npix_view, int filter, int boundary,
                                                 float *out)
{
    const long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= npix_total) return;
    const int b = (int)(o / npix_view);
    const float *tb = tex + (tex_B == 1 ? 0 : (size_t)b * TH * TW * C);
    const float2 q = __ldg(reinterpret_cast<const float2 *>(uv) + o);
    float *dst = out + o * C;
    for (int c0 = 0; c0 < C; c0 += 4) {  // channel groups of four
        float acc[4];
        sample_texture<4>(tb + c0, TH, TW, C, min(4, C - c0), q.x, q.y, filter, boundary, acc);
        for (int a = 0; a < 4 && c0 + a < C; ++a) dst[c0 + a] = acc[a];
    }
}

// ------------------------------------------------------------------------------------------
// vertex normals (mesh.py:85-119)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_face_normals_scatter(const float *v_pos, int V, const int32_t *tri, int F,
                                                              float *acc)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= F) return;
    const int i0 = __ldg(tri + 3 * (size_t)t), i1 = __ldg(tri + 3 * (size_t)t + 1), i2 = __ldg(tri + 3 * (size_t)t + 2);
    if ((unsigned)i0 >= (unsigned)V || (unsigned)i1 >= (unsigned)V || (unsigned)i2 >= (unsigned)V) return;
    const float *p0 = v_pos + 3 * (size_t)i0, *p1 = v_pos + 3 * (size_t)i1, *p2 = v_pos + 3 * (size_t)i2;
    const float ax = __ldg(p1) - __ldg(p0), ay = __ldg(p1 + 1) - __ldg(p0 + 1), az = __ldg(p1 + 2) - __ldg(p0 + 2);
    const float bx = __ldg(p2) - __ldg(p0), by = __ldg(p2 + 1) - __ldg(p0 + 1), bz = __ldg(p2 + 2) - __ldg(p0 + 2);
    const float nx = ay * bz - az * by, ny = az * bx - ax * bz, nz = ax * by - ay * bx;
    const int idx[3] = { i0, i1, i2 };
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float *d = acc + 3 * (size_t)idx[k];
        atomicAdd(d, nx); atomicAdd(d + 1, ny); atomicAdd(d + 2, nz);
    }
}

__global__ void __launch_bounds__(256) k_normalize_vertex_normals(float *acc, int V)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    float x = acc[3 * (size_t)v], y = acc[3 * (size_t)v + 1], z = acc[3 * (size_t)v + 2];
    const float sq = (x * x + y * y) + z * z;
    if (!(sq > 1e-20f)) { x = 0.0f; y = 0.0f; z = 1.0f; }
    const float n = sqrtf((x * x + y * y) + z * z);
    const float d = fmaxf(n, 1e-12f);
    acc[3 * (size_t)v] = x / d; acc[3 * (size_t)v + 1] = y / d; acc[3 * (size_t)v + 2] = z / d;
}

// ------------------------------------------------------------------------------------------
// fused render() shading
// ------------------------------------------------------------------------------------------
struct ShadeParams {
    wr_render_args a;
    const unsigned long long *packed;
    uint32_t *range;  // [B,2] ordered-uint (min over all pixels, max over covered pixels); [B,+2] covered count
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

// One thread per pixel.  grid = (ceil(W/256), H, B).
__global__ void __launch_bounds__(256) k_shade(ShadeParams P)
{
    const wr_render_args &A = P.a;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    const int b = blockIdx.z;
    const int W = A.W, H = A.H;
    const bool live = c < W;
    const size_t o = ((size_t)b * H + r) * W + (live ? c : 0);

    __shared__ float s_m[32];  // mvp (16) + w2c row 2 (4)
    if (threadIdx.x < 16) s_m[threadIdx.x] = A.mvp[16 * b + threadIdx.x];
    else if (threadIdx.x < 20) s_m[threadIdx.x] = A.w2c[16 * b + 8 + (threadIdx.x - 16)];
    __syncthreads();

    bool covered = false;
    float px = 0.f, py = 0.f, pz = 0.f;
    float nx = A.normal_bg[0], ny = A.normal_bg[1], nz = A.normal_bg[2];
    float4 rast = make_float4(0.f, 0.f, 0.f, 0.f);
    int id = -1;
    float u = 0.f, v = 0.f, w = 0.f;
    if (live) {
        const unsigned long long pk = P.packed[o];
        if (pk != WR_EMPTY_PIXEL) {
            covered = true;
            id = (int)(uint32_t)(pk & 0xFFFFFFFFull);
            const int i0 = __ldg(A.tri + 3 * (size_t)id), i1 = __ldg(A.tri + 3 * (size_t)id + 1),
                      i2 = __ldg(A.tri + 3 * (size_t)id + 2);
            const float *q0 = A.v_pos + 3 * (size_t)i0, *q1 = A.v_pos + 3 * (size_t)i1, *q2 = A.v_pos + 3 * (size_t)i2;
            const float x0 = __ldg(q0), y0 = __ldg(q0 + 1), z0 = __ldg(q0 + 2);
            const float x1 = __ldg(q1), y1 = __ldg(q1 + 1), z1 = __ldg(q1 + 2);
            const float x2 = __ldg(q2), y2 = __ldg(q2 + 1), z2 = __ldg(q2 + 2);
            const float *m = s_m;
            // clip-space vertices, utils.py:127-129 in the contract's operation order
            const float c0x = ((m[0] * x0 + m[1] * y0) + m[2] * z0) + m[3];
            const float c0y = ((m[4] * x0 + m[5] * y0) + m[6] * z0) + m[7];
            const float c0z = ((m[8] * x0 + m[9] * y0) + m[10] * z0) + m[11];
            const float c0w = ((m[12] * x0 + m[13] * y0) + m[14] * z0) + m[15];
            const float c1x = ((m[0] * x1 + m[1] * y1) + m[2] * z1) + m[3];
            const float c1y = ((m[4] * x1 + m[5] * y1) + m[6] * z1) + m[7];
            const float c1z = ((m[8] * x1 + m[9] * y1) + m[10] * z1) + m[11];
            const float c1w = ((m[12] * x1 + m[13] * y1) + m[14] * z1) + m[15];
            const float c2x = ((m[0] * x2 + m[1] * y2) + m[2] * z2) + m[3];
            const float c2y = ((m[4] * x2 + m[5] * y2) + m[6] * z2) + m[7];
            const float c2z = ((m[8] * x2 + m[9] * y2) + m[10] * z2) + m[11];
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
            const float zc = ((c0z * a0) + (c1z * a1)) + (c2z * a2);
            const float wc = ((c0w * a0) + (c1w * a1)) + (c2w * a2);
            const float zw = zc / wc;
            u = (b0 >= 0.0f) ? (b0 > 1.0f ? 1.0f : b0) : 0.0f;
            v = (b1 >= 0.0f) ? (b1 > 1.0f ? 1.0f : b1) : 0.0f;
            w = (1.0f - u) - v;
            rast = make_float4(u, v, (zw >= -1.0f) ? (zw > 1.0f ? 1.0f : zw) : -1.0f, (float)(id + 1));
            px = ((x0 * u) + (x1 * v)) + (x2 * w);
            py = ((y0 * u) + (y1 * v)) + (y2 * w);
            pz = ((z0 * u) + (z1 * v)) + (z2 * w);
            if (A.out_normal) {
                int j0 = i0, j1 = i1, j2 = i2;
                if (A.tri_nrm) {
                    j0 = __ldg(A.tri_nrm + 3 * (size_t)id); j1 = __ldg(A.tri_nrm + 3 * (size_t)id + 1);
                    j2 = __ldg(A.tri_nrm + 3 * (size_t)id + 2);
                }
                float ix = 0.f, iy = 0.f, iz = 0.f;
                if ((unsigned)j0 < (unsigned)A.Vn && (unsigned)j1 < (unsigned)A.Vn && (unsigned)j2 < (unsigned)A.Vn) {
                    const float *n0 = A.v_nrm + 3 * (size_t)j0, *n1 = A.v_nrm + 3 * (size_t)j1, *n2 = A.v_nrm + 3 * (size_t)j2;
                    ix = ((__ldg(n0) * u) + (__ldg(n1) * v)) + (__ldg(n2) * w);
                    iy = ((__ldg(n0 + 1) * u) + (__ldg(n1 + 1) * v)) + (__ldg(n2 + 1) * w);
                    iz = ((__ldg(n0 + 2) * u) + (__ldg(n1 + 2) * v)) + (__ldg(n2 + 2) * w);
                }
                const float ln = sqrtf((ix * ix + iy * iy) + iz * iz);
                const float dn = fmaxf(ln, 1e-12f);
                nx = ix / dn; ny = iy / dn; nz = iz / dn;
            }
        }
    }

    if (live) {
        if (A.out_mask) A.out_mask[o] = covered ? 1 : 0;
        if (A.out_tri_id) A.out_tri_id[o] = id;
        if (A.out_rast) reinterpret_cast<float4 *>(A.out_rast)[o] = rast;
        if (A.out_pos) { float *d = A.out_pos + 3 * o; d[0] = px; d[1] = py; d[2] = pz; }
        if (A.out_normal) { float *d = A.out_normal + 3 * o; d[0] = nx; d[1] = ny; d[2] = nz; }
        if (A.out_attr) {
            float *d = A.out_attr + (size_t)A.TC * o;
            if (covered) {
                int t0 = -1, t1 = -1, t2 = -1;
                t0 = __ldg(A.tri_tex + 3 * (size_t)id); t1 = __ldg(A.tri_tex + 3 * (size_t)id + 1);
                t2 = __ldg(A.tri_tex + 3 * (size_t)id + 2);
                float tu = 0.f, tv = 0.f;
                if ((unsigned)t0 < (unsigned)A.Vt && (unsigned)t1 < (unsigned)A.Vt && (unsigned)t2 < (unsigned)A.Vt) {
                    const float *e0 = A.v_tex + 2 * (size_t)t0, *e1 = A.v_tex + 2 * (size_t)t1, *e2 = A.v_tex + 2 * (size_t)t2;
                    tu = ((__ldg(e0) * u) + (__ldg(e1) * v)) + (__ldg(e2) * w);
                    tv = ((__ldg(e0 + 1) * u) + (__ldg(e1 + 1) * v)) + (__ldg(e2 + 1) * w);
                }
                for (int c0 = 0; c0 < A.TC; c0 += 4) {
                    float acc[4];
                    sample_texture<4>(A.texture + c0, A.TH, A.TW, A.TC, min(4, A.TC - c0), tu, tv, A.tex_filter, 0, acc);
                    for (int k = 0; k < 4 && c0 + k < A.TC; ++k) d[c0 + k] = acc[k];
                }
            } else {
                for (int k = 0; k < A.TC; ++k) d[k] = A.attr_bg;
            }
        }
    }

    if (A.out_depth) {
        // view depth = -(w2c * (p,1)).z (render.py:248-249, utils.py:132-139); background uses p = 0
        const float *m2 = s_m + 16;
        const float zv = ((m2[0] * px + m2[1] * py) + m2[2] * pz) + m2[3];
        const float d = -zv;
        if (A.depth_mode == WR_DEPTH_SIMPLE) {
            if (live) A.out_depth[o] = covered ? apply_simple(d, A.depth_p0, A.depth_p1, A.depth_clamp) : A.depth_bg;
        } else {
            if (live) A.out_depth[o] = d;
            // per-view reductions for the second pass
            float lo = live ? d : INFINITY;
            float hi = (live && covered) ? d : -INFINITY;
            lo = warp_min(lo); hi = warp_max(hi);
            __shared__ float s_lo[8], s_hi[8];
            const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
            if (lane == 0) { s_lo[wid] = lo; s_hi[wid] = hi; }
            __syncthreads();
            if (threadIdx.x == 0) {
                for (int k = 1; k < 8; ++k) { lo = fminf(lo, s_lo[k]); hi = fmaxf(hi, s_hi[k]); }
                atomicMin(P.range + 4 * b, wr_float_ordered(lo));
                if (hi > -INFINITY) atomicMax(P.range + 4 * b + 1, wr_float_ordered(hi));
            }
        }
    }
}

__global__ void k_init_range(uint32_t *range, int B)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    range[4 * b] = 0xFFFFFFFFu;  // ordered +max
    range[4 * b + 1] = 0u;       // ordered -max: "no covered pixel"
}

// Second depth pass (render.py:250-257): background <- per-view min, then the normaliser.
__global__ void __launch_bounds__(256) k_depth_finalize(float *depth, const uint8_t *mask_or_null,
                                                        const unsigned long long *packed, const uint32_t *range,
                                                        long long npix_view, int mode, float p0, float p1, float bg)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (i >= npix_view) return;
    const size_t o = (size_t)b * npix_view + i;
    const bool covered = packed[o] != WR_EMPTY_PIXEL;
    const float lo = wr_ordered_float(range[4 * b]);
    const uint32_t hik = range[4 * b + 1];
    const float hi = hik == 0u ? lo : wr_ordered_float(hik);  // no covered pixel: filled image is constant lo
    float d = covered ? depth[o] : lo;
    if (mode == WR_DEPTH_CONTROLNET || mode == WR_DEPTH_ZERO123PP) {
        float n = (d - lo) / ((hi - lo) + 1e-5f);
        n = fminf(fmaxf(n, 0.0f), 1.0f);
        if (mode == WR_DEPTH_CONTROLNET) {
            n = 1.0f - n;
            n = n * p1 + p0;
        }
        d = covered ? n : bg;
    }
    depth[o] = d;
    (void)mask_or_null;
}

}  // namespace

extern "C" int wr_interpolate(wr_ctx *ctx, const float *attr, int attr_B, int V, int A, const float *rast, int B,
                              int H, int W, const int32_t *tri, int F, float *out, void *stream_)
{
    if (!ctx || B < 0 || H <= 0 || W <= 0 || A < 0 || V < 0 || F < 0) return WR_ERR_INVALID_ARGUMENT;
    if (attr_B != 1 && attr_B != B) return WR_ERR_INVALID_ARGUMENT;
    if (B == 0 || A == 0) return WR_OK;
    if (!rast || !out || (V > 0 && !attr) || (F > 0 && !tri)) return WR_ERR_INVALID_ARGUMENT;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) return wr_set_cuda_error(ctx, e, "cudaSetDevice");
    const long long npv = (long long)H * W, total = npv * B;
    k_interpolate<<<wr_div_up(total, 256), 256, 0, stream>>>(attr, attr_B, V, A, rast, total, npv, tri, F, out);
    WR_CHECK_LAUNCH(ctx, "k_interpolate");
    return WR_OK;
}

extern "C" int wr_texture(wr_ctx *ctx, const float *tex, int tex_B, int TH, int TW, int C, const float *uv, int B,
                          int H, int W, int filter, int boundary, float *out, void *stream_)
{
    if (!ctx || B < 0 || H <= 0 || W <= 0 || TH <= 0 || TW <= 0 || C <= 0) return WR_ERR_INVALID_ARGUMENT;
    if (tex_B != 1 && tex_B != B) return WR_ERR_INVALID_ARGUMENT;
    if (filter < 0 || filter > 1 || boundary < 0 || boundary > 2) return WR_ERR_UNSUPPORTED;
    if (B == 0) return WR_OK;
    if (!tex || !uv || !out) return WR_ERR_INVALID_ARGUMENT;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) return wr_set_cuda_error(ctx, e, "cudaSetDevice");
    const long long npv = (long long)H * W, total = npv * B;
    k_texture<<<wr_div_up(total, 256), 256, 0, stream>>>(tex, tex_B, TH, TW, C, uv, total, npv, filter, boundary, out);
    WR_CHECK_LAUNCH(ctx, "k_texture");
    return WR_OK;
}

extern "C" int wr_vertex_normals(wr_ctx *ctx, const float *v_pos, int V, const int32_t *tri, int F, float *v_nrm,
                                 void *stream_)
{
    if (!ctx || V < 0 || F < 0) return WR_ERR_INVALID_ARGUMENT;
    if (V == 0) return WR_OK;
    if (!v_pos || !v_nrm || (F > 0 && !tri)) return WR_ERR_INVALID_ARGUMENT;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) return wr_set_cuda_error(ctx, e, "cudaSetDevice");
    e = cudaMemsetAsync(v_nrm, 0, (size_t)V * 3 * sizeof(float), stream);
    if (e != cudaSuccess) return wr_set_cuda_error(ctx, e, "memset normals");
    if (F > 0) {
        k_face_normals_scatter<<<wr_div_up(F, 256), 256, 0, stream>>>(v_pos, V, tri, F, v_nrm);
        WR_CHECK_LAUNCH(ctx, "k_face_normals_scatter");
    }
    k_normalize_vertex_normals<<<wr_div_up(V, 256), 256, 0, stream>>>(v_nrm, V);
    WR_CHECK_LAUNCH(ctx, "k_normalize_vertex_normals");
    return WR_OK;
}

extern "C" int wr_render(wr_ctx *ctx, const wr_render_args *args, void *stream_)
{
    if (!ctx || !args) return WR_ERR_INVALID_ARGUMENT;
    const wr_render_args &A = *args;
    if (A.B < 0 || A.V < 0 || A.F < 0 || A.H <= 0 || A.W <= 0 || A.H > 8192 || A.W > 8192) return WR_ERR_INVALID_ARGUMENT;
    if (A.F >= (1 << 30)) return WR_ERR_UNSUPPORTED;
    if (A.B == 0) return WR_OK;
    if (!A.mvp || (A.V > 0 && !A.v_pos) || (A.F > 0 && !A.tri)) return WR_ERR_INVALID_ARGUMENT;
    if (A.out_depth && !A.w2c) return WR_ERR_INVALID_ARGUMENT;
    if (A.out_normal && !A.v_nrm) return WR_ERR_INVALID_ARGUMENT;
    if (A.out_attr && (!A.v_tex || !A.tri_tex || !A.texture || A.TH <= 0 || A.TW <= 0 || A.TC <= 0)) return WR_ERR_INVALID_ARGUMENT;
    if (A.depth_mode < WR_DEPTH_NONE || A.depth_mode > WR_DEPTH_SIMPLE) return WR_ERR_INVALID_ARGUMENT;
    if (A.tex_filter < 0 || A.tex_filter > 1) return WR_ERR_UNSUPPORTED;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) return wr_set_cuda_error(ctx, e, "cudaSetDevice");

    VtxSrc src;
    src.pos = A.v_pos; src.mvp = A.mvp; src.V = A.V; src.batched = 0;
    RasterResult res;
    int rc = wr_run_raster(ctx, src, A.B, A.tri, A.F, nullptr, A.H, A.W, 0, &res, nullptr, stream);
    if (rc != WR_OK) return rc;

    ShadeParams P;
    P.a = A;
    P.packed = res.packed;
    P.range = reinterpret_cast<uint32_t *>(res.view_stats);
    const bool two_pass = A.out_depth && A.depth_mode != WR_DEPTH_SIMPLE;
    if (two_pass) {
        k_init_range<<<wr_div_up(A.B, 64), 64, 0, stream>>>(P.range, A.B);
        WR_CHECK_LAUNCH(ctx, "k_init_range");
    }
    k_shade<<<dim3(wr_div_up(A.W, 256), A.H, A.B), 256, 0, stream>>>(P);
    WR_CHECK_LAUNCH(ctx, "k_shade");
    if (two_pass) {
        const long long npv = (long long)A.H * A.W;
        k_depth_finalize<<<dim3(wr_div_up(npv, 256), A.B), 256, 0, stream>>>(A.out_depth, A.out_mask, res.packed, P.range,
                                                                             npv, A.depth_mode, A.depth_p0, A.depth_p1,
                                                                             A.depth_bg);
        WR_CHECK_LAUNCH(ctx, "k_depth_finalize");
    }
    return WR_OK;
}
