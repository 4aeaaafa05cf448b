// Per-pixel operators: dr.interpolate, dr.texture, and the vertex-normal kernels.
// Reference: render.py:64-120 (operators), mesh.py:85-119 (normals).  The fused render() pass is render.cu.
// Operation order of every expression follows DESIGN.md section 3.4-3.6 / 4 (same as oracle/).
#include "common.cuh"
#include "texture.cuh"

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

