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
//
// The reference splats the face normals with scatter_add_ (float atomics on the GPU: the sum order, and with it the
// last bits of every normal, changes from run to run and from process to process).  Here the splat is EXACT and
// therefore order independent: every face normal component -- computed in float exactly as there -- is converted
// to 64-bit fixed point and added with integer atomics; the per-vertex sum is then rounded to float once.  Runs,
// ranks and GPUs all see the same normals (a multi-GPU bake is then reproducible texel for texel: a validity test
// on a normal-derived cosine can no longer flip between ranks), and each component is the correctly rounded sum.
// Scale: 2^(50 - e) with 2^e >= 8 M^2 >= any face normal component (M = largest |coordinate|): 12 bits of headroom
// for the faces around a vertex, 50 bits below the largest possible component.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_vertex_extent(const float *v_pos, long long n3, unsigned *extent_bits)
{
    float m = 0.0f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n3; i += (long long)gridDim.x * blockDim.x) {
        const float a = fabsf(__ldg(v_pos + i));
        if (a <= 3.402823466e38f) m = fmaxf(m, a);   // finite values only
    }
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xFFFFFFFFu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.0f) atomicMax(extent_bits, __float_as_uint(m));   // non-negative floats order like their bits
}

__device__ __forceinline__ int normal_scale_exp(unsigned extent_bits)
{
    int e;
    frexpf(__uint_as_float(extent_bits), &e);   // M < 2^e
    return 50 - (2 * e + 3);                    // 8 M^2 < 2^(2e + 3)
}

__global__ void __launch_bounds__(256) k_face_normals_scatter(const float *v_pos, int V, const int32_t *tri, int F,
                                                              long long *acc, const unsigned *extent_bits)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= F) return;
    const int i0 = __ldg(tri + 3 * (size_t)t), i1 = __ldg(tri + 3 * (size_t)t + 1), i2 = __ldg(tri + 3 * (size_t)t + 2);
    if ((unsigned)i0 >= (unsigned)V || (unsigned)i1 >= (unsigned)V || (unsigned)i2 >= (unsigned)V) return;
    const float *p0 = v_pos + 3 * (size_t)i0, *p1 = v_pos + 3 * (size_t)i1, *p2 = v_pos + 3 * (size_t)i2;
    const float ax = __ldg(p1) - __ldg(p0), ay = __ldg(p1 + 1) - __ldg(p0 + 1), az = __ldg(p1 + 2) - __ldg(p0 + 2);
    const float bx = __ldg(p2) - __ldg(p0), by = __ldg(p2 + 1) - __ldg(p0 + 1), bz = __ldg(p2 + 2) - __ldg(p0 + 2);
    const float n[3] = { ay * bz - az * by, az * bx - ax * bz, ax * by - ay * bx };
    const int se = normal_scale_exp(__ldg(extent_bits));
    long long q[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        // a non-finite component (non-finite vertex) poisons the vertices of this face only, as in the reference;
        // it is carried as the most negative value, which the finishing pass turns back into NaN
        const double d = ldexp((double)n[k], se);
        q[k] = (fabs(d) < 9.0e18) ? __double2ll_rn(d) : (long long)0x8000000000000000ull;
    }
    const int idx[3] = { i0, i1, i2 };
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        unsigned long long *d = reinterpret_cast<unsigned long long *>(acc + 3 * (size_t)idx[k]);
        atomicAdd(d, (unsigned long long)q[0]); atomicAdd(d + 1, (unsigned long long)q[1]); atomicAdd(d + 2, (unsigned long long)q[2]);
    }
}

__global__ void __launch_bounds__(256) k_normalize_vertex_normals(const long long *acc, const unsigned *extent_bits, int V,
                                                                  float *v_nrm)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    const int se = normal_scale_exp(__ldg(extent_bits));
    float c[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const long long q = acc[3 * (size_t)v + k];
        // |sum| beyond 2^62 can only come from a poisoned (non-finite) contribution
        c[k] = (q > -(1ll << 62) && q < (1ll << 62)) ? (float)ldexp((double)q, -se) : __int_as_float(0x7FC00000);
    }
    float x = c[0], y = c[1], z = c[2];
    const float sq = (x * x + y * y) + z * z;
    if (!(sq > 1e-20f)) { x = 0.0f; y = 0.0f; z = 1.0f; }
    const float n = sqrtf((x * x + y * y) + z * z);
    const float d = fmaxf(n, 1e-12f);
    v_nrm[3 * (size_t)v] = x / d; v_nrm[3 * (size_t)v + 1] = y / d; v_nrm[3 * (size_t)v + 2] = z / d;
}

// ------------------------------------------------------------------------------------------
// vertex tangents (mesh.py:121-167): per-face tangent from the UV gradients, averaged over the faces of a
// vertex, normalised, made perpendicular to the vertex normal, normalised again.
// acc [V,4]: (sum of face tangents, face count).  Operation order of every expression: oracle/render_oracle.py
// vertex_tangents (only the order of the per-vertex sum is free, as for the normals).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_face_tangents_scatter(const float *v_pos, int V, const int32_t *tri,
                                                               const float *v_tex, int Vt, const int32_t *tri_tex, int F,
                                                               float4 *acc)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= F) return;
    const int i0 = __ldg(tri + 3 * (size_t)t), i1 = __ldg(tri + 3 * (size_t)t + 1), i2 = __ldg(tri + 3 * (size_t)t + 2);
    const int j0 = __ldg(tri_tex + 3 * (size_t)t), j1 = __ldg(tri_tex + 3 * (size_t)t + 1), j2 = __ldg(tri_tex + 3 * (size_t)t + 2);
    if ((unsigned)i0 >= (unsigned)V || (unsigned)i1 >= (unsigned)V || (unsigned)i2 >= (unsigned)V) return;
    if ((unsigned)j0 >= (unsigned)Vt || (unsigned)j1 >= (unsigned)Vt || (unsigned)j2 >= (unsigned)Vt) return;
    const float *p0 = v_pos + 3 * (size_t)i0, *p1 = v_pos + 3 * (size_t)i1, *p2 = v_pos + 3 * (size_t)i2;
    const float *t0 = v_tex + 2 * (size_t)j0, *t1 = v_tex + 2 * (size_t)j1, *t2 = v_tex + 2 * (size_t)j2;
    const float u1 = __ldg(t1) - __ldg(t0), v1 = __ldg(t1 + 1) - __ldg(t0 + 1);
    const float u2 = __ldg(t2) - __ldg(t0), v2 = __ldg(t2 + 1) - __ldg(t0 + 1);
    float den = u1 * v2 - v1 * u2;
    den = den > 0.0f ? fmaxf(den, 1e-6f) : fminf(den, -1e-6f);   // NaN stays NaN, as torch.clamp keeps it
    float tg[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float e1 = __ldg(p1 + k) - __ldg(p0 + k), e2 = __ldg(p2 + k) - __ldg(p0 + k);
        tg[k] = (e1 * v2 - e2 * v1) / den;
    }
    const int idx[3] = { i0, i1, i2 };
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float *d = reinterpret_cast<float *>(acc + idx[k]);
        atomicAdd(d, tg[0]); atomicAdd(d + 1, tg[1]); atomicAdd(d + 2, tg[2]); atomicAdd(d + 3, 1.0f);
    }
}

__device__ __forceinline__ void normalize3(float &x, float &y, float &z)
{
    const float d = fmaxf(sqrtf((x * x + y * y) + z * z), 1e-12f);   // F.normalize: x / max(|x|, eps)
    x = x / d; y = y / d; z = z / d;
}

__global__ void __launch_bounds__(256) k_finish_vertex_tangents(const float4 *acc, const float *v_nrm, int V, float *out)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    const float4 a = acc[v];
    float x = a.x / a.w, y = a.y / a.w, z = a.z / a.w;   // a vertex without a face: 0 / 0, as in the reference
    normalize3(x, y, z);
    const float nx = __ldg(v_nrm + 3 * (size_t)v), ny = __ldg(v_nrm + 3 * (size_t)v + 1), nz = __ldg(v_nrm + 3 * (size_t)v + 2);
    const float d = (x * nx + y * ny) + z * nz;
    x = x - d * nx; y = y - d * ny; z = z - d * nz;
    normalize3(x, y, z);
    out[3 * (size_t)v] = x; out[3 * (size_t)v + 1] = y; out[3 * (size_t)v + 2] = z;
}

// ------------------------------------------------------------------------------------------
// Normal maps of the views -> UV tangent space (mvadapter/test/utils/pipeline_texture.py:358-396): the view's
// normal image is read in the geometry tangent frame of its camera (a per-view axis `gt`, Gram-Schmidt against
// the rendered normal), turned into a world-space normal and then expressed in the (tangent, bitangent, normal)
// frame rendered from the mesh; output in [0, 1].  One thread per pixel.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void cross3(float ax, float ay, float az, float bx, float by, float bz, float &x, float &y, float &z)
{
    x = ay * bz - az * by; y = az * bx - ax * bz; z = ax * by - ay * bx;
}

__global__ void __launch_bounds__(256) k_tangent_space_normals(const float *normal, const float *tangent, const float *image,
                                                               const float *view_axis, long long total, long long npix_view,
                                                               float *out)
{
    const long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= total) return;
    const int b = (int)(o / npix_view);
    const float nx = normal[3 * o], ny = normal[3 * o + 1], nz = normal[3 * o + 2];
    const float tx = tangent[3 * o], ty = tangent[3 * o + 1], tz = tangent[3 * o + 2];
    // UV tangent frame: rows T, B = N x T, N, each normalised
    float bx, by, bz;
    cross3(nx, ny, nz, tx, ty, tz, bx, by, bz);
    float Tx = tx, Ty = ty, Tz = tz, Bx = bx, By = by, Bz = bz, Nx = nx, Ny = ny, Nz = nz;
    normalize3(Tx, Ty, Tz); normalize3(Bx, By, Bz); normalize3(Nx, Ny, Nz);
    // geometry tangent frame: GB = N x gt, GT = GB x N, rows GT, GB, N normalised
    const float gx = view_axis[3 * b], gy = view_axis[3 * b + 1], gz = view_axis[3 * b + 2];
    float hx, hy, hz, ux, uy, uz;
    cross3(nx, ny, nz, gx, gy, gz, hx, hy, hz);
    cross3(hx, hy, hz, nx, ny, nz, ux, uy, uz);
    normalize3(ux, uy, uz); normalize3(hx, hy, hz);
    // world-space normal = sum_j m_j * row_j of the geometry frame
    const float m0 = image[3 * o] * 2.0f - 1.0f, m1 = image[3 * o + 1] * 2.0f - 1.0f, m2 = image[3 * o + 2] * 2.0f - 1.0f;
    float wx = (m0 * ux + m1 * hx) + m2 * Nx, wy = (m0 * uy + m1 * hy) + m2 * Ny, wz = (m0 * uz + m1 * hz) + m2 * Nz;
    normalize3(wx, wy, wz);
    // components in the UV tangent frame
    float rx = (wx * Tx + wy * Ty) + wz * Tz, ry = (wx * Bx + wy * By) + wz * Bz, rz = (wx * Nx + wy * Ny) + wz * Nz;
    normalize3(rx, ry, rz);
    out[3 * o] = fminf(fmaxf(rx * 0.5f + 0.5f, 0.0f), 1.0f);
    out[3 * o + 1] = fminf(fmaxf(ry * 0.5f + 0.5f, 0.0f), 1.0f);
    out[3 * o + 2] = fminf(fmaxf(rz * 0.5f + 0.5f, 0.0f), 1.0f);
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
    // scratch: [V,3] int64 fixed-point sums + the extent word
    const size_t acc_bytes = wr_align256((size_t)V * 3 * sizeof(long long));
    int rc = wr_scratch_reserve(ctx, acc_bytes + 256, stream);
    if (rc != WR_OK) return rc;
    ctx->clean_bytes = 0;
    long long *acc = static_cast<long long *>(ctx->scratch);
    unsigned *extent = reinterpret_cast<unsigned *>(static_cast<char *>(ctx->scratch) + acc_bytes);
    e = cudaMemsetAsync(acc, 0, acc_bytes + 256, stream);
    if (e != cudaSuccess) return wr_set_cuda_error(ctx, e, "memset normals");
    k_vertex_extent<<<min(wr_div_up((long long)V * 3, 256 * 8), ctx->sm_count * 8), 256, 0, stream>>>(v_pos, (long long)V * 3, extent);
    WR_CHECK_LAUNCH(ctx, "k_vertex_extent");
    if (F > 0) {
        k_face_normals_scatter<<<wr_div_up(F, 256), 256, 0, stream>>>(v_pos, V, tri, F, acc, extent);
        WR_CHECK_LAUNCH(ctx, "k_face_normals_scatter");
    }
    k_normalize_vertex_normals<<<wr_div_up(V, 256), 256, 0, stream>>>(acc, extent, V, v_nrm);
    WR_CHECK_LAUNCH(ctx, "k_normalize_vertex_normals");
    return WR_OK;
}

extern "C" int wr_vertex_tangents(wr_ctx *ctx, const float *v_pos, int V, const int32_t *tri, const float *v_tex, int Vt,
                                  const int32_t *tri_tex, int F, const float *v_nrm, float *v_tang, void *stream_)
{
    if (!ctx || V < 0 || F < 0 || Vt < 0) return WR_ERR_INVALID_ARGUMENT;
    if (V == 0) return WR_OK;
    if (!v_pos || !v_nrm || !v_tang || (F > 0 && (!tri || !tri_tex || !v_tex))) return WR_ERR_INVALID_ARGUMENT;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) return wr_set_cuda_error(ctx, e, "cudaSetDevice");
    int rc = wr_scratch_reserve(ctx, (size_t)V * sizeof(float4), stream);
    if (rc != WR_OK) return rc;
    ctx->clean_bytes = 0;  // the accumulators overwrite the head of the scratch
    float4 *acc = static_cast<float4 *>(ctx->scratch);
    e = cudaMemsetAsync(acc, 0, (size_t)V * sizeof(float4), stream);
    if (e != cudaSuccess) return wr_set_cuda_error(ctx, e, "memset tangents");
    if (F > 0) {
        k_face_tangents_scatter<<<wr_div_up(F, 256), 256, 0, stream>>>(v_pos, V, tri, v_tex, Vt, tri_tex, F, acc);
        WR_CHECK_LAUNCH(ctx, "k_face_tangents_scatter");
    }
    k_finish_vertex_tangents<<<wr_div_up(V, 256), 256, 0, stream>>>(acc, v_nrm, V, v_tang);
    WR_CHECK_LAUNCH(ctx, "k_finish_vertex_tangents");
    return WR_OK;
}

extern "C" int wr_tangent_space_normals(wr_ctx *ctx, const float *normal, const float *tangent, const float *image,
                                        const float *view_axis, int B, int H, int W, float *out, void *stream_)
{
    if (!ctx || B < 0 || H <= 0 || W <= 0) return WR_ERR_INVALID_ARGUMENT;
    if (B == 0) return WR_OK;
    if (!normal || !tangent || !image || !view_axis || !out) return WR_ERR_INVALID_ARGUMENT;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) return wr_set_cuda_error(ctx, e, "cudaSetDevice");
    const long long npv = (long long)H * W, total = npv * B;
    k_tangent_space_normals<<<wr_div_up(total, 256), 256, 0, stream>>>(normal, tangent, image, view_axis, total, npv, out);
    WR_CHECK_LAUNCH(ctx, "k_tangent_space_normals");
    return WR_OK;
}
