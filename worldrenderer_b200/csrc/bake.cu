// Texture unprojection (UV bake).  Reference: uv.py:72-184 (uv_render_geometry), uv.py:193-222
// (uv_render_attr), uv.py:248-348 (validity + exponential blend), uv.py:385-468 (uv_blend, non-Poisson
// branch), driven by projection.py:54-204.  The reference materialises ~10 [Nv,Huv,Wuv,*] tensors; here
// one kernel walks the views per texel and keeps everything in registers.
//
//   k_view_aoi_sobel   per view pixel: camera-space normal -> aoi_cos, zero-padded Sobel magnitude
//   k_dilate_pack      per view pixel: d x d max-pool of the gradient; packs two float4 maps
//                      geo = (pos.xyz, aoi_cos), attr = (rgb, depth_grad) so that a bilinear sample is
//                      4 taps x 2 x 16-byte loads instead of 4 taps x (12 + 4 + 4 + 12) bytes
//   k_uv_unproject     per texel: project into every view, gather, validity, weight, accumulate
//   k_uv_finalize      stitch with the existing texture (after the optional all-reduce)
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256) k_view_aoi_sobel(const float *normal, const uint8_t *mask, const float *depth,
                                                        const float *w2c, int H, int W, int want_grad, float *aoi_out,
                                                        float *g_out)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y, b = blockIdx.z;
    if (c >= W) return;
    const size_t o = ((size_t)b * H + r) * W + c;
    const float *n = normal + 3 * o;
    const float nx = n[0], ny = n[1], nz = n[2];
    float aoi;
    if (mask[o]) {
        const float *R = w2c + 16 * b;
        float x = (R[0] * nx + R[1] * ny) + R[2] * nz;
        float y = (R[4] * nx + R[5] * ny) + R[6] * nz;
        float z = (R[8] * nx + R[9] * ny) + R[10] * nz;
        const float ln = sqrtf((x * x + y * y) + z * z);
        z = z / fmaxf(ln, 1e-12f);
        aoi = z;
    } else {
        aoi = nz;  // uv.py:112: background keeps the render's normal (normal_background)
    }
    aoi = fminf(fmaxf(aoi, 0.0f), 1.0f);
    aoi_out[o] = aoi;
    if (want_grad) {
        const float *d = depth + (size_t)b * H * W;
        auto at = [&](int rr, int cc) -> float {
            return (rr >= 0 && rr < H && cc >= 0 && cc < W) ? __ldg(d + (size_t)rr * W + cc) : 0.0f;
        };
        const float s00 = at(r - 1, c - 1), s01 = at(r - 1, c), s02 = at(r - 1, c + 1);
        const float s10 = at(r, c - 1), s12 = at(r, c + 1);
        const float s20 = at(r + 1, c - 1), s21 = at(r + 1, c), s22 = at(r + 1, c + 1);
        const float gx = ((((s00 - s02) + 2.0f * s10) - 2.0f * s12) + s20) - s22;
        const float gy = ((((s00 + 2.0f * s01) + s02) - s20) - 2.0f * s21) - s22;
        g_out[o] = sqrtf(gx * gx + gy * gy);
    }
}

__global__ void __launch_bounds__(256) k_dilate_pack(const float *aoi, const float *g, const float *position,
                                                     const float *images, int H, int W, int dilation, float *depth_grad,
                                                     float *geo_map, float *attr_map)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y, b = blockIdx.z;
    if (c >= W) return;
    const size_t o = ((size_t)b * H + r) * W + c;
    float dg = 0.0f;
    if (dilation > 0) {
        const int pad = dilation / 2;
        const float *gv = g + (size_t)b * H * W;
        dg = -INFINITY;
        for (int dy = -pad; dy <= pad; ++dy) {
            const int rr = r + dy;
            if (rr < 0 || rr >= H) continue;
            for (int dx = -pad; dx <= pad; ++dx) {
                const int cc = c + dx;
                if (cc < 0 || cc >= W) continue;
                dg = fmaxf(dg, __ldg(gv + (size_t)rr * W + cc));
            }
        }
        if (depth_grad) depth_grad[o] = dg;
    }
    if (geo_map) {
        const float *p = position + 3 * o;
        reinterpret_cast<float4 *>(geo_map)[o] = make_float4(p[0], p[1], p[2], aoi[o]);
    }
    if (attr_map) {
        float rr = 0.f, gg = 0.f, bb = 0.f;
        if (images) { const float *im = images + 3 * o; rr = im[0]; gg = im[1]; bb = im[2]; }
        reinterpret_cast<float4 *>(attr_map)[o] = make_float4(rr, gg, bb, dg);
    }
}

// F.grid_sample(mode="bilinear", padding_mode="zeros", align_corners=False) tap set (uv.py:143-169)
struct Bilinear {
    int x0, y0;
    float w00, w10, w01, w11;
    bool ok;
};

__device__ __forceinline__ Bilinear make_bilinear(float gx, float gy, int W, int H)
{
    Bilinear t;
    const float ix = ((gx + 1.0f) * (float)W - 1.0f) / 2.0f;
    const float iy = ((gy + 1.0f) * (float)H - 1.0f) / 2.0f;
    t.ok = isfinite(ix) && isfinite(iy);
    const float x0f = floorf(ix), y0f = floorf(iy);
    const float tx = ix - x0f, ty = iy - y0f;
    t.x0 = t.ok ? (int)fminf(fmaxf(x0f, -1e9f), 1e9f) : -2;
    t.y0 = t.ok ? (int)fminf(fmaxf(y0f, -1e9f), 1e9f) : -2;
    t.w00 = (1.0f - tx) * (1.0f - ty);
    t.w10 = tx * (1.0f - ty);
    t.w01 = (1.0f - tx) * ty;
    t.w11 = tx * ty;
    return t;
}

__device__ __forceinline__ float4 sample4(const float4 *map, const Bilinear &t, int W, int H)
{
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!t.ok) return acc;
    const int xs[4] = { t.x0, t.x0 + 1, t.x0, t.x0 + 1 };
    const int ys[4] = { t.y0, t.y0, t.y0 + 1, t.y0 + 1 };
    const float ws[4] = { t.w00, t.w10, t.w01, t.w11 };
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (xs[k] >= 0 && xs[k] < W && ys[k] >= 0 && ys[k] < H) {
            const float4 v = __ldg(map + (size_t)ys[k] * W + xs[k]);
            acc.x = acc.x + v.x * ws[k];
            acc.y = acc.y + v.y * ws[k];
            acc.z = acc.z + v.z * ws[k];
            acc.w = acc.w + v.w * ws[k];
        }
    }
    return acc;
}

__device__ __forceinline__ float sample1(const float *map, const Bilinear &t, int W, int H)
{
    float acc = 0.f;
    if (!t.ok) return acc;
    const int xs[4] = { t.x0, t.x0 + 1, t.x0, t.x0 + 1 };
    const int ys[4] = { t.y0, t.y0, t.y0 + 1, t.y0 + 1 };
    const float ws[4] = { t.w00, t.w10, t.w01, t.w11 };
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (xs[k] >= 0 && xs[k] < W && ys[k] >= 0 && ys[k] < H) acc = acc + __ldg(map + (size_t)ys[k] * W + xs[k]) * ws[k];
    return acc;
}

__global__ void __launch_bounds__(256) k_uv_unproject(wr_unproject_args A, int materialise)
{
    extern __shared__ float s_cam[];  // [Nv,16] mvp, then [Nv] exponent
    for (int i = threadIdx.x; i < A.Nv * 16; i += blockDim.x) s_cam[i] = A.mvp[i];
    float *s_expo = s_cam + A.Nv * 16;
    for (int i = threadIdx.x; i < A.Nv; i += blockDim.x)
        s_expo[i] = A.view_weight ? A.alpha / A.view_weight[i] : A.alpha;
    __syncthreads();

    const long long ntex = (long long)A.Hu * A.Wu;
    const long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= ntex) return;
    const bool inside = A.uv_mask[o] != 0;
    float sr = 0.f, sg = 0.f, sb = 0.f, sw = 0.f;
    int nvalid = 0;

    if (inside || materialise) {
        const float *up = A.uv_pos + 3 * o;
        const float ux = up[0], uy = up[1], uz = up[2];
        const long long npix = (long long)A.H * A.W;
        bool valid0 = false;
        for (int v = 0; v < A.Nv; ++v) {
            const float *m = s_cam + 16 * v;
            const float cx = ((m[0] * ux + m[1] * uy) + m[2] * uz) + m[3];
            const float cy = ((m[4] * ux + m[5] * uy) + m[6] * uz) + m[7];
            const float cw = ((m[12] * ux + m[13] * uy) + m[14] * uz) + m[15];
            const float gx = cx / cw, gy = cy / cw;  // uv.py:90, no w > 0 guard
            const Bilinear t = make_bilinear(gx, gy, A.W, A.H);
            const float4 geo = sample4(reinterpret_cast<const float4 *>(A.geo_map) + v * npix, t, A.W, A.H);
            float4 att = make_float4(0.f, 0.f, 0.f, 0.f);
            if (A.attr_map) att = sample4(reinterpret_cast<const float4 *>(A.attr_map) + v * npix, t, A.W, A.H);
            const float dx = geo.x - ux, dy = geo.y - uy, dz = geo.z - uz;
            const float err = sqrtf((dx * dx + dy * dy) + dz * dz);
            bool valid = (err < A.pos_error_eps) && (geo.w > A.aoi_cos_thresh);
            if (A.use_depth_grad) valid = valid && (att.w < A.depth_grad_thresh);
            valid = valid && inside;
            float mp = 0.f;
            if (A.view_masks) {
                mp = sample1(A.view_masks + v * npix, t, A.W, A.H);
                valid = valid && (mp > A.mask_thresh);
            }
            if (A.first_view_dominate) {
                if (v == 0) valid0 = valid;
                else valid = valid && !valid0;
            }
            float wgt = geo.w * (valid ? 1.0f : 0.0f);
            wgt = powf(wgt, s_expo[v]);
            sw = sw + wgt;
            sr = sr + att.x * wgt; sg = sg + att.y * wgt; sb = sb + att.z * wgt;
            nvalid += valid ? 1 : 0;
            if (materialise) {
                const size_t ov = (size_t)v * ntex + o;
                if (A.uv_pos_ndc) { A.uv_pos_ndc[2 * ov] = gx; A.uv_pos_ndc[2 * ov + 1] = gy; }
                if (A.uv_pos_proj) { A.uv_pos_proj[3 * ov] = geo.x; A.uv_pos_proj[3 * ov + 1] = geo.y; A.uv_pos_proj[3 * ov + 2] = geo.z; }
                if (A.uv_pos_error) A.uv_pos_error[ov] = err;
                if (A.uv_aoi_cos) A.uv_aoi_cos[ov] = geo.w;
                if (A.uv_depth_grad) A.uv_depth_grad[ov] = att.w;
                if (A.uv_attr_proj) { A.uv_attr_proj[3 * ov] = att.x; A.uv_attr_proj[3 * ov + 1] = att.y; A.uv_attr_proj[3 * ov + 2] = att.z; }
                if (A.uv_mask_proj) A.uv_mask_proj[ov] = mp;
                if (A.uv_valid) A.uv_valid[ov] = valid ? 1 : 0;
                if (A.uv_weight) A.uv_weight[ov] = wgt;
            }
        }
        if (materialise && A.uv_weight) {  // ExponentialBlend "linear": w / clamp(sum w, 1e-5), clamp [0,1]
            const float den = fmaxf(sw, 1e-5f);
            for (int v = 0; v < A.Nv; ++v) {
                const size_t ov = (size_t)v * ntex + o;
                A.uv_weight[ov] = fminf(fmaxf(A.uv_weight[ov] / den, 0.0f), 1.0f);
            }
        }
    }

    if (A.accum) {
        float *a = A.accum + 5 * o;
        if (A.accumulate) { a[0] += sr; a[1] += sg; a[2] += sb; a[3] += sw; a[4] += (float)nvalid; }
        else { a[0] = sr; a[1] = sg; a[2] = sb; a[3] = sw; a[4] = (float)nvalid; }
    }
    if (A.out_attr || A.out_valid_any) {
        const float den = fmaxf(sw, 1e-5f);
        const float va = nvalid > 0 ? 1.0f : 0.0f;
        if (A.out_valid_any) A.out_valid_any[o] = nvalid > 0 ? 1 : 0;
        if (A.out_attr) {
            float o0 = 0.f, o1 = 0.f, o2 = 0.f;
            if (A.old_attr) { o0 = A.old_attr[3 * o]; o1 = A.old_attr[3 * o + 1]; o2 = A.old_attr[3 * o + 2]; }
            A.out_attr[3 * o] = (sr / den) * va + o0 * (1.0f - va);
            A.out_attr[3 * o + 1] = (sg / den) * va + o1 * (1.0f - va);
            A.out_attr[3 * o + 2] = (sb / den) * va + o2 * (1.0f - va);
        }
    }
}

__global__ void __launch_bounds__(256) k_uv_finalize(const float *accum, const float *old_attr, long long ntex,
                                                     float *out_attr, uint8_t *out_valid_any)
{
    const long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= ntex) return;
    const float *a = accum + 5 * o;
    const float den = fmaxf(a[3], 1e-5f);
    const bool any = a[4] > 0.5f;
    const float va = any ? 1.0f : 0.0f;
    if (out_valid_any) out_valid_any[o] = any ? 1 : 0;
    if (out_attr) {
        float o0 = 0.f, o1 = 0.f, o2 = 0.f;
        if (old_attr) { o0 = old_attr[3 * o]; o1 = old_attr[3 * o + 1]; o2 = old_attr[3 * o + 2]; }
        out_attr[3 * o] = (a[0] / den) * va + o0 * (1.0f - va);
        out_attr[3 * o + 1] = (a[1] / den) * va + o1 * (1.0f - va);
        out_attr[3 * o + 2] = (a[2] / den) * va + o2 * (1.0f - va);
    }
}


// Generic F.grid_sample(bilinear, zeros, align_corners=False) for channels-last maps (uv.py:200-218).
__global__ void __launch_bounds__(256) k_grid_sample(const float *map, int H, int W, int C, const float *ndc,
                                                     long long nsamp_view, long long nsamp_total, float *out)
{
    const long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= nsamp_total) return;
    const int b = (int)(o / nsamp_view);
    const float2 g = __ldg(reinterpret_cast<const float2 *>(ndc) + o);
    const Bilinear t = make_bilinear(g.x, g.y, W, H);
    const float *mv = map + (size_t)b * H * W * C;
    float *dst = out + o * C;
    const int xs[4] = { t.x0, t.x0 + 1, t.x0, t.x0 + 1 };
    const int ys[4] = { t.y0, t.y0, t.y0 + 1, t.y0 + 1 };
    const float ws[4] = { t.w00, t.w10, t.w01, t.w11 };
    for (int ch = 0; ch < C; ++ch) {
        float acc = 0.f;
        if (t.ok) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (xs[k] >= 0 && xs[k] < W && ys[k] >= 0 && ys[k] < H)
                    acc = acc + __ldg(mv + ((size_t)ys[k] * W + xs[k]) * C + ch) * ws[k];
        }
        dst[ch] = acc;
    }
}

}  // namespace

extern "C" int wr_grid_sample(wr_ctx *ctx, const float *map, int B, int H, int W, int C, const float *ndc, int Hs,
                              int Ws, float *out, void *stream_)
{
    if (!ctx || B < 0 || H <= 0 || W <= 0 || C <= 0 || Hs <= 0 || Ws <= 0) return WR_ERR_INVALID_ARGUMENT;
    if (B == 0) return WR_OK;
    if (!map || !ndc || !out) return WR_ERR_INVALID_ARGUMENT;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) return wr_set_cuda_error(ctx, e, "cudaSetDevice");
    const long long nv = (long long)Hs * Ws, total = nv * B;
    k_grid_sample<<<wr_div_up(total, 256), 256, 0, stream>>>(map, H, W, C, ndc, nv, total, out);
    WR_CHECK_LAUNCH(ctx, "k_grid_sample");
    return WR_OK;
}

extern "C" int wr_view_prep(wr_ctx *ctx, const float *normal, const uint8_t *mask, const float *depth,
                            const float *position, const float *w2c, const float *images, int B, int H, int W,
                            int dilation, float *aoi_cos, float *depth_grad, float *geo_map, float *attr_map,
                            void *stream_)
{
    if (!ctx || B < 0 || H <= 0 || W <= 0 || dilation < 0) return WR_ERR_INVALID_ARGUMENT;
    if (dilation > 0 && (dilation & 1) == 0) return WR_ERR_UNSUPPORTED;  // even max-pool changes the map size
    if (B == 0) return WR_OK;
    if (!normal || !mask || !w2c || (dilation > 0 && !depth) || (geo_map && !position)) return WR_ERR_INVALID_ARGUMENT;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) return wr_set_cuda_error(ctx, e, "cudaSetDevice");
    const size_t n = (size_t)B * H * W;
    float *aoi_buf = aoi_cos, *g_buf = nullptr;
    size_t need = 0;
    if (!aoi_buf) need += wr_align256(n * sizeof(float));
    if (dilation > 0) need += wr_align256(n * sizeof(float));
    if (need) {
        int rc = wr_scratch_reserve(ctx, need, stream);
        if (rc != WR_OK) return rc;
        ctx->clean_bytes = 0;  // the temporaries overwrite the raster's packed buffer
        char *p = static_cast<char *>(ctx->scratch);
        if (!aoi_buf) { aoi_buf = reinterpret_cast<float *>(p); p += wr_align256(n * sizeof(float)); }
        if (dilation > 0) g_buf = reinterpret_cast<float *>(p);
    }
    const dim3 grid(wr_div_up(W, 256), H, B);
    // dilation 1 is the identity max-pool: write the gradient straight through the second kernel as well
    wr_stage_begin(ctx);
    wr_stage(ctx, stream, "k_view_aoi_sobel");
    k_view_aoi_sobel<<<grid, 256, 0, stream>>>(normal, mask, depth, w2c, H, W, dilation > 0, aoi_buf, g_buf);
    WR_CHECK_LAUNCH(ctx, "k_view_aoi_sobel");
    if (depth_grad || geo_map || attr_map) {
        wr_stage(ctx, stream, "k_dilate_pack");
        k_dilate_pack<<<grid, 256, 0, stream>>>(aoi_buf, g_buf, position, images, H, W, dilation, depth_grad, geo_map,
                                               attr_map);
        WR_CHECK_LAUNCH(ctx, "k_dilate_pack");
    }
    wr_stage(ctx, stream, "end");
    return WR_OK;
}

extern "C" int wr_uv_unproject(wr_ctx *ctx, const wr_unproject_args *args, void *stream_)
{
    if (!ctx || !args) return WR_ERR_INVALID_ARGUMENT;
    const wr_unproject_args &A = *args;
    if (A.Hu <= 0 || A.Wu <= 0 || A.Nv < 0 || A.H <= 0 || A.W <= 0) return WR_ERR_INVALID_ARGUMENT;
    if (!A.uv_pos || !A.uv_mask || (A.Nv > 0 && (!A.mvp || !A.geo_map))) return WR_ERR_INVALID_ARGUMENT;
    if (A.use_depth_grad && !A.attr_map) return WR_ERR_INVALID_ARGUMENT;
    const size_t smem = (size_t)A.Nv * 17 * sizeof(float);
    if (smem > 48 * 1024) return WR_ERR_UNSUPPORTED;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) return wr_set_cuda_error(ctx, e, "cudaSetDevice");
    const int materialise = (A.uv_pos_ndc || A.uv_pos_proj || A.uv_pos_error || A.uv_aoi_cos || A.uv_depth_grad ||
                             A.uv_attr_proj || A.uv_mask_proj || A.uv_valid || A.uv_weight) ? 1 : 0;
    const long long ntex = (long long)A.Hu * A.Wu;
    wr_stage_begin(ctx);
    wr_stage(ctx, stream, "k_uv_unproject");
    k_uv_unproject<<<wr_div_up(ntex, 256), 256, smem, stream>>>(A, materialise);
    WR_CHECK_LAUNCH(ctx, "k_uv_unproject");
    wr_stage(ctx, stream, "end");
    return WR_OK;
}

extern "C" int wr_uv_finalize(wr_ctx *ctx, const float *accum, const float *old_attr, int Hu, int Wu, float *out_attr,
                              uint8_t *out_valid_any, void *stream_)
{
    if (!ctx || !accum || Hu <= 0 || Wu <= 0) return WR_ERR_INVALID_ARGUMENT;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) return wr_set_cuda_error(ctx, e, "cudaSetDevice");
    const long long ntex = (long long)Hu * Wu;
    k_uv_finalize<<<wr_div_up(ntex, 256), 256, 0, stream>>>(accum, old_attr, ntex, out_attr, out_valid_any);
    WR_CHECK_LAUNCH(ctx, "k_uv_finalize");
    return WR_OK;
}
