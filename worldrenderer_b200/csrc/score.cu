// View scoring of the SmartPainter loop (reference smart_paint.py:98-159).
//
// The reference renders ~108 candidate views of the mesh textured with a "score map", derives the angle-of-
// incidence cosine per pixel from the normal map in a handful of full-tensor torch ops (:118-137) and then walks
// the views in Python with two `.sum().item()` host round trips per view (:148-158).  Here the cosine comes
// straight out of the fused shading kernel (wr_render's out_geo = (pos, aoi_cos), same formula as uv.py:108-119)
// and the per-view sums are two small kernels with a fixed summation order; the host reads B numbers once.
//
//   count_b = #{ attr < lo  and  aoi > aoi_min }
//   fsum_b  = sum over { attr > lo and aoi > aoi_min } of max((aoi - attr) - margin, 0)
//   score_b = (count_b + fsum_b) / (H * W)           (formed on the host in double, like the reference's Python)
#include "common.cuh"

namespace {

constexpr int kScoreBlocks = 64;  // partial sums per view

__global__ void __launch_bounds__(256) k_view_score_partial(const float *attr, int C, const float *geo, long long npix,
                                                            float lo, float aoi_min, float margin, int *cnt_part,
                                                            float *sum_part)
{
    wr_pdl_wait();  // dependent launch behind the shading kernel that wrote attr / geo
    const int b = blockIdx.y;
    const float *av = attr + (size_t)b * npix * C;
    const float4 *gv = reinterpret_cast<const float4 *>(geo) + (size_t)b * npix;
    int cnt = 0;
    float fs = 0.0f;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < npix; i += (long long)kScoreBlocks * 256) {
        const float a = __ldg(av + i * C);
        const float aoi = __ldg(gv + i).w;
        if (aoi > aoi_min) {
            if (a < lo) cnt += 1;
            if (a > lo) fs = fs + fmaxf((aoi - a) - margin, 0.0f);
        }
    }
    // fixed-order reduction: butterfly inside the warp, then warp 0 adds the eight warp sums in order
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, o);
        fs = fs + __shfl_xor_sync(0xFFFFFFFFu, fs, o);
    }
    __shared__ int s_cnt[8];
    __shared__ float s_fs[8];
    if ((threadIdx.x & 31) == 0) { s_cnt[threadIdx.x >> 5] = cnt; s_fs[threadIdx.x >> 5] = fs; }
    __syncthreads();
    if (threadIdx.x == 0) {
        int c = 0;
        float f = 0.0f;
        for (int w = 0; w < 8; ++w) { c += s_cnt[w]; f = f + s_fs[w]; }
        cnt_part[b * kScoreBlocks + blockIdx.x] = c;
        sum_part[b * kScoreBlocks + blockIdx.x] = f;
    }
}

__global__ void __launch_bounds__(32) k_view_score_final(const int *cnt_part, const float *sum_part, int B, int *count,
                                                        float *fsum)
{
    wr_pdl_wait();
    const int b = blockIdx.x * 32 + threadIdx.x;
    if (b >= B) return;
    int c = 0;
    float f = 0.0f;
    for (int k = 0; k < kScoreBlocks; ++k) { c += cnt_part[b * kScoreBlocks + k]; f = f + sum_part[b * kScoreBlocks + k]; }
    count[b] = c;
    fsum[b] = f;
}

}  // namespace

extern "C" int wr_view_scores(wr_ctx *ctx, const float *attr, int C, const float *geo, int B, int H, int W, float lo,
                              float aoi_min, float margin, int32_t *count, float *fsum, void *stream_)
{
    if (!ctx || B < 0 || H <= 0 || W <= 0 || C <= 0) return WR_ERR_INVALID_ARGUMENT;
    if (B == 0) return WR_OK;
    if (!attr || !geo || !count || !fsum) return WR_ERR_INVALID_ARGUMENT;
    if (reinterpret_cast<uintptr_t>(geo) & 15u) return WR_ERR_INVALID_ARGUMENT;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) return wr_set_cuda_error(ctx, e, "cudaSetDevice");
    // partial sums live behind the raster buffers' clean prefix, like the blend scratch
    const size_t keep = wr_align256(ctx->clean_bytes);
    const size_t part_bytes = wr_align256((size_t)B * kScoreBlocks * sizeof(int));
    int rc = wr_scratch_reserve(ctx, keep + 2 * part_bytes, stream);
    if (rc != WR_OK) return rc;
    char *base = static_cast<char *>(ctx->scratch) + wr_align256(ctx->clean_bytes);
    int *cnt_part = reinterpret_cast<int *>(base);
    float *sum_part = reinterpret_cast<float *>(base + part_bytes);
    wr_stage_begin(ctx);
    wr_stage(ctx, stream, "k_view_score");
    wr_launch(k_view_score_partial, dim3(kScoreBlocks, B), dim3(256), stream, !ctx->profiling, attr, C, geo,
              (long long)H * W, lo, aoi_min, margin, cnt_part, sum_part);
    WR_CHECK_LAUNCH(ctx, "k_view_score_partial");
    wr_launch(k_view_score_final, dim3(wr_div_up(B, 32)), dim3(32), stream, !ctx->profiling, (const int *)cnt_part,
              (const float *)sum_part, B, (int *)count, fsum);
    WR_CHECK_LAUNCH(ctx, "k_view_score_final");
    wr_stage(ctx, stream, "end");
    return WR_OK;
}
