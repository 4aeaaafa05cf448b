// Shared declarations of libwr_b200 (sm_100a).  The raster contract implemented here is
// DESIGN.md section 3; every float expression on the coverage / depth-key path is written as a
// sequence of individually rounded binary32 operations and this library is compiled with
// -fmad=false, so coverage and triangle ids cannot depend on FMA contraction.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "../../include/wr_b200.h"

#define WR_COORD_LIMIT 4194304.0f  // 2^22 sub-pixel units
#define WR_GUARD_BAND 16.0f
#define WR_EMPTY_PIXEL 0xFFFFFFFFFFFFFFFFull

// flags of a snapped vertex
#define WR_SV_OK 1u        // w > 0 and snapped coordinates representable
#define WR_SV_FINITE 2u    // all four clip coordinates finite
#define WR_SV_OC_SHIFT 8   // six outcode bits: x<-w, x>w, y<-w, y>w, z<-w, z>w

#define WR_QUEUE_SLOW 0x80000000u  // queue entry flag: triangle needs geometric clipping

struct __align__(16) SnapVert {
    int x, y;        // 1/16 pixel units relative to the sample of pixel (0, 0): round(ndc * 8 * size) + 8 * size - 8,
                     // so that pixel (c, r) samples (16 c, 16 r) and its column is x >> 4 without an offset
    float zw;        // z/w
    uint32_t flags;
};

// Where clip-space vertices come from: a clip-space tensor (dr.rasterize) or world positions
// plus a per-view mvp (fused render, restating utils.py:127-129 in a fixed operation order).
struct VtxSrc {
    const float *pos;   // clip [B,V,4] / [V,4], or world [V,3]
    const float *mvp;   // [B,16] row major, or nullptr when pos is already clip space
    int V;
    int batched;        // clip mode: pos has a leading view dimension
};

__device__ __forceinline__ float4 wr_load_clip(const VtxSrc &s, int b, int v)
{
    if (s.mvp) {
        const float *p = s.pos + 3 * (size_t)v;
        const float x = __ldg(p), y = __ldg(p + 1), z = __ldg(p + 2);
        const float *m = s.mvp + 16 * b;
        float4 c;
        c.x = ((m[0] * x + m[1] * y) + m[2] * z) + m[3];
        c.y = ((m[4] * x + m[5] * y) + m[6] * z) + m[7];
        c.z = ((m[8] * x + m[9] * y) + m[10] * z) + m[11];
        c.w = ((m[12] * x + m[13] * y) + m[14] * z) + m[15];
        return c;
    }
    const float4 *p = reinterpret_cast<const float4 *>(s.pos) + (s.batched ? (size_t)b * s.V : 0) + v;
    return __ldg(p);
}

__device__ __forceinline__ uint32_t wr_depth_key(float zw)
{
    uint32_t u = __float_as_uint(zw);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// order-preserving float <-> uint map for atomicMin / atomicMax on floats
__device__ __forceinline__ uint32_t wr_float_ordered(float f) { return wr_depth_key(f); }
__device__ __forceinline__ float wr_ordered_float(uint32_t k)
{
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
}

#define WR_MAX_STAGES 16

struct wr_ctx {
    int device;
    int sm_count;
    void *scratch;
    size_t scratch_bytes;
    char last_error[256];
    size_t clean_bytes;  // leading bytes of scratch (the packed depth/id buffer) known to be all 0xFF
    // optional per-stage timing (bench.py): events recorded on the launch stream between kernels
    int profiling;
    int n_marks;
    cudaEvent_t marks[WR_MAX_STAGES + 1];
    const char *mark_names[WR_MAX_STAGES + 1];
};

// Marks the START of stage `name` (and the end of the previous one) on `stream` when profiling is on.
static inline void wr_stage(wr_ctx *ctx, cudaStream_t stream, const char *name)
{
    if (!ctx->profiling || ctx->n_marks > WR_MAX_STAGES) return;
    ctx->mark_names[ctx->n_marks] = name;
    cudaEventRecord(ctx->marks[ctx->n_marks], stream);
    ctx->n_marks++;
}
static inline void wr_stage_begin(wr_ctx *ctx) { ctx->n_marks = 0; }

// result of the coverage / visibility stages: packed (depth_key << 32 | triangle id) per pixel
struct RasterResult {
    unsigned long long *packed;  // [B,H,W]
    size_t packed_bytes;
    int32_t *view_stats;         // [B,4] zero-initialised ints the caller may use
    int filled;                  // the FillJob handed to wr_run_raster has been written by the set-up kernel
};

// The kernel that reads `packed` resets every pixel it consumes to WR_EMPTY_PIXEL; after its launch the
// buffer is known clean again and the next call can skip the clear.
static inline void wr_raster_consumed(wr_ctx *ctx, const RasterResult *res) { ctx->clean_bytes = res->packed_bytes; }

// Optional by-product of the vertex pass of a fused render: 16-byte records of the world positions and the
// normals, so that the shading kernel gathers one LDG.128 per vertex attribute instead of three LDG.32.
struct VertexPack {
    const float *v_nrm;   // in: [Vn,3] or nullptr
    int Vn;               // in
    size_t offset;        // in: byte offset of the records inside the extra scratch (256-byte aligned)
    float4 *pos4;         // out: [V]  (x, y, z, 0)
    float4 *nrm4;         // out: [Vn] (x, y, z, 0), nullptr without normals
};
static inline size_t wr_vertex_pack_bytes(int V, int Vn, bool normals)
{
    return (((size_t)(V > 0 ? V : 1) * 16 + 255) & ~(size_t)255) + (normals ? (size_t)(Vn > 0 ? Vn : 1) * 16 : 0);
}

// Background prefill.  The shading pass of a render is bound by its output writes (33 B per pixel, most of them
// background constants), the triangle set-up pass before it by instruction issue with the memory system idle.  A
// FillJob describes the constant background of every output map as up to five segments of 16-byte words with a
// 3-word period (a [.., 3] f32 map repeats every 3 words); the set-up kernel's blocks each write one contiguous
// share of it before they start their own work, so the background goes out to HBM UNDER the set-up pass, and the
// shading kernel then stores covered pixels only.
#define WR_FILL_MAX_SEGS 5
struct FillSeg {
    uint4 *ptr;                  // 16-byte aligned
    unsigned n16;                // 16-byte words (< 2^32)
    unsigned chunk;              // words per participating block (set by wr_fill_plan)
    int uniform;                 // pat[0] == pat[1] == pat[2]
    uint4 pat[3];                // word w of the segment holds pat[w % 3]
};
struct FillJob {
    int nseg;
    unsigned long long total16;
    unsigned stride;             // block b participates iff b % stride == 0, as share b / stride (set by wr_fill_plan)
    unsigned shares;
    FillSeg seg[WR_FILL_MAX_SEGS];
};

// Splits the job over the blocks of the kernel that will carry it: every `stride`-th block writes a share of at
// least ~2048 words (8 per thread of a 256-thread block), so the other blocks pay one compare.
static inline void wr_fill_plan(FillJob *J, unsigned nblocks)
{
    if (J->nseg == 0 || nblocks == 0) return;
    unsigned long long want = J->total16 / 2048;
    if (want < 1) want = 1;
    J->stride = want >= nblocks ? 1u : (unsigned)(nblocks / want);
    J->shares = (nblocks + J->stride - 1) / J->stride;
    for (int s = 0; s < J->nseg; ++s) J->seg[s].chunk = (J->seg[s].n16 + J->shares - 1) / J->shares;
}

// Build switch of the background prefill: measured slower than letting the shading pass write the background
// (DESIGN.md 5b), so it is off and the set-up kernels do not even test for a job (3 % of k_setup_mv's instructions).
#ifndef WR_PREFILL
#define WR_PREFILL 0
#endif

#ifdef __CUDACC__
__device__ __forceinline__ void wr_fill_share(const FillJob &J, unsigned bid)
{
    if (!WR_PREFILL) return;
    if (J.nseg == 0) return;
    if (bid % J.stride != 0) return;
    const unsigned share = bid / J.stride;
#pragma unroll 1
    for (int s = 0; s < J.nseg; ++s) {
        const unsigned chunk = J.seg[s].chunk, n16 = J.seg[s].n16;
        const unsigned long long lo64 = (unsigned long long)share * chunk;
        if (lo64 >= n16) continue;
        const unsigned lo = (unsigned)lo64;
        const unsigned hi = (n16 - lo < chunk) ? n16 : lo + chunk;
        uint4 *dst = J.seg[s].ptr;
        const uint4 p0 = J.seg[s].pat[0];
        if (J.seg[s].uniform) {
            for (unsigned w = lo + threadIdx.x; w < hi; w += blockDim.x) __stcs(dst + w, p0);
        } else {
            const uint4 p1 = J.seg[s].pat[1], p2 = J.seg[s].pat[2];
            unsigned w = lo + threadIdx.x;
            unsigned m = w % 3u;
            const unsigned step = blockDim.x % 3u;
            for (; w < hi; w += blockDim.x) {
                __stcs(dst + w, m == 0 ? p0 : (m == 1 ? p1 : p2));
                m += step;
                if (m >= 3u) m -= 3u;
            }
        }
    }
}
#endif

int wr_scratch_reserve(wr_ctx *ctx, size_t bytes, cudaStream_t stream);
int wr_set_cuda_error(wr_ctx *ctx, cudaError_t e, const char *where);
int wr_run_raster(wr_ctx *ctx, const VtxSrc &src, int B, const int32_t *tri, int F, const int32_t *tri_ranges,
                  int H, int W, size_t extra_bytes, RasterResult *res, void **extra, cudaStream_t stream,
                  VertexPack *pack = nullptr, const FillJob *fill = nullptr);

// Programmatic dependent launch (sm_90+): a kernel launched with the attribute may be scheduled while its
// predecessor on the stream is still draining; it must execute wr_pdl_wait() before touching anything the
// predecessor wrote (the wait also covers the predecessor's predecessors, every kernel of the chain waits
// first thing).  wr_pdl_trigger() in the predecessor allows the dependent's blocks to be placed as soon as all
// of the predecessor's blocks have been issued.  Used to hide the launch latency between the five short kernels
// of a render step; disabled while per-stage events are being recorded.
#ifndef WR_PDL
#define WR_PDL 1
#endif
#ifdef __CUDACC__
__device__ __forceinline__ void wr_pdl_wait()
{
#if WR_PDL
    asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}
__device__ __forceinline__ void wr_pdl_trigger()
{
#if WR_PDL
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
#endif

// 1.0f / (float)a for an integer a != 0 (the doubled triangle area): |(float)a| lies in [1, 2^31], where IEEE division
// never leaves the compiler's in-line path -- MUFU.RCP, one Newton step in two FFMAs, correctly rounded.  Written out
// here, the sequence is the same minus the exponent-range test and the out-of-line call that guard it (6 of 11
// instructions per covered sample of the set-up pass).  Same bits as `1.0f / x` (and as the CPU oracle's division).
#ifndef WR_RCP_FAST
#define WR_RCP_FAST 1
#endif
#ifdef __CUDACC__
__device__ __forceinline__ float wr_rcp_int(int a)
{
    const float x = __int2float_rn(a);
#if WR_RCP_FAST
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    const float e = __fmaf_rn(x, r, -1.0f);
    return __fmaf_rn(r, -e, r);
#else
    return 1.0f / x;
#endif
}
#endif

template <typename... KArgs, typename... Args>
static inline void wr_launch_s(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                               bool dependent, Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (WR_PDL && dependent) ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);  // errors surface through WR_CHECK_LAUNCH
}
template <typename... KArgs, typename... Args>
static inline void wr_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, cudaStream_t stream, bool dependent,
                             Args... args)
{
    wr_launch_s(kernel, grid, block, 0, stream, dependent, args...);
}

#define WR_CHECK_LAUNCH(ctx, where)                                  \
    do {                                                             \
        cudaError_t e__ = cudaGetLastError();                        \
        if (e__ != cudaSuccess) return wr_set_cuda_error(ctx, e__, where); \
    } while (0)

static inline size_t wr_align256(size_t x) { return (x + 255) & ~(size_t)255; }
static inline int wr_div_up(long long a, long long b) { return (int)((a + b - 1) / b); }

#ifdef __CUDACC__
// IEEE sqrtf / division with the zero operand peeled off.  The correctly rounded sequences the compiler emits check
// their operands and leave through an out-of-line slow path (a call of ~30 instructions, taken by the whole warp when
// one lane needs it) for zeros and subnormals -- and exact zeros are the COMMON case in a bake: the Sobel gradient of a
// flat depth region, a texel outside every chart (position 0 -> clip x = 0), a texel no view is valid for (sum 0).
// Measured with ncu on config C: every warp of the unprojection took that path twice per view.
// sqrt(+-0) = +-0 and +-0 / d = +-0 for d > 0, so returning the operand is bit-exact; NaN goes to the operation.
__device__ __forceinline__ float wr_sqrt_z(float s)
{
    float r = s;
    if (s != 0.0f) r = sqrtf(s);
    return r;
}
__device__ __forceinline__ float wr_div_zpos(float x, float d)   // d > 0 required
{
    float r = x;
    if (x != 0.0f) r = x / d;
    return r;
}
#endif
