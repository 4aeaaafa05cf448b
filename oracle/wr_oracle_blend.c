/*
 * wr_oracle_blend.c -- CPU ORACLE (test infrastructure, NOT product code) for the two atlas
 * post-processing steps behind uv_blend (mvadapter/utils/mesh_utils/uv.py:426-461):
 *
 *   wro_poisson_blend  <- PoissonBlendingSolver.__call__   blend.py:214-324
 *                         Jacobi step                       blend.py:60-71 (CUDA), :176-179 (torch)
 *   wro_inpaint        <- uv_padding -> inpaint_cvc         uv.py:373-382, cv_ops.py:11-35
 *
 * Poisson blending: PINNED.  The reference's own solver runs on CPU with its "torch-native"
 * backend (blend.py:172-183); oracle/gen_golden.py records its outputs in
 * tests/golden/poisson.npz and tests/test_oracle_golden.py holds this file to them (1e-5: the
 * conv2d / sum(-1) summation orders of the reference are library-defined, the order below is the
 * left-to-right order of the reference's CUDA kernel, blend.py:69).
 *
 * Seam inpainting: PARITY UNPINNED and SUBSTITUTED.  The reference calls cvcuda.inpaint, a
 * third-party GPU-only operator (CV-CUDA, not vendored, not in this image, no version pin in
 * requirements.txt) and holds no test or golden image for it.  What is kept from the reference
 * is its call contract (cv_ops.py:23-35): float images are quantised with (x*255) truncated to
 * uint8, the mask is "non-zero = fill", the result is uint8 / 255.  The fill itself is the
 * deterministic algorithm stated at wro_inpaint below (nearest known pixel by jump flooding,
 * then an inverse-square-distance average of the known pixels around it), shared bit for bit
 * with csrc/blend.cu.
 *
 * Every float expression is a sequence of individually rounded binary32 operations
 * (-ffp-contract=off here, -fmad=false in the kernels).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

enum { GRAD_SRC = 0, GRAD_MAX = 1, GRAD_AVG = 2 };

/* zero-padded fetch (F.conv2d(..., padding=1), blend.py:245-249) */
static inline float px(const float *img, int H, int W, int C, int r, int c, int ch)
{
    if (r < 0 || r >= H || c < 0 || c >= W) return 0.0f;
    return img[((size_t)r * W + c) * C + ch];
}

/*
 * src, tgt: [H,W,C] f32; mask: [H,W] u8 (non-zero = solve here; the caller has already applied
 * the > 0.5 threshold of blend.py:229-232).  out: [H,W,C] = tgt outside the mask, the clamped
 * Jacobi iterate inside.  num_iters Jacobi sweeps exactly (the "torch-native" backend; the
 * pointer-swapping backends return sweep num_iters & ~1, blend.py:88-100, 166-169 -- the host
 * layer rounds down for them).
 */
int wro_poisson_blend(const float *src, const uint8_t *mask_in, const float *tgt, int H, int W, int C,
                      int num_iters, int grad_mode, float *out, int nthreads)
{
    if (H <= 0 || W <= 0 || C <= 0 || num_iters < 0) return -1;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#else
    (void)nthreads;
#endif
    const size_t n = (size_t)H * W;
    uint8_t *m = (uint8_t *)malloc(n);
    float *B = (float *)malloc(n * C * sizeof(float));
    float *X = (float *)calloc(n * C, sizeof(float));
    float *Y = (float *)calloc(n * C, sizeof(float));
    if (!m || !B || !X || !Y) { free(m); free(B); free(X); free(Y); return -2; }
    /* blend.py:233-236: the image border never belongs to the solve region */
    for (int r = 0; r < H; ++r)
        for (int c = 0; c < W; ++c)
            m[(size_t)r * W + c] = (mask_in[(size_t)r * W + c] != 0) && r > 0 && r < H - 1 && c > 0 && c < W - 1;

    static const int dr[4] = {-1, 1, 0, 0}, dc[4] = {0, 0, -1, 1}; /* up, down, left, right: blend.py:289-297 */
#pragma omp parallel for schedule(static)
    for (int r = 0; r < H; ++r)
        for (int c = 0; c < W; ++c) {
            const size_t p = (size_t)r * W + c;
            for (int ch = 0; ch < C; ++ch) {
                float lap;
                if (grad_mode == GRAD_SRC) { /* blend.py:243-250, kernel [[0,-1,0],[-1,4,-1],[0,-1,0]] */
                    lap = 4.0f * px(src, H, W, C, r, c, ch);
                    for (int k = 0; k < 4; ++k) lap = lap - px(src, H, W, C, r + dr[k], c + dc[k], ch);
                } else { /* blend.py:251-281: four one-sided differences, per direction max-|.| or mean */
                    lap = 0.0f;
                    for (int k = 0; k < 4; ++k) {
                        const float ds = px(src, H, W, C, r, c, ch) - px(src, H, W, C, r + dr[k], c + dc[k], ch);
                        const float dt = px(tgt, H, W, C, r, c, ch) - px(tgt, H, W, C, r + dr[k], c + dc[k], ch);
                        const float pick = (grad_mode == GRAD_MAX) ? (fabsf(ds) > fabsf(dt) ? ds : dt) : (ds + dt) * 0.5f;
                        lap = (k == 0) ? pick : lap + pick;
                    }
                }
                /* blend.py:283-287: sum of the four neighbours of the target that lie outside the region */
                float fq = 0.0f;
                for (int k = 0; k < 4; ++k) {
                    const int rr = r + dr[k], cc = c + dc[k];
                    float v = 0.0f;
                    if (rr >= 0 && rr < H && cc >= 0 && cc < W && !m[(size_t)rr * W + cc]) v = tgt[((size_t)rr * W + cc) * C + ch];
                    fq = (k == 0) ? v : fq + v;
                }
                B[p * C + ch] = lap + fq;                     /* blend.py:299 */
                X[p * C + ch] = m[p] ? tgt[p * C + ch] : 0.0f; /* blend.py:298; index 0 of the reference's X is the zero every outside neighbour maps to */
            }
        }

    for (int it = 0; it < num_iters; ++it) {
#pragma omp parallel for schedule(static)
        for (int r = 1; r < H - 1; ++r)
            for (int c = 1; c < W - 1; ++c) {
                const size_t p = (size_t)r * W + c;
                if (!m[p]) continue;
                for (int ch = 0; ch < C; ++ch) {
                    const float up = X[(p - W) * C + ch], dn = X[(p + W) * C + ch];
                    const float lf = X[(p - 1) * C + ch], rt = X[(p + 1) * C + ch];
                    Y[p * C + ch] = ((((up + dn) + lf) + rt) + B[p * C + ch]) * 0.25f; /* blend.py:69 */
                }
            }
        float *t = X; X = Y; Y = t;
    }
#pragma omp parallel for schedule(static)
    for (size_t p = 0; p < n; ++p)
        for (int ch = 0; ch < C; ++ch) {
            float v = tgt[p * C + ch];
            if (m[p]) { /* blend.py:317-321 */
                v = X[p * C + ch];
                v = v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v);
            }
            out[p * C + ch] = v;
        }
    free(m); free(B); free(X); free(Y);
    return 0;
}

/* ---------------------------------------------------------------------------------------------
 * Seam inpainting (substitute for cvcuda.inpaint, see the header).
 *
 *   known(p)  = mask[p] == 0.  If no pixel is known the image is returned unchanged.
 *   seed(p)   = a nearest known pixel of p found by jump flooding: seed0(p) = p if known else NONE;
 *               for step = 2^(ceil(log2(max(H,W))) - 1), ..., 2, 1 and then once more with step 1:
 *                 best = seed_prev(p); for dy in (-step, 0, step), dx in (-step, 0, step) (that order):
 *                   s = seed_prev(p + (dx,dy)) when in bounds and not NONE;
 *                   s replaces best if best is NONE, or |p-s|^2 < |p-best|^2, or equal and s < best
 *                   (seeds compared as y*W + x).
 *   out(p)    = img(p) for known p, otherwise with q = seed(p), D = |p-q|^2 (integer):
 *                 over the known pixels t with |t-q|^2 <= radius^2, rows then columns ascending:
 *                   w = float(1 + D) / float(1 + |p-t|^2);  acc_c += w * float(img_c(t));  ws += w
 *                 out_c = (uint8) min(255, rint(acc_c / ws))       (round half to even)
 * ------------------------------------------------------------------------------------------- */
#define SEED_NONE (-1)

static inline int64_t dist2(int W, int p_r, int p_c, int32_t s)
{
    const int64_t dy = p_r - s / W, dx = p_c - s % W;
    return dy * dy + dx * dx;
}

int wro_inpaint(const uint8_t *img, const uint8_t *mask, int H, int W, int C, int radius, uint8_t *out, int nthreads)
{
    if (H <= 0 || W <= 0 || C <= 0 || C > 4 || radius < 0 || (int64_t)H * W >= (1ll << 31)) return -1;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#else
    (void)nthreads;
#endif
    const size_t n = (size_t)H * W;
    int32_t *sa = (int32_t *)malloc(n * sizeof(int32_t)), *sb = (int32_t *)malloc(n * sizeof(int32_t));
    if (!sa || !sb) { free(sa); free(sb); return -2; }
    size_t n_known = 0;
    for (size_t p = 0; p < n; ++p) {
        sa[p] = mask[p] == 0 ? (int32_t)p : SEED_NONE;
        n_known += mask[p] == 0;
    }
    memcpy(out, img, n * C);
    if (n_known == 0 || n_known == n) { free(sa); free(sb); return 0; }

    int top = 1;
    while (top < (H > W ? H : W)) top <<= 1;
    int steps[40], ns = 0;
    for (int s = top >> 1; s >= 1; s >>= 1) steps[ns++] = s;
    steps[ns++] = 1;
    for (int i = 0; i < ns; ++i) {
        const int step = steps[i];
#pragma omp parallel for schedule(static)
        for (int r = 0; r < H; ++r)
            for (int c = 0; c < W; ++c) {
                int32_t best = sa[(size_t)r * W + c];
                int64_t bd = best == SEED_NONE ? 0 : dist2(W, r, c, best);
                for (int j = -1; j <= 1; ++j)
                    for (int k = -1; k <= 1; ++k) {
                        const int rr = r + j * step, cc = c + k * step;
                        if (rr < 0 || rr >= H || cc < 0 || cc >= W) continue;
                        const int32_t s = sa[(size_t)rr * W + cc];
                        if (s == SEED_NONE) continue;
                        const int64_t d = dist2(W, r, c, s);
                        if (best == SEED_NONE || d < bd || (d == bd && s < best)) { best = s; bd = d; }
                    }
                sb[(size_t)r * W + c] = best;
            }
        int32_t *t = sa; sa = sb; sb = t;
    }

#pragma omp parallel for schedule(static)
    for (int r = 0; r < H; ++r)
        for (int c = 0; c < W; ++c) {
            const size_t p = (size_t)r * W + c;
            if (mask[p] == 0) continue;
            const int32_t q = sa[p];
            if (q == SEED_NONE) continue; /* cannot happen once a known pixel exists; kept for safety */
            const int qr = q / W, qc = q % W;
            const float num = (float)(1 + dist2(W, r, c, q));
            float acc[4] = {0.f, 0.f, 0.f, 0.f}, ws = 0.0f;
            for (int tr = qr - radius; tr <= qr + radius; ++tr)
                for (int tc = qc - radius; tc <= qc + radius; ++tc) {
                    if (tr < 0 || tr >= H || tc < 0 || tc >= W) continue;
                    if ((tr - qr) * (tr - qr) + (tc - qc) * (tc - qc) > radius * radius) continue;
                    const size_t t = (size_t)tr * W + tc;
                    if (mask[t] != 0) continue;
                    const int64_t dy = r - tr, dx = c - tc;
                    const float w = num / (float)(1 + dy * dy + dx * dx);
                    for (int ch = 0; ch < C; ++ch) acc[ch] = acc[ch] + w * (float)img[t * C + ch];
                    ws = ws + w;
                }
            for (int ch = 0; ch < C; ++ch) {
                float v = rintf(acc[ch] / ws);
                v = v > 255.0f ? 255.0f : v;
                out[p * C + ch] = (uint8_t)v;
            }
        }
    free(sa); free(sb);
    return 0;
}
