// Stand-alone use of the C ABI (include/wr_b200.h) without Python or PyTorch: the host side a C / C++ caller
// of the reference's render path would write.  Reads a little-endian scene file, renders mask / position /
// depth / normal with wr_render, writes the maps back.  tests/test_gpu_c_abi.py drives it and compares the
// result with the Python package bit for bit.
//
//   scene file : int32 V, F, B, H, W ; float v_pos[V*3] ; int32 tri[F*3] ; float mvp[B*16] ; float w2c[B*16]
//   result file: uint8 mask[B*H*W] ; float pos[B*H*W*3] ; float depth[B*H*W] ; float normal[B*H*W*3]
//
// Build (see examples/build.py): g++ c_abi_render.cpp -I../include -I$CUDA/include -L.. -lwr_b200 -lcudart
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "wr_b200.h"

#define CK(call)                                                                             \
    do {                                                                                     \
        cudaError_t e_ = (call);                                                             \
        if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #call, cudaGetErrorString(e_)); return 2; } \
    } while (0)
#define WR(call)                                                                             \
    do {                                                                                     \
        int s_ = (call);                                                                     \
        if (s_ != WR_OK) { fprintf(stderr, "%s: %s (%s)\n", #call, wr_status_string(s_), wr_ctx_last_error(ctx)); return 3; } \
    } while (0)

template <typename T>
static bool read_vec(FILE *f, std::vector<T> &v, size_t n)
{
    v.resize(n);
    return n == 0 || fread(v.data(), sizeof(T), n, f) == n;
}

int main(int argc, char **argv)
{
    if (argc != 3) { fprintf(stderr, "usage: %s scene.bin result.bin\n", argv[0]); return 1; }
    FILE *f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 1; }
    int32_t hdr[5];
    if (fread(hdr, sizeof(int32_t), 5, f) != 5) return 1;
    const int V = hdr[0], F = hdr[1], B = hdr[2], H = hdr[3], W = hdr[4];
    std::vector<float> v_pos, mvp, w2c;
    std::vector<int32_t> tri;
    if (!read_vec(f, v_pos, (size_t)V * 3) || !read_vec(f, tri, (size_t)F * 3) || !read_vec(f, mvp, (size_t)B * 16) ||
        !read_vec(f, w2c, (size_t)B * 16)) { fprintf(stderr, "short scene file\n"); return 1; }
    fclose(f);

    if (wr_version() != WR_B200_ABI_VERSION) {   // a library built from another header: different argument structs
        fprintf(stderr, "libwr_b200.so has ABI version %d, this program was compiled against %d\n", wr_version(),
                WR_B200_ABI_VERSION);
        return 4;
    }
    wr_ctx *ctx = nullptr;
    int st = wr_ctx_create(0, &ctx);
    if (st != WR_OK) { fprintf(stderr, "wr_ctx_create: %s\n", wr_status_string(st)); return 3; }
    cudaStream_t stream;
    CK(cudaStreamCreate(&stream));

    const size_t npix = (size_t)B * H * W;
    float *d_pos_in, *d_nrm_in, *d_mvp, *d_w2c, *d_pos, *d_depth, *d_normal;
    int32_t *d_tri;
    uint8_t *d_mask;
    CK(cudaMalloc(&d_pos_in, v_pos.size() * 4)); CK(cudaMalloc(&d_nrm_in, v_pos.size() * 4));
    CK(cudaMalloc(&d_tri, tri.size() * 4)); CK(cudaMalloc(&d_mvp, mvp.size() * 4)); CK(cudaMalloc(&d_w2c, w2c.size() * 4));
    CK(cudaMalloc(&d_mask, npix)); CK(cudaMalloc(&d_pos, npix * 12)); CK(cudaMalloc(&d_depth, npix * 4));
    CK(cudaMalloc(&d_normal, npix * 12));
    CK(cudaMemcpyAsync(d_pos_in, v_pos.data(), v_pos.size() * 4, cudaMemcpyHostToDevice, stream));
    CK(cudaMemcpyAsync(d_tri, tri.data(), tri.size() * 4, cudaMemcpyHostToDevice, stream));
    CK(cudaMemcpyAsync(d_mvp, mvp.data(), mvp.size() * 4, cudaMemcpyHostToDevice, stream));
    CK(cudaMemcpyAsync(d_w2c, w2c.data(), w2c.size() * 4, cudaMemcpyHostToDevice, stream));

    // mesh.py:85-119, then render.py:220-286 with the default DepthControlNetNormalization
    WR(wr_vertex_normals(ctx, d_pos_in, V, d_tri, F, d_nrm_in, stream));
    wr_render_args a = {};
    a.v_pos = d_pos_in; a.tri = d_tri; a.V = V; a.F = F;
    a.v_nrm = d_nrm_in; a.tri_nrm = nullptr; a.Vn = V;
    a.mvp = d_mvp; a.w2c = d_w2c; a.B = B; a.H = H; a.W = W;
    a.depth_mode = WR_DEPTH_CONTROLNET; a.depth_p0 = 0.25f; a.depth_p1 = 0.75f; a.depth_bg = 0.0f;
    a.out_mask = d_mask; a.out_pos = d_pos; a.out_depth = d_depth; a.out_normal = d_normal;
    WR(wr_render(ctx, &a, stream));

    std::vector<uint8_t> mask(npix);
    std::vector<float> pos(npix * 3), depth(npix), normal(npix * 3);
    CK(cudaMemcpyAsync(mask.data(), d_mask, npix, cudaMemcpyDeviceToHost, stream));
    CK(cudaMemcpyAsync(pos.data(), d_pos, npix * 12, cudaMemcpyDeviceToHost, stream));
    CK(cudaMemcpyAsync(depth.data(), d_depth, npix * 4, cudaMemcpyDeviceToHost, stream));
    CK(cudaMemcpyAsync(normal.data(), d_normal, npix * 12, cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));

    FILE *o = fopen(argv[2], "wb");
    if (!o) { perror(argv[2]); return 1; }
    fwrite(mask.data(), 1, npix, o);
    fwrite(pos.data(), 4, npix * 3, o);
    fwrite(depth.data(), 4, npix, o);
    fwrite(normal.data(), 4, npix * 3, o);
    fclose(o);
    size_t covered = 0;
    for (uint8_t m : mask) covered += m;
    printf("rendered %d views %dx%d of %d faces: %zu covered pixels, scratch %llu bytes\n", B, H, W, F, covered,
           (unsigned long long)wr_ctx_scratch_bytes(ctx));
    wr_ctx_destroy(ctx);
    return 0;
}
