// Does cudaMemsetAsync on a side stream run under an issue-bound kernel?  (nvcc -O3 -arch=sm_100a memset_overlap.cu)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void spin(unsigned *out, int iters)
{
    unsigned x = threadIdx.x + blockIdx.x * blockDim.x, y = x * 7u + 3u;
    for (int i = 0; i < iters; ++i) { x = x * 1664525u + 1013904223u; y ^= x >> 7; y += __popc(x); }
    if (y == 0x12345678u) out[0] = y;
}
__global__ void fillk(uint4 *p, size_t n, uint4 v)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) __stcs(p + i, v);
}
int main()
{
    const size_t bytes = 118u << 20;
    void *buf; unsigned *out;
    cudaMalloc(&buf, bytes); cudaMalloc(&out, 4);
    cudaStream_t s1, s2; cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking); cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking);
    cudaEvent_t e0, e1, f0, f1; cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&f0); cudaEventCreate(&f1);
    auto ms = [&](cudaEvent_t a, cudaEvent_t b) { float t; cudaEventElapsedTime(&t, a, b); return t * 1e3f; };
    for (int rep = 0; rep < 3; ++rep) {
        // compute alone
        cudaEventRecord(e0, s1); spin<<<148 * 8, 256, 0, s1>>>(out, 3000); cudaEventRecord(e1, s1); cudaDeviceSynchronize();
        const float t_spin = ms(e0, e1);
        cudaEventRecord(f0, s2); cudaMemsetAsync(buf, 0, bytes, s2); cudaEventRecord(f1, s2); cudaDeviceSynchronize();
        const float t_set = ms(f0, f1);
        cudaEventRecord(f0, s2); fillk<<<148 * 4, 256, 0, s2>>>((uint4 *)buf, bytes / 16, make_uint4(1, 2, 3, 4)); cudaEventRecord(f1, s2); cudaDeviceSynchronize();
        const float t_fillk = ms(f0, f1);
        // both
        cudaEventRecord(e0, s1); cudaEventRecord(f0, s2);
        spin<<<148 * 8, 256, 0, s1>>>(out, 3000); cudaMemsetAsync(buf, 0, bytes, s2);
        cudaEventRecord(e1, s1); cudaEventRecord(f1, s2); cudaDeviceSynchronize();
        const float b_spin = ms(e0, e1), b_set = ms(f0, f1);
        cudaEventRecord(e0, s1); cudaEventRecord(f0, s2);
        spin<<<148 * 8, 256, 0, s1>>>(out, 3000); fillk<<<148 * 2, 256, 0, s2>>>((uint4 *)buf, bytes / 16, make_uint4(1, 2, 3, 4));
        cudaEventRecord(e1, s1); cudaEventRecord(f1, s2); cudaDeviceSynchronize();
        const float c_spin = ms(e0, e1), c_fill = ms(f0, f1);
        printf("alone: spin %.1f us, memset %.1f us (%.0f GB/s), fill kernel %.1f us (%.0f GB/s) | together: spin %.1f memset %.1f | spin %.1f fillk(296 blocks) %.1f\n",
               t_spin, t_set, bytes / t_set * 1e-3, t_fillk, bytes / t_fillk * 1e-3, b_spin, b_set, c_spin, c_fill);
    }
    return 0;
}
