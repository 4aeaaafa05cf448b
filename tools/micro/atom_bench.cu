// Microbenchmark: throughput of 64-bit atomicMin (RED.MIN.64) vs plain 64-bit stores on an L2-resident buffer.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_red(unsigned long long *buf, size_t n, int reps) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    for (int r = 0; r < reps; ++r) { size_t j = (i + (size_t)r * 977) % n; atomicMin(buf + j, (unsigned long long)(j * 2654435761ull + r)); }
}
__global__ void k_st(unsigned long long *buf, size_t n, int reps) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    for (int r = 0; r < reps; ++r) { size_t j = (i + (size_t)r * 977) % n; buf[j] = j * 2654435761ull + r; }
}
__global__ void k_red32(unsigned int *buf, size_t n, int reps) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    for (int r = 0; r < reps; ++r) { size_t j = (i + (size_t)r * 977) % n; atomicMin(buf + j, (unsigned int)(j * 2654435761u + r)); }
}
int main() {
    const size_t n = 3538944;  // 6 x 768 x 768 entries = 28 MB (the packed buffer of config B)
    unsigned long long *buf; cudaMalloc(&buf, n * 8); cudaMemset(buf, 0xFF, n * 8);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const int reps = 4; const int threads = 256; const int blocks = (int)((n + threads - 1) / threads);
    for (int which = 0; which < 3; ++which) {
        float best = 1e9;
        for (int it = 0; it < 5; ++it) {
            cudaMemset(buf, 0xFF, n * 8);
            cudaEventRecord(a);
            if (which == 0) k_red<<<blocks, threads>>>(buf, n, reps);
            else if (which == 1) k_st<<<blocks, threads>>>(buf, n, reps);
            else k_red32<<<blocks, threads>>>((unsigned int *)buf, n, reps);
            cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
        }
        printf("%s: %.1f us for %.1f M ops -> %.1f G ops/s\n", which == 0 ? "atomicMin u64" : which == 1 ? "store u64" : "atomicMin u32", best * 1e3, n * reps / 1e6, n * reps / best / 1e6);
    }
    return 0;
}
