// What does HBM sustain for in-place updates of CHUNKS of C contiguous bytes at scattered addresses?
// Each warp handles one chunk at a time (read float4 per lane, modify, write back), chunks visited in a
// hashed order so that concurrently running warps touch unrelated DRAM pages.  Compare with C = whole array.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o chunk_bw chunk_bw.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned mix(unsigned x) { x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16; return x; }
// chunk_f4 = float4 elements per chunk (C / 16); n_chunks power of two; order: 0 sequential, 1 hashed (bijective: odd multiplier + xor)
template <int UNROLL>
__global__ void k(float4* a, unsigned n_chunks, unsigned chunk_f4, int hashed) {
    const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31, n_warps = (gridDim.x * blockDim.x) >> 5;
    for (unsigned c = warp; c < n_chunks; c += n_warps) {
        unsigned cc = hashed ? ((c * 0x9E3779B1u) ^ 0x5bd1e995u) & (n_chunks - 1) : c;
        float4* p = a + (size_t)cc * chunk_f4;
        for (unsigned i = lane; i < chunk_f4; i += 32 * UNROLL) {
            float4 v[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; u++) if (i + u * 32 < chunk_f4) v[u] = p[i + u * 32];
#pragma unroll
            for (int u = 0; u < UNROLL; u++) if (i + u * 32 < chunk_f4) { v[u].x += 1.f; p[i + u * 32] = v[u]; }
        }
    }
}
int main() {
    const size_t bytes = 1ull << 30;  // 1 GiB
    float4* a; cudaMalloc(&a, bytes); cudaMemset(a, 0, bytes);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int hashed = 0; hashed < 2; hashed++)
    for (unsigned C : {64u, 128u, 256u, 512u, 1024u, 2048u, 4096u, 8192u, 32768u}) {
        unsigned chunk_f4 = C / 16, n_chunks = (unsigned)(bytes / C);
        float best = 1e9;
        for (int rep = 0; rep < 5; rep++) {
            cudaEventRecord(e0);
            k<4><<<148 * 16, 256>>>(a, n_chunks, chunk_f4, hashed);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        printf("%s chunks of %6u B: %7.1f GB/s (read+write)\n", hashed ? "scattered " : "sequential", C, 2.0 * bytes / best / 1e6);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
