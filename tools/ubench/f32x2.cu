// Microbenchmark: issue rate of packed fp32 (FFMA2/FADD2/FMUL2) vs scalar FFMA on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2 f32x2.cu && ./f32x2
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
template <int MODE>
__global__ void k(float* out, long long* cyc, int iters) {
    float s = threadIdx.x * 1e-3f;
    u64 a[8]; float f[16];
    for (int i = 0; i < 8; i++) { a[i] = ((u64)__float_as_uint(s + i) << 32) | __float_as_uint(s - i); }
    for (int i = 0; i < 16; i++) f[i] = s + i;
    u64 m = ((u64)__float_as_uint(1.0001f) << 32) | __float_as_uint(0.9999f);
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 8; i++) a[i] = fma2(a[i], m, m);
        } else if (MODE == 1) {
#pragma unroll
            for (int i = 0; i < 16; i++) f[i] = fma1(f[i], 1.0001f, 0.5f);
        } else if (MODE == 2) {
#pragma unroll
            for (int i = 0; i < 8; i++) a[i] = add2(a[i], m);
        } else {
#pragma unroll
            for (int i = 0; i < 8; i++) f[i] = fma1(f[i], 1.0001f, 0.5f);
        }
    }
    long long t1 = clock64();
    float r = 0;
    for (int i = 0; i < 8; i++) r += __uint_as_float((unsigned)a[i]) + __uint_as_float((unsigned)(a[i] >> 32));
    for (int i = 0; i < 16; i++) r += f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    float* out; long long* cyc; cudaMalloc(&out, 4 << 20); cudaMalloc(&cyc, 8);
    const int iters = 4096;
    const char* names[] = {"FFMA2 x8/iter", "FFMA x16/iter", "FADD2 x8/iter", "FFMA x8/iter"};
    for (int threads : {32, 128, 256, 512, 1024}) {
        for (int mode = 0; mode < 4; mode++) {
            long long h = 0;
            for (int rep = 0; rep < 2; rep++) {
                if (mode == 0) k<0><<<148, threads>>>(out, cyc, iters);
                if (mode == 1) k<1><<<148, threads>>>(out, cyc, iters);
                if (mode == 2) k<2><<<148, threads>>>(out, cyc, iters);
                if (mode == 3) k<3><<<148, threads>>>(out, cyc, iters);
                cudaDeviceSynchronize();
            }
            cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            int n = (mode == 1) ? 16 : 8;
            double warps_per_smsp = threads / 32.0 / 4.0;
            printf("threads/SM %4d  %-14s  %.2f cycles per warp-instruction per SMSP (%.1f warps/SMSP)\n", threads, names[mode],
                   (double)h / iters / n / (warps_per_smsp < 1 ? 1 : warps_per_smsp), warps_per_smsp);
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
