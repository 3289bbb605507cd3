// Experiment helper: occupy n SMs (one CTA each, all of the SM's shared memory) for a bounded time, so that
// another kernel runs on the remaining SMs only.  nvcc -shared -Xcompiler -fPIC -gencode arch=compute_100a,code=sm_100a
#include <cuda_runtime.h>
__global__ void hog(unsigned long long max_ns, unsigned* sink) {
    extern __shared__ unsigned char s[];
    unsigned long long t0, t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    do {
        __nanosleep(2000);
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    } while (t - t0 < max_ns);
    if (threadIdx.x == 0 && s[0] == 123) *sink = 1;
}
extern "C" int hog_start(void* stream, int n_ctas, int max_ms) {
    static unsigned* sink = nullptr;
    if (!sink) cudaMalloc(&sink, 4);
    const int smem = 227 * 1024;
    if (cudaFuncSetAttribute(hog, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return -1;
    hog<<<n_ctas, 32, smem, (cudaStream_t)stream>>>((unsigned long long)max_ms * 1000000ull, sink);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}
