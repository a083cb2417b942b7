// development probe: how many CTAs that each tcgen05.alloc 64 columns does an SM hold, by the occupancy API and in practice
#include <cstdio>
#include <cuda_runtime.h>
template <bool TM>
__global__ void __launch_bounds__(128, 7) k(unsigned long long *out, int spin)
{
    extern __shared__ unsigned char sm[];
    unsigned *slot = reinterpret_cast<unsigned *>(sm);
    unsigned base = 0;
    if (TM) {
        if (threadIdx.x < 32) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(slot)), "r"(64) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        base = *slot;
    }
    unsigned long long t0 = clock64();
    while (clock64() - t0 < (unsigned long long)spin) {}
    if (threadIdx.x == 0) { unsigned smid; asm("mov.u32 %0, %%smid;" : "=r"(smid)); out[blockIdx.x] = ((unsigned long long)smid << 32) | base; }
    if (TM) { __syncthreads(); if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(64) : "memory"); }
}
template <bool TM> void run(const char *name)
{
    auto kern = k<TM>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 30224);
    int nb = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, 128, 30224);
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, kern);
    unsigned long long *out; cudaMalloc(&out, 8 * 148 * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int ctas = 1; ctas <= 8; ctas++) {
        cudaEventRecord(e0);
        kern<<<148 * ctas, 128, 30224>>>(out, 2000000);
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("%s regs %d occupancy API %d: grid 148x%d -> %.3f ms (%s)\n", name, fa.numRegs, nb, ctas, ms, cudaGetErrorString(e));
    }
}
int main() { run<false>("plain"); run<true>("tmem64"); return 0; }
