// L2 -> shared-memory streaming bandwidth with 1-D TMA bulk copies (developer microbenchmark).
// Every CTA streams chunks of an L2-resident buffer into a 2-stage smem ring and discards them.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(const char* src, size_t bytes_total, int chunk, int iters)
{
    extern __shared__ __align__(128) unsigned char sm[];
    uint64_t* bar = (uint64_t*)sm;
    char* buf = (char*)sm + 128;
    if (threadIdx.x == 0) {
        for (int b = 0; b < 2; ++b) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar[b])));
        asm volatile("fence.mbarrier_init.release.cluster;");
        asm volatile("fence.proxy.async.shared::cta;");
    }
    __syncthreads();
    const size_t nchunks = bytes_total / chunk;
    if (threadIdx.x == 0) {
        for (int it = 0; it < iters; ++it) {
            const int b = it & 1;
            if (it >= 2) {  // wait for the copy issued two iterations ago
                uint32_t done;
                do { asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p;}" : "=r"(done) : "r"(s32(&bar[b])), "r"((uint32_t)(((it - 2) >> 1) & 1)) : "memory"); } while (!done);
            }
            const size_t c = ((size_t)blockIdx.x * 977 + (size_t)it * 131) % nchunks;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar[b])), "r"((uint32_t)chunk) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(buf + (size_t)b * chunk)), "l"(src + c * chunk), "r"((uint32_t)chunk), "r"(s32(&bar[b])) : "memory");
        }
        for (int it = iters; it < iters + 2; ++it) {
            const int b = it & 1;
            uint32_t done;
            do { asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p;}" : "=r"(done) : "r"(s32(&bar[b])), "r"((uint32_t)(((it - 2) >> 1) & 1)) : "memory"); } while (!done);
        }
    }
}
int main()
{
    const size_t total = 64ull << 20;  // 64 MiB: L2 resident (126 MB L2)
    char* d; cudaMalloc(&d, total); cudaMemset(d, 1, total);
    for (int chunk : {16384, 32768, 65536}) {
        for (int ctas_per_sm : {1, 2}) {
            const int grid = 148 * ctas_per_sm, iters = 4000;
            const size_t smem = 128 + 2 * (size_t)chunk;
            cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            k<<<grid, 64, smem>>>(d, total, chunk, 200);  // warm
            cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
            cudaEventRecord(a); k<<<grid, 64, smem>>>(d, total, chunk, iters); cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b);
            printf("chunk %6d B, %d CTA/SM: %.1f GB/s L2->smem (err=%s)\n", chunk, ctas_per_sm, (double)grid * iters * chunk / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
        }
    }
    return 0;
}
