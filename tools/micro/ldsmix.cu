// What bounds the forward kernel's shared-memory pipe (developer microbenchmark, profiles/r2_ldsmix.json).
// ctr_fwd_kernel reads 32-image pixel records with LDS.128 from one strip buffer while the TMA engine fills the other.
// This kernel reproduces the two ingredients in isolation, one 640-thread CTA per SM like the kernel:
//   pattern "linear"   lane l reads the l-th 16-byte vector of a row (the conflict-free ceiling of ldsbw.cu)
//   pattern "records"  a quarter-warp = 2 rays x 4 lanes; every ray reads its own 128-byte record (pseudo-random
//                      position), each lane the two halves of its 32-byte block in parity-swizzled order
//   tma = 0 / 1        a producer warp keeps refilling the OTHER half of shared memory with 1-D bulk copies from an
//                      L2-resident buffer as fast as the mbarrier round trip allows
// Output: bytes / clk / SM the 19 consumer warps achieve, and the bulk-copy rate that ran beside them.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ldsmix ldsmix.cu && ./ldsmix [out.json]
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

constexpr int kThreads = 640, kHalf = 96 * 1024, kCopy = 8 * 1024;   // two 96 KB halves; bulk copies of 8 KB

__global__ void __launch_bounds__(kThreads, 1) k(const float* __restrict__ src, float* out, long long* cycles, long long* tma_bytes,
                                                  int iters, int pattern, int tma)
{
    extern __shared__ __align__(128) unsigned char sm[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm);
    volatile int* stop = reinterpret_cast<volatile int*>(sm + 64);
    float* rd = reinterpret_cast<float*>(sm + 128);              // read half
    float* wr = rd + kHalf / 4;                                   // half the TMA engine writes
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < kHalf / 4; i += kThreads) rd[i] = (float)(i & 1023) * 1e-3f;
    if (tid == 0) {
        *stop = 0;
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (warp == kThreads / 32 - 1) {                              // producer warp
        if (lane == 0 && tma) {
            long long moved = 0;
            uint32_t phase = 0;
            const float* s = src + (size_t)blockIdx.x * (kHalf / 4);
            while (!*stop) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"((uint32_t)kHalf) : "memory");
                for (int c = 0; c < kHalf / kCopy; ++c)
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                     s32(reinterpret_cast<unsigned char*>(wr) + c * kCopy)),
                                 "l"(reinterpret_cast<const unsigned char*>(s) + c * kCopy), "r"((uint32_t)kCopy), "r"(s32(bar))
                                 : "memory");
                uint32_t done = 0;
                while (!done)
                    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                                 : "=r"(done) : "r"(s32(bar)), "r"(phase) : "memory");
                phase ^= 1;
                moved += kHalf;
            }
            tma_bytes[blockIdx.x] = moved;
        }
        return;
    }
    // consumers
    const float4* base = reinterpret_cast<const float4*>(rd);
    const int nrec = kHalf / 128;                                 // 128-byte records
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    uint32_t rng = (uint32_t)(warp * 9781 + blockIdx.x * 6271 + 12345);
    const int ray = lane >> 2, l4 = lane & 3, par = ray & 1;
    int row = warp;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            float4 v0, v1;
            if (pattern == 0) {
                v0 = base[(row * 32 + lane) % (kHalf / 16)];
                v1 = base[((row + 3) * 32 + lane) % (kHalf / 16)];
                row += 7;
                if (row >= kHalf / 512) row -= kHalf / 512;
            } else {
                rng = rng * 1664525u + 1013904223u;               // same for all lanes of the warp
                const int rec = (int)(((rng >> 8) + (uint32_t)ray * 2654435761u) % (uint32_t)nrec);
                v0 = base[rec * 8 + 2 * l4 + par];
                v1 = base[rec * 8 + 2 * l4 + (par ^ 1)];
            }
            acc[0] += v0.x + v1.x; acc[1] += v0.y + v1.y; acc[2] += v0.z + v1.z; acc[3] += v0.w + v1.w;
        }
    }
    const long long t1 = clock64();
    out[blockIdx.x * kThreads + tid] = acc[0] + acc[1] + acc[2] + acc[3];
    asm volatile("bar.sync 1, %0;" ::"r"(kThreads - 32) : "memory");
    if (tid == 0) { cycles[blockIdx.x] = t1 - t0; *stop = 1; }
}

int main(int argc, char** argv)
{
    const int grid = 148, iters = 20000;
    float *out, *src; long long *cyc, *tb;
    cudaMalloc(&out, sizeof(float) * grid * kThreads);
    cudaMalloc(&src, (size_t)grid * kHalf);
    cudaMemset(src, 0, (size_t)grid * kHalf);
    cudaMalloc(&cyc, sizeof(long long) * grid);
    cudaMalloc(&tb, sizeof(long long) * grid);
    const size_t smem = 128 + 2 * kHalf;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    FILE* js = argc > 1 ? fopen(argv[1], "w") : nullptr;
    if (js) fprintf(js, "{\n  \"what\": \"LDS.128 rate of 19 consumer warps per SM with / without concurrent TMA bulk fills (tools/micro/ldsmix.cu)\"");
    const char* pn[2] = {"linear", "records"};
    for (int pattern = 0; pattern < 2; ++pattern)
        for (int tma = 0; tma < 2; ++tma) {
            cudaMemset(tb, 0, sizeof(long long) * grid);
            k<<<grid, kThreads, smem>>>(src, out, cyc, tb, 200, pattern, tma);
            k<<<grid, kThreads, smem>>>(src, out, cyc, tb, iters, pattern, tma);
            cudaDeviceSynchronize();
            long long h[148], t[148];
            cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
            cudaMemcpy(t, tb, sizeof(t), cudaMemcpyDeviceToHost);
            double cavg = 0, tavg = 0;
            for (int i = 0; i < grid; ++i) { cavg += (double)h[i] / grid; tavg += (double)t[i] / grid; }
            const double bytes = (double)(kThreads - 32) * iters * 4 * 32;
            printf("%-8s tma=%d: LDS %6.1f B/clk/SM   TMA fill %5.1f B/clk/SM   (%s)\n", pn[pattern], tma, bytes / cavg, tavg / cavg,
                   cudaGetErrorString(cudaGetLastError()));
            if (js) fprintf(js, ",\n  \"%s_tma%d\": {\"lds_bytes_per_clk_per_sm\": %.2f, \"tma_fill_bytes_per_clk_per_sm\": %.2f}", pn[pattern], tma,
                            bytes / cavg, tavg / cavg);
        }
    if (js) { fprintf(js, "\n}\n"); fclose(js); }
    return 0;
}
