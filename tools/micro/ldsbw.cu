// Shared-memory read bandwidth of one SM and of the chip (developer microbenchmark; replaces the DERIVED
// "128 B/clk/SM" ceiling of bench.py's on-chip roofline with a measured one -- profiles/r2_ldsbw.json).
// Every warp issues conflict-free LDS.{32,64,128} in an unrolled loop (lane l reads the l-th vector of a row, rows
// advance with the iteration, so neither the compiler nor the hardware can elide a load); the sums keep the loads
// alive.  Cycles come from clock64() inside the kernel (bytes / clk / SM), GB/s from CUDA events over all SMs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ldsbw ldsbw.cu && ./ldsbw [out.json]
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

template <typename V>
__global__ void __launch_bounds__(1024) k(float* out, long long* cycles, int iters)
{
    extern __shared__ __align__(16) unsigned char sm[];
    constexpr int VW = sizeof(V) / 4;                      // words per vector
    const int nvec = 48 * 1024 / sizeof(V);                // 48 KB window
    float* smf = reinterpret_cast<float*>(sm);
    for (int i = threadIdx.x; i < nvec * VW; i += blockDim.x) smf[i] = (float)(i & 1023) * 1e-3f;
    __syncthreads();
    const V* base = reinterpret_cast<const V*>(sm);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    int row = warp;
    const int rows = nvec / 32;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const V v = base[row * 32 + lane];             // 32 lanes x sizeof(V) contiguous: conflict-free
            const float* f = reinterpret_cast<const float*>(&v);
#pragma unroll
            for (int w = 0; w < VW; ++w) acc[w] += f[w];
            row += 7;
            if (row >= rows) row -= rows;
        }
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc[0] + acc[1] + acc[2] + acc[3];
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <typename V>
static void run(const char* name, FILE* js, bool last)
{
    const int grid = 148, threads = 1024, iters = 4000;
    float* out; long long* cyc;
    cudaMalloc(&out, sizeof(float) * grid * threads);
    cudaMalloc(&cyc, sizeof(long long) * grid);
    const size_t smem = 48 * 1024;
    k<V><<<grid, threads, smem>>>(out, cyc, 100);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a); k<V><<<grid, threads, smem>>>(out, cyc, iters); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double cmax = 0, cavg = 0;
    for (int i = 0; i < grid; ++i) { cavg += (double)h[i] / grid; if ((double)h[i] > cmax) cmax = (double)h[i]; }
    const double bytes_per_cta = (double)threads * iters * 8 * sizeof(V);
    const double per_clk = bytes_per_cta / cavg, gbs = bytes_per_cta * grid / ms / 1e6;
    printf("%-8s %6.1f B/clk/SM (thread-0 cycles, mean over %d SMs; max %.0f)   %8.0f GB/s over the chip (%.3f ms, err=%s)\n", name,
           per_clk, grid, cmax, gbs, ms, cudaGetErrorString(cudaGetLastError()));
    if (js) fprintf(js, "  \"%s_bytes_per_clk_per_sm\": %.2f,\n  \"%s_chip_GBs\": %.1f%s\n", name, per_clk, name, gbs, last ? "" : ",");
    cudaFree(out); cudaFree(cyc);
}

int main(int argc, char** argv)
{
    FILE* js = argc > 1 ? fopen(argv[1], "w") : nullptr;
    if (js) fprintf(js, "{\n  \"what\": \"conflict-free shared-memory loads, 1024 threads per SM, 148 SMs (tools/micro/ldsbw.cu)\",\n");
    run<float>("lds32", js, false);
    run<float2>("lds64", js, false);
    run<float4>("lds128", js, true);
    if (js) { fprintf(js, "}\n"); fclose(js); }
    return 0;
}
