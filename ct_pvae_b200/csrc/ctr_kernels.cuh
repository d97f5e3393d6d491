// ctr_kernels.cuh -- sm_100a kernels of the Radon path and their launchers.
//
//   K0  ctr_pack_image_kernel   [B,X,Y] -> batch-interleaved, halo-padded packs (row-major + transposed),
//                               4, 8, 16 or 32 images per pixel record
//   K0' ctr_pack_sino_kernel    [B,A,W] -> [G,A,NB/4,W+2,4] sinogram planes with zero halo bins
//   K1  ctr_fwd_kernel          ray-driven forward projector (project_tf_fast / project_tf_low_mem,
//                               /root/reference/ctvae/forward_functions.py:80-123, :49-78), optionally with
//                               the fused measurement log-likelihood epilogue (helper_functions.py:355-368)
//   K2  ctr_bp_kernel<EXACT>    pixel-driven gather adjoint, the exact transpose of K1 (no atomics)
//   K2' ctr_bp_kernel<TF>       TensorFlow's registered gradient of the projector graph
//   K3a ctr_fbp_filter_kernel   circular row filter in shared memory (fbp_tensorflow.py:49-50); the ramp filter's
//       ctr_fbp_filter_sparse_kernel  kernel vanishes at the even offsets: odd taps only, parity-split rows
//   K3b ctr_bp_kernel<FBP>      linear-interpolating back-projection of iradon (fbp_tensorflow.py:52-74)
//   K3  ctr_fbp_fused_kernel    iradon as one thread-block-cluster kernel (opt-in)
//
// Data movement: image strips (K1) and sinogram bin windows (K2/K3b) are staged in
// shared memory by the TMA engine with 1-D bulk copies (cp.async.bulk, SASS UBLKCP)
// completing on mbarriers.  K1 and K2/K3b run a producer warp / consumer warps ring (full +
// empty mbarriers per buffer, no CTA-wide barrier in the loop): strips for K1, batches of 8
// angles for K2/K3b.  The packs exist so that every staged block is a contiguous,
// 16-byte aligned range in HBM (whole packed rows, or -- wide detectors -- the column
// window of every row that the CTA's rays cross) and so that one 128-bit shared-memory
// load serves 4 images: the per-sample geometry (coordinates, floor, weights, address)
// is computed once per 4, 8 or 16 (K1) or 16-32 (K2) images.  K1's big-batch shape reads
// 32-image records with parity-swizzled 8-image lanes (conflict-free quarter-warps) and,
// with tall windowed strips, keeps the previous sample's bottom row in registers; the
// nearest-neighbour projector reads them with two lanes per ray x 16 images (rotated loads).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "ctr_core.h"

// ------------------------------------------------------------------------------------------ PTX glue
namespace ctr {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
// make mbarrier.init visible to the async proxy (TMA) before the first copy is issued
__device__ __forceinline__ void fence_barrier_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    const uint32_t addr = smem_u32(bar);
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!done);
}
// 1-D TMA bulk copy global -> shared, completion bytes counted on `bar`.
// dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

inline std::atomic<long long>& launch_counter()
{
    static std::atomic<long long> c{0};
    return c;
}

constexpr int kFwdNB = 4;   // images interleaved per forward CTA (one LDS.128 per tap)
constexpr int kBpTW = 32, kBpAB = 8;
constexpr int kFwdMaxThreads = 768;                  // block size incl. the producer warp
constexpr int kFwdMaxConsumers = kFwdMaxThreads - 32;  // 736 = 23 warps

__host__ __device__ static inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

// ------------------------------------------------------------------------------------------ K0 packs
// [B,X,Y] -> pixel records of REC images, halo-padded: pk0 [G][X+2][up0][REC] (row-major) and pk1 [G][Y+2][up1][REC]
// (transposed).  One CTA = a 16 x 16 tile of the packed frame x all REC images of one record:
//   load   thread (r, c) of the tile reads its pixel from each of the REC images -- every warp load is two 64-byte row
//          segments, and all REC loads of a thread are in flight before the first is used (the kernel is pure data
//          movement: r1's version, 4 images per CTA and 16-byte stores 128 bytes apart, ran at 23 % of the HBM rate);
//   store  through a shared-memory tile [REC][17-pixel rows] (+ padding: conflict-free in both directions): the lanes of
//          a warp write whole records, 512 contiguous bytes per store instruction, for both packs.
// grid (ceil(max(up0, Y+2) / 16), ceil(max(X+2, up1) / 16), G), block 256.
constexpr int kPackT = 16;                       // tile edge (pixels)
constexpr int kPackStride = 16 * 17 + 17;        // floats between images in the tile: == 1 (mod 32)
template <int REC>
__global__ void __launch_bounds__(256) ctr_pack_image_kernel(const float* __restrict__ img, int B, int X, int Y,
                                                             float* __restrict__ pk0, float* __restrict__ pk1,
                                                             int up0, int up1)   // packed row lengths (pixels) of the two packs
{
    extern __shared__ __align__(128) float tile[];                // [REC][kPackStride], pixel (r, c) at r * 17 + c
    const int g = blockIdx.z;
    const int pr0 = blockIdx.y * kPackT, pc0 = blockIdx.x * kPackT;
    const int tid = threadIdx.x, tr = tid >> 4, tc = tid & 15;
    {
        const int r = pr0 + tr - 1, c = pc0 + tc - 1;            // image coordinates of packed (pr0 + tr, pc0 + tc)
        const bool inside = r >= 0 && r < X && c >= 0 && c < Y;
        const float* src = img + ((size_t)g * REC * X + (inside ? r : 0)) * Y + (inside ? c : 0);
        float v[REC];
#pragma unroll
        for (int n = 0; n < REC; ++n) v[n] = (inside && g * REC + n < B) ? __ldg(src + (size_t)n * X * Y) : 0.f;
#pragma unroll
        for (int n = 0; n < REC; ++n) tile[n * kPackStride + tr * 17 + tc] = v[n];
    }
    __syncthreads();
    constexpr int Q = REC / 4;                                   // 16-byte chunks per record
    for (int idx = tid; idx < kPackT * kPackT * Q; idx += 256) {
        const int q = idx % Q, pix = idx / Q;
        if (pk0) {                                               // consecutive pixels of a tile row
            const int r = pix >> 4, c = pix & 15;
            const int pr = pr0 + r, pc = pc0 + c;
            if (pr < X + 2 && pc < up0) {                        // columns Y+2 .. up0-1 are zero padding
                const float* t = tile + (4 * q) * kPackStride + r * 17 + c;
                *reinterpret_cast<float4*>(pk0 + (((size_t)g * (X + 2) + pr) * up0 + pc) * REC + 4 * q) =
                    make_float4(t[0], t[kPackStride], t[2 * kPackStride], t[3 * kPackStride]);
            }
        }
        if (pk1) {                                               // consecutive pixels of a tile column
            const int c = pix >> 4, r = pix & 15;
            const int pr = pr0 + r, pc = pc0 + c;
            if (pc < Y + 2 && pr < up1) {
                const float* t = tile + (4 * q) * kPackStride + r * 17 + c;
                *reinterpret_cast<float4*>(pk1 + (((size_t)g * (Y + 2) + pc) * up1 + pr) * REC + 4 * q) =
                    make_float4(t[0], t[kPackStride], t[2 * kPackStride], t[3 * kPackStride]);
            }
        }
    }
}

inline cudaError_t launch_pack_image(const float* img, int B, int X, int Y, float* pk0, float* pk1, int rec, int up0, int up1,
                                     cudaStream_t st)
{
    const int G = (B + rec - 1) / rec;
    const int wpix = up0 > Y + 2 ? up0 : Y + 2, hpix = X + 2 > up1 ? X + 2 : up1;
    dim3 grid((wpix + kPackT - 1) / kPackT, (hpix + kPackT - 1) / kPackT, G), block(256);
    const size_t smem = (size_t)rec * kPackStride * sizeof(float);
    switch (rec) {
        case 4: ctr_pack_image_kernel<4><<<grid, block, smem, st>>>(img, B, X, Y, pk0, pk1, up0, up1); break;
        case 8: ctr_pack_image_kernel<8><<<grid, block, smem, st>>>(img, B, X, Y, pk0, pk1, up0, up1); break;
        case 16: ctr_pack_image_kernel<16><<<grid, block, smem, st>>>(img, B, X, Y, pk0, pk1, up0, up1); break;
        case 32: ctr_pack_image_kernel<32><<<grid, block, smem, st>>>(img, B, X, Y, pk0, pk1, up0, up1); break;
        default: return cudaErrorInvalidValue;
    }
    launch_counter()++;
    return cudaGetLastError();
}

// [B,A,W] -> [G][A][NB/4 planes][W+2][4]: zero halo bins, 4 images interleaved per 16-byte bin.
// grid (ceil((W+2)/128), A, G), block 128.
template <int NB>
__global__ void __launch_bounds__(128) ctr_pack_sino_kernel(const float* __restrict__ y, int B, int A, int W,
                                                            float* __restrict__ spk)
{
    const int jp = blockIdx.x * blockDim.x + threadIdx.x;
    if (jp >= W + 2) return;
    const int a = blockIdx.y, g = blockIdx.z, j = jp - 1;
    // all NB loads are issued before the first store (latency-bound kernel, like the image pack)
    float v[NB];
    const bool jok = j >= 0 && j < W;
#pragma unroll
    for (int n = 0; n < NB; ++n) {
        const int b = g * NB + n;
        const bool ok = jok && b < B;
        v[n] = ok ? __ldg(y + ((size_t)(ok ? b : 0) * A + a) * W + (ok ? j : 0)) : 0.f;
    }
#pragma unroll
    for (int h = 0; h < NB / 4; ++h) {
        float* dst = spk + ((((size_t)g * A + a) * (NB / 4) + h) * (W + 2) + jp) * 4;
        *reinterpret_cast<float4*>(dst) = make_float4(v[4 * h], v[4 * h + 1], v[4 * h + 2], v[4 * h + 3]);
    }
}

// ------------------------------------------------------------------------------------------ K1 forward
struct FwdParams {
    const float* pk[2];     // packed images per class  [G][Vp][Up][NB*DEPTH]
    CtrClassGeom geom[2];
    const CtrRay* rays;     // class-sorted ray table [A]
    const CtrChunk* chunks; // [gridDim.x] rays, class, strip height and column window of every CTA column
                            // (angle-subset calls: the per-ray table, indexed through sel and pos)
    const int* sel;         // angle subset: CTA column x serves angle sel[x] and writes sinogram row x; null = whole plan
    const int* pos;         // [A] angle -> row of the ray table
    int H, W, A, B;
    int jwd, ns;            // consumer threads per angle slot (JW bins x DEPTH groups) and angle slots; block = jwd*ns + 32
    int stages;             // strip buffers in the shared-memory ring (2..4)
    float* sino;            // [B][A][W]   (EPI 0: ray sums; EPI 1: d loglik / d proj, the adjoint's cotangent)
    // fused measurement log-likelihood epilogue (EPI 1), helper_functions.py:355-368
    const float* mask;      // [B][A_all]
    const float* meas;      // [B][A_all][W]  measured sinogram (proj_sample)
    const int* amap;        // [A] plan angle -> column of mask/meas (the angles_i gather), or null = identity
    int A_all;
    float pnm, sqrt_reg;    // poisson_noise_multiplier, sqrt_reg
    float* partial;         // [jchunks*gridDim.x][G*NB*DEPTH] per-CTA log-likelihood sums
};

// One CTA = (angle chunk of NS*KA same-class angles) x (image group of NB) x (detector chunk of JW bins).
// thread (tx, ty): detector bin j = blockIdx.y*JW + tx, angles ty*KA .. ty*KA+KA-1 of the chunk.
// grid = (ray chunks, detector chunks, image super-groups), decoded centre-first (see the kernel): the CTAs in flight at
// any time work on one or two detector chunks of every super-group, whose column windows are re-read from L2.
// All threads walk the image group's strips in lock step; thread 0 drives the TMA double buffer.
// DEPTH > 1 is the depth-first variant: DEPTH image groups share one pixel record and the
// lanes of a quarter-warp are (8/DEPTH rays) x (DEPTH groups), so a quarter-warp's LDS.128
// touches 8/DEPTH records of DEPTH*16 contiguous bytes -- shared-memory bank conflicts drop
// from ~1.7x (8 rays spaced 1/cos(theta) > 1 chunks apart) to ~1.2x at DEPTH = 4.
// The vertical-reuse march keeps two register sets of records alive across loop trips: its CTAs are capped at
// 640 threads so that ptxas may use 96 registers instead of 80 (no spills; 704 threads still get 80: registers
// are granted per warp in blocks that make 88 x 22 warps overflow the file).
// 16 images per lane (NB = 16, two lanes per ray; the nearest-neighbour projector): 96 registers, so CTAs of 640
// threads like the reuse march.
constexpr int kFwdReuseThreads = 640, kFwdWideThreads = 640;
__host__ __device__ constexpr int fwd_max_threads(int NB, int REUSE) { return NB == 16 ? kFwdWideThreads : (REUSE ? kFwdReuseThreads : kFwdMaxThreads); }
template <int NB, int KA, int INTERP, int EPI, int DEPTH, int REUSE = 0>
__global__ void __launch_bounds__(fwd_max_threads(NB, REUSE), 1) ctr_fwd_kernel(const FwdParams p)
{
    constexpr int REC = NB * DEPTH;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // block = JWD * NS consumer threads (JWD = JW bins x DEPTH image groups per angle slot) + one producer warp
    const int JWD = p.jwd, NS = p.ns, JW = JWD / DEPTH;
    const int nconsumers = JWD * NS;
    const int tid = threadIdx.x;
    const bool producer = tid >= nconsumers;                 // last warp: drives the TMA strip pipeline
    const int ty = producer ? 0 : tid / JWD;
    const int tx = (tid - ty * JWD) / DEPTH, gsub = (tid - ty * JWD) % DEPTH;
    // 8 images per lane (32-image records): the two 16-byte halves of a lane's block are read in an order
    // that alternates with the ray's parity, which makes every quarter-warp load conflict-free (ctr_ldv8_swz)
    // 16 images per lane: four loads per record, rotated by the ray's index mod 4 (ctr_ldv16_rot)
    const int swz = (NB == 16) ? (tx & 3) * 4 : (NB == 8) ? (tx & 1) * 4 : 0;
    const int NA = NS * KA;

    const int S = p.stages;                                                 // ring depth (<= 4)
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);                 // [S] strip landed (TMA complete_tx)
    uint64_t* empty = full + 4;                                             // [S] strip consumed (one arrive per consumer warp)
    CtrRay* rays_s = reinterpret_cast<CtrRay*>(smem_raw + 128);              // NA rays
    const int rays_bytes = round_up(NA * (int)sizeof(CtrRay), 128);

    const CtrChunk ch = p.sel ? p.chunks[p.pos[p.sel[blockIdx.x]]] : p.chunks[blockIdx.x];
    const int cls = ch.cls, first = ch.first, cnt = ch.cnt;
    // Launch order: detector chunks centre first, image super-groups inside.  The rays through the middle of the image
    // cross the most pixels and the chunks at the detector's ends see little of it, so the expensive CTAs start early
    // and the last wave is made of cheap ones (longest-processing-time-first).  Matters for grids of a few waves -- a
    // rank's angle block: 90 of C4's 720 angles = 480 CTAs, 0.898 -> 0.847 ms on the slowest block, 6.75 -> 6.37 ms
    // summed over the 8 blocks -- and is free at 25 waves (5.65 ms either way).
    const int L = blockIdx.y + gridDim.y * blockIdx.z, jr = L / (int)gridDim.z, g = L - jr * (int)gridDim.z;
    const int jz = (jr & 1) ? ((int)gridDim.y - 1) / 2 + (jr + 1) / 2 : ((int)gridDim.y - 1) / 2 - jr / 2;
    const CtrClassGeom geom = cls ? p.geom[1] : p.geom[0];  // static indices: stays in registers
    const int R = ch.R;
    // column-windowed strips (ch.wc > 0): a strip row holds ch.wc pixels starting at column cst[b]
    const int Us = ch.wc > 0 ? ch.wc : geom.Up;              // row stride of the strip buffers (pixels)
    const int strip_floats = (R + 1) * Us * REC;
    int* cst = reinterpret_cast<int*>(smem_raw + 64);                      // [S] window start of the strip in buffer b
    float* buf0 = reinterpret_cast<float*>(smem_raw + 128 + rays_bytes);   // S strip buffers, strip_floats apart
    const int K = (geom.Vp + R - 1) / R;
    const float* pkg = (cls ? p.pk[1] : p.pk[0]) + (size_t)g * geom.Vp * geom.Up * REC;

    for (int k = tid; k < cnt; k += blockDim.x) rays_s[k] = p.rays[first + k];
    if (tid == 0) {
        for (int b = 0; b < S; ++b) {
            mbar_init(&full[b], 1);
            mbar_init(&empty[b], nconsumers / 32);
        }
        fence_barrier_init();
    }
    __syncthreads();

    // ---- producer warp: strip k goes to buffer k%S as soon as every consumer warp has released
    // strip k-S.  No CTA-wide barrier in the loop: a warp that finishes a strip early moves on to the
    // next (already resident) one, so the ragged ends of the strips overlap instead of idling the SM.
    if (producer) {
        const int lane = tid & 31;
        if (ch.wc == 0) {
            if (lane == 0) {
                for (int k = 0, b = 0, use = 0; k < K; ++k) {      // b = k % S, use = k / S
                    if (use > 0) mbar_wait(&empty[b], (uint32_t)((use - 1) & 1));
                    const int rows = min(R + 1, geom.Vp - k * R);
                    const uint32_t bytes = (uint32_t)rows * geom.Up * REC * 4u;
                    mbar_arrive_expect_tx(&full[b], bytes);
                    bulk_g2s(buf0 + (size_t)b * strip_floats, pkg + (size_t)k * R * geom.Up * REC, bytes, &full[b]);
                    if (++b == S) { b = 0; ++use; }
                }
            }
        } else {
            // windowed: the whole warp takes part, lane r copies packed row r of the strip (R + 1 <= 32).
            // The window start follows the CTA's rays: lane q < cnt holds the line family of ray q.
            const float jlo = (float)(jz * JW), jhi = (float)min(p.W - 1, jz * JW + JW - 1);
            CtrWinCoef co{0.f, 0.f, 0.f};
            if (lane < cnt) co = ctr_win_coef(rays_s[lane]);
            const uint32_t row_bytes = (uint32_t)Us * REC * 4u;
            for (int k = 0, b = 0, use = 0; k < K; ++k) {
                float umin = 3.0e38f, umax = -3.0e38f;
                if (lane < cnt)
                    ctr_win_range(co, jlo, jhi, (float)(k * R + geom.offv) - 0.5f, (float)(k * R + geom.offv + R), geom.ulo,
                                  geom.uhi, umin, umax);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) umin = fminf(umin, __shfl_xor_sync(0xffffffffu, umin, o));
                const int c0 = ctr_win_start(umin, geom.offu, geom.Up, Us);
                const int rows = min(R + 1, geom.Vp - k * R);
                if (lane == 0) {
                    if (use > 0) mbar_wait(&empty[b], (uint32_t)((use - 1) & 1));
                    cst[b] = c0;
                    mbar_arrive_expect_tx(&full[b], (uint32_t)rows * row_bytes);   // release: publishes cst[b]
                }
                __syncwarp();
                if (lane < rows)
                    bulk_g2s(buf0 + (size_t)b * strip_floats + (size_t)lane * Us * REC,
                             pkg + ((size_t)(k * R + lane) * geom.Up + c0) * REC, row_bytes, &full[b]);
                if (++b == S) { b = 0; ++use; }
            }
        }
    } else {
        // ---- consumers.  per-ray state: next step and steps left (coefficients are re-read per strip)
        const int j = jz * JW + tx, lbase = ty * KA;
        float ri[KA];
        int rn[KA];
        float acc[KA][NB];
#pragma unroll
        for (int q = 0; q < KA; ++q) {
            const int la = lbase + q;
            ri[q] = 0.f;
            rn[q] = 0;
            if (la < cnt && j < p.W) {
                CtrRayState s;
                ctr_ray_begin(rays_s[la], geom, j, p.H, s);
                ri[q] = s.fi;
                rn[q] = s.n;
            }
#pragma unroll
            for (int n = 0; n < NB; ++n) acc[q][n] = 0.f;
        }

        for (int k = 0, b = 0, use = 0; k < K; ++k) {
            mbar_wait(&full[b], (uint32_t)(use & 1));
            const float* strip = buf0 + (size_t)b * strip_floats + gsub * NB;
            const float vend = (float)((k + 1) * R + geom.offv);
            const int rbase = k * R + geom.offv;
            const int offu = geom.offu + (ch.wc > 0 ? cst[b] : 0);   // first packed column held by this strip
#pragma unroll
            for (int q = 0; q < KA; ++q) {
                if (rn[q] > 0) {
                    const CtrRay r = rays_s[lbase + q];
                    CtrRayState s;
                    s.pu = CTR_MUL(r.u0, (float)j);
                    s.pv = CTR_MUL(r.v0, (float)j);
                    s.fi = ri[q];
                    s.n = rn[q];
                    s.dfi = (r.v1 >= 0.f) ? 1.f : -1.f;
                    if (REUSE && NB >= 8 && INTERP == CTR_BILINEAR) ctr_march_reuse<NB, REC>(strip, Us, vend, rbase, offu, r, s, acc[q], swz);
                    else ctr_march<NB, INTERP, REC>(strip, Us, vend, rbase, offu, r, s, acc[q], swz);
                    ri[q] = s.fi;
                    rn[q] = s.n;
                }
            }
            __syncwarp();
            if ((tid & 31) == 0) mbar_arrive(&empty[b]);       // this warp is done reading the buffer
            if (++b == S) { b = 0; ++use; }
        }

        float lsum[NB];
#pragma unroll
        for (int n = 0; n < NB; ++n) lsum[n] = 0.f;
#pragma unroll
        for (int q = 0; q < KA; ++q) {
            const int la = lbase + q;
            if (la < cnt && j < p.W) {
                // sinogram row; column of mask / measurement (a subset gathers them at the plan's angle index, like
                // the reference's tf.gather at angles_i, helper_functions.py:355-357)
                const int a = p.sel ? (int)blockIdx.x : rays_s[la].angle;
                const int ao = p.sel ? rays_s[la].angle : ((EPI && p.amap) ? p.amap[a] : a);
#pragma unroll
                for (int n = 0; n < NB; ++n) {
                    const int b = (g * DEPTH + gsub) * NB + ctr_img_of_reg<NB>(n, swz);   // the loads are swizzled / rotated
                    if (b < p.B) {
                        float outv = acc[q][n];
                        if (EPI) {
                            const size_t ma = (size_t)b * p.A_all + ao;
                            float lp;
                            ctr_loglik_term(acc[q][n], __ldg(p.mask + ma), __ldg(p.meas + ma * p.W + j), p.pnm, p.sqrt_reg, lp, outv);
                            lsum[n] += lp;
                        }
                        p.sino[((size_t)b * p.A + a) * p.W + j] = outv;
                    }
                }
            }
        }
        if (EPI) {
            if (NB == 8 && swz) {   // back to image order before lanes of different parity are combined
#pragma unroll
                for (int n = 0; n < 4; ++n) { const float t = lsum[n]; lsum[n] = lsum[n + 4]; lsum[n + 4] = t; }
            }
            if (NB == 16) {         // undo the rotation: register n holds image (n + swz) & 15
                if (swz & 4) {      // rotate up by 4
#pragma unroll
                    for (int n = 0; n < 4; ++n) {
                        const float t = lsum[(12 + n) % NB];
                        lsum[(12 + n) % NB] = lsum[(8 + n) % NB]; lsum[(8 + n) % NB] = lsum[(4 + n) % NB];
                        lsum[(4 + n) % NB] = lsum[n % NB]; lsum[n % NB] = t;
                    }
                }
                if (swz & 8) {      // rotate by 8
#pragma unroll
                    for (int n = 0; n < 8; ++n) { const float t = lsum[n % NB]; lsum[n % NB] = lsum[(n + 8) % NB]; lsum[(n + 8) % NB] = t; }
                }
            }
            // warp-level part of the deterministic log-likelihood reduction (lanes of equal gsub)
#pragma unroll
            for (int n = 0; n < NB; ++n) {
                float v = lsum[n];
#pragma unroll
                for (int o = 16; o >= DEPTH; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
                lsum[n] = v;
            }
        }
        if (EPI) {
            // every consumer warp has left the strip buffers only after this barrier (named barrier 1,
            // consumers only: the producer warp has nothing to add and may already have exited)
            asm volatile("bar.sync 1, %0;" ::"r"(nconsumers) : "memory");
            float* red = buf0;
            const int lane = tid & 31, warp = tid >> 5, nwarps = nconsumers >> 5;
            if (lane < DEPTH) {
#pragma unroll
                for (int n = 0; n < NB; ++n) red[(warp * DEPTH + lane) * NB + n] = lsum[n];
            }
            asm volatile("bar.sync 1, %0;" ::"r"(nconsumers) : "memory");
            if (tid < REC) {   // tid = gsub * NB + n
                float v = 0.f;
                for (int w = 0; w < nwarps; ++w) v += red[w * REC + tid];
                const size_t cta = (size_t)jz * gridDim.x + blockIdx.x;
                p.partial[cta * ((size_t)gridDim.z * REC) + (size_t)g * REC + tid] = v;
            }
        }
    }
}

// loglik[b] = sum over CTAs of partial[cta][b]  (float64 accumulation, fixed order -> deterministic)
__global__ void __launch_bounds__(128) ctr_loglik_reduce_kernel(const float* __restrict__ partial, int nctas, int stride, int B,
                                                                float* __restrict__ loglik)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double acc = 0.0;
    for (int c = 0; c < nctas; ++c) acc += (double)partial[(size_t)c * stride + b];
    loglik[b] = (float)acc;
}

// ------------------------------------------------------------------------------------------ K2 / K2' / K3b
struct BpParams {
    const float* spk;      // packed sinogram [G][A][NB/4][W+2][4]
    const float* table;    // [A][8] forward (EXACT) or inverted (TF) transforms; unused for FBP
    const double* cs;      // [A][2] cos/sin(theta) for FBP
    float* out;            // [B][X][Y]
    int B, A, X, Y, H, W, padx, pady;
    int win;               // bins staged per angle (<= W+2)
    float scale;           // 1, or pi/(2A) for FBP
    // angle subset (training's angle minibatch, helper_functions.py:355-357): row a of the packed sinogram is angle
    // sel[a] of the plan's tables; null = identity
    const int* sel;
    // angle-sharded exchange: with nranks > 1 the tile of image b is stored straight into the exchange
    // buffer of the rank that owns b, slot `rank` (peer memory over NVLink), instead of `out`
    CtrExchange xg;
};

// bins a 32 x TH pixel tile can touch for one angle: |u| extent sqrt(31^2+(TH-1)^2) + 6 bins of slack
__host__ __device__ constexpr int bp_win(int TH) { return TH <= 8 ? 40 : 44; }
// angles staged per batch: 4 for the small 32x4 tiles (4 CTAs per SM must fit in shared memory)
__host__ __device__ constexpr int bp_ab(int TH, int NB = 16, int MINB = 2) { return (TH <= 4 || (NB >= 32 && MINB >= 3)) ? 4 : kBpAB; }

// One CTA = 32 x TH pixel tile x image group of NB (TH consumer warps) + one producer warp.
// Angles are processed in batches of AB: lanes 0..AB-1 of the producer warp each own one
// angle of the batch -- they compute the bin window the tile needs, publish its start + the
// angle's coefficients to shared memory, and issue the window's TMA bulk copies (one per
// 4-image plane) -- while the consumer warps gather from the other batch buffer and release
// it warp by warp.  No atomics anywhere: each thread owns its pixel's NB accumulators, and
// the per-pixel geometry is shared by the NB images.
template <int NB, int TH, int MINB, int MODE, int INTERP>
__global__ void __launch_bounds__(kBpTW * (TH + 1), MINB) ctr_bp_kernel(const BpParams p)
{
    constexpr int TW = kBpTW, AB = bp_ab(TH, NB, MINB), NBP = NB / 4;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);            // [2] batch landed (AB producer lanes + their TMA bytes)
    uint64_t* empty = full + 2;                                        // [2] batch consumed (one arrive per consumer warp)
    int* jb = reinterpret_cast<int*>(smem_raw + 32);                   // [2][AB]
    float* tbl = reinterpret_cast<float*>(smem_raw + 128);             // [2][AB][8]
    double* css = reinterpret_cast<double*>(smem_raw + 128 + 2 * AB * 8 * 4);  // [2][AB][2]
    float* wins = reinterpret_cast<float*>(smem_raw + 128 + 2 * AB * 8 * 4 + 2 * AB * 2 * 8);  // [2][AB][NBP][win][4]

    const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * TW + tx;
    const int c0 = blockIdx.x * TW, r0 = blockIdx.y * TH, g = blockIdx.z;
    const int c = c0 + tx, r = r0 + ty;
    const int Wp2 = p.W + 2;
    const int win = p.win;
    const int pstride = win * 4;
    const int nbatch = (p.A + AB - 1) / AB;
    const float px = (float)(c + p.pady), py = (float)(r + p.padx);
    // FBP pixel coordinates (fbp_tensorflow.py:52-53): x' = row - x_size/2, y' = col - y_size/2
    const double xpr = (double)r - 0.5 * (double)p.X, ypr = (double)c - 0.5 * (double)p.Y, half_w = 0.5 * (double)p.W;

    if (tid == 0) {
        mbar_init(&full[0], AB);
        mbar_init(&full[1], AB);
        mbar_init(&empty[0], TH);
        mbar_init(&empty[1], TH);
        fence_barrier_init();
    }
    __syncthreads();

    auto produce = [&](int nb, int lane) {
        const int s = nb & 1;
        const int a = nb * AB + lane;
        uint64_t* bar = &full[s];
        if (a >= p.A) { mbar_arrive(bar); return; }
        const int at = p.sel ? p.sel[a] : a;      // row of the plan's tables
        float cu[4];
        if (MODE == CTR_ADJ_FBP) {
            const double co = p.cs[2 * at], si = p.cs[2 * at + 1];
            css[(s * AB + lane) * 2 + 0] = co;
            css[(s * AB + lane) * 2 + 1] = si;
            int q = 0;
            for (int cy = 0; cy < 2; ++cy)
                for (int cx = 0; cx < 2; ++cx) {
                    const double xr = (double)(r0 + cy * (TH - 1)) - 0.5 * (double)p.X;
                    const double yr = (double)(c0 + cx * (TW - 1)) - 0.5 * (double)p.Y;
                    cu[q++] = (float)(yr * co - xr * si + 0.5 * (double)p.W);
                }
        } else {
            float t[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) t[q] = p.table[8 * at + q];
#pragma unroll
            for (int q = 0; q < 8; ++q) tbl[(s * AB + lane) * 8 + q] = t[q];
            int q = 0;
            for (int cy = 0; cy < 2; ++cy)
                for (int cx = 0; cx < 2; ++cx) {
                    const float qx = (float)(c0 + cx * (TW - 1) + p.pady), qy = (float)(r0 + cy * (TH - 1) + p.padx);
                    float uj, vi;
                    if (MODE == CTR_ADJ_EXACT) ctr_adj_centre(t, qx, qy, uj, vi);
                    else uj = CTR_ADD(CTR_ADD(CTR_MUL(t[0], qx), CTR_MUL(t[1], qy)), t[2]);
                    cu[q++] = uj;
                }
        }
        const int start = ctr_window_start(cu[0], cu[1], cu[2], cu[3], Wp2, win);
        jb[s * AB + lane] = start;
        const uint32_t bytes = (uint32_t)win * 16u;
        mbar_arrive_expect_tx(bar, bytes * NBP);   // release: publishes jb/tbl/css to the waiters
#pragma unroll
        for (int h = 0; h < NBP; ++h)
            bulk_g2s(wins + ((size_t)(s * AB + lane) * NBP + h) * pstride,
                     p.spk + ((((size_t)g * p.A + a) * NBP + h) * Wp2 + start) * 4, bytes, bar);
    };

    // ---- producer warp (threadIdx.y == TH): lane k < AB owns angle k of every batch.  A batch buffer is
    // refilled as soon as every consumer warp has released it -- no CTA-wide barrier in the angle loop (r1 ncu:
    // one warp in four sat at the __syncthreads of the previous version).
    if (ty == TH) {
        if (tx < AB) {
            for (int nb = 0; nb < nbatch; ++nb) {
                if (nb >= 2) mbar_wait(&empty[nb & 1], (uint32_t)(((nb >> 1) - 1) & 1));
                produce(nb, tx);
            }
        }
        return;
    }

    float acc[NB];
#pragma unroll
    for (int n = 0; n < NB; ++n) acc[n] = 0.f;

    for (int nb = 0; nb < nbatch; ++nb) {
        const int s = nb & 1;
        mbar_wait(&full[s], (uint32_t)((nb >> 1) & 1));
        const int na = min(AB, p.A - nb * AB);
        for (int k = 0; k < na; ++k) {
            const float* ywin = wins + (size_t)(s * AB + k) * NBP * pstride;
            const int start = jb[s * AB + k];
            if (MODE == CTR_ADJ_FBP) {
                ctr_adj_fbp<NB>(&css[(s * AB + k) * 2], p.W, half_w, xpr, ypr, ywin, pstride, start, acc);
            } else {
                float t[8];
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const float4 tv = reinterpret_cast<const float4*>(tbl + (s * AB + k) * 8)[q];
                    t[4 * q] = tv.x; t[4 * q + 1] = tv.y; t[4 * q + 2] = tv.z; t[4 * q + 3] = tv.w;
                }
                if (MODE == CTR_ADJ_EXACT) ctr_adj_exact<NB, INTERP>(t, p.H, p.W, px, py, ywin, pstride, start, acc);
                else ctr_adj_tf<NB, INTERP>(t, p.H, p.W, px, py, ywin, pstride, start, acc);
            }
        }
        __syncwarp();
        if (tx == 0) mbar_arrive(&empty[s]);       // this warp is done reading the batch buffers
    }

    if (r < p.X && c < p.Y) {
        if (p.xg.nranks <= 1) {
#pragma unroll
            for (int n = 0; n < NB; ++n) {
                const int b = g * NB + n;
                if (b < p.B) p.out[((size_t)b * p.X + r) * p.Y + c] = acc[n] * p.scale;
            }
        } else {
            // fused exchange: this rank's partial of image b goes to slot `rank` of b's owner (a 128-byte row
            // segment per warp and image, written over NVLink while other tiles are still being computed)
#pragma unroll
            for (int n = 0; n < NB; ++n) {
                const int b = g * NB + n;
                if (b < p.B) {
                    const int owner = ctr_xg_owner(b, p.xg.Bs);
                    p.xg.peer[owner][ctr_xg_index(p.xg.rank, b - owner * p.xg.Bs, p.xg.Bs, r, c, p.X, p.Y)] = acc[n] * p.scale;
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------ exchange: barrier + sum
// Second half of the fused angle-sharded adjoint.  Every rank's ctr_bp_kernel has stored its partial images into the
// owners' exchange buffers; this kernel (stream-ordered after it) tells every peer "my stores are done", waits until
// every peer has said the same, and sums the nranks slots of this rank's images in rank order (deterministic).
//   flags      this rank's flag words [nranks]: flags[s] = last epoch rank s has completed
//   peer_flags the same array of every rank (peer memory)
// A peer that does not arrive within timeout_ns raises *err (ctr_comm_check reports CTR_ECOMM) instead of hanging.
struct XchgParams {
    unsigned* peer_flags[CTR_MAX_RANKS];
    unsigned* flags;
    int* err;
    const float* slots;     // [nranks][n] this rank's exchange buffer (current parity)
    float* out;             // [n] = Bs * X * Y
    size_t n;
    unsigned epoch;
    int nranks, rank;
    unsigned long long timeout_ns;
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p)
{
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__global__ void __launch_bounds__(512) ctr_xchg_sum_kernel(const XchgParams p)
{
    const int tid = threadIdx.x;
    if (blockIdx.x == 0 && tid < p.nranks) {
        __threadfence_system();                              // the adjoint kernel's peer stores (stream order) before the flag
        st_release_sys(p.peer_flags[tid] + p.rank, p.epoch);
    }
    if (tid < p.nranks) {
        const unsigned long long t0 = global_ns();
        while ((int)(ld_acquire_sys(p.flags + tid) - p.epoch) < 0) {
            if (global_ns() - t0 > p.timeout_ns) { atomicExch(p.err, 1 + tid); break; }
            __nanosleep(200);
        }
    }
    __syncthreads();
    const size_t stride = (size_t)gridDim.x * blockDim.x, i0 = (size_t)blockIdx.x * blockDim.x + tid;
    if ((p.n & 3) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0) {   // 16-byte aligned slots and result: 128-bit accesses
        const size_t n4 = p.n / 4;
        const float4* s4 = reinterpret_cast<const float4*>(p.slots);
        float4* o4 = reinterpret_cast<float4*>(p.out);
        for (size_t i = i0; i < n4; i += stride) {
            float4 a = __ldcg(s4 + i);
            for (int s = 1; s < p.nranks; ++s) {
                const float4 b = __ldcg(s4 + (size_t)s * n4 + i);
                a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
            }
            o4[i] = a;
        }
    } else {
        for (size_t i = i0; i < p.n; i += stride) {
            float a = __ldcg(p.slots + i);
            for (int s = 1; s < p.nranks; ++s) a += __ldcg(p.slots + (size_t)s * p.n + i);
            p.out[i] = a;
        }
    }
}

// ------------------------------------------------------------------------------------------ K3a filter
// rf[n] = sum_k s[k] * h[(n-k) mod P]  ==  real(ifft(fft(s) * filter_1d))   for real s
// (fbp_tensorflow.py:49-50; no zero padding, so the convolution is circular).
// grid (A, G), block = ceil(P/2) threads rounded up to a warp (at most 256).  smem: NB rows interleaved [P][NB] +
// doubled kernel h2[2P].  A thread computes TWO adjacent bins for all NB images: per k it reads the NB sinogram values
// once (broadcast LDS.128) and ONE new kernel value -- the second bin's h at step k is the first bin's h at step k-1 --
// so the loop is 2*NB FFMA per NB/4 + 1 shared-memory loads (r1's one-bin-per-thread loop was bound by those loads, and
// 72 of its 256 threads had no bin at P = 184: 0.48 -> see DESIGN.md section 4).  k ascending, fmaf(h, s, acc): the
// same sums, bit for bit, as the fused kernel's filter stage.
// Output goes straight into the plane-layout sinogram pack K3b reads.

// Row (angle a) of the NB images of filter group g inside a pack whose image groups hold nbg >= NB images (the gather
// may take 32 images per thread while the filter works on 16): [nbg-group][A][nbg/4 planes][P+2][4].
template <int NB>
__device__ __forceinline__ float* filt_dst_row(float* spk, int g, int a, int A, int P, int nbg)
{
    const int first = g * NB, grp = first / nbg, plane0 = (first - grp * nbg) / 4;
    return spk + (((size_t)grp * A + a) * (nbg / 4) + plane0) * (size_t)(P + 2) * 4;
}

template <int NB>
__global__ void __launch_bounds__(256) ctr_fbp_filter_kernel(const float* __restrict__ sino, const float* __restrict__ h,
                                                             int B, int A, int P, int nbg, float* __restrict__ spk)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* s = reinterpret_cast<float*>(smem_raw);   // [P][NB]
    float* h2 = s + (size_t)P * NB;                   // [2P] (+ 1 pad)
    const int a = blockIdx.x, g = blockIdx.y;
    for (int idx = threadIdx.x; idx < P * NB; idx += blockDim.x) {
        const int n = idx / P, k = idx - n * P;       // coalesced along k per image
        const int b = g * NB + n;
        s[k * NB + n] = (b < B) ? __ldg(sino + ((size_t)b * A + a) * P + k) : 0.f;
    }
    for (int m = threadIdx.x; m <= 2 * P; m += blockDim.x) h2[m] = (m < 2 * P) ? __ldg(h + (m >= P ? m - P : m)) : 0.f;
    __syncthreads();
    float* dst_row = filt_dst_row<NB>(spk, g, a, A, P, nbg);
    for (int n0 = 2 * threadIdx.x; n0 < P; n0 += 2 * blockDim.x) {
        float a0[NB], a1[NB];
#pragma unroll
        for (int n = 0; n < NB; ++n) { a0[n] = 0.f; a1[n] = 0.f; }
        const float* hp = h2 + n0 + P;
        float h_hi = hp[1];                           // h2[n0 + 1 + P - k] at k = 0
        for (int k = 0; k < P; ++k) {
            const float h_lo = hp[-k];
            float sv[NB];
            ctr_ldv<NB>(s + (size_t)k * NB, sv);
#pragma unroll
            for (int n = 0; n < NB; ++n) {
                a0[n] = fmaf(h_lo, sv[n], a0[n]);
                a1[n] = fmaf(h_hi, sv[n], a1[n]);
            }
            h_hi = h_lo;
        }
#pragma unroll
        for (int q = 0; q < NB / 4; ++q) {
            float* d = dst_row + ((size_t)q * (P + 2) + n0 + 1) * 4;
            *reinterpret_cast<float4*>(d) = make_float4(a0[4 * q], a0[4 * q + 1], a0[4 * q + 2], a0[4 * q + 3]);
            if (n0 + 1 < P) *reinterpret_cast<float4*>(d + 4) = make_float4(a1[4 * q], a1[4 * q + 1], a1[4 * q + 2], a1[4 * q + 3]);
        }
    }
    if (threadIdx.x < 2 * (NB / 4)) {  // halo bins (never read by the FBP gather; keep them defined)
        const int q = threadIdx.x >> 1;
        *reinterpret_cast<float4*>(dst_row + ((size_t)q * (P + 2) + ((threadIdx.x & 1) ? P + 1 : 0)) * 4) =
            make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

// ---- K3a', the ramp filter.  real(ifft(ramp)) is zero at every even offset except 0 (P even: the detector of
// pad_phantom always is), so an output bin of parity pn only needs the input bins of the OTHER parity plus its own
// bin times h[0]: half the multiply-adds of the dense loop above.  The row is staged de-interleaved by parity,
// s[par][P/2][NB]; a thread owns BPT outputs of ONE parity (n = 2(a0+q)+pn) and a warp is parity-uniform, so per input
// bin the whole warp reads the same NB values (broadcast LDS.128) and every thread slides its window of odd taps
// along by one register (one new LDS.32 per step): BPT*NB FFMA per NB/4 + 1 loads.
//   hs      [1 + P + kFiltBPTMax] odd taps, doubled and shifted by one: hs[1 + i] = h[(2i + 1) mod P], i < P
//   sum     k ascending over the other parity, fmaf(h, s, acc), then fmaf(h[0], s[n], acc): the fused kernel's filter
//           stage calls the same routine, so the two paths stay bit-identical.
constexpr int kFiltBPTMax = 4;
template <int NBQ, int BPT>
__device__ __forceinline__ void ctr_filter_sparse_taps(const float* __restrict__ so, const float* __restrict__ ss, int sstride,
                                                       const float* __restrict__ hs, float h0, int Ph, int a0, int pn,
                                                       float (&acc)[BPT][NBQ])
{
    const float* hp = hs + 1 + (a0 - (1 - pn) + Ph);   // tap of output a0 at k2 = 0; output q uses hp[q - k2]
    float hr[BPT];
#pragma unroll
    for (int q = 0; q < BPT; ++q) {
        hr[q] = hp[q];
#pragma unroll
        for (int n = 0; n < NBQ; ++n) acc[q][n] = 0.f;
    }
    for (int k2 = 0; k2 < Ph; ++k2) {
        float sv[NBQ];
        ctr_ldv<NBQ>(so + (size_t)k2 * sstride, sv);
#pragma unroll
        for (int q = 0; q < BPT; ++q)
#pragma unroll
            for (int n = 0; n < NBQ; ++n) acc[q][n] = fmaf(hr[q], sv[n], acc[q][n]);
#pragma unroll
        for (int q = BPT - 1; q > 0; --q) hr[q] = hr[q - 1];
        hr[0] = hp[-k2 - 1];                           // hs[0] is the pad the last step reads
    }
#pragma unroll
    for (int q = 0; q < BPT; ++q) {
        if (a0 + q >= Ph) continue;
        float sv[NBQ];
        ctr_ldv<NBQ>(ss + (size_t)(a0 + q) * sstride, sv);
#pragma unroll
        for (int n = 0; n < NBQ; ++n) acc[q][n] = fmaf(h0, sv[n], acc[q][n]);
    }
}

// grid (A, G), block = 2 parities x wper warps (<= 256 threads): one sinogram row of NB images per CTA.  (Several rows
// per CTA were measured and lost -- C5, 1000 x 180 rows of 184 bins: 4 rows 0.265 ms, 2 rows 0.228, 1 row 0.208: small
// CTAs overlap one CTA's global loads with another's multiply-adds.)
template <int NB, int BPT>
__global__ void __launch_bounds__(256) ctr_fbp_filter_sparse_kernel(const float* __restrict__ sino, const float* __restrict__ hs,
                                                                    float h0, int B, int A, int P, int wper, int nbg,
                                                                    float* __restrict__ spk)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int Ph = P / 2;
    float* s = reinterpret_cast<float*>(smem_raw);            // [2][Ph][NB]
    float* hsm = s + (size_t)P * NB;                           // [1 + P + kFiltBPTMax]
    const int a = blockIdx.x, g = blockIdx.y, tid = threadIdx.x;
    // stage the row of the NB images: lanes = (2 neighbouring quads of bins) x (16 images) -- a full 32-byte sector
    // per image row and warp load, and shared-memory stores with at most 2-way bank conflicts
    if ((P & 3) == 0 && (reinterpret_cast<uintptr_t>(sino) & 15) == 0) {
        const int Pq = P / 4;
        for (int idx = tid; idx < Pq * NB; idx += blockDim.x) {
            const int n = idx % NB, kq = idx / NB;
            const int b = g * NB + n;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (b < B) v = __ldg(reinterpret_cast<const float4*>(sino + ((size_t)b * A + a) * P) + kq);
            float* d = s + (size_t)(2 * kq) * NB + n;                            // bins 4kq, 4kq+2 -> even[2kq], even[2kq+1]
            d[0] = v.x; d[NB] = v.z;
            d[(size_t)Ph * NB] = v.y; d[(size_t)Ph * NB + NB] = v.w;             // bins 4kq+1, 4kq+3 -> odd[2kq], odd[2kq+1]
        }
    } else {
        for (int idx = tid; idx < P * NB; idx += blockDim.x) {
            const int n = idx % NB, k = idx / NB;
            const int b = g * NB + n;
            s[((size_t)(k & 1) * Ph + (k >> 1)) * NB + n] = (b < B) ? __ldg(sino + ((size_t)b * A + a) * P + k) : 0.f;
        }
    }
    for (int m = tid; m < 1 + P + kFiltBPTMax; m += blockDim.x) hsm[m] = __ldg(hs + m);
    __syncthreads();
    const int warp = tid >> 5, lane = tid & 31;
    const int pn = warp / wper, a0 = ((warp % wper) * 32 + lane) * BPT;
    float* dst_row = filt_dst_row<NB>(spk, g, a, A, P, nbg);
    if (a0 < Ph) {
        float acc[BPT][NB];
        ctr_filter_sparse_taps<NB, BPT>(s + (size_t)(1 - pn) * Ph * NB, s + (size_t)pn * Ph * NB, NB, hsm, h0, Ph, a0, pn, acc);
#pragma unroll
        for (int q = 0; q < BPT; ++q) {
            if (a0 + q >= Ph) continue;
            const int n0 = 2 * (a0 + q) + pn;
#pragma unroll
            for (int h = 0; h < NB / 4; ++h)
                *reinterpret_cast<float4*>(dst_row + ((size_t)h * (P + 2) + n0 + 1) * 4) =
                    make_float4(acc[q][4 * h], acc[q][4 * h + 1], acc[q][4 * h + 2], acc[q][4 * h + 3]);
        }
    }
    if (warp == 0 && lane < 2 * (NB / 4)) {  // halo bins (never read by the FBP gather; keep them defined)
        const int h = lane >> 1;
        *reinterpret_cast<float4*>(dst_row + ((size_t)h * (P + 2) + ((lane & 1) ? P + 1 : 0)) * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

// launch shape of the ramp filter: outputs per thread (fewest idle lanes, then fewest accumulators) and warps per parity
struct FiltShape { int bpt, wper; size_t smem; };
inline FiltShape fbp_filter_sparse_shape(int P, int NB)
{
    FiltShape f{2, 1, 0};
    const int Ph = P / 2;
    int best = 1 << 30;
    for (int bpt : {3, 2, 4}) {
        if (bpt * NB > 48) continue;                                  // accumulators per thread
        const int tasks = (Ph + bpt - 1) / bpt, wp = (tasks + 31) / 32;
        if (2 * wp * 32 > 256) continue;
        const int idle = wp * 32 * bpt - Ph;
        if (idle < best) { best = idle; f.bpt = bpt; f.wper = wp; }
    }
    f.smem = ((size_t)P * NB + 1 + P + kFiltBPTMax) * sizeof(float);
    return f;
}

template <int NB>
inline cudaError_t launch_fbp_filter_sparse(const float* sino, const float* hs, float h0, int B, int A, int P, int nbg, float* spk,
                                            cudaStream_t st)
{
    const FiltShape f = fbp_filter_sparse_shape(P, NB);
    const int G = (B + NB - 1) / NB;
    dim3 grid(A, G), block(2 * f.wper * 32);
    cudaError_t e = cudaSuccess;
#define CTR_FILT_CASE(BPT)                                                                                                     \
    case BPT:                                                                                                                  \
        e = cudaFuncSetAttribute(ctr_fbp_filter_sparse_kernel<NB, BPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)f.smem); \
        if (e != cudaSuccess) return e;                                                                                        \
        ctr_fbp_filter_sparse_kernel<NB, BPT><<<grid, block, f.smem, st>>>(sino, hs, h0, B, A, P, f.wper, nbg, spk);           \
        break;
    switch (f.bpt) {
        CTR_FILT_CASE(2)
        CTR_FILT_CASE(3)
        CTR_FILT_CASE(4)
        default: return cudaErrorInvalidValue;
    }
#undef CTR_FILT_CASE
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------ K3 fused FBP
// iradon in ONE kernel (fbp_tensorflow.py:49-74): the circular row filter runs in shared memory and its output never
// leaves the chip.  A thread-block CLUSTER of CL CTAs serves one group of 16 sinograms:
//   filter   the rows of an angle batch are dealt out to the CTAs as "row quads" (angle k, 4 images): each CTA loads its
//            quads' raw rows, convolves them with the spatial kernel (no redundant work anywhere in the cluster) and
//            stores every filtered bin into the batch buffer of ALL CTAs of the cluster through distributed shared
//            memory (st.shared::cluster; 16 bytes per bin and quad: the plane layout the gather reads);
//   gather   every CTA owns 1/CL of the image's pixels (4 per thread, 64 accumulators); it back-projects the batch from
//            its own shared memory exactly like ctr_bp_kernel<FBP> (same ctr_adj_fbp, same angle order: bit-identical).
// The batch buffers are double buffered: F(b+1) then BP(b), one cluster barrier per batch.
// Fits images of up to 8 * 2048 pixels (128 x 128, the foam dataset: BASELINE configs[4]); larger ones take the
// two-kernel path (ctr_fbp_filter_kernel + ctr_bp_kernel<FBP>).
constexpr int kFusedNB = 16, kFusedPPT = 4, kFusedThreads = 512, kFusedPxPerCta = kFusedThreads * kFusedPPT;
struct FbpFusedParams {
    const float* sino;     // [B][A][P]
    const float* h;        // [P] spatial kernel real(ifft(filter_1d))
    const float* hs;       // ramp filter (sparse != 0): the odd taps, doubled (see ctr_filter_sparse_taps), and h[0]
    float h0;
    int sparse;
    const double* cs;      // [A][2] cos / sin(theta)
    float* out;            // [B][X][Y]
    int B, A, P, X, Y;
    int AB;                // angles per batch (1, 2, 4 or 8; chosen so the buffers fit shared memory)
    float scale;           // pi / (2 A_total)
    CtrExchange xg;
};

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_nctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// 16-byte store into the same shared-memory offset of CTA `rank` of this cluster
__device__ __forceinline__ void st_cluster_f4(const void* local, uint32_t rank, float4 v)
{
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(local)), "r"(rank));
    asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(remote), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__host__ __device__ inline size_t fbp_fused_smem(int P, int AB, int CL)
{
    const int quads = (AB * (kFusedNB / 4) + CL - 1) / CL;                  // row quads a CTA filters per batch
    return 2ull * AB * (kFusedNB / 4) * P * 16 + (size_t)quads * P * 16 + 2ull * P * 4 + 16 + 4 * (1 + kFiltBPTMax);
}

__global__ void __launch_bounds__(kFusedThreads, 1) ctr_fbp_fused_kernel(const FbpFusedParams p)
{
    constexpr int NB = kFusedNB, NBP = NB / 4, PPT = kFusedPPT;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int P = p.P, AB = p.AB;
    const int CL = (int)cluster_nctarank(), q = (int)cluster_ctarank();
    const int g = blockIdx.x / CL, tid = threadIdx.x;
    const int quads = (AB * NBP + CL - 1) / CL;
    float* filt = reinterpret_cast<float*>(smem_raw);                        // [2][AB][NBP][P][4]
    const int stage_floats = AB * NBP * P * 4;
    float* raw = filt + 2 * stage_floats;                                    // [quads][P][4]
    float* h2 = raw + quads * P * 4;                                         // [2P], or [1 + P + kFiltBPTMax] odd taps (ramp)
    if (p.sparse) for (int m = tid; m < 1 + P + kFiltBPTMax; m += kFusedThreads) h2[m] = __ldg(p.hs + m);
    else for (int m = tid; m < 2 * P; m += kFusedThreads) h2[m] = __ldg(p.h + (m >= P ? m - P : m));
    const int Ph = P / 2, nbp = (Ph + 1) / 2;                                // ramp: pair tasks per parity

    // this thread's pixels (consecutive lanes = consecutive pixels of a row: neighbouring bins in the gather)
    const int npix = p.X * p.Y, per = (npix + CL - 1) / CL;
    const int px_lo = q * per, px_hi = min(npix, px_lo + per);
    double xpr[PPT], ypr[PPT];
    float acc[PPT][NB];
#pragma unroll
    for (int i = 0; i < PPT; ++i) {
        const int pi = min(px_lo + tid + i * kFusedThreads, npix - 1);
        const int r = pi / p.Y, c = pi - r * p.Y;
        xpr[i] = (double)r - 0.5 * (double)p.X;                              // fbp_tensorflow.py:52-53
        ypr[i] = (double)c - 0.5 * (double)p.Y;
#pragma unroll
        for (int n = 0; n < NB; ++n) acc[i][n] = 0.f;
    }
    const int nbatch = (p.A + AB - 1) / AB;
    const double half_p = 0.5 * (double)P;
    const int nb2 = (P + 1) / 2;                                             // filter tasks per quad: 2 bins each

    auto filter_batch = [&](int b) {
        float* dst = filt + (b & 1) * stage_floats;
        // raw rows of this CTA's quads: quad t = (angle k = t / NBP, plane h = t % NBP), t = q + CL * ql
        // (ramp filter: de-interleaved by bin parity, [ql][parity][P/2][4])
        for (int idx = tid; idx < quads * 4 * P; idx += kFusedThreads) {
            const int kk = idx % P, n4 = (idx / P) & 3, ql = idx / (4 * P);
            const int t = q + CL * ql, k = t / NBP, hpl = t - k * NBP;
            const int a = b * AB + k, bimg = g * NB + 4 * hpl + n4;
            const int slot = p.sparse ? (kk & 1) * Ph + (kk >> 1) : kk;
            raw[(ql * P + slot) * 4 + n4] = (t < AB * NBP && a < p.A && bimg < p.B) ? __ldg(p.sino + ((size_t)bimg * p.A + a) * P + kk) : 0.f;
        }
        __syncthreads();
        if (p.sparse) {
            for (int task = tid; task < quads * 2 * nbp; task += kFusedThreads) {
                const int ql = task / (2 * nbp), rem = task - ql * 2 * nbp, pn = rem / nbp, a0 = (rem - pn * nbp) * 2;
                const int t = q + CL * ql, k = t / NBP, hpl = t - k * NBP;
                if (t >= AB * NBP || b * AB + k >= p.A) continue;
                const float* sq = raw + (size_t)ql * P * 4;
                float acc[2][4];
                ctr_filter_sparse_taps<4, 2>(sq + (size_t)(1 - pn) * Ph * 4, sq + (size_t)pn * Ph * 4, 4, h2, p.h0, Ph, a0, pn, acc);
                float* d0 = dst + ((size_t)(k * NBP + hpl) * P + 2 * a0 + pn) * 4;
                for (int r = 0; r < CL; ++r) {
                    st_cluster_f4(d0, (uint32_t)r, make_float4(acc[0][0], acc[0][1], acc[0][2], acc[0][3]));
                    if (a0 + 1 < Ph) st_cluster_f4(d0 + 8, (uint32_t)r, make_float4(acc[1][0], acc[1][1], acc[1][2], acc[1][3]));
                }
            }
            return;
        }
        for (int task = tid; task < quads * nb2; task += kFusedThreads) {
            const int ql = task / nb2, n0 = (task - ql * nb2) * 2;
            const int t = q + CL * ql, k = t / NBP, hpl = t - k * NBP;
            if (t >= AB * NBP || b * AB + k >= p.A) continue;
            // rf[n] = sum_k s[k] * h[(n - k) mod P], k ascending like ctr_fbp_filter_kernel (bit-identical sums)
            const float* sq = raw + (size_t)ql * P * 4;
            const float* hp = h2 + n0 + P;
            float a0[4] = {0.f, 0.f, 0.f, 0.f}, a1[4] = {0.f, 0.f, 0.f, 0.f};
            float h_hi = hp[1];                                              // h2[n0 + 1 + P - k] at k = 0
            for (int kk = 0; kk < P; ++kk) {
                const float h_lo = hp[-kk];
                const float4 sv = *reinterpret_cast<const float4*>(sq + kk * 4);
                a0[0] = fmaf(h_lo, sv.x, a0[0]); a0[1] = fmaf(h_lo, sv.y, a0[1]); a0[2] = fmaf(h_lo, sv.z, a0[2]); a0[3] = fmaf(h_lo, sv.w, a0[3]);
                a1[0] = fmaf(h_hi, sv.x, a1[0]); a1[1] = fmaf(h_hi, sv.y, a1[1]); a1[2] = fmaf(h_hi, sv.z, a1[2]); a1[3] = fmaf(h_hi, sv.w, a1[3]);
                h_hi = h_lo;
            }
            float* d0 = dst + ((size_t)(k * NBP + hpl) * P + n0) * 4;
            for (int r = 0; r < CL; ++r) {
                st_cluster_f4(d0, (uint32_t)r, make_float4(a0[0], a0[1], a0[2], a0[3]));
                if (n0 + 1 < P) st_cluster_f4(d0 + 4, (uint32_t)r, make_float4(a1[0], a1[1], a1[2], a1[3]));
            }
        }
    };

    __syncthreads();
    filter_batch(0);
    cluster_sync_all();
    for (int b = 0; b < nbatch; ++b) {
        if (b + 1 < nbatch) filter_batch(b + 1);      // into the other stage, while nobody reads it
        const float* src = filt + (b & 1) * stage_floats;
        const int na = min(AB, p.A - b * AB);
        for (int k = 0; k < na; ++k) {
            double cs[2];
            cs[0] = __ldg(p.cs + 2 * (b * AB + k));
            cs[1] = __ldg(p.cs + 2 * (b * AB + k) + 1);
            const float* ywin = src + (size_t)k * NBP * P * 4;
#pragma unroll
            for (int i = 0; i < PPT; ++i) ctr_adj_fbp<NB>(cs, P, half_p, xpr[i], ypr[i], ywin, P * 4, 1, acc[i]);
        }
        cluster_sync_all();                            // F(b+1) has landed everywhere; BP(b) has finished everywhere
    }

#pragma unroll
    for (int i = 0; i < PPT; ++i) {
        const int pi = px_lo + tid + i * kFusedThreads;
        if (pi >= px_hi) continue;
        const int r = pi / p.Y, c = pi - r * p.Y;
#pragma unroll
        for (int n = 0; n < NB; ++n) {
            const int bimg = g * NB + n;
            if (bimg >= p.B) continue;
            const float v = acc[i][n] * p.scale;
            if (p.xg.nranks <= 1) {
                p.out[((size_t)bimg * p.X + r) * p.Y + c] = v;
            } else {
                const int owner = ctr_xg_owner(bimg, p.xg.Bs);
                p.xg.peer[owner][ctr_xg_index(p.xg.rank, bimg - owner * p.xg.Bs, p.xg.Bs, r, c, p.X, p.Y)] = v;
            }
        }
    }
}

// cluster size and angle batch for an image of X x Y and a detector of P bins; CL = 0: does not fit the fused kernel
inline void fbp_fused_shape(int X, int Y, int P, int smem_optin, int& CL, int& AB)
{
    CL = 0; AB = 0;
    const long long npix = (long long)X * Y;
    for (int c : {1, 2, 4, 8})
        if (npix <= (long long)c * kFusedPxPerCta) { CL = c; break; }
    if (!CL) return;
    for (int ab : {8, 4, 2, 1})
        if (fbp_fused_smem(P, ab, CL) <= (size_t)smem_optin) { AB = ab; break; }
    if (!AB) CL = 0;
}

inline cudaError_t launch_fbp_fused(const FbpFusedParams& p, int CL, cudaStream_t st)
{
    const size_t smem = fbp_fused_smem(p.P, p.AB, CL);
    cudaError_t e = cudaFuncSetAttribute(ctr_fbp_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int G = (p.B + kFusedNB - 1) / kFusedNB;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(G * CL));
    cfg.blockDim = dim3(kFusedThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, ctr_fbp_fused_kernel, p);
    launch_counter()++;
    return e != cudaSuccess ? e : cudaGetLastError();
}

// ------------------------------------------------------------------------------------------ launchers
struct FwdConfig {
    int JW, NS, KA, R, jchunks, depth, stages;
    int lanes;      // lanes per ray (each owns 4 * depth / lanes images of the pixel record)
    int reuse;      // 1: bilinear march keeps the previous bottom row in registers (ctr_march_reuse; 8-image lanes)
    int windowed;   // 1: column-windowed strips, per-chunk R and window (R and smem are filled in by ctr_plan_create)
    size_t smem;
    static constexpr int fixed_bytes(int NA) { return 128 + (NA * (int)sizeof(CtrRay) + 127) / 128 * 128; }
    int angles_per_cta() const { return NS * KA; }
};

// Tuning history (r1/r2 measurements, all on B200): two ring stages beat three or four (+3 %), pairing a central
// with an edge bin per thread lost to SIMT inefficiency at the shadow edge (C2 0.645 vs 0.600 ms), an i-synchronous
// quarter-warp march lost to its sit-out trips (C4 slice 3.95 vs 2.50 ms), 8-image records on a wide detector split
// in two were flat.  Those variants are gone from the library; DESIGN.md section 7 keeps the numbers.
constexpr int kFwdStages = 2;

// Shape of the 4-image-record forward CTA (any detector, any batch): JW detector bins x NS angle slots x KA angles
// per slot, and the largest strip height R whose double buffer fits the shared-memory budget.
inline FwdConfig fwd_config(int W, const CtrClassGeom geom[2], int smem_budget)
{
    FwdConfig c{};
    c.depth = 1;
    c.lanes = 1;
    c.KA = 2;
    c.JW = round_up(W, 32);
    if (c.JW > kFwdMaxConsumers) c.JW = kFwdMaxConsumers;
    c.jchunks = (W + c.JW - 1) / c.JW;
    c.NS = (c.JW >= 128) ? 1 : 128 / c.JW;
    const int threads = c.JW * c.NS + 32;
    const int ctas_per_sm = threads <= 256 ? 4 : (threads <= 512 ? 2 : 1);
    if (smem_budget > (228 * 1024) / ctas_per_sm - 1024) smem_budget = (228 * 1024) / ctas_per_sm - 1024;
    const int fixed = FwdConfig::fixed_bytes(c.NS * c.KA);
    const int Upmax = geom[0].Up > geom[1].Up ? geom[0].Up : geom[1].Up;
    const int Vpmax = geom[0].Vp > geom[1].Vp ? geom[0].Vp : geom[1].Vp;
    const int row_bytes = Upmax * kFwdNB * 4;
    c.stages = kFwdStages;
    int rows = (smem_budget - fixed) / (c.stages * row_bytes);   // rows per buffer = R + 1
    if (rows > Vpmax) rows = Vpmax;
    if (rows > 33) rows = 33;   // bigger strips only lengthen the un-overlapped first load
    c.R = rows - 1;
    if (c.R < 1) c.R = 0;  // caller treats 0 as "image too wide for the strip buffers"
    c.smem = (size_t)fixed + (size_t)c.stages * (size_t)(c.R + 1) * row_bytes;
    return c;
}

// Depth-first shapes: several image groups of 4 share one pixel record, the lanes of a quarter-warp are
// (rays) x (groups).  rec32 == false: 16-image records with 4 lanes per ray (detectors of <= 184 bins) or 8-image
// records with 2 lanes per ray (<= 368 bins).  rec32 == true: 32-image records, 4 lanes per ray x 8 images per lane,
// parity-swizzled loads (conflict-free quarter-warps).  Detectors too wide for whole-row strips get COLUMN-WINDOWED
// strips: the detector is cut into chunks of JW bins, NS angle slots share a CTA (neighbouring angles need nearly the
// same window), and every strip holds only the columns those rays cross; R, the windows and the shared-memory size
// then depend on the angles and are filled in by ctr_plan_create (ctr_h_build_chunks).
// win_ns > 0 forces the number of angle slots of the windowed shape (the plan retries with fewer slots when the
// angles sharing a CTA are too far apart for a window that fits).
constexpr int kFwdDepth = 4;
inline FwdConfig fwd_config_depth(int W, const CtrClassGeom geom[2], int smem_budget, bool rec32, int win_ns = 0)
{
    FwdConfig c{};
    c.lanes = (round_up(W, 8) * kFwdDepth <= kFwdMaxConsumers) ? kFwdDepth : 2;
    c.stages = kFwdStages;
    c.depth = c.lanes;                                   // 4 images per lane ...
    if (rec32) {                                         // ... or 8 (32-image records), four lanes per ray only
        if (c.lanes != 4 && round_up(W, 8) * c.lanes <= kFwdMaxConsumers) return c;   // R = 0: mid-size detectors keep 8-image records
        c.lanes = 4;
        c.depth = 8;
    }
    c.NS = 1;
    c.KA = 2;
    c.R = 0;
    c.smem = 0;
    c.jchunks = 1;
    c.JW = round_up(W, 32 / c.lanes);
    if (c.JW * c.lanes > kFwdMaxConsumers) {
        c.lanes = 4;
        c.depth = rec32 ? 8 : 4;
        c.NS = rec32 ? 4 : 2;   // r1 sweep at 64 x 512^2 x 720: 6.60 ms (NS 4) vs 6.80 (2) / 6.63 (8) with 32-image records
        if (win_ns > 0) c.NS = win_ns;
        // r1: the vertical-reuse march pays with the tall strips of the windowed shape (C4 6.57 -> 6.24 ms),
        // not with the 5-row strips of whole-row 32-image records (C2 0.521 -> 0.532 ms)
        c.reuse = rec32 ? 1 : 0;
        const int maxc = c.reuse ? kFwdReuseThreads - 32 : kFwdMaxConsumers;
        { const int q = c.NS >= 4 ? 2 : 8 / c.NS; c.JW = maxc / (c.lanes * c.NS) / q * q; }   // whole warps per CTA
        c.jchunks = (W + c.JW - 1) / c.JW;
        c.JW = round_up((W + c.jchunks - 1) / c.jchunks, c.NS >= 4 ? 2 : 8 / c.NS);   // even out the detector chunks
        c.windowed = 1;
        return c;   // R == 0 until the plan has sized the windows
    }
    c.reuse = 0;   // whole-row strips: the reuse march needs 96 registers, i.e. CTAs of <= 640 threads
    const int fixed = FwdConfig::fixed_bytes(c.NS * c.KA);
    const int Upmax = geom[0].Up > geom[1].Up ? geom[0].Up : geom[1].Up;
    const int Vpmax = geom[0].Vp > geom[1].Vp ? geom[0].Vp : geom[1].Vp;
    const int row_bytes = Upmax * kFwdNB * c.depth * 4;
    int rows = (smem_budget - fixed) / (c.stages * row_bytes);
    if (rows > Vpmax) rows = Vpmax;
    if (rows > 33) rows = 33;
    c.R = rows - 1;
    if (c.R < 1) c.R = 0;
    c.smem = (size_t)fixed + (size_t)c.stages * (size_t)(c.R + 1) * row_bytes;
    return c;
}

// 32-image records read by TWO lanes per ray, 16 images each (rotated loads, ctr_ldv16_rot): the per-sample geometry
// (coordinates, rounding, address, loop: ~50 instructions) is paid once per 16 images.  Serves the NEAREST-neighbour
// projector (project_tf_fast's default), whose samples read one record each and are therefore bound by that geometry:
// r2, 64 x 512^2 x 720: 3.22 -> 2.12 ms, 256 x 128^2 x 180: 0.300 -> 0.264 ms.  The bilinear projector is bound by its
// four shared-memory records per sample and lost with this shape (5.71 -> 5.76..6.35 ms over NS = 2..12 and CTAs of
// 384 / 480 / 512 threads: its reuse march needs 156 registers), so it keeps the 8-image lanes.
// Wide detectors: column-windowed strips with NS angle slots x KA = 2 angles per CTA (sweep at C4: NS 4 / 6 / 8 / 12 /
// 16 -> 2.25 / 2.17 / 2.12 / 2.15 / 2.15 ms); narrow ones: whole-row strips.
constexpr int kFwdWideNS = 8;
inline FwdConfig fwd_config_wide(int W, const CtrClassGeom geom[2], int smem_budget, int win_ns = 0)
{
    FwdConfig c{};
    c.lanes = 2;
    c.depth = 8;
    c.stages = kFwdStages;
    c.KA = 2;
    c.NS = 1;
    c.jchunks = 1;
    const int maxc = kFwdWideThreads - 32;
    c.JW = round_up(W, 16);
    if (c.JW * c.lanes > maxc) {
        c.NS = win_ns > 0 ? win_ns : kFwdWideNS;
        int q = 16;                                          // JW * NS must be a multiple of 16: whole warps per CTA
        for (int d = 2; d <= 16; d *= 2) if (c.NS % d == 0) q = 16 / d;
        c.JW = maxc / (c.lanes * c.NS) / q * q;
        if (c.JW < q) return c;                              // R = 0: too many angle slots
        c.jchunks = (W + c.JW - 1) / c.JW;
        c.JW = round_up((W + c.jchunks - 1) / c.jchunks, q);
        c.windowed = 1;
        return c;                                            // R == 0 until the plan has sized the windows
    }
    c.reuse = 0;
    const int fixed = FwdConfig::fixed_bytes(c.NS * c.KA);
    const int Upmax = geom[0].Up > geom[1].Up ? geom[0].Up : geom[1].Up;
    const int Vpmax = geom[0].Vp > geom[1].Vp ? geom[0].Vp : geom[1].Vp;
    const int row_bytes = Upmax * kFwdNB * c.depth * 4;
    int rows = (smem_budget - fixed) / (c.stages * row_bytes);
    if (rows > Vpmax) rows = Vpmax;
    if (rows > 33) rows = 33;
    c.R = rows - 1;
    if (c.R < 1) c.R = 0;
    c.smem = (size_t)fixed + (size_t)c.stages * (size_t)(c.R + 1) * row_bytes;
    return c;
}

template <int NBL, int INTERP, int EPI, int LANES, int REUSE>
inline cudaError_t launch_fwd_one(const FwdParams& p, const FwdConfig& c, dim3 grid, dim3 block, cudaStream_t st)
{
    cudaError_t e = cudaFuncSetAttribute(ctr_fwd_kernel<NBL, 2, INTERP, EPI, LANES, REUSE>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem);
    if (e != cudaSuccess) return e;
    ctr_fwd_kernel<NBL, 2, INTERP, EPI, LANES, REUSE><<<grid, block, c.smem, st>>>(p);
    launch_counter()++;
    return cudaGetLastError();
}

// G counts pixel records (super-groups of kFwdNB * depth images) along the batch
template <int INTERP, int EPI>
inline cudaError_t launch_fwd_ka(const FwdParams& p, const FwdConfig& c, int G, int chunks, cudaStream_t st)
{
    dim3 grid(chunks, c.jchunks, G), block(c.JW * c.lanes * c.NS + 32);   // + the producer warp
    if (c.KA != 2) return cudaErrorInvalidValue;
    if (c.lanes == 2 && c.depth == 8)   // 32-image records, 16 per lane: nearest only (bilinear is bound by shared-memory loads, see fwd_config_wide)
        return INTERP == CTR_NEAREST ? launch_fwd_one<16, CTR_NEAREST, EPI, 2, 0>(p, c, grid, block, st) : cudaErrorInvalidValue;
    if (c.lanes == 4 && c.depth == 8 && c.reuse && INTERP == CTR_BILINEAR) return launch_fwd_one<8, INTERP, EPI, 4, 1>(p, c, grid, block, st);
    if (c.lanes == 4 && c.depth == 8) return launch_fwd_one<8, INTERP, EPI, 4, 0>(p, c, grid, block, st);   // 32-image records
    if (c.lanes == 4) return launch_fwd_one<kFwdNB, INTERP, EPI, 4, 0>(p, c, grid, block, st);             // 16-image records
    if (c.lanes == 2) return launch_fwd_one<kFwdNB, INTERP, EPI, 2, 0>(p, c, grid, block, st);             // 8-image records
    return launch_fwd_one<kFwdNB, INTERP, EPI, 1, 0>(p, c, grid, block, st);                               // 4-image records
}

inline size_t bp_smem_bytes(int win, int NB, int AB)
{
    return 128 + 2 * kBpAB * 8 * 4 + 2 * kBpAB * 2 * 8 + 2ull * AB * (size_t)win * NB * 4;
}

// Three shapes: the per-pixel geometry (~130 instructions for the exact adjoint) is shared by
// all images of a thread, so big batches take 32 images per thread, medium ones 16 (both on
// 32x8-pixel tiles), tiny ones 8 on 32x16 tiles (fewer idle accumulator lanes).
// (r1: 32 images pay off only for the geometry-heavy exact adjoint -- C4 5.63 -> 4.64 ms;
// the 2-tap FBP gather loses occupancy and stays at 16; three resident 32-image CTAs or 32x4 / 32x5 / 32x6
// tiles were measured and lost: 0.47 / 0.377 / 0.371 / 0.376 vs 0.371 ms at C2.)
// X, Y: when the 32-image shape would not even give every SM one CTA (small images x small batch, e.g. the
// 32..64-image chunks of the host pipeline at 128^2) the 16-image shape has twice the CTAs: 0.105 vs 0.141 ms
// at 32 x 128^2 x 180.  Pass X = 0 for "size for the worst case" (workspace queries).
inline int bp_nb_for_batch(int B, int mode, int X = 0, int Y = 0)
{
    if (B <= 8) return 8;
    if (B < 24) return 16;
    const long long ctas32 = (X > 0 && Y > 0) ? (long long)((Y + kBpTW - 1) / kBpTW) * ((X + 7) / 8) * ((B + 31) / 32) : -1;
    // 2-tap gathers (TF-compat gradient, FBP): 32 images only pay on grids of several waves (TF-compat C4 3.15 -> 2.90 ms,
    // C2 unchanged).  The FBP row filter always works on 16 images and writes into the gather's groups (filt_dst_row).
    if (mode == CTR_ADJ_TF || mode == CTR_ADJ_FBP)
        return ctas32 >= 1024 ? 32 : 16;
    if (ctas32 >= 0 && ctas32 < 148) return 16;
    return 32;
}

template <int NB, int TH, int MINB, int MODE, int INTERP>
inline cudaError_t launch_bp_cfg(BpParams p, cudaStream_t st)
{
    const int G = (p.B + NB - 1) / NB;
    dim3 grid((p.Y + kBpTW - 1) / kBpTW, (p.X + TH - 1) / TH, G), block(kBpTW, TH + 1);   // + the producer warp
    if (p.win > bp_win(TH)) p.win = bp_win(TH);
    const size_t smem = bp_smem_bytes(p.win, NB, bp_ab(TH, NB, MINB));
    cudaError_t e = cudaFuncSetAttribute(ctr_bp_kernel<NB, TH, MINB, MODE, INTERP>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    ctr_bp_kernel<NB, TH, MINB, MODE, INTERP><<<grid, block, smem, st>>>(p);
    launch_counter()++;
    return cudaGetLastError();
}

template <int MODE, int INTERP>
inline cudaError_t launch_bp(const BpParams& p, int nb, cudaStream_t st)
{
    if (nb == 32) return launch_bp_cfg<32, 8, 2, MODE, INTERP>(p, st);
    if (nb == 16) return launch_bp_cfg<16, 8, 3, MODE, INTERP>(p, st);
    return launch_bp_cfg<8, 16, 2, MODE, INTERP>(p, st);
}

}  // namespace ctr
