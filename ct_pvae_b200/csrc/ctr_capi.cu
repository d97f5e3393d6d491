// ctr_capi.cu -- the C ABI declared in include/ctradon.h: plans, argument checks,
// workspace carving and kernel launches.  No torch, no Python, no CPU compute path.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>   // types only: the library is opened at run time (ctr_comm_nccl_*), there is no link-time dependency

#include <cmath>
#include <cstdio>
#include <atomic>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/ctradon.h"
#include "ctr_core.h"
#include "ctr_host.h"
#include "ctr_kernels.cuh"

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg)
{
    g_err = msg;
    return code;
}
int fail_cuda(cudaError_t e, const char* what)
{
    g_err = std::string(what) + ": " + cudaGetErrorString(e);
    return CTR_ECUDA;
}
#define CTR_CUDA(call)                                   \
    do {                                                 \
        cudaError_t e__ = (call);                        \
        if (e__ != cudaSuccess) return fail_cuda(e__, #call); \
    } while (0)

// switch to the plan's device for the duration of a call
struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    cudaError_t err = cudaSuccess;
    explicit DeviceGuard(int dev)
    {
        err = cudaGetDevice(&prev);
        if (err == cudaSuccess && prev != dev) err = cudaSetDevice(dev);
        ok = (err == cudaSuccess);
    }
    ~DeviceGuard()
    {
        int cur = -1;
        if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
    }
};

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---- optional per-kernel event timing (ctr_profile_*) ----
struct ProfRec { int id; cudaEvent_t e0, e1; };
std::mutex g_prof_mu;
std::vector<ProfRec> g_prof;
std::atomic<int> g_prof_on{0};

struct ProfScope {
    int id; cudaStream_t st; cudaEvent_t e0 = nullptr, e1 = nullptr; bool on;
    ProfScope(int id_, cudaStream_t st_) : id(id_), st(st_), on(g_prof_on.load() != 0)
    {
        if (!on) return;
        if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) { on = false; return; }
        cudaEventRecord(e0, st);
    }
    ~ProfScope()
    {
        if (!on) return;
        cudaEventRecord(e1, st);
        std::lock_guard<std::mutex> lk(g_prof_mu);
        g_prof.push_back({id, e0, e1});
    }
};

}  // namespace

constexpr int kShapes = 4;
struct ctr_plan {
    int device = 0;
    int A = 0, X = 0, Y = 0, pad = 0, H = 0, W = 0, padx = 0, pady = 0;
    std::vector<float> t, tinv;
    std::vector<CtrRay> rays;
    int n_cls[2] = {0, 0};
    CtrClassGeom geom[2];
    // forward kernel shapes: [0] 4 images per pixel record (any detector, any batch), [1] depth-first 8/16-image
    // records, [2] 32-image records with 8 images per lane, [3] 32-image records with 16 images per lane (two lanes
    // per ray); [1]..[3] are column-windowed on wide detectors
    // (fc.R == 0: shape unavailable for this geometry)
    struct Shape {
        ctr::FwdConfig fc;
        std::vector<CtrChunk> chunks;   // CTA columns
        CtrChunk* d_chunks = nullptr;
        // one chunk per ray, in table order: the CTA columns of an angle-subset call (ctr_radon_*_sel), where the
        // angles sharing a CTA are not known when the plan is made
        std::vector<CtrChunk> singles;
        CtrChunk* d_singles = nullptr;
        size_t smem_single = 0;
    } shape[kShapes];
    std::vector<int> pos;      // angle -> its row in the class-sorted ray table
    int* d_pos = nullptr;
    int sm_count = 148;
    void* d_block = nullptr;   // one device allocation: [t | tinv | rays | chunk tables]
    float* d_t = nullptr;
    float* d_tinv = nullptr;
    CtrRay* d_rays = nullptr;
};

struct ctr_fbp_plan {
    int device = 0;
    int A = 0, P = 0, x_size = 0, y_size = 0;
    double* d_cs = nullptr;   // [A][2]
    float* d_h = nullptr;     // [P] spatial kernel
    // ramp filter: the kernel vanishes at every even offset but 0 -> half the taps (ctr_fbp_filter_sparse_kernel)
    int sparse = 0;
    float h0 = 0.f;
    float* d_hs = nullptr;    // [1 + P + kFiltBPTMax] odd taps, doubled
    int fused_cl = 0, fused_ab = 0;   // cluster size / angle batch of the single-kernel path (0: image too large for it)
    int use_fused = 0;                // ctr_fbp_plan_set_fused
};

extern "C" {

int ctr_version(void) { return CTR_VERSION; }
const char* ctr_last_error(void) { return g_err.c_str(); }
long long ctr_launch_count(void) { return ctr::launch_counter().load(); }

static const char* kKernelNames[CTR_K_COUNT] = {"ctr_pack_image_kernel", "ctr_pack_sino_kernel", "ctr_fwd_kernel",
                                                "ctr_bp_kernel<exact>", "ctr_bp_kernel<tf_compat>",
                                                "ctr_fbp_filter_kernel", "ctr_bp_kernel<fbp>", "ctr_xchg_sum_kernel", "ctr_fbp_fused_kernel"};
const char* ctr_kernel_name(int id) { return (id >= 0 && id < CTR_K_COUNT) ? kKernelNames[id] : ""; }

int ctr_profile_enable(int on)
{
    g_prof_on.store(on ? 1 : 0);
    return CTR_OK;
}

int ctr_profile_reset(void)
{
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (auto& r : g_prof) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
    g_prof.clear();
    return CTR_OK;
}

int ctr_profile_read(int id, double* total_ms, long long* launches)
{
    if (id < 0 || id >= CTR_K_COUNT) return fail(CTR_EINVAL, "ctr_profile_read: bad kernel id");
    std::lock_guard<std::mutex> lk(g_prof_mu);
    double tot = 0.0;
    long long n = 0;
    for (auto& r : g_prof) {
        if (r.id != id) continue;
        CTR_CUDA(cudaEventSynchronize(r.e1));
        float ms = 0.f;
        CTR_CUDA(cudaEventElapsedTime(&ms, r.e0, r.e1));
        tot += ms;
        ++n;
    }
    if (total_ms) *total_ms = tot;
    if (launches) *launches = n;
    return CTR_OK;
}

int ctr_device_synchronize(int device)
{
    DeviceGuard guard(device);
    if (!guard.ok) return fail_cuda(guard.err, "cudaSetDevice");
    CTR_CUDA(cudaDeviceSynchronize());
    return CTR_OK;
}

int ctr_num_proj_pix(int X, int Y) { return (X > 0 && Y > 0) ? ctr_h_num_proj_pix(X, Y) : fail(CTR_EINVAL, "X, Y must be positive"); }

int ctr_frame(int X, int Y, int pad, int* H, int* W, int* padx, int* pady)
{
    if (X <= 0 || Y <= 0) return fail(CTR_EINVAL, "X, Y must be positive");
    int h, w, px, py;
    ctr_h_frame(X, Y, pad, h, w, px, py);
    if (H) *H = h;
    if (W) *W = w;
    if (padx) *padx = px;
    if (pady) *pady = py;
    return CTR_OK;
}

int ctr_make_transforms(const double* theta, int A, int H, int W, float* out)
{
    if (!theta || !out || A < 0 || H <= 0 || W <= 0) return fail(CTR_EINVAL, "ctr_make_transforms: bad argument");
    ctr_h_make_transforms(theta, A, H, W, out);
    return CTR_OK;
}

int ctr_invert_transforms(const float* t, int A, float* out)
{
    if (!t || !out || A < 0) return fail(CTR_EINVAL, "ctr_invert_transforms: bad argument");
    ctr_h_invert_transforms(t, A, out);
    return CTR_OK;
}

int ctr_filter_to_spatial(const double* fr, const double* fi, int P, double* out)
{
    if (!fr || !out || P <= 0) return fail(CTR_EINVAL, "ctr_filter_to_spatial: bad argument");
    // real part of the inverse DFT, O(P^2) in float64 (P is a detector width)
    const double w = 2.0 * M_PI / (double)P;
    for (int n = 0; n < P; ++n) {
        double acc = 0.0;
        for (int k = 0; k < P; ++k) {
            const double ph = w * (double)(((long long)n * k) % P);
            acc += fr[k] * std::cos(ph) - (fi ? fi[k] * std::sin(ph) : 0.0);
        }
        out[n] = acc / (double)P;
    }
    return CTR_OK;
}

// ------------------------------------------------------------------------------------------ plans
int ctr_plan_create(const double* theta, int A, int X, int Y, int pad, int device, ctr_plan** out)
{
    if (!out) return fail(CTR_EINVAL, "ctr_plan_create: out is NULL");
    *out = nullptr;
    if (!theta || A <= 0 || X <= 0 || Y <= 0) return fail(CTR_EINVAL, "ctr_plan_create: need theta, A>0, X>0, Y>0");
    if ((long long)X * Y > (1ll << 28)) return fail(CTR_EUNSUPPORTED, "ctr_plan_create: image too large");
    ctr_plan* p = new (std::nothrow) ctr_plan();
    if (!p) return fail(CTR_EINVAL, "ctr_plan_create: out of host memory");
    p->device = device; p->A = A; p->X = X; p->Y = Y; p->pad = pad ? 1 : 0;
    ctr_h_frame(X, Y, p->pad, p->H, p->W, p->padx, p->pady);
    p->t.resize((size_t)A * 8);
    p->tinv.resize((size_t)A * 8);
    ctr_h_make_transforms(theta, A, p->H, p->W, p->t.data());
    ctr_h_invert_transforms(p->t.data(), A, p->tinv.data());
    ctr_h_class_geom(X, Y, p->padx, p->pady, p->geom);
    std::vector<int> seg;
    ctr_h_build_rays(p->t.data(), A, p->rays, p->n_cls[0], &seg);
    p->n_cls[1] = A - p->n_cls[0];

    DeviceGuard guard(device);
    if (!guard.ok) { int rc = fail_cuda(guard.err, "cudaSetDevice"); delete p; return rc; }
    int smem_optin = 0;
    cudaError_t e = cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    if (e != cudaSuccess) { int rc = fail_cuda(e, "cudaDeviceGetAttribute"); delete p; return rc; }
    if (cudaDeviceGetAttribute(&p->sm_count, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || p->sm_count <= 0) p->sm_count = 148;
    p->shape[0].fc = ctr::fwd_config(p->W, p->geom, smem_optin - 2048);
    p->shape[1].fc = ctr::fwd_config_depth(p->W, p->geom, smem_optin - 2048, false);
    p->shape[2].fc = ctr::fwd_config_depth(p->W, p->geom, smem_optin - 2048, true);
    p->shape[3].fc = ctr::fwd_config_wide(p->W, p->geom, smem_optin - 2048);
    if (p->shape[0].fc.R < 1) { delete p; return fail(CTR_EUNSUPPORTED, "ctr_plan_create: image rows too wide for the shared-memory strips"); }
    // CTA columns: chunks of consecutive table entries, strip height and (wide detectors) column windows
    for (int k = 0; k < kShapes; ++k) {
        ctr_plan::Shape& sh = p->shape[k];
        ctr::FwdConfig& fc = sh.fc;
        size_t strip_bytes = 0;
        if (fc.windowed) {
            const int rmax = 16;   // r1: taller windowed strips only widen the windows (C4: R <= 10 or 3 stages +3 %)
            // widely spaced angles (sparse-angle minibatches): retry with fewer angle slots per CTA
            for (int ns = fc.NS; ns >= 1; ns /= 2) {
                if (ns != fc.NS)
                    fc = (k == 3) ? ctr::fwd_config_wide(p->W, p->geom, smem_optin - 2048, ns)
                                  : ctr::fwd_config_depth(p->W, p->geom, smem_optin - 2048, k == 2, ns);
                if (fc.JW < 1 || !fc.windowed) { fc.R = 0; continue; }
                const int NA = fc.angles_per_cta(), fixed = ctr::FwdConfig::fixed_bytes(NA);
                const size_t budget = (size_t)(smem_optin - 2048 - fixed);
                if (ctr_h_build_chunks(p->rays, seg, p->geom, NA, p->W, fc.JW, fc.jchunks, ctr::kFwdNB * fc.depth * 4, fc.stages,
                                       budget, true, 0, 4, rmax, sh.chunks, strip_bytes)) {
                    fc.R = 1;   // available; the strip height is per chunk
                    fc.smem = (size_t)fixed + (size_t)fc.stages * strip_bytes;
                    break;
                }
                fc.R = 0;       // some CTA's rays are too far apart for a window that fits
                sh.chunks.clear();
            }
        } else if (fc.R >= 1) {
            ctr_h_build_chunks(p->rays, seg, p->geom, fc.angles_per_cta(), p->W, fc.JW, fc.jchunks, ctr::kFwdNB * fc.depth * 4,
                               fc.stages, 0, false, fc.R, 1, fc.R, sh.chunks, strip_bytes);
        }
    }
    // angle-subset calls: one CTA column per ray (same strip / window sizing rules, one ray per window)
    for (int k = 0; k < kShapes; ++k) {
        ctr_plan::Shape& sh = p->shape[k];
        const ctr::FwdConfig& fc = sh.fc;
        if (fc.R < 1) continue;
        size_t strip_bytes = 0;
        const int fixed = ctr::FwdConfig::fixed_bytes(fc.angles_per_cta());
        bool ok;
        if (fc.windowed)
            ok = ctr_h_build_chunks(p->rays, seg, p->geom, 1, p->W, fc.JW, fc.jchunks, ctr::kFwdNB * fc.depth * 4, fc.stages,
                                    (size_t)(smem_optin - 2048 - fixed), true, 0, 4, 16, sh.singles, strip_bytes);
        else
            ok = ctr_h_build_chunks(p->rays, seg, p->geom, 1, p->W, fc.JW, fc.jchunks, ctr::kFwdNB * fc.depth * 4, fc.stages, 0, false,
                                    fc.R, 1, fc.R, sh.singles, strip_bytes);
        if (!ok || (int)sh.singles.size() != A) { sh.singles.clear(); continue; }   // subset calls fall back to another shape
        sh.smem_single = (size_t)fixed + (size_t)fc.stages * strip_bytes;
    }
    p->pos.assign(A, 0);
    for (int k = 0; k < A; ++k) p->pos[p->rays[k].angle] = k;
    // one allocation + one upload for all tables
    const size_t tb = (size_t)A * 8 * sizeof(float), rb = (size_t)A * sizeof(CtrRay), pb = align_up((size_t)A * sizeof(int), 16);
    size_t cb[2 * kShapes], ctot = 0;
    for (int k = 0; k < kShapes; ++k) {
        cb[k] = p->shape[k].chunks.size() * sizeof(CtrChunk);
        cb[kShapes + k] = p->shape[k].singles.size() * sizeof(CtrChunk);
        ctot += cb[k] + cb[kShapes + k];
    }
    std::vector<unsigned char> host(2 * tb + rb + pb + ctot);
    std::memcpy(host.data(), p->t.data(), tb);
    std::memcpy(host.data() + tb, p->tinv.data(), tb);
    std::memcpy(host.data() + 2 * tb, p->rays.data(), rb);
    std::memcpy(host.data() + 2 * tb + rb, p->pos.data(), (size_t)A * sizeof(int));
    for (size_t k = 0, off = 2 * tb + rb + pb; k < 2 * kShapes; off += cb[k], ++k) {
        const std::vector<CtrChunk>& v = k < kShapes ? p->shape[k].chunks : p->shape[k - kShapes].singles;
        if (cb[k]) std::memcpy(host.data() + off, v.data(), cb[k]);
    }
    if ((e = cudaMalloc(&p->d_block, host.size())) != cudaSuccess ||
        (e = cudaMemcpy(p->d_block, host.data(), host.size(), cudaMemcpyHostToDevice)) != cudaSuccess) {
        int rc = fail_cuda(e, "ctr_plan_create: table upload");
        cudaFree(p->d_block);
        delete p;
        return rc;
    }
    p->d_t = (float*)p->d_block;
    p->d_tinv = (float*)((char*)p->d_block + tb);
    p->d_rays = (CtrRay*)((char*)p->d_block + 2 * tb);
    p->d_pos = (int*)((char*)p->d_block + 2 * tb + rb);
    for (size_t k = 0, off = 2 * tb + rb + pb; k < 2 * kShapes; off += cb[k], ++k) {
        if (k < kShapes) p->shape[k].d_chunks = (CtrChunk*)((char*)p->d_block + off);
        else p->shape[k - kShapes].d_singles = (CtrChunk*)((char*)p->d_block + off);
    }
    *out = p;
    return CTR_OK;
}

int ctr_plan_destroy(ctr_plan* p)
{
    if (!p) return CTR_OK;
    DeviceGuard guard(p->device);
    cudaFree(p->d_block);
    delete p;
    return CTR_OK;
}

int ctr_plan_info(const ctr_plan* p, int* A, int* X, int* Y, int* H, int* W, int* padx, int* pady)
{
    if (!p) return fail(CTR_EINVAL, "ctr_plan_info: plan is NULL");
    if (A) *A = p->A;
    if (X) *X = p->X;
    if (Y) *Y = p->Y;
    if (H) *H = p->H;
    if (W) *W = p->W;
    if (padx) *padx = p->padx;
    if (pady) *pady = p->pady;
    return CTR_OK;
}

int ctr_plan_tables(const ctr_plan* p, float* fwd, float* inv)
{
    if (!p) return fail(CTR_EINVAL, "ctr_plan_tables: plan is NULL");
    if (fwd) std::memcpy(fwd, p->t.data(), p->t.size() * sizeof(float));
    if (inv) std::memcpy(inv, p->tinv.data(), p->tinv.size() * sizeof(float));
    return CTR_OK;
}

// which forward shape serves a batch of B: the deepest pixel record the batch fills reasonably
// (32 images from 17 up, 8/16 from 3 lanes' worth up), else the 4-image records
static const ctr_plan::Shape& shape_for(const ctr_plan* p, int B, bool subset = false, int interp = CTR_INTERP_BILINEAR)
{
    const ctr_plan::Shape& s16 = p->shape[1];
    const ctr_plan::Shape& s32 = p->shape[2];
    // angle-subset calls run one ray per CTA column (the `singles` tables)
    const bool ok16 = s16.fc.R >= 1 && B >= 3 * s16.fc.depth && (!subset || !s16.singles.empty());
    const bool ok32 = s32.fc.R >= 1 && B > 16 && (!subset || !s32.singles.empty());
    // nearest neighbour: one record per sample, so the per-sample geometry dominates and 16 images per lane win
    // (r2, 64 x 512^2 x 720: 3.22 -> 2.15 ms; 256 x 128^2 x 180: 0.300 -> 0.265 ms)
    const ctr_plan::Shape& sw = p->shape[3];
    const bool okw = interp == CTR_INTERP_NEAREST && sw.fc.R >= 1 && B > 16 && (!subset || !sw.singles.empty());
    if (!subset && ok16 && ok32 && s16.fc.lanes == 4 && !s16.fc.windowed && !s32.fc.windowed) {
        // Both run one CTA per SM; a 32-image CTA does twice the work of a 16-image one in 1.84x the time
        // (r1: 93 vs 51 us per wave at 128^2 x 180).  Small batches (the chunks of the host pipeline) leave the last
        // wave mostly empty, so count waves: 64 images -> 182 CTAs = 2 waves (32) vs 364 CTAs = 3 waves (16).
        auto waves = [&](const ctr_plan::Shape& sh) {
            const long long rec = ctr::kFwdNB * sh.fc.depth, G = (B + rec - 1) / rec;
            const long long ctas = G * (long long)sh.chunks.size() * sh.fc.jchunks;
            return (double)((ctas + p->sm_count - 1) / p->sm_count);
        };
        return (waves(s32) * 1.84 <= waves(s16)) ? (okw ? sw : s32) : s16;
    }
    if (okw) return sw;
    if (ok32) return s32;
    if (ok16) return s16;
    return p->shape[0];
}
static const ctr::FwdConfig& fwd_cfg_for(const ctr_plan* p, int B, bool subset = false, int interp = CTR_INTERP_BILINEAR)
{
    return shape_for(p, B, subset, interp).fc;
}

// one packed copy of the batch; sized for the deeper of the full-plan and the angle-subset shape of this batch size
static size_t pack_bytes(const ctr_plan* p, int B)
{
    size_t best = 0;
    for (int k = 0; k < 4; ++k) {
        const size_t rec = (size_t)ctr::kFwdNB * fwd_cfg_for(p, B, (k & 1) != 0, (k & 2) ? CTR_INTERP_NEAREST : CTR_INTERP_BILINEAR).depth;
        const size_t G = ((size_t)B + rec - 1) / rec;
        const size_t px0 = (size_t)p->geom[0].Vp * p->geom[0].Up, px1 = (size_t)p->geom[1].Vp * p->geom[1].Up;
        best = std::max(best, align_up(G * std::max(px0, px1) * rec * sizeof(float), 256));
    }
    return best;
}

// human-readable description of the forward shape a batch of B would run (tests, bench, tuning)
int ctr_plan_describe(const ctr_plan* p, int B, char* buf, size_t n)
{
    if (!p || !buf || n == 0 || B <= 0) return fail(CTR_EINVAL, "ctr_plan_describe: bad argument");
    const ctr::FwdConfig& fc = fwd_cfg_for(p, B);
    const std::vector<CtrChunk>& ch = shape_for(p, B).chunks;
    int rlo = 1 << 30, rhi = 0, wlo = 1 << 30, whi = 0, nwin = 0;
    for (const CtrChunk& c : ch) {
        rlo = std::min(rlo, c.R); rhi = std::max(rhi, c.R);
        if (c.wc > 0) { wlo = std::min(wlo, c.wc); whi = std::max(whi, c.wc); ++nwin; }
    }
    const ctr::FwdConfig& fn = fwd_cfg_for(p, B, false, CTR_INTERP_NEAREST);   // the nearest-neighbour projector's shape
    snprintf(buf, n, "images_per_record=%d windowed=%d window_chunks=%d/%d JW=%d jchunks=%d NS=%d KA=%d stages=%d R=%d..%d wc=%d..%d smem=%zu "
             "adjoint_images_per_thread: exact=%d tf_compat=%d nearest: images_per_lane=%d JW=%d NS=%d",
             ctr::kFwdNB * fc.depth, fc.windowed, nwin, (int)ch.size(), fc.JW, fc.jchunks, fc.NS, fc.KA, fc.stages, rlo, rhi,
             nwin ? wlo : 0, whi, fc.smem, ctr::bp_nb_for_batch(B, CTR_ADJ_EXACT, p->X, p->Y), ctr::bp_nb_for_batch(B, CTR_ADJ_TF, p->X, p->Y),
             ctr::kFwdNB * fn.depth / fn.lanes, fn.JW, fn.NS);
    return CTR_OK;
}

size_t ctr_forward_workspace_bytes(const ctr_plan* p, int B)
{
    if (!p || B <= 0) return 0;
    return 2 * pack_bytes(p, B);
}

// sinogram pack bytes; the exact adjoint may use wider image groups than the 2-tap gathers,
// so size for the larger padding
static size_t spk_bytes(int B, int A, int W)
{
    size_t best = 0;
    for (int mode : {CTR_ADJ_EXACT, CTR_ADJ_TF}) {
        const size_t NB = (size_t)ctr::bp_nb_for_batch(B, mode);
        const size_t G = ((size_t)B + NB - 1) / NB;
        best = std::max(best, G * (size_t)A * (size_t)(W + 2) * NB * sizeof(float));
    }
    return align_up(best, 256);
}

size_t ctr_adjoint_workspace_bytes(const ctr_plan* p, int B)
{
    if (!p || B <= 0) return 0;
    return spk_bytes(B, p->A, p->W);
}

// ------------------------------------------------------------------------------------------ forward
static int adjoint_impl(const ctr_plan* p, const float* dsino, float* dimg, int B, int interp, int mode, float scale,
                        const int* sel, int n_sel, const CtrExchange* xg, void* ws, size_t ws_bytes, void* stream, const char* who);
static size_t spk_bytes(int B, int A, int W);

struct LoglikArgs {
    const float* mask; const float* meas; const int* amap; int A_all; float pnm, sqrt_reg; float* loglik;
};

static size_t loglik_partial_bytes(const ctr_plan* p, int B)
{
    size_t best = 0;
    for (int k = 0; k < 4; ++k) {
        const int subset = k & 1, interp = (k & 2) ? CTR_INTERP_NEAREST : CTR_INTERP_BILINEAR;
        const ctr::FwdConfig& fc = fwd_cfg_for(p, B, subset != 0, interp);
        const size_t chunks = subset ? (size_t)p->A : shape_for(p, B, false, interp).chunks.size();   // a subset has at most A columns
        const size_t rec = (size_t)ctr::kFwdNB * fc.depth;
        const size_t G = ((size_t)B + rec - 1) / rec;
        best = std::max(best, align_up(chunks * fc.jchunks * G * rec * sizeof(float), 256));
    }
    return best;
}

static int forward_impl(const ctr_plan* p, const float* img, float* out, int B, int interp, const LoglikArgs* ll,
                        const int* sel, int n_sel, void* ws, size_t ws_bytes, void* stream, const char* who)
{
    if (!p || !img || !out) return fail(CTR_EINVAL, std::string(who) + ": NULL plan or buffer");
    if (sel && (n_sel <= 0 || n_sel > p->A)) return fail(CTR_EINVAL, std::string(who) + ": the angle subset must have 1..A entries");
    if (B <= 0) return fail(CTR_EINVAL, std::string(who) + ": B must be positive");
    if (interp != CTR_INTERP_NEAREST && interp != CTR_INTERP_BILINEAR) return fail(CTR_EINVAL, std::string(who) + ": bad interp");
    const size_t need = ll ? ctr_loglik_workspace_bytes(p, B) : ctr_forward_workspace_bytes(p, B);
    if (!ws || ws_bytes < need) return fail(CTR_EWORKSPACE, std::string(who) + ": workspace too small");
    if (((uintptr_t)ws & 255) != 0) return fail(CTR_EINVAL, std::string(who) + ": workspace must be 256-byte aligned");
    DeviceGuard guard(p->device);
    if (!guard.ok) return fail_cuda(guard.err, "cudaSetDevice");
    cudaStream_t st = (cudaStream_t)stream;
    const ctr_plan::Shape& sh = shape_for(p, B, sel != nullptr, interp);
    ctr::FwdConfig fc = sh.fc;
    if (sel) {
        if (sh.singles.empty()) return fail(CTR_EUNSUPPORTED, std::string(who) + ": no per-ray strip table for this geometry");
        fc.smem = sh.smem_single;
    }
    const int rec = ctr::kFwdNB * fc.depth;
    const int G = (B + rec - 1) / rec;            // pixel records along the batch (super-groups)
    float* pk0 = p->n_cls[0] ? (float*)ws : nullptr;
    float* pk1 = p->n_cls[1] ? (float*)((char*)ws + pack_bytes(p, B)) : nullptr;
    {
        ProfScope prof(CTR_K_PACK_IMAGE, st);
        CTR_CUDA(ctr::launch_pack_image(img, B, p->X, p->Y, pk0, pk1, rec, p->geom[0].Up, p->geom[1].Up, st));
    }
    ctr::FwdParams fp;
    fp.pk[0] = pk0; fp.pk[1] = pk1;
    fp.geom[0] = p->geom[0]; fp.geom[1] = p->geom[1];
    fp.rays = p->d_rays;
    fp.chunks = sel ? sh.d_singles : sh.d_chunks;
    fp.sel = sel;
    fp.pos = p->d_pos;
    fp.jwd = fc.JW * fc.lanes;
    fp.ns = fc.NS;
    fp.stages = fc.stages;

    const int chunks = sel ? n_sel : (int)sh.chunks.size();
    fp.H = p->H; fp.W = p->W; fp.A = sel ? n_sel : p->A; fp.B = B;
    fp.sino = out;
    fp.mask = nullptr; fp.meas = nullptr; fp.amap = nullptr; fp.A_all = p->A; fp.pnm = 1.f; fp.sqrt_reg = 0.f; fp.partial = nullptr;
    cudaError_t e;
    if (!ll) {
        ProfScope prof(CTR_K_FORWARD, st);
        e = (interp == CTR_INTERP_NEAREST) ? ctr::launch_fwd_ka<CTR_NEAREST, 0>(fp, fc, G, chunks, st)
                                           : ctr::launch_fwd_ka<CTR_BILINEAR, 0>(fp, fc, G, chunks, st);
        if (e != cudaSuccess) return fail_cuda(e, "ctr_fwd_kernel launch");
        return CTR_OK;
    }
    fp.mask = ll->mask; fp.meas = ll->meas; fp.amap = ll->amap; fp.A_all = ll->A_all; fp.pnm = ll->pnm; fp.sqrt_reg = ll->sqrt_reg;
    fp.partial = (float*)((char*)ws + 2 * pack_bytes(p, B));
    {
        ProfScope prof(CTR_K_FORWARD, st);
        e = (interp == CTR_INTERP_NEAREST) ? ctr::launch_fwd_ka<CTR_NEAREST, 1>(fp, fc, G, chunks, st)
                                           : ctr::launch_fwd_ka<CTR_BILINEAR, 1>(fp, fc, G, chunks, st);
    }
    if (e != cudaSuccess) return fail_cuda(e, "ctr_fwd_kernel<loglik> launch");
    ctr::ctr_loglik_reduce_kernel<<<(B + 127) / 128, 128, 0, st>>>(fp.partial, chunks * fc.jchunks, G * rec, B, ll->loglik);
    ctr::launch_counter()++;
    CTR_CUDA(cudaGetLastError());
    return CTR_OK;
}

int ctr_radon_forward(const ctr_plan* p, const float* img, float* sino, int B, int interp, void* ws, size_t ws_bytes,
                      void* stream)
{
    return forward_impl(p, img, sino, B, interp, nullptr, nullptr, 0, ws, ws_bytes, stream, "ctr_radon_forward");
}

int ctr_radon_forward_sel(const ctr_plan* p, const float* img, float* sino, int B, int interp, const int* sel, int n_sel,
                          void* ws, size_t ws_bytes, void* stream)
{
    if (!sel) return fail(CTR_EINVAL, "ctr_radon_forward_sel: the angle subset is NULL");
    return forward_impl(p, img, sino, B, interp, nullptr, sel, n_sel, ws, ws_bytes, stream, "ctr_radon_forward_sel");
}

int ctr_radon_adjoint_sel(const ctr_plan* p, const float* dsino, float* dimg, int B, int interp, int mode, float scale,
                          const int* sel, int n_sel, void* ws, size_t ws_bytes, void* stream)
{
    if (!sel) return fail(CTR_EINVAL, "ctr_radon_adjoint_sel: the angle subset is NULL");
    return adjoint_impl(p, dsino, dimg, B, interp, mode, scale, sel, n_sel, nullptr, ws, ws_bytes, stream, "ctr_radon_adjoint_sel");
}

int ctr_radon_loglik_sel(const ctr_plan* p, const float* img, const float* mask, const float* meas, const int* sel, int n_sel,
                         float pnm, float sqrt_reg, float* loglik, float* dproj, int B, int interp, void* ws, size_t ws_bytes,
                         void* stream)
{
    if (!mask || !meas || !loglik || !sel) return fail(CTR_EINVAL, "ctr_radon_loglik_sel: NULL mask, measurement, subset or output");
    if (!(pnm > 0.f) || !(sqrt_reg >= 0.f)) return fail(CTR_EINVAL, "ctr_radon_loglik_sel: pnm must be > 0 and sqrt_reg >= 0");
    LoglikArgs ll{mask, meas, nullptr, p ? p->A : 0, pnm, sqrt_reg, loglik};
    return forward_impl(p, img, dproj, B, interp, &ll, sel, n_sel, ws, ws_bytes, stream, "ctr_radon_loglik_sel");
}

size_t ctr_loglik_workspace_bytes(const ctr_plan* p, int B)
{
    if (!p || B <= 0) return 0;
    return 2 * pack_bytes(p, B) + loglik_partial_bytes(p, B);
}

int ctr_radon_loglik(const ctr_plan* p, const float* img, const float* mask, const float* meas, const int* angle_map,
                     int A_all, float pnm, float sqrt_reg, float* loglik, float* dproj, int B, int interp, void* ws,
                     size_t ws_bytes, void* stream)
{
    if (!mask || !meas || !loglik) return fail(CTR_EINVAL, "ctr_radon_loglik: NULL mask, measurement or output");
    if (A_all < (p ? p->A : 0) && !angle_map) return fail(CTR_EINVAL, "ctr_radon_loglik: A_all smaller than the plan's angle count");
    if (!(pnm > 0.f) || !(sqrt_reg >= 0.f)) return fail(CTR_EINVAL, "ctr_radon_loglik: pnm must be > 0 and sqrt_reg >= 0");
    LoglikArgs ll{mask, meas, angle_map, A_all, pnm, sqrt_reg, loglik};
    return forward_impl(p, img, dproj, B, interp, &ll, nullptr, 0, ws, ws_bytes, stream, "ctr_radon_loglik");
}

// ------------------------------------------------------------------------------------------ adjoint
int ctr_radon_adjoint(const ctr_plan* p, const float* dsino, float* dimg, int B, int interp, int mode, void* ws,
                      size_t ws_bytes, void* stream)
{
    return ctr_radon_adjoint_scaled(p, dsino, dimg, B, interp, mode, 1.0f, ws, ws_bytes, stream);
}

// The adjoint launch sequence (pack + gather kernel) shared by the plain, subset and sharded entry points.
//   sel / n_sel : angle subset (device int32 indices into the plan's angles) or null = all of the plan's angles
//   xg          : angle-sharded exchange target (nranks > 1) or null
static int adjoint_impl(const ctr_plan* p, const float* dsino, float* dimg, int B, int interp, int mode, float scale,
                        const int* sel, int n_sel, const CtrExchange* xg, void* ws, size_t ws_bytes, void* stream, const char* who)
{
    if (!p || !dsino || (!dimg && !xg)) return fail(CTR_EINVAL, std::string(who) + ": NULL plan or buffer");
    if (B <= 0) return fail(CTR_EINVAL, std::string(who) + ": B must be positive");
    if (interp != CTR_INTERP_NEAREST && interp != CTR_INTERP_BILINEAR) return fail(CTR_EINVAL, std::string(who) + ": bad interp");
    if (mode != CTR_ADJOINT_EXACT && mode != CTR_ADJOINT_TF_COMPAT) return fail(CTR_EINVAL, std::string(who) + ": bad mode");
    if (sel && (n_sel <= 0 || n_sel > p->A)) return fail(CTR_EINVAL, std::string(who) + ": the angle subset must have 1..A entries");
    const int A = sel ? n_sel : p->A;
    if (!ws || ws_bytes < spk_bytes(B, A, p->W)) return fail(CTR_EWORKSPACE, std::string(who) + ": workspace too small");
    if (((uintptr_t)ws & 255) != 0) return fail(CTR_EINVAL, std::string(who) + ": workspace must be 256-byte aligned");
    DeviceGuard guard(p->device);
    if (!guard.ok) return fail_cuda(guard.err, "cudaSetDevice");
    cudaStream_t st = (cudaStream_t)stream;
    const int NBb = ctr::bp_nb_for_batch(B, mode == CTR_ADJOINT_EXACT ? CTR_ADJ_EXACT : CTR_ADJ_TF, p->X, p->Y);
    const int G = (B + NBb - 1) / NBb;
    float* spk = (float*)ws;
    {
        dim3 grid((p->W + 2 + 127) / 128, A, G), block(128);
        ProfScope prof(CTR_K_PACK_SINO, st);
        if (NBb == 32) ctr::ctr_pack_sino_kernel<32><<<grid, block, 0, st>>>(dsino, B, A, p->W, spk);
        else if (NBb == 16) ctr::ctr_pack_sino_kernel<16><<<grid, block, 0, st>>>(dsino, B, A, p->W, spk);
        else ctr::ctr_pack_sino_kernel<8><<<grid, block, 0, st>>>(dsino, B, A, p->W, spk);
        ctr::launch_counter()++;
        CTR_CUDA(cudaGetLastError());
    }
    ctr::BpParams bp{};
    bp.spk = spk;
    bp.table = (mode == CTR_ADJOINT_EXACT) ? p->d_t : p->d_tinv;
    bp.cs = nullptr;
    bp.out = dimg;
    bp.B = B; bp.A = A; bp.X = p->X; bp.Y = p->Y; bp.H = p->H; bp.W = p->W; bp.padx = p->padx; bp.pady = p->pady;
    bp.win = p->W + 2;   // clamped to the tile's window size by the launcher
    bp.scale = scale;
    bp.sel = sel;
    if (xg) bp.xg = *xg; else bp.xg.nranks = 1;
    cudaError_t e;
    ProfScope prof(mode == CTR_ADJOINT_EXACT ? CTR_K_ADJ_EXACT : CTR_K_ADJ_TF, st);
    if (mode == CTR_ADJOINT_EXACT)
        e = (interp == CTR_INTERP_NEAREST) ? ctr::launch_bp<CTR_ADJ_EXACT, CTR_NEAREST>(bp, NBb, st)
                                           : ctr::launch_bp<CTR_ADJ_EXACT, CTR_BILINEAR>(bp, NBb, st);
    else
        e = (interp == CTR_INTERP_NEAREST) ? ctr::launch_bp<CTR_ADJ_TF, CTR_NEAREST>(bp, NBb, st)
                                           : ctr::launch_bp<CTR_ADJ_TF, CTR_BILINEAR>(bp, NBb, st);
    if (e != cudaSuccess) return fail_cuda(e, "ctr_bp_kernel launch");
    return CTR_OK;
}

int ctr_radon_adjoint_scaled(const ctr_plan* p, const float* dsino, float* dimg, int B, int interp, int mode, float scale,
                             void* ws, size_t ws_bytes, void* stream)
{
    return adjoint_impl(p, dsino, dimg, B, interp, mode, scale, nullptr, 0, nullptr, ws, ws_bytes, stream, "ctr_radon_adjoint");
}

// ------------------------------------------------------------------------------------------ FBP
int ctr_fbp_plan_create(const double* theta, int A, int P, int x_size, int y_size, const double* fr, const double* fi,
                        int device, ctr_fbp_plan** out)
{
    if (!out) return fail(CTR_EINVAL, "ctr_fbp_plan_create: out is NULL");
    *out = nullptr;
    if (!theta || !fr || A <= 0 || P <= 0 || x_size <= 0 || y_size <= 0)
        return fail(CTR_EINVAL, "ctr_fbp_plan_create: bad argument");
    if ((size_t)P * (32 + 2) * 4 > 200 * 1024) return fail(CTR_EUNSUPPORTED, "ctr_fbp_plan_create: P too large for the smem filter");
    ctr_fbp_plan* p = new (std::nothrow) ctr_fbp_plan();
    if (!p) return fail(CTR_EINVAL, "ctr_fbp_plan_create: out of host memory");
    p->device = device; p->A = A; p->P = P; p->x_size = x_size; p->y_size = y_size;
    std::vector<double> cs((size_t)A * 2), hd(P);
    for (int a = 0; a < A; ++a) { cs[2 * a] = std::cos(theta[a]); cs[2 * a + 1] = std::sin(theta[a]); }
    ctr_filter_to_spatial(fr, fi, P, hd.data());
    std::vector<float> hf(P);
    for (int k = 0; k < P; ++k) hf[k] = (float)hd[k];
    // "ramp" (and the identity filter): real(ifft(filter)) is zero at the even offsets 2, 4, ... up to the rounding of
    // the transform (measured 3e-17 relative); those taps are dropped and the row filter does half the work
    std::vector<float> hs;
    if (P % 2 == 0) {
        double hmax = 0.0, emax = 0.0;
        for (int k = 0; k < P; ++k) hmax = std::max(hmax, std::fabs(hd[k]));
        for (int k = 2; k < P; k += 2) emax = std::max(emax, std::fabs(hd[k]));
        if (emax <= 1e-12 * hmax) {
            p->sparse = 1;
            p->h0 = hf[0];
            hs.assign((size_t)1 + P + ctr::kFiltBPTMax, 0.f);
            for (int i = 0; i < P + ctr::kFiltBPTMax; ++i) hs[1 + i] = hf[(2 * i + 1) % P];
        }
    }
    DeviceGuard guard(device);
    if (!guard.ok) { int rc = fail_cuda(guard.err, "cudaSetDevice"); delete p; return rc; }
    {
        int smem_optin = 0;
        if (cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device) == cudaSuccess)
            ctr::fbp_fused_shape(x_size, y_size, P, smem_optin - 1024, p->fused_cl, p->fused_ab);
    }
    cudaError_t e;
    if ((e = cudaMalloc(&p->d_cs, cs.size() * sizeof(double))) != cudaSuccess ||
        (e = cudaMalloc(&p->d_h, hf.size() * sizeof(float))) != cudaSuccess ||
        (e = cudaMemcpy(p->d_cs, cs.data(), cs.size() * sizeof(double), cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = cudaMemcpy(p->d_h, hf.data(), hf.size() * sizeof(float), cudaMemcpyHostToDevice)) != cudaSuccess ||
        (p->sparse && ((e = cudaMalloc(&p->d_hs, hs.size() * sizeof(float))) != cudaSuccess ||
                       (e = cudaMemcpy(p->d_hs, hs.data(), hs.size() * sizeof(float), cudaMemcpyHostToDevice)) != cudaSuccess))) {
        int rc = fail_cuda(e, "ctr_fbp_plan_create: table upload");
        cudaFree(p->d_cs); cudaFree(p->d_h); cudaFree(p->d_hs);
        delete p;
        return rc;
    }
    *out = p;
    return CTR_OK;
}

int ctr_fbp_plan_destroy(ctr_fbp_plan* p)
{
    if (!p) return CTR_OK;
    DeviceGuard guard(p->device);
    cudaFree(p->d_cs); cudaFree(p->d_h); cudaFree(p->d_hs);
    delete p;
    return CTR_OK;
}

int ctr_fbp_plan_set_fused(ctr_fbp_plan* p, int on)
{
    if (!p) return fail(CTR_EINVAL, "ctr_fbp_plan_set_fused: plan is NULL");
    p->use_fused = on ? 1 : 0;
    return (on && p->fused_cl == 0) ? 1 : CTR_OK;   // 1: accepted, but this geometry only has the two-kernel path
}

size_t ctr_fbp_workspace_bytes(const ctr_fbp_plan* p, int B)
{
    if (!p || B <= 0) return 0;
    return spk_bytes(B, p->A, p->P);
}

// A_total: the angle count the pi/(2A) scale refers to (the plan's own A, or the total over all angle shards)
static int fbp_impl(const ctr_fbp_plan* p, const float* sino, int A, int A_total, float* recon, int B, const CtrExchange* xg,
                    void* ws, size_t ws_bytes, void* stream)
{
    if (!p || !sino || (!recon && !xg)) return fail(CTR_EINVAL, "ctr_fbp: NULL plan or buffer");
    if (A != p->A)
        return fail(CTR_EINVAL, "The given ``theta`` does not match the number of projections in ``radon_image``.");
    if (B <= 0) return fail(CTR_EINVAL, "ctr_fbp: B must be positive");
    if (!ws || ws_bytes < ctr_fbp_workspace_bytes(p, B)) return fail(CTR_EWORKSPACE, "ctr_fbp: workspace too small");
    if (((uintptr_t)ws & 255) != 0) return fail(CTR_EINVAL, "ctr_fbp: workspace must be 256-byte aligned");
    DeviceGuard guard(p->device);
    if (!guard.ok) return fail_cuda(guard.err, "cudaSetDevice");
    cudaStream_t st = (cudaStream_t)stream;
    if (p->use_fused && p->fused_cl > 0) {
        // single kernel: the filtered rows stay in (distributed) shared memory
        ctr::FbpFusedParams fp{};
        fp.sino = sino; fp.h = p->d_h; fp.cs = p->d_cs; fp.out = recon;
        fp.hs = p->d_hs; fp.h0 = p->h0; fp.sparse = p->sparse;
        fp.B = B; fp.A = p->A; fp.P = p->P; fp.X = p->x_size; fp.Y = p->y_size; fp.AB = p->fused_ab;
        fp.scale = (float)(M_PI / (2.0 * (double)A_total));
        if (xg) fp.xg = *xg; else fp.xg.nranks = 1;
        ProfScope prof(CTR_K_FBP_FUSED, st);
        const cudaError_t e = ctr::launch_fbp_fused(fp, p->fused_cl, st);
        if (e != cudaSuccess) return fail_cuda(e, "ctr_fbp_fused_kernel launch");
        return CTR_OK;
    }
    const int NBb = ctr::bp_nb_for_batch(B, CTR_ADJ_FBP, p->x_size, p->y_size);
    const int NBf = NBb >= 16 ? 16 : 8;                                    // images per row-filter CTA
    const int G = (B + NBf - 1) / NBf;
    float* spk = (float*)ws;
    {
        ProfScope prof(CTR_K_FBP_FILTER, st);
        if (p->sparse) {
            const cudaError_t e = (NBf == 16)
                ? ctr::launch_fbp_filter_sparse<16>(sino, p->d_hs, p->h0, B, p->A, p->P, NBb, spk, st)
                : ctr::launch_fbp_filter_sparse<8>(sino, p->d_hs, p->h0, B, p->A, p->P, NBb, spk, st);
            if (e != cudaSuccess) return fail_cuda(e, "ctr_fbp_filter_sparse_kernel launch");
        } else {
            const size_t smem = ((size_t)p->P * (NBf + 2) + 1) * sizeof(float);
            const int pairs = (p->P + 1) / 2;                              // two adjacent bins per thread
            dim3 grid(p->A, G), block(std::min(256, (pairs + 31) / 32 * 32));
            if (NBf == 16) {
                CTR_CUDA(cudaFuncSetAttribute(ctr::ctr_fbp_filter_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                ctr::ctr_fbp_filter_kernel<16><<<grid, block, smem, st>>>(sino, p->d_h, B, p->A, p->P, NBb, spk);
            } else {
                CTR_CUDA(cudaFuncSetAttribute(ctr::ctr_fbp_filter_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                ctr::ctr_fbp_filter_kernel<8><<<grid, block, smem, st>>>(sino, p->d_h, B, p->A, p->P, NBb, spk);
            }
        }
        ctr::launch_counter()++;
        CTR_CUDA(cudaGetLastError());
    }
    ctr::BpParams bp{};
    if (xg) bp.xg = *xg; else bp.xg.nranks = 1;
    bp.spk = spk; bp.table = nullptr; bp.cs = p->d_cs; bp.out = recon;
    bp.B = B; bp.A = p->A; bp.X = p->x_size; bp.Y = p->y_size; bp.H = p->P; bp.W = p->P; bp.padx = 0; bp.pady = 0;
    bp.win = p->P + 2;
    bp.scale = (float)(M_PI / (2.0 * (double)A_total));
    cudaError_t e;
    {
        ProfScope prof(CTR_K_FBP_BP, st);
        e = ctr::launch_bp<CTR_ADJ_FBP, CTR_BILINEAR>(bp, NBb, st);
    }
    if (e != cudaSuccess) return fail_cuda(e, "ctr_bp_kernel<FBP> launch");
    return CTR_OK;
}

int ctr_fbp(const ctr_fbp_plan* p, const float* sino, int A, float* recon, int B, void* ws, size_t ws_bytes, void* stream)
{
    return fbp_impl(p, sino, A, p ? p->A : A, recon, B, nullptr, ws, ws_bytes, stream);
}

// ------------------------------------------------------------------------------------------ host pipeline
// Host-buffer entry points: the batch is cut into chunks that flow through a ring of device
// staging slots on three streams (copy-in / kernels / copy-out), so PCIe runs in both
// directions while the kernels of another chunk execute.  Successive calls on the same pipe
// append to the same ring: the copy-out of one call overlaps the copy-in of the next.
// Pageable host memory (a plain NumPy array, what the reference's callers pass) is staged through
// page-locked buffers that belong to the slots: worker threads copy chunk k+1 into its slot's pinned
// buffer while chunk k crosses PCIe; results land in the slot's pinned buffer and are copied out to
// the caller's array by ctr_hostpipe_wait (or when the slot comes round again).
struct ctr_hostpipe {
    static constexpr int kSlots = 6;   // r1 trace: with 3 the next call's copy-in stalls on slots still queued for kernels
    const ctr_plan* plan = nullptr;
    int device = 0, chunk = 0;
    size_t buf_bytes = 0, ws_bytes = 0;
    cudaStream_t s_in = nullptr, s_comp = nullptr, s_out = nullptr;
    struct Slot {
        float* d_in = nullptr; float* d_out = nullptr; void* ws = nullptr;
        float* h_in = nullptr; float* h_out = nullptr;        // page-locked staging, allocated on first pageable use
        cudaEvent_t in_done = nullptr, comp_done = nullptr, out_done = nullptr;
        float* user_out = nullptr; size_t user_out_bytes = 0;  // pending copy h_out -> caller's pageable array
    } slot[kSlots];
    long long seq = 0;     // chunks issued so far
    std::mutex mu;
    // diagnostics (ctr_hostpipe_trace): timestamps of every chunk's stages, printed by ctr_hostpipe_wait
    struct Trace { int kind, n; cudaEvent_t in0, in1, c0, c1, o0, o1; };
    std::vector<Trace> trace;
    cudaEvent_t t0 = nullptr;
    bool tracing = false;
};

static void hostpipe_free(ctr_hostpipe* hp)
{
    for (auto& s : hp->slot) {
        cudaFree(s.d_in); cudaFree(s.d_out); cudaFree(s.ws);
        if (s.h_in) cudaFreeHost(s.h_in);
        if (s.h_out) cudaFreeHost(s.h_out);
        if (s.in_done) cudaEventDestroy(s.in_done);
        if (s.comp_done) cudaEventDestroy(s.comp_done);
        if (s.out_done) cudaEventDestroy(s.out_done);
    }
    if (hp->s_in) cudaStreamDestroy(hp->s_in);
    if (hp->s_comp) cudaStreamDestroy(hp->s_comp);
    if (hp->s_out) cudaStreamDestroy(hp->s_out);
    delete hp;
}

int ctr_hostpipe_create(const ctr_plan* plan, int chunk, ctr_hostpipe** out)
{
    if (!out) return fail(CTR_EINVAL, "ctr_hostpipe_create: out is NULL");
    *out = nullptr;
    if (!plan || chunk <= 0) return fail(CTR_EINVAL, "ctr_hostpipe_create: need a plan and chunk > 0");
    DeviceGuard guard(plan->device);
    if (!guard.ok) return fail_cuda(guard.err, "cudaSetDevice");
    ctr_hostpipe* hp = new (std::nothrow) ctr_hostpipe();
    if (!hp) return fail(CTR_EINVAL, "ctr_hostpipe_create: out of host memory");
    hp->plan = plan; hp->device = plan->device; hp->chunk = chunk;
    const size_t img_b = (size_t)chunk * plan->X * plan->Y * sizeof(float), sino_b = (size_t)chunk * plan->A * plan->W * sizeof(float);
    hp->buf_bytes = align_up(std::max(img_b, sino_b), 256);
    hp->ws_bytes = std::max(ctr_forward_workspace_bytes(plan, chunk), ctr_adjoint_workspace_bytes(plan, chunk));
    cudaError_t e = cudaSuccess;
    auto ok = [&](cudaError_t r) { if (e == cudaSuccess) e = r; return e == cudaSuccess; };
    ok(cudaStreamCreateWithFlags(&hp->s_in, cudaStreamNonBlocking));
    ok(cudaStreamCreateWithFlags(&hp->s_comp, cudaStreamNonBlocking));
    ok(cudaStreamCreateWithFlags(&hp->s_out, cudaStreamNonBlocking));
    for (auto& s : hp->slot) {
        ok(cudaMalloc((void**)&s.d_in, hp->buf_bytes));
        ok(cudaMalloc((void**)&s.d_out, hp->buf_bytes));
        ok(cudaMalloc(&s.ws, hp->ws_bytes));
        ok(cudaEventCreateWithFlags(&s.in_done, cudaEventDisableTiming));
        ok(cudaEventCreateWithFlags(&s.comp_done, cudaEventDisableTiming));
        ok(cudaEventCreateWithFlags(&s.out_done, cudaEventDisableTiming));
    }
    if (e != cudaSuccess) { hostpipe_free(hp); return fail_cuda(e, "ctr_hostpipe_create"); }
    *out = hp;
    return CTR_OK;
}

int ctr_hostpipe_destroy(ctr_hostpipe* hp)
{
    if (!hp) return CTR_OK;
    DeviceGuard guard(hp->device);
    cudaStreamSynchronize(hp->s_in); cudaStreamSynchronize(hp->s_comp); cudaStreamSynchronize(hp->s_out);
    hostpipe_free(hp);
    return CTR_OK;
}

// is this host pointer page-locked (cudaHostAlloc / cudaHostRegister / torch pin_memory)?
static bool host_is_pinned(const void* p)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { (void)cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost;
}

// host-to-host copy between a caller's pageable array and a slot's pinned buffer, split over a few threads
// (one core moves ~10 GB/s, less than one PCIe Gen5 x16 direction)
static void host_copy(void* dst, const void* src, size_t bytes)
{
    constexpr size_t kMinPart = 1u << 20;
    int parts = (int)std::min<size_t>(4, bytes / kMinPart);
    if (parts <= 1) { std::memcpy(dst, src, bytes); return; }
    const size_t per = align_up((bytes + parts - 1) / parts, 4096);
    std::vector<std::thread> th;
    for (int k = 1; k < parts; ++k) {
        const size_t lo = (size_t)k * per;
        if (lo >= bytes) break;
        const size_t n = std::min(per, bytes - lo);
        th.emplace_back([=] { std::memcpy((char*)dst + lo, (const char*)src + lo, n); });
    }
    std::memcpy(dst, src, std::min(per, bytes));
    for (auto& t : th) t.join();
}

// deliver a slot's staged result to the caller's pageable array (if one is pending)
static int hostpipe_drain_slot(ctr_hostpipe::Slot& s)
{
    if (!s.user_out) return CTR_OK;
    CTR_CUDA(cudaEventSynchronize(s.out_done));
    host_copy(s.user_out, s.h_out, s.user_out_bytes);
    s.user_out = nullptr;
    return CTR_OK;
}

// kind 0: forward (in = images [B,X,Y], out = sinograms [B,A,W]); kind 1: adjoint (the mirror)
static int hostpipe_run_locked(ctr_hostpipe* hp, int kind, const float* in_host, float* out_host, int B, int interp, int mode)
{
    const ctr_plan* p = hp->plan;
    const size_t img_n = (size_t)p->X * p->Y, sino_n = (size_t)p->A * p->W;
    const size_t in_n = kind == 0 ? img_n : sino_n, out_n = kind == 0 ? sino_n : img_n;
    const bool in_pinned = host_is_pinned(in_host), out_pinned = host_is_pinned(out_host);
    for (int lo = 0; lo < B; lo += hp->chunk) {
        const int n = std::min(hp->chunk, B - lo);
        ctr_hostpipe::Slot& s = hp->slot[hp->seq % ctr_hostpipe::kSlots];
        const bool reused = hp->seq >= ctr_hostpipe::kSlots;
        int rc = hostpipe_drain_slot(s);                       // a pageable result of the slot's previous chunk
        if (rc != CTR_OK) return rc;
        const float* src = in_host + (size_t)lo * in_n;
        const size_t in_bytes = (size_t)n * in_n * sizeof(float), out_bytes = (size_t)n * out_n * sizeof(float);
        if (!in_pinned) {
            if (!s.h_in) CTR_CUDA(cudaHostAlloc((void**)&s.h_in, hp->buf_bytes, cudaHostAllocDefault));
            if (reused) CTR_CUDA(cudaEventSynchronize(s.in_done));   // the previous upload from this staging buffer
            host_copy(s.h_in, src, in_bytes);
            src = s.h_in;
        }
        if (!out_pinned && !s.h_out) CTR_CUDA(cudaHostAlloc((void**)&s.h_out, hp->buf_bytes, cudaHostAllocDefault));
        if (reused) CTR_CUDA(cudaStreamWaitEvent(hp->s_in, s.comp_done, 0));     // staging input consumed
        const bool tracing = hp->tracing;
        ctr_hostpipe::Trace tr{kind, n, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
        if (tracing) {
            for (cudaEvent_t* ev : {&tr.in0, &tr.in1, &tr.c0, &tr.c1, &tr.o0, &tr.o1}) cudaEventCreate(ev);
            if (!hp->t0) { cudaEventCreate(&hp->t0); cudaEventRecord(hp->t0, hp->s_in); }
            cudaEventRecord(tr.in0, hp->s_in);
        }
        CTR_CUDA(cudaMemcpyAsync(s.d_in, src, in_bytes, cudaMemcpyHostToDevice, hp->s_in));
        CTR_CUDA(cudaEventRecord(s.in_done, hp->s_in));
        if (tracing) cudaEventRecord(tr.in1, hp->s_in);
        CTR_CUDA(cudaStreamWaitEvent(hp->s_comp, s.in_done, 0));
        if (reused) CTR_CUDA(cudaStreamWaitEvent(hp->s_comp, s.out_done, 0));    // staging output copied out
        if (tracing) cudaEventRecord(tr.c0, hp->s_comp);
        rc = kind == 0 ? ctr_radon_forward(p, s.d_in, s.d_out, n, interp, s.ws, hp->ws_bytes, hp->s_comp)
                       : ctr_radon_adjoint(p, s.d_in, s.d_out, n, interp, mode, s.ws, hp->ws_bytes, hp->s_comp);
        if (rc != CTR_OK) return rc;
        CTR_CUDA(cudaEventRecord(s.comp_done, hp->s_comp));
        if (tracing) cudaEventRecord(tr.c1, hp->s_comp);
        CTR_CUDA(cudaStreamWaitEvent(hp->s_out, s.comp_done, 0));
        if (tracing) cudaEventRecord(tr.o0, hp->s_out);
        float* dst = out_pinned ? out_host + (size_t)lo * out_n : s.h_out;
        CTR_CUDA(cudaMemcpyAsync(dst, s.d_out, out_bytes, cudaMemcpyDeviceToHost, hp->s_out));
        CTR_CUDA(cudaEventRecord(s.out_done, hp->s_out));
        if (!out_pinned) { s.user_out = out_host + (size_t)lo * out_n; s.user_out_bytes = out_bytes; }
        if (tracing) { cudaEventRecord(tr.o1, hp->s_out); hp->trace.push_back(tr); }
        ++hp->seq;
    }
    return CTR_OK;
}

static int hostpipe_run(ctr_hostpipe* hp, int kind, const float* in_host, float* out_host, int B, int interp, int mode,
                        const char* who)
{
    if (!hp || !in_host || !out_host) return fail(CTR_EINVAL, std::string(who) + ": NULL pipe or buffer");
    if (B <= 0) return fail(CTR_EINVAL, std::string(who) + ": B must be positive");
    DeviceGuard guard(hp->device);
    if (!guard.ok) return fail_cuda(guard.err, "cudaSetDevice");
    std::lock_guard<std::mutex> lk(hp->mu);
    const int rc = hostpipe_run_locked(hp, kind, in_host, out_host, B, interp, mode);
    if (rc != CTR_OK) {
        // chunks enqueued before the failure still read / write the caller's buffers: let all three streams run
        // dry before the error (and the right to release those buffers) goes back to the caller
        const std::string msg = g_err;
        cudaStreamSynchronize(hp->s_in); cudaStreamSynchronize(hp->s_comp); cudaStreamSynchronize(hp->s_out);
        for (auto& s : hp->slot) s.user_out = nullptr;
        (void)cudaGetLastError();
        g_err = msg;
    }
    return rc;
}

int ctr_hostpipe_forward(ctr_hostpipe* hp, const float* img_host, float* sino_host, int B, int interp)
{
    if (interp != CTR_INTERP_NEAREST && interp != CTR_INTERP_BILINEAR) return fail(CTR_EINVAL, "ctr_hostpipe_forward: bad interp");
    return hostpipe_run(hp, 0, img_host, sino_host, B, interp, 0, "ctr_hostpipe_forward");
}

int ctr_hostpipe_adjoint(ctr_hostpipe* hp, const float* dsino_host, float* dimg_host, int B, int interp, int mode)
{
    if (interp != CTR_INTERP_NEAREST && interp != CTR_INTERP_BILINEAR) return fail(CTR_EINVAL, "ctr_hostpipe_adjoint: bad interp");
    if (mode != CTR_ADJOINT_EXACT && mode != CTR_ADJOINT_TF_COMPAT) return fail(CTR_EINVAL, "ctr_hostpipe_adjoint: bad mode");
    return hostpipe_run(hp, 1, dsino_host, dimg_host, B, interp, mode, "ctr_hostpipe_adjoint");
}

int ctr_hostpipe_trace(ctr_hostpipe* hp, int on)
{
    if (!hp) return fail(CTR_EINVAL, "ctr_hostpipe_trace: pipe is NULL");
    std::lock_guard<std::mutex> lk(hp->mu);
    hp->tracing = on != 0;
    return CTR_OK;
}

int ctr_hostpipe_wait(ctr_hostpipe* hp)
{
    if (!hp) return fail(CTR_EINVAL, "ctr_hostpipe_wait: pipe is NULL");
    DeviceGuard guard(hp->device);
    if (!guard.ok) return fail_cuda(guard.err, "cudaSetDevice");
    std::lock_guard<std::mutex> lk(hp->mu);
    CTR_CUDA(cudaStreamSynchronize(hp->s_out));   // every result of every call issued so far has left the device
    for (auto& s : hp->slot) {                     // ... and pageable results go from the staging buffers to the caller
        const int rc = hostpipe_drain_slot(s);
        if (rc != CTR_OK) return rc;
    }
    if (!hp->trace.empty()) {
        for (auto& t : hp->trace) {
            float v[6];
            cudaEvent_t evs[6] = {t.in0, t.in1, t.c0, t.c1, t.o0, t.o1};
            for (int q = 0; q < 6; ++q) { cudaEventElapsedTime(&v[q], hp->t0, evs[q]); cudaEventDestroy(evs[q]); }
            fprintf(stderr, "[hostpipe] %s n=%d  in %.3f-%.3f  kernels %.3f-%.3f  out %.3f-%.3f ms\n", t.kind ? "adj" : "fwd", t.n,
                    v[0], v[1], v[2], v[3], v[4], v[5]);
        }
        hp->trace.clear();
        cudaEventDestroy(hp->t0);
        hp->t0 = nullptr;
    }
    return CTR_OK;
}

int ctr_hostpipe_done(ctr_hostpipe* hp)
{
    if (!hp) return fail(CTR_EINVAL, "ctr_hostpipe_done: pipe is NULL");
    DeviceGuard guard(hp->device);
    std::lock_guard<std::mutex> lk(hp->mu);
    for (auto& s : hp->slot)
        if (s.user_out) return 0;                  // a pageable result still waits for ctr_hostpipe_wait to deliver it
    const cudaError_t e = cudaStreamQuery(hp->s_out);
    if (e == cudaSuccess) return 1;
    if (e == cudaErrorNotReady) return 0;
    return fail_cuda(e, "cudaStreamQuery");
}

// ------------------------------------------------------------------------------------------ angle-sharded exchange
// One ctr_comm per rank (one rank per GPU).  Every rank owns an allocation
//   [ flags: CTR_MAX_RANKS x u32 | err: i32 @256 | exchange buffers @4096: 2 parities x `bytes` ]
// that every other rank maps over NVLink peer memory (CUDA IPC between processes, cudaDeviceEnablePeerAccess inside
// one process).  The adjoint / FBP kernels of an angle shard store their partial images straight into the owners'
// buffers; ctr_xchg_sum_kernel then exchanges "done" flags and sums the slots (see ctr_kernels.cuh).
// Two parities: call k+1 writes the other half while a slower rank may still be summing call k; the flag exchange of
// call k+1 keeps anyone from starting call k+2 before everybody has finished summing call k.
namespace {
constexpr size_t kXgHeader = 4096;
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*CommAbort)(ncclComm_t) = nullptr;
    ncclResult_t (*CommGetAsyncError)(ncclComm_t, ncclResult_t*) = nullptr;
    ncclResult_t (*ReduceScatter)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
std::mutex g_nccl_mu;
NcclApi g_nccl;

// libnccl.so.2 is opened on first use: a process that already loaded it (torch) shares that copy
const NcclApi* nccl_api()
{
    std::lock_guard<std::mutex> lk(g_nccl_mu);
    if (g_nccl.lib) return &g_nccl;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return nullptr;
    NcclApi a;
    a.lib = h;
#define CTR_SYM(field, name) *(void**)(&a.field) = dlsym(h, name); if (!a.field) { dlclose(h); return nullptr; }
    CTR_SYM(GetUniqueId, "ncclGetUniqueId")
    CTR_SYM(CommInitRank, "ncclCommInitRank")
    CTR_SYM(CommInitAll, "ncclCommInitAll")
    CTR_SYM(CommDestroy, "ncclCommDestroy")
    CTR_SYM(CommAbort, "ncclCommAbort")
    CTR_SYM(CommGetAsyncError, "ncclCommGetAsyncError")
    CTR_SYM(ReduceScatter, "ncclReduceScatter")
    CTR_SYM(AllReduce, "ncclAllReduce")
    CTR_SYM(GetErrorString, "ncclGetErrorString")
#undef CTR_SYM
    g_nccl = a;
    return &g_nccl;
}

struct CommHandle {                 // what ctr_comm_export writes (CTR_COMM_HANDLE_BYTES)
    cudaIpcMemHandle_t mem;         // 64 bytes
    unsigned long long bytes;
    int rank, nranks;
    unsigned magic;
    char pad_[CTR_COMM_HANDLE_BYTES - 64 - 8 - 4 - 4 - 4];
};
static_assert(sizeof(CommHandle) == CTR_COMM_HANDLE_BYTES, "handle layout");
constexpr unsigned kCommMagic = 0x43545243u;   // "CTRC"
}  // namespace

struct ctr_comm {
    int nranks = 1, rank = 0, device = 0;
    size_t bytes = 0;                            // one parity of the exchange buffer (256-byte multiple)
    void* base = nullptr;
    void* peer_base[CTR_MAX_RANKS] = {};
    bool ipc_opened[CTR_MAX_RANKS] = {};
    bool connected = false;
    unsigned epoch = 0;
    unsigned long long timeout_ns = 10ull * 1000 * 1000 * 1000;
    std::mutex mu;
    ncclComm_t nccl = nullptr;
    float* nccl_partial = nullptr;               // [B][X][Y] partial of the NCCL algorithm (grown on demand)
    size_t nccl_partial_bytes = 0;
    int sm_count = 148;
};

static int comm_alloc(int nranks, int rank, int device, size_t bytes, ctr_comm** out)
{
    if (!out) return fail(CTR_EINVAL, "ctr_comm_create: out is NULL");
    *out = nullptr;
    if (nranks < 1 || nranks > CTR_MAX_RANKS || rank < 0 || rank >= nranks)
        return fail(CTR_EINVAL, "ctr_comm_create: need 1 <= nranks <= 16 and 0 <= rank < nranks");
    if (bytes == 0) return fail(CTR_EINVAL, "ctr_comm_create: exchange bytes must be positive");
    DeviceGuard guard(device);
    if (!guard.ok) return fail_cuda(guard.err, "cudaSetDevice");
    ctr_comm* c = new (std::nothrow) ctr_comm();
    if (!c) return fail(CTR_EINVAL, "ctr_comm_create: out of host memory");
    c->nranks = nranks; c->rank = rank; c->device = device;
    c->bytes = align_up(bytes, 256);
    if (cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || c->sm_count <= 0) c->sm_count = 148;
    cudaError_t e = cudaMalloc(&c->base, kXgHeader + 2 * c->bytes);
    if (e == cudaSuccess) e = cudaMemset(c->base, 0, kXgHeader);
    if (e != cudaSuccess) { cudaFree(c->base); delete c; return fail_cuda(e, "ctr_comm_create: exchange buffer"); }
    c->peer_base[rank] = c->base;
    c->connected = (nranks == 1);
    *out = c;
    return CTR_OK;
}

int ctr_comm_create(int nranks, int rank, int device, size_t exchange_bytes, ctr_comm** out)
{
    return comm_alloc(nranks, rank, device, exchange_bytes, out);
}

int ctr_comm_export(const ctr_comm* c, void* handle_out)
{
    if (!c || !handle_out) return fail(CTR_EINVAL, "ctr_comm_export: NULL argument");
    DeviceGuard guard(c->device);
    if (!guard.ok) return fail_cuda(guard.err, "cudaSetDevice");
    CommHandle h;
    std::memset(&h, 0, sizeof(h));
    CTR_CUDA(cudaIpcGetMemHandle(&h.mem, c->base));
    h.bytes = c->bytes; h.rank = c->rank; h.nranks = c->nranks; h.magic = kCommMagic;
    std::memcpy(handle_out, &h, sizeof(h));
    return CTR_OK;
}

int ctr_comm_connect(ctr_comm* c, const void* handles)
{
    if (!c || !handles) return fail(CTR_EINVAL, "ctr_comm_connect: NULL argument");
    DeviceGuard guard(c->device);
    if (!guard.ok) return fail_cuda(guard.err, "cudaSetDevice");
    std::lock_guard<std::mutex> lk(c->mu);
    const CommHandle* hs = (const CommHandle*)handles;
    for (int s = 0; s < c->nranks; ++s) {
        if (hs[s].magic != kCommMagic || hs[s].rank != s || hs[s].nranks != c->nranks)
            return fail(CTR_EINVAL, "ctr_comm_connect: handle " + std::to_string(s) + " is not rank " + std::to_string(s) + "'s export");
        if (hs[s].bytes != c->bytes) return fail(CTR_EINVAL, "ctr_comm_connect: ranks disagree on the exchange size");
    }
    for (int s = 0; s < c->nranks; ++s) {
        if (s == c->rank || c->peer_base[s]) continue;
        cudaError_t e = cudaIpcOpenMemHandle(&c->peer_base[s], hs[s].mem, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) { c->peer_base[s] = nullptr; return fail_cuda(e, "ctr_comm_connect: cudaIpcOpenMemHandle (is NVLink/PCIe peer access available?)"); }
        c->ipc_opened[s] = true;
    }
    c->connected = true;
    return CTR_OK;
}

int ctr_comm_create_all(int nranks, const int* devices, size_t exchange_bytes, ctr_comm** out)
{
    if (!devices || !out) return fail(CTR_EINVAL, "ctr_comm_create_all: NULL argument");
    if (nranks < 1 || nranks > CTR_MAX_RANKS) return fail(CTR_EINVAL, "ctr_comm_create_all: need 1 <= nranks <= 16");
    for (int r = 0; r < nranks; ++r) out[r] = nullptr;
    int rc = CTR_OK;
    for (int r = 0; r < nranks && rc == CTR_OK; ++r) rc = comm_alloc(nranks, r, devices[r], exchange_bytes, &out[r]);
    for (int r = 0; r < nranks && rc == CTR_OK; ++r) {
        DeviceGuard guard(devices[r]);
        if (!guard.ok) { rc = fail_cuda(guard.err, "cudaSetDevice"); break; }
        for (int s = 0; s < nranks; ++s) {
            if (s == r) continue;
            if (devices[s] != devices[r]) {
                int can = 0;
                cudaDeviceCanAccessPeer(&can, devices[r], devices[s]);
                if (!can) { rc = fail(CTR_EUNSUPPORTED, "ctr_comm_create_all: no peer access between the devices"); break; }
                const cudaError_t e = cudaDeviceEnablePeerAccess(devices[s], 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { rc = fail_cuda(e, "cudaDeviceEnablePeerAccess"); break; }
                (void)cudaGetLastError();
            }
            out[r]->peer_base[s] = out[s]->base;
        }
        out[r]->connected = true;
    }
    if (rc != CTR_OK)
        for (int r = 0; r < nranks; ++r) { ctr_comm_destroy(out[r]); out[r] = nullptr; }
    return rc;
}

int ctr_comm_set_timeout_ms(ctr_comm* c, int ms)
{
    if (!c || ms <= 0) return fail(CTR_EINVAL, "ctr_comm_set_timeout_ms: bad argument");
    c->timeout_ns = (unsigned long long)ms * 1000000ull;
    return CTR_OK;
}

int ctr_comm_info(const ctr_comm* c, int* nranks, int* rank, int* connected, int* has_nccl, size_t* exchange_bytes)
{
    if (!c) return fail(CTR_EINVAL, "ctr_comm_info: comm is NULL");
    if (nranks) *nranks = c->nranks;
    if (rank) *rank = c->rank;
    if (connected) *connected = c->connected ? 1 : 0;
    if (has_nccl) *has_nccl = c->nccl ? 1 : 0;
    if (exchange_bytes) *exchange_bytes = c->bytes;
    return CTR_OK;
}

int ctr_comm_nccl_unique_id(void* id_out)
{
    if (!id_out) return fail(CTR_EINVAL, "ctr_comm_nccl_unique_id: NULL argument");
    const NcclApi* n = nccl_api();
    if (!n) return fail(CTR_EUNSUPPORTED, "libnccl.so.2 could not be opened");
    ncclUniqueId id;
    const ncclResult_t r = n->GetUniqueId(&id);
    if (r != ncclSuccess) return fail(CTR_ECOMM, std::string("ncclGetUniqueId: ") + n->GetErrorString(r));
    std::memcpy(id_out, &id, sizeof(id));
    return CTR_OK;
}

int ctr_comm_nccl_init(ctr_comm* c, const void* id_in)
{
    if (!c || !id_in) return fail(CTR_EINVAL, "ctr_comm_nccl_init: NULL argument");
    const NcclApi* n = nccl_api();
    if (!n) return fail(CTR_EUNSUPPORTED, "libnccl.so.2 could not be opened");
    DeviceGuard guard(c->device);
    if (!guard.ok) return fail_cuda(guard.err, "cudaSetDevice");
    ncclUniqueId id;
    std::memcpy(&id, id_in, sizeof(id));
    const ncclResult_t r = n->CommInitRank(&c->nccl, c->nranks, id, c->rank);
    if (r != ncclSuccess) { c->nccl = nullptr; return fail(CTR_ECOMM, std::string("ncclCommInitRank: ") + n->GetErrorString(r)); }
    return CTR_OK;
}

// CTR_OK, or CTR_ECOMM when a peer missed a flag exchange (its process died or hung) or NCCL reports an
// asynchronous error (ncclCommGetAsyncError).  Synchronises with nothing but a 4-byte read.
int ctr_comm_check(ctr_comm* c)
{
    if (!c) return fail(CTR_EINVAL, "ctr_comm_check: comm is NULL");
    DeviceGuard guard(c->device);
    if (!guard.ok) return fail_cuda(guard.err, "cudaSetDevice");
    int err = 0;
    CTR_CUDA(cudaMemcpy(&err, (char*)c->base + 256, sizeof(int), cudaMemcpyDeviceToHost));
    if (err != 0)
        return fail(CTR_ECOMM, "angle-sharded exchange: rank " + std::to_string(err - 1) + " did not arrive within the timeout (rank " +
                                   std::to_string(c->rank) + " waited)");
    if (c->nccl) {
        const NcclApi* n = nccl_api();
        ncclResult_t ar = ncclSuccess;
        const ncclResult_t r = n->CommGetAsyncError(c->nccl, &ar);
        if (r != ncclSuccess) return fail(CTR_ECOMM, std::string("ncclCommGetAsyncError: ") + n->GetErrorString(r));
        if (ar != ncclSuccess && ar != ncclInProgress) return fail(CTR_ECOMM, std::string("NCCL asynchronous error: ") + n->GetErrorString(ar));
    }
    return CTR_OK;
}

int ctr_comm_destroy(ctr_comm* c)
{
    if (!c) return CTR_OK;
    DeviceGuard guard(c->device);
    cudaDeviceSynchronize();
    if (c->nccl) {
        const NcclApi* n = nccl_api();
        if (n) n->CommDestroy(c->nccl);
    }
    for (int s = 0; s < c->nranks; ++s)
        if (c->ipc_opened[s] && c->peer_base[s]) cudaIpcCloseMemHandle(c->peer_base[s]);
    cudaFree(c->nccl_partial);
    cudaFree(c->base);
    delete c;
    return CTR_OK;
}

// shared front half of the sharded calls: argument checks, the epoch's exchange target
static int exchange_begin(ctr_comm* c, int B, size_t img_floats, int algo, CtrExchange* xg, const char* who)
{
    if (!c) return fail(CTR_EINVAL, std::string(who) + ": comm is NULL");
    if (algo != CTR_EXCHANGE_P2P && algo != CTR_EXCHANGE_NCCL) return fail(CTR_EINVAL, std::string(who) + ": bad exchange algorithm");
    if (B <= 0 || B % c->nranks != 0)
        return fail(CTR_EINVAL, std::string(who) + ": the batch must be a positive multiple of the rank count (the result is left batch-sharded)");
    if (algo == CTR_EXCHANGE_NCCL) {
        if (c->nranks > 1 && !c->nccl) return fail(CTR_EINVAL, std::string(who) + ": ctr_comm_nccl_init has not been called");
        return CTR_OK;
    }
    if (!c->connected) return fail(CTR_EINVAL, std::string(who) + ": ctr_comm_connect has not been called");
    if ((size_t)B * img_floats * sizeof(float) > c->bytes) return fail(CTR_EWORKSPACE, std::string(who) + ": exchange buffer smaller than B*X*Y*4 bytes");
    // the epoch is committed by exchange_finish_p2p, once the back-projection kernel has been enqueued: a call that
    // fails before that leaves the ranks' epochs in step
    const unsigned epoch = c->epoch + 1;
    for (int s = 0; s < c->nranks; ++s) xg->peer[s] = (float*)((char*)c->peer_base[s] + kXgHeader + (size_t)(epoch & 1) * c->bytes);
    xg->nranks = c->nranks; xg->rank = c->rank; xg->Bs = B / c->nranks;
    return CTR_OK;
}

// shared back half (P2P): flag exchange + fixed-order sum of this rank's slots into `out`
static int exchange_finish_p2p(ctr_comm* c, const CtrExchange& xg, float* out, size_t n, cudaStream_t st)
{
    ctr::XchgParams xp{};
    for (int s = 0; s < c->nranks; ++s) xp.peer_flags[s] = (unsigned*)c->peer_base[s];
    xp.flags = (unsigned*)c->base;
    xp.err = (int*)((char*)c->base + 256);
    xp.slots = xg.peer[c->rank];
    xp.out = out;
    xp.n = n;
    xp.epoch = ++c->epoch;
    xp.nranks = c->nranks; xp.rank = c->rank;
    xp.timeout_ns = c->timeout_ns;
    size_t blocks = (n / 4 + 511) / 512;
    if (blocks > (size_t)c->sm_count * 2) blocks = (size_t)c->sm_count * 2;
    if (blocks < 1) blocks = 1;
    ProfScope prof(CTR_K_XCHG_SUM, st);
    ctr::ctr_xchg_sum_kernel<<<(unsigned)blocks, 512, 0, st>>>(xp);
    ctr::launch_counter()++;
    CTR_CUDA(cudaGetLastError());
    return CTR_OK;
}

static int exchange_finish_nccl(ctr_comm* c, const float* partial, float* out, size_t n_shard, cudaStream_t st, const char* who)
{
    const NcclApi* n = nccl_api();
    const ncclResult_t r = n->ReduceScatter(partial, out, n_shard, ncclFloat, ncclSum, c->nccl, st);
    if (r != ncclSuccess) return fail(CTR_ECOMM, std::string(who) + ": ncclReduceScatter: " + n->GetErrorString(r));
    ncclResult_t ar = ncclSuccess;
    if (n->CommGetAsyncError(c->nccl, &ar) == ncclSuccess && ar != ncclSuccess && ar != ncclInProgress)
        return fail(CTR_ECOMM, std::string(who) + ": NCCL asynchronous error: " + n->GetErrorString(ar));
    return CTR_OK;
}

static int nccl_partial_buffer(ctr_comm* c, size_t bytes)
{
    if (c->nccl_partial_bytes >= bytes) return CTR_OK;
    cudaFree(c->nccl_partial);
    c->nccl_partial = nullptr; c->nccl_partial_bytes = 0;
    CTR_CUDA(cudaMalloc((void**)&c->nccl_partial, bytes));
    c->nccl_partial_bytes = bytes;
    return CTR_OK;
}

size_t ctr_adjoint_sharded_workspace_bytes(const ctr_comm* c, const ctr_plan* p, int B)
{
    (void)c;
    return ctr_adjoint_workspace_bytes(p, B);
}

int ctr_radon_adjoint_sharded(ctr_comm* c, const ctr_plan* p, const float* dsino_local, float* dimg_shard, int B, int interp,
                              int mode, int algo, void* ws, size_t ws_bytes, void* stream)
{
    if (!p || !dimg_shard) return fail(CTR_EINVAL, "ctr_radon_adjoint_sharded: NULL plan or result");
    const size_t img = (size_t)p->X * p->Y;
    if (c && c->nranks == 1)
        return adjoint_impl(p, dsino_local, dimg_shard, B, interp, mode, 1.0f, nullptr, 0, nullptr, ws, ws_bytes, stream, "ctr_radon_adjoint_sharded");
    std::unique_lock<std::mutex> lk;
    if (c) lk = std::unique_lock<std::mutex>(c->mu);
    CtrExchange xg{};
    int rc = exchange_begin(c, B, img, algo, &xg, "ctr_radon_adjoint_sharded");
    if (rc != CTR_OK) return rc;
    DeviceGuard guard(c->device);
    if (!guard.ok) return fail_cuda(guard.err, "cudaSetDevice");
    if (c->device != p->device) return fail(CTR_EINVAL, "ctr_radon_adjoint_sharded: plan and comm are on different devices");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n_shard = (size_t)(B / c->nranks) * img;
    if (algo == CTR_EXCHANGE_P2P) {
        rc = adjoint_impl(p, dsino_local, nullptr, B, interp, mode, 1.0f, nullptr, 0, &xg, ws, ws_bytes, stream, "ctr_radon_adjoint_sharded");
        if (rc != CTR_OK) return rc;
        return exchange_finish_p2p(c, xg, dimg_shard, n_shard, st);
    }
    if ((rc = nccl_partial_buffer(c, (size_t)B * img * sizeof(float))) != CTR_OK) return rc;
    rc = adjoint_impl(p, dsino_local, c->nccl_partial, B, interp, mode, 1.0f, nullptr, 0, nullptr, ws, ws_bytes, stream, "ctr_radon_adjoint_sharded");
    if (rc != CTR_OK) return rc;
    return exchange_finish_nccl(c, c->nccl_partial, dimg_shard, n_shard, st, "ctr_radon_adjoint_sharded");
}

int ctr_fbp_sharded(ctr_comm* c, const ctr_fbp_plan* p, const float* sino_local, int A_local, int A_total, float* recon_shard,
                    int B, int algo, void* ws, size_t ws_bytes, void* stream)
{
    if (!p || !recon_shard) return fail(CTR_EINVAL, "ctr_fbp_sharded: NULL plan or result");
    if (A_total < A_local || A_total <= 0) return fail(CTR_EINVAL, "ctr_fbp_sharded: A_total must cover the local angles");
    const size_t img = (size_t)p->x_size * p->y_size;
    if (c && c->nranks == 1) return fbp_impl(p, sino_local, A_local, A_total, recon_shard, B, nullptr, ws, ws_bytes, stream);
    std::unique_lock<std::mutex> lk;
    if (c) lk = std::unique_lock<std::mutex>(c->mu);
    CtrExchange xg{};
    int rc = exchange_begin(c, B, img, algo, &xg, "ctr_fbp_sharded");
    if (rc != CTR_OK) return rc;
    DeviceGuard guard(c->device);
    if (!guard.ok) return fail_cuda(guard.err, "cudaSetDevice");
    if (c->device != p->device) return fail(CTR_EINVAL, "ctr_fbp_sharded: plan and comm are on different devices");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n_shard = (size_t)(B / c->nranks) * img;
    if (algo == CTR_EXCHANGE_P2P) {
        if ((rc = fbp_impl(p, sino_local, A_local, A_total, nullptr, B, &xg, ws, ws_bytes, stream)) != CTR_OK) return rc;
        return exchange_finish_p2p(c, xg, recon_shard, n_shard, st);
    }
    if ((rc = nccl_partial_buffer(c, (size_t)B * img * sizeof(float))) != CTR_OK) return rc;
    if ((rc = fbp_impl(p, sino_local, A_local, A_total, c->nccl_partial, B, nullptr, ws, ws_bytes, stream)) != CTR_OK) return rc;
    return exchange_finish_nccl(c, c->nccl_partial, recon_shard, n_shard, st, "ctr_fbp_sharded");
}

// ------------------------------------------------------------------------------------------ DLPack
static int dl_check(const DLTensor* t, const char* name, int min_ndim, int max_ndim, int device)
{
    if (!t || !t->data) return fail(CTR_EINVAL, std::string(name) + ": NULL tensor");
    if (t->device.device_type != kDLCUDA) return fail(CTR_EINVAL, std::string(name) + ": not a CUDA tensor (there is no CPU path)");
    if (t->device.device_id != device) return fail(CTR_EINVAL, std::string(name) + ": tensor is on a different device than the plan");
    if (t->ndim < min_ndim || t->ndim > max_ndim) return fail(CTR_EINVAL, std::string(name) + ": wrong rank");
    if (t->strides) {
        int64_t expect = 1;
        for (int d = t->ndim - 1; d >= 0; --d) {
            if (t->shape[d] != 1 && t->strides[d] != expect) return fail(CTR_EINVAL, std::string(name) + ": not compact row-major");
            expect *= t->shape[d];
        }
    }
    return CTR_OK;
}
static int dl_check_f32(const DLTensor* t, const char* name, int device)
{
    int rc = dl_check(t, name, 3, 4, device);
    if (rc) return rc;
    if (t->dtype.code != kDLFloat || t->dtype.bits != 32 || t->dtype.lanes != 1) return fail(CTR_EINVAL, std::string(name) + ": dtype must be float32");
    if (t->ndim == 4 && t->shape[3] != 1) return fail(CTR_EINVAL, std::string(name) + ": trailing channel axis must have size 1");
    return CTR_OK;
}
static void* dl_ptr(const DLTensor* t) { return (char*)t->data + t->byte_offset; }
static size_t dl_bytes(const DLTensor* t)
{
    size_t n = 1;
    for (int d = 0; d < t->ndim; ++d) n *= (size_t)t->shape[d];
    return n * (size_t)((t->dtype.bits * t->dtype.lanes + 7) / 8);
}

int ctr_radon_forward_dl(const ctr_plan* p, const DLTensor* img, DLTensor* sino, int interp, DLTensor* ws, void* stream)
{
    if (!p) return fail(CTR_EINVAL, "ctr_radon_forward_dl: plan is NULL");
    int rc;
    if ((rc = dl_check_f32(img, "img", p->device)) || (rc = dl_check_f32(sino, "sino", p->device)) ||
        (rc = dl_check(ws, "workspace", 1, 8, p->device)))
        return rc;
    const int64_t B = img->shape[0];
    if (img->shape[1] != p->X || img->shape[2] != p->Y) return fail(CTR_EINVAL, "img: shape does not match the plan's X, Y");
    if (sino->shape[0] != B || sino->shape[1] != p->A || sino->shape[2] != p->W) return fail(CTR_EINVAL, "sino: expected [B, A, W]");
    return ctr_radon_forward(p, (const float*)dl_ptr(img), (float*)dl_ptr(sino), (int)B, interp, dl_ptr(ws), dl_bytes(ws), stream);
}

int ctr_radon_adjoint_dl(const ctr_plan* p, const DLTensor* dsino, DLTensor* dimg, int interp, int mode, DLTensor* ws, void* stream)
{
    if (!p) return fail(CTR_EINVAL, "ctr_radon_adjoint_dl: plan is NULL");
    int rc;
    if ((rc = dl_check_f32(dsino, "dsino", p->device)) || (rc = dl_check_f32(dimg, "dimg", p->device)) ||
        (rc = dl_check(ws, "workspace", 1, 8, p->device)))
        return rc;
    const int64_t B = dsino->shape[0];
    if (dsino->shape[1] != p->A || dsino->shape[2] != p->W) return fail(CTR_EINVAL, "dsino: expected [B, A, W]");
    if (dimg->shape[0] != B || dimg->shape[1] != p->X || dimg->shape[2] != p->Y) return fail(CTR_EINVAL, "dimg: expected [B, X, Y]");
    return ctr_radon_adjoint(p, (const float*)dl_ptr(dsino), (float*)dl_ptr(dimg), (int)B, interp, mode, dl_ptr(ws), dl_bytes(ws), stream);
}

int ctr_fbp_dl(const ctr_fbp_plan* p, const DLTensor* sino, DLTensor* recon, DLTensor* ws, void* stream)
{
    if (!p) return fail(CTR_EINVAL, "ctr_fbp_dl: plan is NULL");
    int rc;
    if ((rc = dl_check_f32(sino, "sinogram", p->device)) || (rc = dl_check_f32(recon, "recon", p->device)) ||
        (rc = dl_check(ws, "workspace", 1, 8, p->device)))
        return rc;
    const int64_t B = sino->shape[0];
    if (sino->shape[2] != p->P) return fail(CTR_EINVAL, "sinogram: last axis must be num_proj_pix");
    if (recon->shape[0] != B || recon->shape[1] != p->x_size || recon->shape[2] != p->y_size) return fail(CTR_EINVAL, "recon: expected [B, x_size, y_size]");
    return ctr_fbp(p, (const float*)dl_ptr(sino), (int)sino->shape[1], (float*)dl_ptr(recon), (int)B, dl_ptr(ws), dl_bytes(ws), stream);
}

}  // extern "C"
