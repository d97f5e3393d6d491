// ctr_emu.cpp -- CPU emulation of the kernels' per-thread logic (tests only).
//
// Runs the SAME ctr_core.h / ctr_host.h code the sm_100a kernels run, with the CTA
// structure (packed layouts, strips, angle chunks, pixel tiles, bin windows)
// replayed by plain loops.  It exists so the geometry logic can be checked against
// the oracle in this GPU-less container; it is not shipped, not imported by
// ct_pvae_b200/, and is no fallback for anything.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../ct_pvae_b200/csrc/ctr_core.h"
#include "../../ct_pvae_b200/csrc/ctr_host.h"

namespace {

constexpr int NB = 4;

// [G][Vp][Up][DEPTH*NB] packs of both classes, zero halos (mirrors ctr_pack_image_kernel):
// DEPTH image groups of NB share one pixel record.
template <int DEPTH>
void pack_images(const float* img, int B, int X, int Y, const CtrClassGeom geom[2], std::vector<float>& pk0,
                 std::vector<float>& pk1)
{
    constexpr int REC = NB * DEPTH;
    const int G = (B + REC - 1) / REC;
    const int up0 = geom[0].Up, up1 = geom[1].Up;      // padded row lengths
    pk0.assign((size_t)G * (X + 2) * up0 * REC, 0.f);
    pk1.assign((size_t)G * (Y + 2) * up1 * REC, 0.f);
    for (int b = 0; b < B; ++b) {
        const int g = b / REC, n = b % REC;
        for (int r = 0; r < X; ++r)
            for (int c = 0; c < Y; ++c) {
                const float v = img[((size_t)b * X + r) * Y + c];
                pk0[(((size_t)g * (X + 2) + r + 1) * up0 + c + 1) * REC + n] = v;
                pk1[(((size_t)g * (Y + 2) + c + 1) * up1 + r + 1) * REC + n] = v;
            }
    }
}

// [G][A][NB/4 planes][W+2][4] with zero halo bins (mirrors ctr_pack_sino_kernel).
template <int NBP>
void pack_sino(const float* y, int B, int A, int W, std::vector<float>& spk)
{
    const int G = (B + NBP - 1) / NBP;
    spk.assign((size_t)G * A * (W + 2) * NBP, 0.f);
    for (int b = 0; b < B; ++b)
        for (int a = 0; a < A; ++a)
            for (int j = 0; j < W; ++j) {
                const int n = b % NBP;
                spk[((((size_t)(b / NBP) * A + a) * (NBP / 4) + n / 4) * (W + 2) + j + 1) * 4 + (n % 4)] =
                    y[((size_t)b * A + a) * W + j];
            }
}

// DEPTH = lanes per ray, NBL = images per lane (4, or 8 with the parity-swizzled loads of 32-image records)
template <int INTERP, int DEPTH, int NBL = NB, bool REUSE = true>
void forward_impl(const float* img, int B, int X, int Y, int H, int W, int padx, int pady,
                  const float* t, int A, int R, float* sino)
{
    constexpr int REC = NBL * DEPTH;
    CtrClassGeom geom[2];
    ctr_h_class_geom(X, Y, padx, pady, geom);
    std::vector<float> pk[2];
    pack_images<REC / NB>(img, B, X, Y, geom, pk[0], pk[1]);
    std::vector<CtrRay> rays;
    int n0 = 0;
    ctr_h_build_rays(t, A, rays, n0);
    const int G = (B + REC - 1) / REC;
    for (int g = 0; g < G; ++g) {
        for (size_t ri = 0; ri < rays.size(); ++ri) {
            const CtrRay& r = rays[ri];
            const CtrClassGeom& cg = geom[r.cls];
            const float* pkg = pk[r.cls].data() + (size_t)g * cg.Vp * cg.Up * REC;
            const int K = (cg.Vp + R - 1) / R;
            for (int j = 0; j < W; ++j)
              for (int gsub = 0; gsub < DEPTH; ++gsub) {   // the DEPTH lanes that share ray j
                CtrRayState s;
                ctr_ray_begin(r, cg, j, H, s);
                float acc[NBL] = {};
                const int swz = (NBL == 16) ? (j & 3) * 4 : (NBL == 8) ? (j & 1) * 4 : 0;
                for (int k = 0; k < K; ++k) {
                    // the strip buffer the TMA bulk copy would have filled: rows [kR, kR+R+1)
                    const int rows = std::min(R + 1, cg.Vp - k * R);
                    std::vector<float> strip((size_t)(R + 1) * cg.Up * REC, -1e30f);  // poison what is not loaded
                    std::memcpy(strip.data(), pkg + (size_t)k * R * cg.Up * REC, sizeof(float) * rows * cg.Up * REC);
                    if (REUSE && NBL >= 8 && INTERP == CTR_BILINEAR)   // the windowed shape's march for 8/16-image lanes
                        ctr_march_reuse<NBL, REC>(strip.data() + gsub * NBL, cg.Up, (float)((k + 1) * R + cg.offv),
                                                  k * R + cg.offv, cg.offu, r, s, acc, swz);
                    else
                        ctr_march<NBL, INTERP, REC>(strip.data() + gsub * NBL, cg.Up, (float)((k + 1) * R + cg.offv),
                                                    k * R + cg.offv, cg.offu, r, s, acc, swz);
                }
                for (int n = 0; n < NBL; ++n) {
                    const int b = (g * DEPTH + gsub) * NBL + ctr_img_of_reg<NBL>(n, swz);
                    if (b < B) sino[((size_t)b * A + r.angle) * W + j] = acc[n];
                }
            }
        }
    }
}

// Column-windowed strips (ctr_fwd_kernel with CtrChunk::wc > 0): CTA = chunk of <= NA neighbouring rays x
// detector chunk of JW bins; per strip the producer warp places a window of wc columns from the CTA's line
// families.  Everything outside the window is NaN here, so a sample that falls outside poisons its ray sum.
// Returns the number of chunks that really got a window (wc > 0), or -1 if the shape does not fit `budget`.
template <int INTERP, int DEPTH, int NBL = NB>
int forward_window_impl(const float* img, int B, int X, int Y, int H, int W, int padx, int pady, const float* t, int A,
                        int JW, int NA, int Rmax, int budget, float* sino)
{
    constexpr int REC = NBL * DEPTH;
    CtrClassGeom geom[2];
    ctr_h_class_geom(X, Y, padx, pady, geom);
    std::vector<float> pk[2];
    pack_images<REC / NB>(img, B, X, Y, geom, pk[0], pk[1]);
    std::vector<CtrRay> rays;
    std::vector<int> seg;
    int n0 = 0;
    ctr_h_build_rays(t, A, rays, n0, &seg);
    const int jchunks = (W + JW - 1) / JW;
    std::vector<CtrChunk> chunks;
    size_t strip_bytes = 0;
    if (!ctr_h_build_chunks(rays, seg, geom, NA, W, JW, jchunks, REC * 4, 2, (size_t)budget, true, 0, 2, Rmax, chunks, strip_bytes))
        return -1;
    int windowed = 0, covered = 0;
    const int G = (B + REC - 1) / REC;
    for (const CtrChunk& ch : chunks) {
        windowed += ch.wc > 0;
        covered += ch.cnt;
        const CtrClassGeom& cg = geom[ch.cls];
        const int R = ch.R, Us = ch.wc > 0 ? ch.wc : cg.Up, K = (cg.Vp + R - 1) / R;
        for (int z = 0; z < jchunks; ++z)
            for (int g = 0; g < G; ++g) {
                const float* pkg = pk[ch.cls].data() + (size_t)g * cg.Vp * cg.Up * REC;
                // consumer state of every (ray, bin, group) thread of the CTA
                const int nb = std::min(JW, W - z * JW);
                std::vector<CtrRayState> st((size_t)ch.cnt * nb);
                std::vector<float> acc((size_t)ch.cnt * nb * REC, 0.f);
                for (int q = 0; q < ch.cnt; ++q)
                    for (int jj = 0; jj < nb; ++jj) ctr_ray_begin(rays[ch.first + q], cg, z * JW + jj, H, st[(size_t)q * nb + jj]);
                for (int k = 0; k < K; ++k) {
                    // producer warp
                    int c0 = 0;
                    if (ch.wc > 0) {
                        const float jlo = (float)(z * JW), jhi = (float)std::min(W - 1, z * JW + JW - 1);
                        float umin = 3.0e38f, umax = -3.0e38f;
                        for (int q = 0; q < ch.cnt; ++q)
                            ctr_win_range(ctr_win_coef(rays[ch.first + q]), jlo, jhi, (float)(k * R + cg.offv) - 0.5f,
                                          (float)(k * R + cg.offv + R), cg.ulo, cg.uhi, umin, umax);
                        c0 = ctr_win_start(umin, cg.offu, cg.Up, Us);
                    }
                    const int rows = std::min(R + 1, cg.Vp - k * R);
                    std::vector<float> strip((size_t)(R + 1) * Us * REC, std::nanf(""));
                    for (int rr = 0; rr < rows; ++rr)
                        std::memcpy(strip.data() + (size_t)rr * Us * REC, pkg + ((size_t)(k * R + rr) * cg.Up + c0) * REC,
                                    sizeof(float) * Us * REC);
                    // consumers
                    for (int q = 0; q < ch.cnt; ++q)
                        for (int jj = 0; jj < nb; ++jj)
                            for (int gsub = 0; gsub < DEPTH; ++gsub) {
                                CtrRayState s = st[(size_t)q * nb + jj];   // the DEPTH lanes of a ray march identically
                                if (NBL >= 8 && INTERP == CTR_BILINEAR)
                                    ctr_march_reuse<NBL, REC>(strip.data() + gsub * NBL, Us, (float)((k + 1) * R + cg.offv),
                                                              k * R + cg.offv, cg.offu + c0, rays[ch.first + q], s,
                                                              &acc[((size_t)q * nb + jj) * REC + gsub * NBL],
                                                              (NBL == 16) ? (jj & 3) * 4 : (jj & 1) * 4);
                                else
                                    ctr_march<NBL, INTERP, REC>(strip.data() + gsub * NBL, Us, (float)((k + 1) * R + cg.offv),
                                                                k * R + cg.offv, cg.offu + c0, rays[ch.first + q], s,
                                                                &acc[((size_t)q * nb + jj) * REC + gsub * NBL],
                                                                (NBL == 16) ? (jj & 3) * 4 : (NBL == 8) ? (jj & 1) * 4 : 0);
                                if (gsub == DEPTH - 1) st[(size_t)q * nb + jj] = s;
                            }
                }
                for (int q = 0; q < ch.cnt; ++q)
                    for (int jj = 0; jj < nb; ++jj)
                        for (int n = 0; n < REC; ++n) {
                            // register n % NBL of lane n / NBL holds image ctr_img_of_reg(n % NBL, swz) of that lane's block
                            const int swz = (NBL == 16) ? (jj & 3) * 4 : (NBL == 8) ? (jj & 1) * 4 : 0;
                            const int b = g * REC + (n / NBL) * NBL + ctr_img_of_reg<NBL>(n % NBL, swz);
                            if (b < B) sino[((size_t)b * A + rays[ch.first + q].angle) * W + z * JW + jj] = acc[((size_t)q * nb + jj) * REC + n];
                        }
            }
    }
    return covered == A ? windowed : -2;
}

template <int MODE, int INTERP, int NBA>
void adjoint_impl(const float* y, int B, int X, int Y, int H, int W, int padx, int pady,
                  const float* table, int A, int TW, int TH, int win, float* out)
{
    constexpr int NB = NBA;  // images per back-projection CTA (shadows the forward's NB)
    std::vector<float> spk;
    pack_sino<NB>(y, B, A, W, spk);
    const int G = (B + NB - 1) / NB;
    const int Wp2 = W + 2;
    const int winc = std::min(win, Wp2);
    for (int g = 0; g < G; ++g)
        for (int r0 = 0; r0 < X; r0 += TH)
            for (int c0 = 0; c0 < Y; c0 += TW) {
                std::vector<float> acc((size_t)TW * TH * NB, 0.f);
                for (int a = 0; a < A; ++a) {
                    const float* t = table + 8 * a;
                    // producer lane: window start from the 4 tile corners
                    float cu[4];
                    int q = 0;
                    for (int cy = 0; cy < 2; ++cy)
                        for (int cx = 0; cx < 2; ++cx) {
                            const float px = (float)(c0 + cx * (TW - 1) + pady), py = (float)(r0 + cy * (TH - 1) + padx);
                            float uj, vi;
                            if (MODE == CTR_ADJ_EXACT) ctr_adj_centre(t, px, py, uj, vi);
                            else uj = CTR_ADD(CTR_ADD(CTR_MUL(t[0], px), CTR_MUL(t[1], py)), t[2]);
                            cu[q++] = uj;
                        }
                    const int start = ctr_window_start(cu[0], cu[1], cu[2], cu[3], Wp2, winc);
                    // NaN guard zones either side: an out-of-window read poisons the output
                    constexpr int GUARD = 16;
                    // one bulk copy per plane; planes sit pstride floats apart in shared memory
                    const int pstride = (winc + 2 * GUARD) * 4;
                    std::vector<float> ywin_buf((size_t)pstride * (NB / 4), std::nanf(""));
                    float* ywin_p = ywin_buf.data() + (size_t)GUARD * 4;
                    for (int h = 0; h < NB / 4; ++h)
                        std::memcpy(ywin_p + (size_t)h * pstride,
                                    spk.data() + ((((size_t)g * A + a) * (NB / 4) + h) * Wp2 + start) * 4,
                                    sizeof(float) * winc * 4);
                    for (int ty = 0; ty < TH; ++ty)
                        for (int tx = 0; tx < TW; ++tx) {
                            const float px = (float)(c0 + tx + pady), py = (float)(r0 + ty + padx);
                            float* ac = &acc[((size_t)ty * TW + tx) * NB];
                            if (MODE == CTR_ADJ_EXACT) ctr_adj_exact<NB, INTERP>(t, H, W, px, py, ywin_p, pstride, start, ac);
                            else ctr_adj_tf<NB, INTERP>(t, H, W, px, py, ywin_p, pstride, start, ac);
                        }
                }
                for (int ty = 0; ty < TH; ++ty)
                    for (int tx = 0; tx < TW; ++tx) {
                        const int r = r0 + ty, c = c0 + tx;
                        if (r >= X || c >= Y) continue;
                        for (int n = 0; n < NB; ++n) {
                            const int b = g * NB + n;
                            if (b < B) out[((size_t)b * X + r) * Y + c] = acc[((size_t)ty * TW + tx) * NB + n];
                        }
                    }
            }
}

}  // namespace

extern "C" {

void emu_make_transforms(const double* theta, int A, int H, int W, float* t) { ctr_h_make_transforms(theta, A, H, W, t); }
void emu_invert_transforms(const float* t, int A, float* o) { ctr_h_invert_transforms(t, A, o); }
int emu_num_proj_pix(int X, int Y) { return ctr_h_num_proj_pix(X, Y); }

void emu_forward(const float* img, int B, int X, int Y, int H, int W, int padx, int pady, const float* t, int A,
                 int interp, int R, float* sino)
{
    if (interp == CTR_NEAREST) forward_impl<CTR_NEAREST, 1>(img, B, X, Y, H, W, padx, pady, t, A, R, sino);
    else forward_impl<CTR_BILINEAR, 1>(img, B, X, Y, H, W, padx, pady, t, A, R, sino);
}

// depth-first pack: 4 image groups share a pixel record (ctr_fwd_kernel<..., DEPTH = 4>)
void emu_forward_depth(const float* img, int B, int X, int Y, int H, int W, int padx, int pady, const float* t, int A,
                       int interp, int R, float* sino)
{
    if (interp == CTR_NEAREST) forward_impl<CTR_NEAREST, 4>(img, B, X, Y, H, W, padx, pady, t, A, R, sino);
    else forward_impl<CTR_BILINEAR, 4>(img, B, X, Y, H, W, padx, pady, t, A, R, sino);
}

// 32-image records, 8 images per lane with parity-swizzled loads (ctr_fwd_kernel<8, 2, ., ., 4>)
void emu_forward_rec32(const float* img, int B, int X, int Y, int H, int W, int padx, int pady, const float* t, int A,
                       int interp, int R, float* sino)
{
    if (interp == CTR_NEAREST) forward_impl<CTR_NEAREST, 4, 8>(img, B, X, Y, H, W, padx, pady, t, A, R, sino);
    else forward_impl<CTR_BILINEAR, 4, 8>(img, B, X, Y, H, W, padx, pady, t, A, R, sino);
}
// same records, plain march (the whole-row 32-image shape of narrow detectors)
void emu_forward_rec32_plain(const float* img, int B, int X, int Y, int H, int W, int padx, int pady, const float* t, int A,
                             int interp, int R, float* sino)
{
    if (interp == CTR_NEAREST) forward_impl<CTR_NEAREST, 4, 8, false>(img, B, X, Y, H, W, padx, pady, t, A, R, sino);
    else forward_impl<CTR_BILINEAR, 4, 8, false>(img, B, X, Y, H, W, padx, pady, t, A, R, sino);
}
int emu_forward_window32(const float* img, int B, int X, int Y, int H, int W, int padx, int pady, const float* t, int A,
                         int interp, int JW, int NA, int Rmax, int budget, float* sino)
{
    if (interp == CTR_NEAREST) return forward_window_impl<CTR_NEAREST, 4, 8>(img, B, X, Y, H, W, padx, pady, t, A, JW, NA, Rmax, budget, sino);
    return forward_window_impl<CTR_BILINEAR, 4, 8>(img, B, X, Y, H, W, padx, pady, t, A, JW, NA, Rmax, budget, sino);
}

// 32-image records, two lanes per ray x 16 images with rotated loads (ctr_fwd_kernel<16, 2, ., ., 2>)
void emu_forward_wide(const float* img, int B, int X, int Y, int H, int W, int padx, int pady, const float* t, int A,
                      int interp, int R, int reuse, float* sino)
{
    if (interp == CTR_NEAREST) forward_impl<CTR_NEAREST, 2, 16>(img, B, X, Y, H, W, padx, pady, t, A, R, sino);
    else if (reuse) forward_impl<CTR_BILINEAR, 2, 16, true>(img, B, X, Y, H, W, padx, pady, t, A, R, sino);
    else forward_impl<CTR_BILINEAR, 2, 16, false>(img, B, X, Y, H, W, padx, pady, t, A, R, sino);
}
int emu_forward_window_wide(const float* img, int B, int X, int Y, int H, int W, int padx, int pady, const float* t, int A,
                            int interp, int JW, int NA, int Rmax, int budget, float* sino)
{
    if (interp == CTR_NEAREST) return forward_window_impl<CTR_NEAREST, 2, 16>(img, B, X, Y, H, W, padx, pady, t, A, JW, NA, Rmax, budget, sino);
    return forward_window_impl<CTR_BILINEAR, 2, 16>(img, B, X, Y, H, W, padx, pady, t, A, JW, NA, Rmax, budget, sino);
}

// column-windowed depth-first strips (16-image records)
int emu_forward_window(const float* img, int B, int X, int Y, int H, int W, int padx, int pady, const float* t, int A,
                       int interp, int JW, int NA, int Rmax, int budget, float* sino)
{
    if (interp == CTR_NEAREST) return forward_window_impl<CTR_NEAREST, 4>(img, B, X, Y, H, W, padx, pady, t, A, JW, NA, Rmax, budget, sino);
    return forward_window_impl<CTR_BILINEAR, 4>(img, B, X, Y, H, W, padx, pady, t, A, JW, NA, Rmax, budget, sino);
}

// mode 0: exact (table = forward transforms); mode 1: tf-compat (table = inverted transforms)
void emu_adjoint(const float* y, int B, int X, int Y, int H, int W, int padx, int pady, const float* table, int A,
                 int interp, int mode, int TW, int TH, int win, float* out)
{
    constexpr int NBA = 16;
    if (mode == CTR_ADJ_EXACT) {
        if (interp == CTR_NEAREST) adjoint_impl<CTR_ADJ_EXACT, CTR_NEAREST, NBA>(y, B, X, Y, H, W, padx, pady, table, A, TW, TH, win, out);
        else adjoint_impl<CTR_ADJ_EXACT, CTR_BILINEAR, NBA>(y, B, X, Y, H, W, padx, pady, table, A, TW, TH, win, out);
    } else {
        if (interp == CTR_NEAREST) adjoint_impl<CTR_ADJ_TF, CTR_NEAREST, NBA>(y, B, X, Y, H, W, padx, pady, table, A, TW, TH, win, out);
        else adjoint_impl<CTR_ADJ_TF, CTR_BILINEAR, NBA>(y, B, X, Y, H, W, padx, pady, table, A, TW, TH, win, out);
    }
}

// Angle-sharded exact adjoint with the fused exchange, all ranks emulated in one loop (the GPU guide forbids ranks
// that wait on one another on one device; the multi-GPU run is tests/test_gpu_multi.py).  y is the FULL cotangent
// [B,A,W]; rank s back-projects its angle block (sharding.shard_range) and stores image b into slot s of
// owner(b) (ctr_xg_owner / ctr_xg_index, the same routines the kernel epilogue uses); every owner then sums its
// slots in rank order.  out [B,X,Y] is the concatenation of the owners' shards.
int emu_adjoint_sharded(const float* y, int B, int X, int Y, int H, int W, int padx, int pady, const float* table, int A,
                        int interp, int nranks, float* out)
{
    if (B % nranks != 0 || A < nranks) return -1;
    const int Bs = B / nranks;
    const size_t slot = (size_t)Bs * X * Y;
    std::vector<std::vector<float>> xbuf(nranks, std::vector<float>((size_t)nranks * slot, std::nanf("")));
    const int base = A / nranks, rem = A % nranks;
    for (int s = 0; s < nranks; ++s) {
        const int lo = s * base + std::min(s, rem), n = base + (s < rem ? 1 : 0);
        std::vector<float> ys((size_t)B * n * W), part((size_t)B * X * Y);
        for (int b = 0; b < B; ++b)
            std::memcpy(&ys[(size_t)b * n * W], y + ((size_t)b * A + lo) * W, sizeof(float) * (size_t)n * W);
        if (interp == CTR_NEAREST) adjoint_impl<CTR_ADJ_EXACT, CTR_NEAREST, 16>(ys.data(), B, X, Y, H, W, padx, pady, table + 8 * lo, n, 32, 8, 40, part.data());
        else adjoint_impl<CTR_ADJ_EXACT, CTR_BILINEAR, 16>(ys.data(), B, X, Y, H, W, padx, pady, table + 8 * lo, n, 32, 8, 40, part.data());
        for (int b = 0; b < B; ++b) {
            const int owner = ctr_xg_owner(b, Bs);
            for (int r = 0; r < X; ++r)
                for (int c = 0; c < Y; ++c)
                    xbuf[owner][ctr_xg_index(s, b - owner * Bs, Bs, r, c, X, Y)] = part[((size_t)b * X + r) * Y + c];
        }
    }
    for (int o = 0; o < nranks; ++o)
        for (size_t i = 0; i < slot; ++i) {
            float a = xbuf[o][i];
            for (int s = 1; s < nranks; ++s) a += xbuf[o][(size_t)s * slot + i];
            out[(size_t)o * slot + i] = a;
        }
    return 0;
}

}  // extern "C"
