// ctr_host.h -- host-side geometry shared by the C-ABI (ctr_capi.cu) and the CPU
// emulation harness (tests/emu/).  Header-only, plain C++17, no CUDA types.
//
// Restates, for the product side (the oracle has its own copy on purpose):
//   ctvae/forward_functions.py:29-36  pad_phantom's detector size and pad offsets
//   ctvae/forward_functions.py:113    tfa.image.rotate(imgs, -theta) transform table
//   tensorflow/python/ops/image_ops.py _image_projective_transform_v3_grad (inverse table)
// Must be compiled with -ffp-contract=off: the table is float32, evaluated in TF's order.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

#include "ctr_core.h"

inline int ctr_h_num_proj_pix(int X, int Y)
{
    const double d = std::sqrt((double)((int64_t)X * X + (int64_t)Y * Y)) + 2.0;
    return (int)(std::ceil(d / 2.0) * 2.0);
}

// Frame (H x W) that gets rotated and where the X x Y image sits in it.
inline void ctr_h_frame(int X, int Y, int pad, int& H, int& W, int& padx, int& pady)
{
    if (!pad) { H = X; W = Y; padx = 0; pady = 0; return; }
    const int P = ctr_h_num_proj_pix(X, Y);
    H = P; W = P;
    padx = (P - X) / 2;  // tf.pad "before" amounts; the odd remainder goes after
    pady = (P - Y) / 2;
}

// angles_to_projective_transforms(-theta, H, W) of tensorflow-addons 0.17.1.
inline void ctr_h_make_transforms(const double* theta, int A, int H, int W, float* t)
{
    const float wm1 = (float)W - 1.0f, hm1 = (float)H - 1.0f;
    for (int a = 0; a < A; ++a) {
        const float ang = (float)(-theta[a]);
        const float c = cosf(ang), s = sinf(ang);
        float* r = t + 8 * a;
        r[0] = c; r[1] = -s; r[2] = (wm1 - (c * wm1 - s * hm1)) / 2.0f;
        r[3] = s; r[4] = c;  r[5] = (hm1 - (s * wm1 + c * hm1)) / 2.0f;
        r[6] = 0.f; r[7] = 0.f;
    }
}

// float32 3x3 inverse by partially pivoted Gauss-Jordan, renormalised by m[2][2]
// (tf.linalg.inv + matrices_to_flat_transforms in TF's gradient).
inline void ctr_h_invert_transforms(const float* t, int A, float* tinv)
{
    for (int a = 0; a < A; ++a) {
        const float* r = t + 8 * a;
        float m[3][6] = {{r[0], r[1], r[2], 1.f, 0.f, 0.f},
                         {r[3], r[4], r[5], 0.f, 1.f, 0.f},
                         {r[6], r[7], 1.f, 0.f, 0.f, 1.f}};
        for (int col = 0; col < 3; ++col) {
            int piv = col;
            for (int k = col + 1; k < 3; ++k)
                if (std::fabs(m[k][col]) > std::fabs(m[piv][col])) piv = k;
            if (piv != col)
                for (int q = 0; q < 6; ++q) { const float tmp = m[col][q]; m[col][q] = m[piv][q]; m[piv][q] = tmp; }
            const float d = m[col][col];
            for (int q = 0; q < 6; ++q) m[col][q] = m[col][q] / d;
            for (int k = 0; k < 3; ++k) {
                if (k == col) continue;
                const float f = m[k][col];
                for (int q = 0; q < 6; ++q) m[k][q] = m[k][q] - f * m[col][q];
            }
        }
        const float w = m[2][5];
        float* o = tinv + 8 * a;
        o[0] = m[0][3] / w; o[1] = m[0][4] / w; o[2] = m[0][5] / w;
        o[3] = m[1][3] / w; o[4] = m[1][4] / w; o[5] = m[1][5] / w;
        o[6] = m[2][3] / w; o[7] = m[2][4] / w;
    }
}

// Packed-image geometry of the two classes (see CtrRay): class 0 strips run over
// image rows (v = y), class 1 over image columns (v = x, transposed pack).
// Packed rows are padded (with zeros) to a multiple of 8 pixels so that a row of 16-byte
// chunks starts at bank group 0 whatever the record depth: samples of a quarter-warp that
// sit in different rows then only collide when they share a column (see tools/bank_sim.py).
inline int ctr_h_row_pixels(int n) { return (n + 2 + 7) / 8 * 8; }

inline void ctr_h_class_geom(int X, int Y, int padx, int pady, CtrClassGeom g[2])
{
    g[0].ulo = (float)(pady - 1); g[0].uhi = (float)(pady + Y);
    g[0].vlo = (float)(padx - 1); g[0].vhi = (float)(padx + X);
    g[0].offu = pady - 1; g[0].offv = padx - 1; g[0].Up = ctr_h_row_pixels(Y); g[0].Vp = X + 2;
    g[1].ulo = (float)(padx - 1); g[1].uhi = (float)(padx + X);
    g[1].vlo = (float)(pady - 1); g[1].vhi = (float)(pady + Y);
    g[1].offu = padx - 1; g[1].offv = pady - 1; g[1].Up = ctr_h_row_pixels(X); g[1].Vp = Y + 2;
}

// Class-sorted ray table: class-0 angles first (n0 of them), then class 1.  Inside a class the
// rays are ordered by (side the detector axis runs to, slope of the rays): CTAs take consecutive
// table entries, and neighbours that look at the image from almost the same direction need
// almost the same column window (ctr_h_build_chunks).  `seg` receives the start of every run of
// equal (class, side) plus the end of the table.
inline void ctr_h_build_rays(const float* t, int A, std::vector<CtrRay>& rays, int& n0, std::vector<int>* seg = nullptr)
{
    rays.clear();
    rays.reserve(A);
    for (int a = 0; a < A; ++a) {
        const float* r = t + 8 * a;
        // x = t0*j + t1*i + t2 ; y = t3*j + t4*i + t5.  Strip axis = the one that
        // advances fastest per step i: |t4| >= |t1| -> rows (class 0).
        const int cls = (std::fabs(r[4]) >= std::fabs(r[1])) ? 0 : 1;
        CtrRay q;
        if (cls == 0) { q.u0 = r[0]; q.u1 = r[1]; q.u2 = r[2]; q.v0 = r[3]; q.v1 = r[4]; q.v2 = r[5]; }
        else          { q.u0 = r[3]; q.u1 = r[4]; q.u2 = r[5]; q.v0 = r[0]; q.v1 = r[1]; q.v2 = r[2]; }
        q.angle = a; q.cls = cls;
        rays.push_back(q);
    }
    auto side = [](const CtrRay& q) { return ctr_win_coef(q).b < 0.f ? 1 : 0; };
    std::stable_sort(rays.begin(), rays.end(), [&](const CtrRay& x, const CtrRay& y) {
        if (x.cls != y.cls) return x.cls < y.cls;
        const int sx = side(x), sy = side(y);
        if (sx != sy) return sx < sy;
        return ctr_win_coef(x).g < ctr_win_coef(y).g;
    });
    n0 = 0;
    for (const CtrRay& q : rays) n0 += (q.cls == 0);
    if (seg) {
        seg->clear();
        for (int k = 0; k < (int)rays.size(); ++k)
            if (k == 0 || rays[k].cls != rays[k - 1].cls || side(rays[k]) != side(rays[k - 1])) seg->push_back(k);
        seg->push_back((int)rays.size());
    }
}

// Column window (pixels) that the CTA serving rays [first, first+cnt) x detector chunks of JW bins
// needs when its strips hold R key rows: the maximum over detector chunks and strips of what the
// device will ask for (same ctr_win_* routines, same float32 arithmetic).
inline int ctr_h_window_pixels(const CtrRay* rays, int cnt, const CtrClassGeom& g, int W, int JW, int jchunks, int R)
{
    int need = 0;
    const int K = (g.Vp + R - 1) / R;
    std::vector<CtrWinCoef> co(cnt);
    for (int q = 0; q < cnt; ++q) co[q] = ctr_win_coef(rays[q]);
    for (int z = 0; z < jchunks; ++z) {
        const float jlo = (float)(z * JW), jhi = (float)std::min(W - 1, z * JW + JW - 1);
        for (int k = 0; k < K; ++k) {
            const float vlo = (float)(k * R + g.offv) - 0.5f, vhi = (float)(k * R + g.offv + R);
            float umin = 3.0e38f, umax = -3.0e38f;
            for (int q = 0; q < cnt; ++q) ctr_win_range(co[q], jlo, jhi, vlo, vhi, g.ulo, g.uhi, umin, umax);
            need = std::max(need, ctr_win_need(umin, umax));
        }
    }
    return need;
}

// Cut the ray table into CTA-sized chunks of <= NA rays that never straddle a segment.
//   windowed == false: whole packed rows, R key rows per strip for everybody.
//   windowed == true : per chunk, the tallest strip (R <= Rmax) whose `stages` buffers of
//                      (R+1) x window pixels fit `strip_budget` bytes; false if some chunk cannot
//                      get R >= Rmin (the caller then keeps the whole-row configuration).
// max_strip_bytes receives the largest single strip buffer of any chunk.
inline bool ctr_h_build_chunks(const std::vector<CtrRay>& rays, const std::vector<int>& seg, const CtrClassGeom geom[2],
                               int NA, int W, int JW, int jchunks, int rec_bytes, int stages, size_t strip_budget,
                               bool windowed, int R, int Rmin, int Rmax, std::vector<CtrChunk>& out, size_t& max_strip_bytes)
{
    out.clear();
    max_strip_bytes = 0;
    for (size_t s = 0; s + 1 < seg.size(); ++s)
        for (int first = seg[s]; first < seg[s + 1]; first += NA) {
            CtrChunk c{};
            c.first = first;
            c.cnt = std::min(NA, seg[s + 1] - first);
            c.cls = rays[first].cls;
            const CtrClassGeom& g = geom[c.cls];
            if (!windowed) {
                c.wc = 0;
                c.R = R;
                max_strip_bytes = std::max(max_strip_bytes, (size_t)(R + 1) * g.Up * rec_bytes);
            } else {
                bool fit = false;
                for (int r = std::min(Rmax, g.Vp); r >= Rmin && !fit; --r) {
                    int wc = (ctr_h_window_pixels(&rays[first], c.cnt, g, W, JW, jchunks, r) + 1) / 2 * 2;
                    if (wc >= g.Up) wc = g.Up;            // the window is the whole row
                    const size_t bytes = (size_t)(r + 1) * wc * rec_bytes;
                    if ((size_t)stages * bytes <= strip_budget) {
                        c.wc = (wc == g.Up) ? 0 : wc;
                        c.R = r;
                        max_strip_bytes = std::max(max_strip_bytes, bytes);
                        fit = true;
                    }
                }
                if (!fit) return false;
            }
            out.push_back(c);
        }
    return true;
}
