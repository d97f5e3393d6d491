// ctr_core.h -- per-thread arithmetic of the Radon kernels, written once as
// __host__ __device__ so the exact same code runs inside the sm_100a kernels
// (ctr_kernels.cu) and inside the CPU emulation harness under tests/emu/ that
// checks the geometry logic without a GPU.  Nothing here touches memory spaces,
// barriers or TMA; that glue lives in ctr_kernels.cu.
//
// Semantics restated (reference file:line in /root/reference):
//   ctvae/forward_functions.py:113   tfa.image.rotate(imgs, -theta)  -> ImageProjectiveTransformV3
//   ctvae/forward_functions.py:114   tf.reduce_sum(imgs_rot, 1)
// Sample coordinates follow TF's ProjectiveGenerator bit for bit: float32,
// ((t0*x_out) + (t1*y_out)) + t2, no FMA contraction (CTR_MUL/CTR_ADD).
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define CTR_HD __host__ __device__ __forceinline__
#else
#define CTR_HD inline
#endif

#if defined(__CUDA_ARCH__)
#define CTR_MUL(a, b) __fmul_rn((a), (b))
#define CTR_ADD(a, b) __fadd_rn((a), (b))
#define CTR_SUB(a, b) __fsub_rn((a), (b))
#else  // host: translation units are compiled with -ffp-contract=off
#define CTR_MUL(a, b) ((a) * (b))
#define CTR_ADD(a, b) ((a) + (b))
#define CTR_SUB(a, b) ((a) - (b))
#endif

enum { CTR_NEAREST = 0, CTR_BILINEAR = 1 };
enum { CTR_ADJ_EXACT = 0, CTR_ADJ_TF = 1, CTR_ADJ_FBP = 2 };

// Per-angle ray coefficients, normalised to the packed-image orientation the angle
// is served from ("class"):
//   u(j,i) = (u0*j + u1*i) + u2   coordinate along a packed row (contiguous axis)
//   v(j,i) = (v0*j + v1*i) + v2   coordinate across packed rows (the strip axis)
// class 0 (|cos| >= |sin|): u = x_in, v = y_in, served from the row-major pack;
// class 1 (|sin| >  |cos|): u = y_in, v = x_in, served from the transposed pack.
// Either way |v1| >= 0.707, so every ray crosses the strips at a steady rate.
struct CtrRay {
    float u0, u1, u2;
    float v0, v1, v2;
    int angle;  // row of the sinogram this ray set writes
    int cls;
};

// Packed-image geometry of one class (frame coordinates -> packed indices).
// A sample is inside the image's footprint iff ulo < u < uhi and vlo < v < vhi.
struct CtrClassGeom {
    float ulo, uhi, vlo, vhi;
    int offu, offv;  // packed index = (int)floor(coord) - off   (off = pad - 1: one halo pixel)
    int Up, Vp;      // packed row length (pixels) and row count, halos included
};

// One forward CTA's share of the class-sorted ray table: rays [first, first+cnt) of class
// `cls`.  wc > 0 selects column-windowed strips: instead of whole packed rows the strip
// buffers hold, for every packed row, the `wc` pixels starting at a per-strip column that
// follows the CTA's rays across the image (see ctr_win_*).  R = key rows per strip.
struct CtrChunk {
    int first, cnt;
    int cls;
    int wc;
    int R;
    int pad_[3];
};

// Where the rays of one angle are, as a function of the detector bin j and the strip
// coordinate v:  u(j, v) = a + b*j + g*v  (eliminate the step index i from CtrRay's two
// lines; |g| <= 1, |b| in [1, 1.42]).  Real arithmetic, used only to PLACE the window
// with two pixels of slack -- sample coordinates themselves stay the exact float32 ones.
struct CtrWinCoef {
    float a, b, g;
};

CTR_HD CtrWinCoef ctr_win_coef(const CtrRay& r)
{
    CtrWinCoef c;
    c.g = r.u1 / r.v1;                                  // IEEE division on both sides (no fast-math)
    c.b = CTR_SUB(r.u0, CTR_MUL(c.g, r.v0));            // explicit roundings: host and device agree bit for bit
    c.a = CTR_SUB(r.u2, CTR_MUL(c.g, r.v2));
    return c;
}

// u-extent of bins [jlo, jhi] x key rows of strip [rbase, rbase+R) (v in [rbase-0.5, rbase+R]
// covers floor- and round-keyed samples), clipped to the image footprint (ulo, uhi).
CTR_HD void ctr_win_range(const CtrWinCoef& c, float jlo, float jhi, float vlo, float vhi, float ulo, float uhi,
                          float& umin, float& umax)
{
    const float bj0 = CTR_MUL(c.b, jlo), bj1 = CTR_MUL(c.b, jhi), gv0 = CTR_MUL(c.g, vlo), gv1 = CTR_MUL(c.g, vhi);
    float lo = CTR_ADD(CTR_ADD(c.a, fminf(bj0, bj1)), fminf(gv0, gv1));
    float hi = CTR_ADD(CTR_ADD(c.a, fmaxf(bj0, bj1)), fmaxf(gv0, gv1));
    lo = fminf(fmaxf(lo, ulo), uhi);
    hi = fminf(fmaxf(hi, ulo), uhi);
    umin = fminf(umin, lo);
    umax = fmaxf(umax, hi);
}

// First packed column of the window: leftmost tap floor(umin) with one pixel of slack,
// kept inside the packed row.  Columns needed: floor(umin)-1 .. floor(umax)+2.
CTR_HD int ctr_win_start(float umin, int offu, int Up, int wc)
{
    int c = (int)floorf(umin) - 1 - offu;
    if (c > Up - wc) c = Up - wc;
    if (c < 0) c = 0;
    return c;
}
CTR_HD int ctr_win_need(float umin, float umax) { return (int)floorf(umax) - (int)floorf(umin) + 4; }

// NB interleaved floats -> registers (128-bit shared-memory loads on the device).
template <int NB>
CTR_HD void ctr_ldv(const float* __restrict__ p, float* __restrict__ out)
{
#if defined(__CUDA_ARCH__)
    static_assert(NB % 4 == 0, "NB must be a multiple of 4 (16-byte vectors)");
#pragma unroll
    for (int q = 0; q < NB / 4; ++q) {
        const float4 t = reinterpret_cast<const float4*>(p)[q];
        out[4 * q + 0] = t.x; out[4 * q + 1] = t.y; out[4 * q + 2] = t.z; out[4 * q + 3] = t.w;
    }
#else
    for (int q = 0; q < NB; ++q) out[q] = p[q];
#endif
}

// Eight images of one lane out of a 32-image (128-byte) pixel record, as two 16-byte loads in an order
// that depends on the ray's parity (swz = 0 or 4 floats): in the same instruction the four lanes of an
// even ray read chunks 0,2,4,6 of their record and those of an odd ray chunks 1,3,5,7, so the eight
// lanes of a quarter-warp always hit eight different bank groups -- conflict-free wherever the two
// rays' pixels are.  out[0..3] holds images (swz..swz+3) of the lane's block, out[4..7] the other half.
CTR_HD void ctr_ldv8_swz(const float* __restrict__ p, int swz, float* __restrict__ out)
{
    ctr_ldv<4>(p + swz, out);
    ctr_ldv<4>(p + (swz ^ 4), out + 4);
}

// Sixteen images of one lane out of a 32-image record (TWO lanes per ray): four 16-byte loads in an order rotated
// by the ray's index mod 4 (rot = 0, 4, 8 or 12 floats).  A quarter-warp is then 4 rays x 2 lanes, and in every
// instruction the rays' lanes read chunks {t+r mod 4} and {4 + (t+r mod 4)}, r = 0..3: eight different bank groups,
// conflict-free wherever the four rays' pixels are.  out[4t..4t+3] holds images ((4t + rot) & 15) + 0..3 of the block.
CTR_HD void ctr_ldv16_rot(const float* __restrict__ p, int rot, float* __restrict__ out)
{
#pragma unroll
    for (int t = 0; t < 4; ++t) ctr_ldv<4>(p + ((4 * t + rot) & 15), out + 4 * t);
}

// which image of the lane's block register n holds under the load order above (swz = 0 for plain loads):
// n ^ swz for the parity swizzle of 8-image lanes, (n + rot) & 15 for the rotation of 16-image lanes -- the same
// expression, since swz and rot are multiples of 4 below NB.
template <int NB>
CTR_HD int ctr_img_of_reg(int n, int swz) { return (n + swz) & (NB - 1); }

// Sinogram windows are staged as NB/4 planes of [bin][4 images] (pstride floats apart):
// with 16-byte bins, lanes that read consecutive bins hit consecutive bank groups.
template <int NB>
CTR_HD void ctr_ld_bins(const float* __restrict__ ywin, int pstride, int idx, float* __restrict__ out)
{
#pragma unroll
    for (int h = 0; h < NB / 4; ++h) ctr_ldv<4>(ywin + (size_t)h * pstride + (size_t)idx * 4, out + 4 * h);
}

CTR_HD float ctr_coord(float p0j, float c1, float fi, float c2)
{
    return CTR_ADD(CTR_ADD(p0j, CTR_MUL(c1, fi)), c2);
}

// std::round (half away from zero), exact for every float (no v+0.5 double rounding).
CTR_HD float ctr_round(float v)
{
    float r = rintf(v);          // nearest-even
    float d = CTR_SUB(v, r);     // exact
    if (d == 0.5f && v > 0.f) r += 1.f;
    if (d == -0.5f && v < 0.f) r -= 1.f;
    return r;
}

// floor / round-half-away of a coordinate as (float, int) in one go.  On the device the
// 1.5*2^23 trick does it on the FP32 pipe (no FRND/F2I on the quarter-rate conversion
// unit): v + 12582912 rounded down (resp. to nearest-even) leaves the integer in the
// mantissa.  Exact for |v| < 2^22, far beyond any frame coordinate.
#define CTR_MAGIC 12582912.0f
#define CTR_MAGIC_BITS 0x4B400000
CTR_HD void ctr_floor_fi(float v, float& f, int& i)
{
#if defined(__CUDA_ARCH__)
    const float t = __fadd_rd(v, CTR_MAGIC);
    f = __fsub_rn(t, CTR_MAGIC);
    i = __float_as_int(t) - CTR_MAGIC_BITS;
#else
    f = floorf(v);
    i = (int)f;
#endif
}
CTR_HD void ctr_round_fi(float v, float& f, int& i)
{
#if defined(__CUDA_ARCH__)
    const float t = __fadd_rn(v, CTR_MAGIC);   // nearest-even
    float r = __fsub_rn(t, CTR_MAGIC);
    int k = __float_as_int(t) - CTR_MAGIC_BITS;
    const float d = __fsub_rn(v, r);            // exact
    if (d == 0.5f && v > 0.f) { r += 1.f; k += 1; }     // ties go away from zero (std::round)
    if (d == -0.5f && v < 0.f) { r -= 1.f; k -= 1; }
    f = r;
    i = k;
#else
    f = ctr_round(v);
    i = (int)f;
#endif
}

// Smallest i in [0,H] with sgn*coord(i) > bound (strict) or >= bound.  float32 rounding is
// monotone, so coord(i) is monotone in i and the predicate flips once: a closed-form estimate
// of the crossing is corrected with the exact float32 predicate (usually 1-2 evaluations; a ray
// almost parallel to the bound may walk further, still at most H steps).
CTR_HD int ctr_search(float p0j, float c1, float c2, float sgn, float bound, bool strict, int H)
{
    const float base = sgn * CTR_ADD(p0j, c2);          // g(i) ~ base + |c1| * i
    const float slope = fabsf(c1);
    if (slope == 0.f)                                   // coordinate does not depend on i: all or nothing
        return (strict ? (base > bound) : (base >= bound)) ? 0 : H;
    int i = 0;
    {
        float x = (bound - base) / slope;               // g(i) > bound  <=>  i > x   (real arithmetic)
        x = fminf(fmaxf(x, -1.f), (float)H);            // also maps +-inf into range; NaN -> -1
        i = (int)floorf(x) + 1;
        if (i > H) i = H;
    }
#define CTR_PRED(ii) (strict ? (sgn * ctr_coord(p0j, c1, (float)(ii), c2) > bound) : (sgn * ctr_coord(p0j, c1, (float)(ii), c2) >= bound))
    while (i > 0 && CTR_PRED(i - 1)) --i;
    while (i < H && !CTR_PRED(i)) ++i;
#undef CTR_PRED
    return i;
}

// [ib, ie): the steps i of ray j whose sample lies strictly inside the footprint
// (everything outside contributes exactly zero to the row sum).
CTR_HD void ctr_ray_interval(const CtrRay& r, const CtrClassGeom& g, int j, int H, int& ib, int& ie)
{
    const float pu = CTR_MUL(r.u0, (float)j), pv = CTR_MUL(r.v0, (float)j);
    const float su = (r.u1 >= 0.f) ? 1.f : -1.f;
    const float sv = (r.v1 >= 0.f) ? 1.f : -1.f;
    const int ib_u = ctr_search(pu, r.u1, r.u2, su, su > 0.f ? g.ulo : -g.uhi, true, H);
    const int ie_u = ctr_search(pu, r.u1, r.u2, su, su > 0.f ? g.uhi : -g.ulo, false, H);
    const int ib_v = ctr_search(pv, r.v1, r.v2, sv, sv > 0.f ? g.vlo : -g.vhi, true, H);
    const int ie_v = ctr_search(pv, r.v1, r.v2, sv, sv > 0.f ? g.vhi : -g.vlo, false, H);
    ib = ib_u > ib_v ? ib_u : ib_v;
    ie = ie_u < ie_v ? ie_u : ie_v;
    if (ie < ib) ie = ib;
}

// A ray in flight: next step (kept as a float: integers < 2^24 are exact and the
// coordinate needs it as a float anyway), steps left, direction of travel (towards larger v).
struct CtrRayState {
    float pu, pv;  // u0*j, v0*j
    float fi, dfi; // next step i and +-1
    int n;
};

CTR_HD void ctr_ray_begin(const CtrRay& r, const CtrClassGeom& g, int j, int H, CtrRayState& s)
{
    int ib, ie;
    ctr_ray_interval(r, g, j, H, ib, ie);
    s.pu = CTR_MUL(r.u0, (float)j);
    s.pv = CTR_MUL(r.v0, (float)j);
    s.n = ie - ib;
    s.dfi = (r.v1 >= 0.f) ? 1.f : -1.f;
    s.fi = (float)((r.v1 >= 0.f) ? ib : ie - 1);
}

// March one ray through one strip.  `strip` holds packed rows [row0p, row0p+rows)
// as [row][Up][NB] floats; this call consumes every remaining sample whose key row
// (floor(v) for bilinear, round(v) for nearest) is < vend.  Keys only grow along
// the march, so successive strips partition the ray exactly.
//   vend  = (float)(row0p + R + offv)   first key row that belongs to the next strip
//   rbase = row0p + offv                key row of strip row 0
// One sample of one ray: coordinates -> (key row, packed offsets, weights).  Split out so the
// march loop can put two samples in flight at once.
template <int INTERP>
struct CtrSample {
    float kvf;           // key row as float (floor / round of v)
    int off;             // (row, col) offset of the first tap in pixels, relative to the strip
    float w00, w01, w10, w11;
};

template <int INTERP>
CTR_HD void ctr_sample(const CtrRay& r, float pu, float pv, float fi, int Up, int rbase, int offu, CtrSample<INTERP>& o)
{
    const float u = ctr_coord(pu, r.u1, fi, r.u2);
    const float v = ctr_coord(pv, r.v1, fi, r.v2);
    float kuf;
    int kvi, kui;
    if (INTERP == CTR_NEAREST) {
        ctr_round_fi(v, o.kvf, kvi);
        ctr_round_fi(u, kuf, kui);
    } else {
        ctr_floor_fi(v, o.kvf, kvi);
        ctr_floor_fi(u, kuf, kui);
        // TF: (x - floor(x)) and (floor(x)+1 - x); the latter equals 1 - (x - floor(x)) exactly
        const float fu = CTR_SUB(u, kuf), gu = CTR_SUB(1.f, fu);
        const float fv = CTR_SUB(v, o.kvf), gv = CTR_SUB(1.f, fv);
        o.w00 = gv * gu; o.w01 = gv * fu; o.w10 = fv * gu; o.w11 = fv * fu;
    }
    o.off = (kvi - rbase) * Up + (kui - offu);
}

template <int NB>
CTR_HD void ctr_ld_rec(const float* __restrict__ p, int swz, float* __restrict__ out)
{
    if (NB == 16) ctr_ldv16_rot(p, swz, out);
    else if (NB == 8) ctr_ldv8_swz(p, swz, out);
    else ctr_ldv<NB>(p, out);
}

template <int NB, int INTERP, int REC>
CTR_HD void ctr_gather(const float* __restrict__ strip, int Up, const CtrSample<INTERP>& o, float* __restrict__ acc, int swz = 0)
{
    const float* p0 = strip + o.off * REC;
    if (INTERP == CTR_NEAREST) {
        float a[NB];
        ctr_ld_rec<NB>(p0, swz, a);
#pragma unroll
        for (int q = 0; q < NB; ++q) acc[q] += a[q];
    } else {
        const float* p1 = p0 + Up * REC;
        float a00[NB], a01[NB], a10[NB], a11[NB];
        ctr_ld_rec<NB>(p0, swz, a00); ctr_ld_rec<NB>(p0 + REC, swz, a01);
        ctr_ld_rec<NB>(p1, swz, a10); ctr_ld_rec<NB>(p1 + REC, swz, a11);
#pragma unroll
        for (int q = 0; q < NB; ++q)
            acc[q] = fmaf(o.w11, a11[q], fmaf(o.w10, a10[q], fmaf(o.w01, a01[q], fmaf(o.w00, a00[q], acc[q]))));
    }
}

// REC = floats per packed pixel record (NB, or NB*DEPTH when DEPTH image groups share a
// record and `strip` already points at this thread's group inside the record).
// (A two-samples-per-trip variant was measured in r1 and was not faster: the loop is bound by
// the shared-memory and issue pipes, not by latency.)
template <int NB, int INTERP, int REC = NB>
CTR_HD void ctr_march(const float* __restrict__ strip, int Up, float vend, int rbase, int offu,
                      const CtrRay& r, CtrRayState& s, float* __restrict__ acc, int swz = 0)
{
    while (s.n > 0) {
        CtrSample<INTERP> a;
        ctr_sample<INTERP>(r, s.pu, s.pv, s.fi, Up, rbase, offu, a);
        if (a.kvf >= vend) break;
        ctr_gather<NB, INTERP, REC>(strip, Up, a, acc, swz);
        s.fi += s.dfi;
        --s.n;
    }
}

// Bilinear march with VERTICAL REUSE.  Rays cross the strips at >= 0.707 rows per step, so about half of
// the samples sit exactly one packed row below the previous one in the same column pair: their top records
// are the previous sample's bottom records.  Two register sets alternate as top / bottom row (loop unrolled
// by two, no register moves); the top loads are skipped when the strip offset says the set already holds
// that row ("off == previous off + Up" <=> row + 1, same column).  A shared-memory wavefront is saved whenever
// every ray of a quarter-warp skips (r1 model: 34-40 % of the top loads).  Same products and summation order
// as ctr_march, hence bit-identical sums.  Reuse never crosses a strip (the state lives in this call).
template <int NB, int REC>
CTR_HD void ctr_march_reuse(const float* __restrict__ strip, int Up, float vend, int rbase, int offu,
                            const CtrRay& r, CtrRayState& s, float* __restrict__ acc, int swz = 0)
{
    float t0[NB], t1[NB], b0[NB], b1[NB];   // set A = (t0, t1), set B = (b0, b1): records at columns c, c+1
    int held = -(1 << 30);                   // strip offset of the row pair the "bottom" set of the last sample holds
#define CTR_REUSE_HALF(TOP0, TOP1, BOT0, BOT1)                                                                      \
    {                                                                                                               \
        CtrSample<CTR_BILINEAR> a;                                                                                  \
        ctr_sample<CTR_BILINEAR>(r, s.pu, s.pv, s.fi, Up, rbase, offu, a);                                          \
        if (a.kvf >= vend) break;                                                                                   \
        const float* p0 = strip + a.off * REC;                                                                      \
        const float* p1 = p0 + Up * REC;                                                                            \
        if (a.off != held) {                                                                                        \
            ctr_ld_rec<NB>(p0, swz, TOP0);                                                                          \
            ctr_ld_rec<NB>(p0 + REC, swz, TOP1);                                                                    \
        }                                                                                                           \
        ctr_ld_rec<NB>(p1, swz, BOT0);                                                                              \
        ctr_ld_rec<NB>(p1 + REC, swz, BOT1);                                                                        \
        held = a.off + Up;                                                                                          \
        for (int q = 0; q < NB; ++q)                                                                                \
            acc[q] = fmaf(a.w11, BOT1[q], fmaf(a.w10, BOT0[q], fmaf(a.w01, TOP1[q], fmaf(a.w00, TOP0[q], acc[q])))); \
        s.fi += s.dfi;                                                                                              \
        --s.n;                                                                                                      \
    }
    while (s.n > 0) {
        CTR_REUSE_HALF(t0, t1, b0, b1)
        if (s.n <= 0) break;
        CTR_REUSE_HALF(b0, b1, t0, t1)
    }
#undef CTR_REUSE_HALF
}

// ---------------------------------------------------------------------------------------------
// Pixel-driven back-projections.  t[] is one row of the [A,8] transform table
// (forward table for EXACT, inverted table for TF).  (px,py) are the pixel's frame
// coordinates as floats.  ywin holds sinogram bins [jbase_p, jbase_p+win) of the
// halo-padded row (packed bin index = j + 1; bins -1 and W are zero) as NB/4 planes
// of [bin][4 images], `pstride` floats apart (see ctr_ld_bins).

// First-order inverse of the forward rotation: which (j,i) lattice point samples closest to (px,py).
CTR_HD void ctr_adj_centre(const float* t, float px, float py, float& uj, float& vi)
{
    const float dx = px - t[2], dy = py - t[5];
    uj = t[0] * dx + t[3] * dy;
    vi = t[1] * dx + t[4] * dy;
}

// Exact transpose of the forward operator (north_star: <Ax,y> == <x,A^T y>).
// Every (j,i) whose sample can touch pixel (px,py) lies in the 3x3 lattice window
// around round(M^-1 p) (rotated 2x2 footprint, half-diagonal 1.414 < 1.5); the
// forward's float32 coordinates and tap weights are re-evaluated for each of them,
// so a superset window is harmless (weight exactly 0).
template <int NB, int INTERP>
CTR_HD void ctr_adj_exact(const float* t, int H, int W, float px, float py,
                          const float* __restrict__ ywin, int pstride, int jbase_p, float* __restrict__ acc)
{
    float uj, vi;
    ctr_adj_centre(t, px, py, uj, vi);
    // clamp in float first: uj/vi can be far outside for pad=False corners
    uj = fminf(fmaxf(uj, 0.f), (float)(W - 1));
    vi = fminf(fmaxf(vi, 0.f), (float)(H - 1));
    const int j0 = (int)rintf(uj), i0 = (int)rintf(vi);
    // nearest: std::round(x) == px  <=>  x in [px-0.5, px+0.5)  (px >= 1), (-0.5, 0.5) for px == 0;
    // both bounds are exact floats, so the test needs no rounding at all
    const float lox = (px >= 1.f) ? px - 0.5f : -0.49999997f, hix = px + 0.5f;
    const float loy = (py >= 1.f) ? py - 0.5f : -0.49999997f, hiy = py + 0.5f;
#pragma unroll
    for (int dj = -1; dj <= 1; ++dj) {
        const int j = j0 + dj;
        const float fj = (float)j;
        const float p0x = CTR_MUL(t[0], fj), p0y = CTR_MUL(t[3], fj);
        float wsum = 0.f;
#pragma unroll
        for (int di = -1; di <= 1; ++di) {
            const int i = i0 + di;
            const float fi = (float)i;
            const float x = ctr_coord(p0x, t[1], fi, t[2]);
            const float y = ctr_coord(p0y, t[4], fi, t[5]);
            float w;
            if (INTERP == CTR_NEAREST) {
                w = (x >= lox && x < hix && y >= loy && y < hiy) ? 1.f : 0.f;
            } else {
                const float wx = fmaxf(0.f, CTR_SUB(1.f, fabsf(CTR_SUB(x, px))));
                const float wy = fmaxf(0.f, CTR_SUB(1.f, fabsf(CTR_SUB(y, py))));
                w = wx * wy;
            }
            if (i >= 0 && i < H) wsum += w;
        }
        // Nearest: a pixel is hit by one sample on average, so most side bins carry no weight; a bin without weight is
        // skipped (it would add exact zeros) and a quarter-warp whose eight pixels all skip saves the shared-memory
        // wavefront (r2, C4: 4.41 -> 4.25 ms).  Bilinear pixels rarely all skip: the branch costs more than it saves
        // (4.38 -> 4.54 ms), so they always read their three bins.
        if (INTERP == CTR_NEAREST && wsum == 0.f) continue;
        // plane by plane (4 images per 16-byte bin): keeps the live load registers at 4 instead of NB
        const float* yb = ywin + (size_t)(j + 1 - jbase_p) * 4;
#pragma unroll
        for (int h = 0; h < NB / 4; ++h) {
            float y4[4];
            ctr_ldv<4>(yb + (size_t)h * pstride, y4);
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[4 * h + q] += wsum * y4[q];
        }
    }
}

// TensorFlow's registered gradient (ImageProjectiveTransformV3 applied to the
// row-broadcast cotangent with the inverted transforms; main_ct_vae.py:471-481).
template <int NB, int INTERP>
CTR_HD void ctr_adj_tf(const float* ti, int H, int W, float px, float py,
                       const float* __restrict__ ywin, int pstride, int jbase_p, float* __restrict__ acc)
{
    const float x = CTR_ADD(CTR_ADD(CTR_MUL(ti[0], px), CTR_MUL(ti[1], py)), ti[2]);
    const float y = CTR_ADD(CTR_ADD(CTR_MUL(ti[3], px), CTR_MUL(ti[4], py)), ti[5]);
    if (INTERP == CTR_NEAREST) {
        float jj, ii;
        int jji, iii;
        ctr_round_fi(x, jj, jji);
        ctr_round_fi(y, ii, iii);
        if (iii >= 0 && iii < H && jji >= 0 && jji < W) {
            float yv[NB];
            ctr_ld_bins<NB>(ywin, pstride, jji + 1 - jbase_p, yv);
#pragma unroll
            for (int q = 0; q < NB; ++q) acc[q] += yv[q];
        }
    } else {
        const float xf = floorf(x), yf = floorf(y);
        if (xf >= -1.f && xf <= (float)(W - 1)) {  // else both column taps are fill
            const float wxf = CTR_SUB(CTR_ADD(xf, 1.f), x), wxc = CTR_SUB(x, xf);
            const float wyf = (yf >= 0.f && yf < (float)H) ? CTR_SUB(CTR_ADD(yf, 1.f), y) : 0.f;
            const float wyc = (yf >= -1.f && yf < (float)(H - 1)) ? CTR_SUB(y, yf) : 0.f;
            float y0[NB], y1[NB];
            ctr_ld_bins<NB>(ywin, pstride, (int)xf + 1 - jbase_p, y0);
            ctr_ld_bins<NB>(ywin, pstride, (int)xf + 2 - jbase_p, y1);
            // TF: (yc-y)*v_f + (y-yf)*v_c with v_f = v_c = the row interpolation (the cotangent is
            // constant along rows), i.e. (wyf + wyc) * (wxf*y0 + wxc*y1); fused here (<= 1 ulp apart)
            const float wy = wyf + wyc;
            const float a0 = wy * wxf, a1 = wy * wxc;
#pragma unroll
            for (int q = 0; q < NB; ++q) acc[q] = fmaf(a1, y1[q], fmaf(a0, y0[q], acc[q]));
        }
    }
}

// The two taps of a linear interpolation along the detector: bins base, base + 1 of the staged window.
struct CtrTap2 {
    int base;
    float w0, w1;
};

// Taps of one pixel for iradon's back-projection (ctvae/fbp_tensorflow.py:52-71): geometry in float64 like the
// reference, interpolation weights in float32.
//   cs[0]=cos(theta), cs[1]=sin(theta); xpr = row - x_size/2, ypr = col - y_size/2;
//   idx = t + P/2 is the fractional bin ((t - x_ref_min)/(x_ref_max - x_ref_min)*(P-1) for the reference's grid).
// tfp's constant extension clamps idx to [0, P-1] and takes bins (below, below+1) with below <= P-2: idx < 0 gives
// bins (0, 1) with alpha 0, idx >= P-1 bins (P-2, P-1) with alpha 1 -- applied here to the integer floor and the
// float32 weight, which is the same thing.  floor() on the device is the 1.5*2^52 trick on the FP64 add pipe (the
// integer lands in the low word): no FRND / F2I / I2F on the quarter-rate conversion unit (r2: the gather's loop had
// five of those per pixel-angle).
CTR_HD CtrTap2 ctr_tap_fbp(const double* cs, int P, double half_p, double xpr, double ypr, int jbase_p)   // half_p = 0.5 * P
{
    const double tt = ypr * cs[0] - xpr * cs[1];
    const double idx = tt + half_p;
#if defined(__CUDA_ARCH__)
    const double tm = __dadd_rd(idx, 6755399441055744.0);     // exact floor for |idx| < 2^31
    int bi = __double2loint(tm);
    const double below = tm - 6755399441055744.0;
#else
    const double below = floor(idx);
    int bi = (int)below;
#endif
    float alpha = (float)(idx - below);
    if (bi < 0) { bi = 0; alpha = 0.f; }
    if (bi > P - 2) { bi = P - 2; alpha = 1.f; }
    if (P < 2) { bi = 0; alpha = 0.f; }                        // a one-bin detector: both taps are bin 0
    CtrTap2 t;
    t.base = bi + 1 - jbase_p;
    t.w0 = 1.f - alpha;
    t.w1 = alpha;
    return t;
}

// Back-projection stage of iradon for one pixel: the two taps of ctr_tap_fbp.  The halo bins of the packed row
// (packed index = j + 1) are never read here.
template <int NB>
CTR_HD void ctr_adj_fbp(const double* cs, int P, double half_p, double xpr, double ypr,
                        const float* __restrict__ ywin, int pstride, int jbase_p, float* __restrict__ acc)
{
    const CtrTap2 t = ctr_tap_fbp(cs, P, half_p, xpr, ypr, jbase_p);
    float yb[NB], ya[NB];
    ctr_ld_bins<NB>(ywin, pstride, t.base, yb);
    ctr_ld_bins<NB>(ywin, pstride, t.base + 1, ya);
#pragma unroll
    for (int q = 0; q < NB; ++q) acc[q] = fmaf(t.w1, ya[q], fmaf(t.w0, yb[q], acc[q]));
}

// Measurement log-likelihood of one ray-sum (ctvae/helper_functions.py:359-368):
//   pm = proj * mask ;  scale = sqrt_reg + sqrt(pm / pnm + sqrt_reg)
//   lp = Normal(loc = pm, scale).log_prob(y)
// and its derivative with respect to proj (what the tape would propagate for d(sum lp)).
CTR_HD void ctr_loglik_term(float proj, float mask, float y, float pnm, float sqrt_reg, float& lp, float& dlp_dproj)
{
    const float pm = proj * mask;
    const float sr = sqrtf(pm / pnm + sqrt_reg);
    const float sc = sqrt_reg + sr;
    const float inv = 1.f / sc;
    const float z = (y - pm) * inv;
    lp = -0.5f * z * z - logf(sc) - 0.91893853320467274f;   // 0.5*log(2*pi)
    const float dsc = 0.5f / (pnm * sr);                     // d scale / d pm
    dlp_dproj = (z * inv + (z * z - 1.f) * inv * dsc) * mask;
}

// ---------------------------------------------------------------------------------------------
// Angle-sharded exchange.  Rank r owns the images [r*Bs, (r+1)*Bs) of the summed back-projection.
// Its exchange buffer holds one slot per source rank: [nranks][Bs][X][Y]; the adjoint kernel of rank s stores
// its partial of image b into slot s of owner(b), and the owner sums its slots in rank order.
#define CTR_MAX_RANKS 16
struct CtrExchange {
    float* peer[CTR_MAX_RANKS];   // exchange buffer (current parity) of every rank, mapped into this rank
    int nranks, rank, Bs;
};
CTR_HD int ctr_xg_owner(int b, int Bs) { return b / Bs; }
CTR_HD size_t ctr_xg_index(int src_rank, int b_local, int Bs, int r, int c, int X, int Y)
{
    return (((size_t)src_rank * Bs + b_local) * X + r) * Y + c;
}

// Sinogram-bin window a pixel tile needs for one angle: packed start bin so that
// [start, start+win) covers every bin any of the tile's pixels can touch.
//   u(px,py) is linear, so its extremes over the tile sit at the 4 corners.
CTR_HD int ctr_window_start(float ua, float ub, float uc, float ud, int Wp2, int win)
{
    float umin = fminf(fminf(ua, ub), fminf(uc, ud));
    // keep the float->int conversion in range for far-out corners (pad=False)
    umin = fminf(fmaxf(umin, -4.f), (float)Wp2);
    int start = (int)floorf(umin) - 2 + 1;  // two bins of slack, +1: packed index
    const int maxs = Wp2 - win;
    if (start > maxs) start = maxs;
    if (start < 0) start = 0;
    return start;
}
