/*
 * radon_oracle.c -- CPU restatement of CT_PVAE's parallel-beam projector path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under ct_pvae_b200/ may import, link or
 * execute this file; it is the checker for tests/, __graft_entry__.smoke() and
 * the cpu_baseline / --impl reference legs of bench.py.
 *
 * PARITY PARTLY PINNED: the reference (vganapati/CT_PVAE) ships no tests, golden
 * vectors or fixtures for this path, and its arithmetic lives in third-party
 * packages that are neither vendored nor installable here:
 *   tensorflow-addons==0.17.1  tfa.image.rotate / angles_to_projective_transforms
 *   tensorflow==2.8.1          ImageProjectiveTransformV3 (+ its registered gradient)
 * Their published algorithms are restated below.  Anchors on the reference's own
 * call sites:
 *   ctvae/forward_functions.py:18-46   pad_phantom          -> orc_num_proj_pix / pad offsets
 *   ctvae/forward_functions.py:80-123  project_tf_fast      -> orc_forward (+ orc_forward_dataflow)
 *   ctvae/forward_functions.py:49-78   project_tf_low_mem   -> orc_forward with interp=1
 *   ctvae/main_ct_vae.py:471-481       tape.gradient        -> orc_adjoint_tf (TF's registered gradient)
 *   (north_star)                       exact transpose      -> orc_adjoint_exact
 * The rotation is pinned by the known-answer vectors of tensorflow-addons' OWN test
 * suite (transform_ops_test.py::test_rotate_even / test_rotate_odd / test_bilinear,
 * transcribed in tests/test_tfa_known_answers.py): nearest exactly, bilinear to the
 * 1e-3 that test asserts.  Still UNPINNED: last-ulp behaviour of the bilinear kernel,
 * TF's registered gradient, tfp's interp_regular_1d_grid.
 * Further known answers (tests/test_oracle.py): the toy dataset's closed
 * form sinograms (scripts/images_to_sinograms.py:54-59), theta=0 column sums,
 * mass conservation, and an independent bilinear implementation
 * (torch grid_sample, align_corners=True, zeros padding).
 *
 * Build: see oracle/Makefile  (-O2 -ffp-contract=off: no FMA contraction, the
 * expression order below is the order TF's CPU kernel evaluates).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_NEAREST 0
#define ORC_BILINEAR 1

int orc_version(void) { return 1; }

/* torchrun exports OMP_NUM_THREADS=1 to every rank; the reference arm of bench.py (rank 0 alone) asks for
 * all host cores again. */
void orc_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* pad_phantom, forward_functions.py:29-30:
 *   num_proj_pix = ceil((sqrt(X^2 + Y^2) + 2) / 2) * 2      (float64) */
int orc_num_proj_pix(int X, int Y)
{
    double d = sqrt((double)((int64_t)X * X + (int64_t)Y * Y)) + 2.0;
    return (int)(ceil(d / 2.0) * 2.0);
}

/* pad_phantom, forward_functions.py:32-36: pad "before" = (P - n) // 2. */
void orc_pad_offsets(int X, int Y, int P, int *padx, int *pady)
{
    *padx = (P - X) / 2;
    *pady = (P - Y) / 2;
}

/*
 * tfa.image.rotate(images, -theta) -> angles_to_projective_transforms(-theta, H, W):
 * angle cast to float32, cos/sin in float32, offsets about ((W-1)/2, (H-1)/2),
 * row = [cos, -sin, x_off, sin, cos, y_off, 0, 0]   (forward_functions.py:113).
 * `theta` is what the caller passed to project_tf_fast; the minus sign is applied here.
 */
void orc_make_transforms(const double *theta, int A, int H, int W, float *t)
{
    const float wm1 = (float)W - 1.0f;
    const float hm1 = (float)H - 1.0f;
    for (int a = 0; a < A; ++a) {
        float ang = (float)(-theta[a]);
        float c = cosf(ang);
        float s = sinf(ang);
        float x_off = (wm1 - (c * wm1 - s * hm1)) / 2.0f;
        float y_off = (hm1 - (s * wm1 + c * hm1)) / 2.0f;
        float *r = t + 8 * a;
        r[0] = c;  r[1] = -s; r[2] = x_off;
        r[3] = s;  r[4] = c;  r[5] = y_off;
        r[6] = 0.f; r[7] = 0.f;
    }
}

/*
 * TF's gradient of ImageProjectiveTransformV3 inverts the 3x3 matrices with
 * tf.linalg.inv (float32, partially pivoted LU) and renormalises by m[2][2]
 * (tensorflow/python/ops/image_ops.py, _image_projective_transform_v3_grad;
 * flat_transforms_to_matrices / matrices_to_flat_transforms).  The elimination
 * order inside Eigen's LU is not observable from here, so this is a plain
 * partial-pivot Gauss-Jordan in float32; entries can differ from TF's by ulps.
 */
void orc_invert_transforms(const float *t, int A, float *tinv)
{
    for (int a = 0; a < A; ++a) {
        const float *r = t + 8 * a;
        float m[3][6] = {
            { r[0], r[1], r[2], 1.f, 0.f, 0.f },
            { r[3], r[4], r[5], 0.f, 1.f, 0.f },
            { r[6], r[7], 1.f,  0.f, 0.f, 1.f },
        };
        for (int col = 0; col < 3; ++col) {
            int piv = col;
            for (int k = col + 1; k < 3; ++k)
                if (fabsf(m[k][col]) > fabsf(m[piv][col])) piv = k;
            if (piv != col)
                for (int q = 0; q < 6; ++q) { float tmp = m[col][q]; m[col][q] = m[piv][q]; m[piv][q] = tmp; }
            float d = m[col][col];
            for (int q = 0; q < 6; ++q) m[col][q] = m[col][q] / d;
            for (int k = 0; k < 3; ++k) {
                if (k == col) continue;
                float f = m[k][col];
                for (int q = 0; q < 6; ++q) m[k][q] = m[k][q] - f * m[col][q];
            }
        }
        float w = m[2][5];
        float *o = tinv + 8 * a;
        o[0] = m[0][3] / w; o[1] = m[0][4] / w; o[2] = m[0][5] / w;
        o[3] = m[1][3] / w; o[4] = m[1][4] / w; o[5] = m[1][5] / w;
        o[6] = m[2][3] / w; o[7] = m[2][4] / w;
    }
}

/* A padded-frame read: the X x Y image sits at (padx, pady) inside the H x W
 * frame (tf.pad CONSTANT zeros, forward_functions.py:45); outside the frame the
 * op's fill_value (0) applies.  Both are zero, so one test on image bounds. */
static inline float frame_read(const float *img, int X, int Y, int padx, int pady,
                               int H, int W, int64_t fy, int64_t fx)
{
    if (fy < 0 || fy >= H || fx < 0 || fx >= W) return 0.f;
    int64_t r = fy - padx, c = fx - pady;
    if (r < 0 || r >= X || c < 0 || c >= Y) return 0.f;
    return img[r * Y + c];
}

/* ProjectiveGenerator::operator() of ImageProjectiveTransformV3 (fill_mode CONSTANT). */
static inline void out_to_in(const float *t, int ox_i, int oy_i, float *ix, float *iy, int *finite)
{
    float ox = (float)ox_i, oy = (float)oy_i;
    float projection = t[6] * ox + t[7] * oy + 1.f;
    if (projection == 0.f) { *finite = 0; *ix = 0.f; *iy = 0.f; return; }
    *finite = 1;
    *ix = (t[0] * ox + t[1] * oy + t[2]) / projection;
    *iy = (t[3] * ox + t[4] * oy + t[5]) / projection;
}

static inline float sample_frame(const float *img, int X, int Y, int padx, int pady, int H, int W,
                                 float x, float y, int interp)
{
    if (interp == ORC_NEAREST) {
        /* nearest_interpolation: std::round (half away from zero) */
        return frame_read(img, X, Y, padx, pady, H, W, (int64_t)roundf(y), (int64_t)roundf(x));
    }
    float yf = floorf(y), xf = floorf(x);
    float yc = yf + 1.f, xc = xf + 1.f;
    float v_f = (xc - x) * frame_read(img, X, Y, padx, pady, H, W, (int64_t)yf, (int64_t)xf)
              + (x - xf) * frame_read(img, X, Y, padx, pady, H, W, (int64_t)yf, (int64_t)xc);
    float v_c = (xc - x) * frame_read(img, X, Y, padx, pady, H, W, (int64_t)yc, (int64_t)xf)
              + (x - xf) * frame_read(img, X, Y, padx, pady, H, W, (int64_t)yc, (int64_t)xc);
    return (yc - y) * v_f + (y - yf) * v_c;
}

/*
 * project_tf_fast / project_tf_low_mem as one fused loop nest:
 *   sino[b, a, j] = sum_{i=0}^{H-1} rotate_a(pad(img_b))[i, j]
 * (forward_functions.py:113-114 / :70-75).  The per-sample values are float32
 * exactly as TF computes them; the row sum is taken in float64 and rounded once
 * (TF's reduction order is unspecified, this is the centre of what it can return).
 * img [B,X,Y], t [A,8], sino [B,A,W].
 */
void orc_forward(const float *img, int B, int X, int Y, int H, int W, int padx, int pady,
                 const float *t, int A, int interp, float *sino)
{
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b) {
        for (int a = 0; a < A; ++a) {
            const float *im = img + (size_t)b * X * Y;
            const float *ta = t + 8 * a;
            float *out = sino + ((size_t)b * A + a) * W;
            for (int j = 0; j < W; ++j) {
                double acc = 0.0;
                for (int i = 0; i < H; ++i) {
                    float x, y; int fin;
                    out_to_in(ta, j, i, &x, &y, &fin);
                    if (!fin) continue;
                    acc += (double)sample_frame(im, X, Y, padx, pady, H, W, x, y, interp);
                }
                out[j] = (float)acc;
            }
        }
    }
}

/*
 * The same operator following the reference's DATAFLOW (the CPU baseline that
 * bench.py times): zero-pad to [H,W], replicate per angle, materialise the
 * rotated stack [A,H,W,b] in float32, then reduce over rows in float32
 * (forward_functions.py:92-94,107-108,113-114,118-121).  `scratch` must hold
 * nthreads * H*W*Bc floats where Bc = channel (batch) count handled at once;
 * here each (angle) task rotates all B channels like TF's channel-last layout.
 * padded [H,W,B] channel-last is built by the caller-visible helper below.
 */
void orc_forward_dataflow(const float *img, int B, int X, int Y, int H, int W, int padx, int pady,
                          const float *t, int A, int interp, float *sino)
{
    /* tf.pad + tf.transpose(perm=[3,1,2,0]) : [1,H,W,B] */
    float *padded = (float *)calloc((size_t)H * W * B, sizeof(float));
    for (int b = 0; b < B; ++b)
        for (int r = 0; r < X; ++r)
            for (int c = 0; c < Y; ++c)
                padded[((size_t)(r + padx) * W + (c + pady)) * B + b] = img[((size_t)b * X + r) * Y + c];
#pragma omp parallel
    {
        float *rot = (float *)malloc((size_t)H * W * B * sizeof(float));
#pragma omp for schedule(dynamic, 1)
        for (int a = 0; a < A; ++a) {
            const float *ta = t + 8 * a;
            /* ImageProjectiveTransformV3 on image n = a of the repeated stack */
            for (int i = 0; i < H; ++i) {
                for (int j = 0; j < W; ++j) {
                    float x, y; int fin;
                    out_to_in(ta, j, i, &x, &y, &fin);
                    float *dst = rot + ((size_t)i * W + j) * B;
                    if (!fin) { memset(dst, 0, sizeof(float) * B); continue; }
                    if (interp == ORC_NEAREST) {
                        int64_t yy = (int64_t)roundf(y), xx = (int64_t)roundf(x);
                        if (yy < 0 || yy >= H || xx < 0 || xx >= W) { memset(dst, 0, sizeof(float) * B); continue; }
                        memcpy(dst, padded + ((size_t)yy * W + xx) * B, sizeof(float) * B);
                    } else {
                        float yf = floorf(y), xf = floorf(x), yc = yf + 1.f, xc = xf + 1.f;
                        int64_t y0 = (int64_t)yf, x0 = (int64_t)xf, y1 = (int64_t)yc, x1 = (int64_t)xc;
                        int v00 = (y0 >= 0 && y0 < H && x0 >= 0 && x0 < W);
                        int v01 = (y0 >= 0 && y0 < H && x1 >= 0 && x1 < W);
                        int v10 = (y1 >= 0 && y1 < H && x0 >= 0 && x0 < W);
                        int v11 = (y1 >= 0 && y1 < H && x1 >= 0 && x1 < W);
                        const float *p00 = v00 ? padded + ((size_t)y0 * W + x0) * B : 0;
                        const float *p01 = v01 ? padded + ((size_t)y0 * W + x1) * B : 0;
                        const float *p10 = v10 ? padded + ((size_t)y1 * W + x0) * B : 0;
                        const float *p11 = v11 ? padded + ((size_t)y1 * W + x1) * B : 0;
                        float wxf = xc - x, wxc = x - xf, wyf = yc - y, wyc = y - yf;
                        for (int b = 0; b < B; ++b) {
                            float vf = wxf * (v00 ? p00[b] : 0.f) + wxc * (v01 ? p01[b] : 0.f);
                            float vc = wxf * (v10 ? p10[b] : 0.f) + wxc * (v11 ? p11[b] : 0.f);
                            dst[b] = wyf * vf + wyc * vc;
                        }
                    }
                }
            }
            /* tf.reduce_sum(axis=1) then transpose to [B,A,W] */
            for (int j = 0; j < W; ++j)
                for (int b = 0; b < B; ++b) {
                    float acc = 0.f;
                    for (int i = 0; i < H; ++i) acc += rot[((size_t)i * W + j) * B + b];
                    sino[((size_t)b * A + a) * W + j] = acc;
                }
        }
        free(rot);
    }
    free(padded);
}

/*
 * Exact transpose of orc_forward's linear map (north_star: <Ax,y> = <x,A^T y>):
 *   g[b, r, c] = sum_{a,j,i} w_{a,j,i}(r,c) * y[b,a,j]
 * with the very same float32 sample coordinates and tap weights as the forward;
 * products and sums are taken in float64 and rounded once.
 * y [B,A,W], g [B,X,Y].
 */
void orc_adjoint_exact(const float *ys, int B, int X, int Y, int H, int W, int padx, int pady,
                       const float *t, int A, int interp, float *g)
{
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < B; ++b) {
        double *acc = (double *)calloc((size_t)X * Y, sizeof(double));
        for (int a = 0; a < A; ++a) {
            const float *ta = t + 8 * a;
            const float *yr = ys + ((size_t)b * A + a) * W;
            for (int j = 0; j < W; ++j) {
                double yv = (double)yr[j];
                for (int i = 0; i < H; ++i) {
                    float x, y; int fin;
                    out_to_in(ta, j, i, &x, &y, &fin);
                    if (!fin) continue;
                    int64_t fy[4], fx[4]; double w[4]; int n;
                    if (interp == ORC_NEAREST) {
                        fy[0] = (int64_t)roundf(y); fx[0] = (int64_t)roundf(x); w[0] = 1.0; n = 1;
                    } else {
                        float yf = floorf(y), xf = floorf(x), yc = yf + 1.f, xc = xf + 1.f;
                        float wxf = xc - x, wxc = x - xf, wyf = yc - y, wyc = y - yf;
                        fy[0] = (int64_t)yf; fx[0] = (int64_t)xf; w[0] = (double)wyf * (double)wxf;
                        fy[1] = (int64_t)yf; fx[1] = (int64_t)xc; w[1] = (double)wyf * (double)wxc;
                        fy[2] = (int64_t)yc; fx[2] = (int64_t)xf; w[2] = (double)wyc * (double)wxf;
                        fy[3] = (int64_t)yc; fx[3] = (int64_t)xc; w[3] = (double)wyc * (double)wxc;
                        n = 4;
                    }
                    for (int k = 0; k < n; ++k) {
                        if (fy[k] < 0 || fy[k] >= H || fx[k] < 0 || fx[k] >= W) continue;
                        int64_t r = fy[k] - padx, c = fx[k] - pady;
                        if (r < 0 || r >= X || c < 0 || c >= Y) continue;
                        acc[r * Y + c] += w[k] * yv;
                    }
                }
            }
        }
        float *gb = g + (size_t)b * X * Y;
        for (size_t p = 0; p < (size_t)X * Y; ++p) gb[p] = (float)acc[p];
        free(acc);
    }
}

/*
 * TensorFlow's registered gradient of the same graph (main_ct_vae.py:471-481):
 *   grad(reduce_sum axis 1) : Z_a[i, j] = y[b,a,j] for every row i of the frame
 *   grad(rotate)            : ImageProjectiveTransformV3(Z, inverse transforms, same interpolation, fill 0)
 *   grad(repeat)            : sum over angles;   grad(pad): crop to the X x Y window.
 * It is a pixel-driven interpolating back-projection, NOT the transpose of orc_forward.
 * y [B,A,W], tinv [A,8] from orc_invert_transforms, g [B,X,Y].
 */
void orc_adjoint_tf(const float *ys, int B, int X, int Y, int H, int W, int padx, int pady,
                    const float *tinv, int A, int interp, float *g)
{
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b) {
        for (int r = 0; r < X; ++r) {
            for (int c = 0; c < Y; ++c) {
                double acc = 0.0;
                for (int a = 0; a < A; ++a) {
                    const float *yr = ys + ((size_t)b * A + a) * W;
                    float x, y; int fin;
                    out_to_in(tinv + 8 * a, c + pady, r + padx, &x, &y, &fin);
                    if (!fin) continue;
#define ZREAD(yy, xx) (((yy) >= 0 && (yy) < H && (xx) >= 0 && (xx) < W) ? yr[(xx)] : 0.f)
                    float v;
                    if (interp == ORC_NEAREST) {
                        int64_t yy = (int64_t)roundf(y), xx = (int64_t)roundf(x);
                        v = ZREAD(yy, xx);
                    } else {
                        float yf = floorf(y), xf = floorf(x), yc = yf + 1.f, xc = xf + 1.f;
                        int64_t y0 = (int64_t)yf, x0 = (int64_t)xf, y1 = (int64_t)yc, x1 = (int64_t)xc;
                        float v_f = (xc - x) * ZREAD(y0, x0) + (x - xf) * ZREAD(y0, x1);
                        float v_c = (xc - x) * ZREAD(y1, x0) + (x - xf) * ZREAD(y1, x1);
                        v = (yc - y) * v_f + (y - yf) * v_c;
                    }
#undef ZREAD
                    acc += (double)v;
                }
                g[((size_t)b * X + r) * Y + c] = (float)acc;
            }
        }
    }
}

/*
 * Back-projection stage of iradon (ctvae/fbp_tensorflow.py:52-74) in float64, on
 * an already filtered sinogram rf[B,A,P] (the FFT filter stage :49-50 is done in
 * numpy by oracle/radon_oracle.py):
 *   t = y'cos(theta) - x'sin(theta), x' = row - x_size/2, y' = col - y_size/2
 *   tfp.math.interp_regular_1d_grid(t, -P/2, P/2-1, rf[:,a,:]) with the default
 *   fill_value='constant_extension' (edge clamp), summed over angles, * pi/(2A).
 */
void orc_iradon_backproject(const double *rf, const double *theta, int B, int A, int P,
                            int x_size, int y_size, double *out)
{
    const double ref_min = 0.0 - (double)P / 2.0;
    const double ref_max = (double)(P - 1) - (double)P / 2.0;
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b) {
        for (int r = 0; r < x_size; ++r) {
            for (int c = 0; c < y_size; ++c) {
                double xpr = (double)r - (double)x_size / 2.0;
                double ypr = (double)c - (double)y_size / 2.0;
                double acc = 0.0;
                for (int a = 0; a < A; ++a) {
                    const double *row = rf + ((size_t)b * A + a) * P;
                    double tt = ypr * cos(theta[a]) - xpr * sin(theta[a]);
                    /* interp_regular_1d_grid */
                    double idx = (tt - ref_min) / (ref_max - ref_min) * (double)(P - 1);
                    if (idx < 0.0) idx = 0.0;
                    if (idx > (double)(P - 1)) idx = (double)(P - 1);
                    double below = floor(idx);
                    double above = below + 1.0;
                    if (above > (double)(P - 1)) above = (double)(P - 1);
                    below = above - 1.0;
                    if (below < 0.0) below = 0.0;
                    double alpha = idx - below;
                    acc += (1.0 - alpha) * row[(int)below] + alpha * row[(int)above];
                }
                out[((size_t)b * x_size + r) * y_size + c] = acc * M_PI / (2.0 * (double)A);
            }
        }
    }
}
