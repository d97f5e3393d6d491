/* capi_smoke.c -- libctradon.so driven from plain C, no Python anywhere: what a host in another language would do.
 *   host mode (default): geometry helpers and error conventions only -- runs without a GPU;
 *   ./capi_smoke gpu   : plan + forward + exact adjoint on device 0 and the adjoint identity <Ax,y> == <x,A^T y>.
 * Build (tests/test_capi.py does this):
 *   gcc -O2 -std=c11 -I include tests/capi/capi_smoke.c -o capi_smoke -L ct_pvae_b200 -lctradon \
 *       -L /usr/local/cuda/lib64 -lcudart -lm -Wl,-rpath,$PWD/ct_pvae_b200 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ctradon.h"

/* the CUDA runtime calls the GPU mode needs, declared by hand so that the file compiles without cuda_runtime.h */
extern int cudaMalloc(void** p, size_t n);
extern int cudaFree(void* p);
extern int cudaMemcpy(void* dst, const void* src, size_t n, int kind);
extern int cudaDeviceSynchronize(void);

#define CHECK(cond, msg)                                                         \
    do {                                                                         \
        if (!(cond)) { fprintf(stderr, "FAIL: %s (%s)\n", msg, ctr_last_error()); return 1; } \
    } while (0)

static int host_checks(void)
{
    CHECK(ctr_version() >= 200, "version");
    CHECK(ctr_num_proj_pix(128, 128) == 184 && ctr_num_proj_pix(512, 512) == 728, "pad_phantom detector size");
    int H, W, px, py;
    CHECK(ctr_frame(128, 128, 1, &H, &W, &px, &py) == CTR_OK && H == 184 && W == 184 && px == 28 && py == 28, "frame");
    const double theta[2] = {0.0, 1.5707963267948966};
    float t[16];
    CHECK(ctr_make_transforms(theta, 2, 2, 2, t) == CTR_OK, "transforms");
    CHECK(t[0] == 1.f && t[1] == 0.f && t[2] == 0.f && t[4] == 1.f, "theta = 0 is the identity transform");
    CHECK(ctr_num_proj_pix(0, 4) == CTR_EINVAL && strlen(ctr_last_error()) > 0, "bad argument -> CTR_EINVAL + message");
    CHECK(ctr_plan_destroy(NULL) == CTR_OK && ctr_comm_destroy(NULL) == CTR_OK, "destroying NULL is a no-op");
    ctr_plan* p = (ctr_plan*)1;
    CHECK(ctr_plan_create(NULL, 0, 4, 4, 0, 0, &p) == CTR_EINVAL && p == NULL, "plan_create rejects NULL theta");
    return 0;
}

static int gpu_checks(void)
{
    enum { B = 20, X = 48, A = 15 };
    double theta[A];
    for (int a = 0; a < A; ++a) theta[a] = 3.141592653589793 * a / A;
    ctr_plan* plan = NULL;
    CHECK(ctr_plan_create(theta, A, X, X, 1, 0, &plan) == CTR_OK, "plan_create");
    int W = 0;
    CHECK(ctr_plan_info(plan, NULL, NULL, NULL, NULL, &W, NULL, NULL) == CTR_OK && W == ctr_num_proj_pix(X, X), "plan_info");
    const size_t ni = (size_t)B * X * X, ns = (size_t)B * A * W;
    float *x = malloc(ni * 4), *y = malloc(ns * 4), *Ax = malloc(ns * 4), *Aty = malloc(ni * 4);
    unsigned s = 12345u;
    for (size_t i = 0; i < ni; ++i) { s = s * 1664525u + 1013904223u; x[i] = (float)(s >> 8) / 16777216.f; }
    for (size_t i = 0; i < ns; ++i) { s = s * 1664525u + 1013904223u; y[i] = (float)(s >> 8) / 16777216.f; }
    float *dx, *dy, *dAx, *dAty;
    void *ws1, *ws2;
    const size_t w1 = ctr_forward_workspace_bytes(plan, B), w2 = ctr_adjoint_workspace_bytes(plan, B);
    CHECK(!cudaMalloc((void**)&dx, ni * 4) && !cudaMalloc((void**)&dy, ns * 4) && !cudaMalloc((void**)&dAx, ns * 4) &&
              !cudaMalloc((void**)&dAty, ni * 4) && !cudaMalloc(&ws1, w1) && !cudaMalloc(&ws2, w2), "cudaMalloc");
    cudaMemcpy(dx, x, ni * 4, 1);
    cudaMemcpy(dy, y, ns * 4, 1);
    CHECK(ctr_radon_forward(plan, dx, dAx, B, CTR_INTERP_BILINEAR, ws1, w1, NULL) == CTR_OK, "forward");
    CHECK(ctr_radon_adjoint(plan, dy, dAty, B, CTR_INTERP_BILINEAR, CTR_ADJOINT_EXACT, ws2, w2, NULL) == CTR_OK, "adjoint");
    CHECK(ctr_radon_forward(plan, dx, dAx, B, CTR_INTERP_BILINEAR, ws1, w1 / 2, NULL) == CTR_EWORKSPACE, "short workspace is refused");
    CHECK(!cudaDeviceSynchronize(), "kernels ran");
    cudaMemcpy(Ax, dAx, ns * 4, 2);
    cudaMemcpy(Aty, dAty, ni * 4, 2);
    double lhs = 0, rhs = 0;
    for (size_t i = 0; i < ns; ++i) lhs += (double)Ax[i] * y[i];
    for (size_t i = 0; i < ni; ++i) rhs += (double)x[i] * Aty[i];
    printf("<Ax,y> = %.9g   <x,A^T y> = %.9g   rel diff %.2e   kernel launches %lld\n", lhs, rhs, fabs(lhs - rhs) / fabs(lhs), ctr_launch_count());
    CHECK(fabs(lhs - rhs) <= 2e-6 * fabs(lhs), "adjoint identity");
    cudaFree(dx); cudaFree(dy); cudaFree(dAx); cudaFree(dAty); cudaFree(ws1); cudaFree(ws2);
    free(x); free(y); free(Ax); free(Aty);
    CHECK(ctr_plan_destroy(plan) == CTR_OK, "plan_destroy");
    return 0;
}

int main(int argc, char** argv)
{
    if (host_checks()) return 1;
    if (argc > 1 && strcmp(argv[1], "gpu") == 0 && gpu_checks()) return 1;
    printf("capi_smoke ok (%s)\n", argc > 1 ? argv[1] : "host");
    return 0;
}
