/*
 * ctradon.h -- C ABI of libctradon.so, the B200 (sm_100a) Radon path of CT_PVAE.
 *
 * The reference exposes no FFI: its boundary is four Python functions
 *   ctvae/forward_functions.py:18   pad_phantom(phantom, dim=3, integrate_vae=False)
 *   ctvae/forward_functions.py:49   project_tf_low_mem(phantom, theta, pad=False)
 *   ctvae/forward_functions.py:80   project_tf_fast(phantom, theta, pad=False, dim=3, integrate_vae=False)
 *   ctvae/fbp_tensorflow.py:14      iradon(sinogram, theta, x_size, y_size, filter_1d)
 * plus TensorFlow's autodiff through project_tf_fast (ctvae/main_ct_vae.py:471-481).
 * The entry points below are what a ctypes binding of those functions calls; each
 * one names the reference lines it replaces.  INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - Every function returns CTR_OK (0) or a negative CTR_E* code and never throws
 *     or aborts; ctr_last_error() gives the thread-local message of the last failure.
 *   - The caller allocates and owns every buffer, inputs, outputs and workspace
 *     alike.  The library borrows pointers for the duration of the call, never
 *     allocates result memory and never calls a DLPack deleter.
 *   - Device buffers are float32, C-contiguous: images [B,X,Y], sinograms [B,A,W]
 *     (the reference's trailing size-1 channel axis is a free reshape).
 *   - Work is enqueued on `stream` (a cudaStream_t passed as void*, NULL = legacy
 *     default stream) and the call returns without synchronising.
 *   - Plans are immutable after creation and may be shared between threads.
 *   - There is no CPU path: without a CUDA device every compute call fails with
 *     CTR_ECUDA.
 */
#ifndef CTRADON_H_
#define CTRADON_H_
#include <stddef.h>
#include <stdint.h>

#include "ctr_dlpack.h"

#ifdef __cplusplus
extern "C" {
#endif

#define CTR_VERSION 200 /* 0.2.0 */

enum {
    CTR_OK = 0,
    CTR_EINVAL = -1,      /* bad argument (shape, enum, NULL, dtype, device, contiguity) */
    CTR_ECUDA = -2,       /* a CUDA runtime call failed; message carries cudaGetErrorString */
    CTR_EWORKSPACE = -3,  /* workspace smaller than ctr_*_workspace_bytes() */
    CTR_EUNSUPPORTED = -4, /* shape outside what the kernels are built for */
    CTR_ECOMM = -5         /* angle-sharded exchange failed: a peer did not arrive in time, or an NCCL error */
};

enum { CTR_INTERP_NEAREST = 0, CTR_INTERP_BILINEAR = 1 };
/* CTR_ADJOINT_EXACT    : exact transpose of the forward operator (<Ax,y> == <x,A^T y>)
 * CTR_ADJOINT_TF_COMPAT: TensorFlow's registered gradient of the reference graph
 *                        (rotate the row-broadcast cotangent by the inverted transforms) */
enum { CTR_ADJOINT_EXACT = 0, CTR_ADJOINT_TF_COMPAT = 1 };

typedef struct ctr_plan ctr_plan;
typedef struct ctr_fbp_plan ctr_fbp_plan;

int ctr_version(void);
const char* ctr_last_error(void);
/* number of kernel launches this process has made through the library (bench.py's gpu_launches) */
long long ctr_launch_count(void);

/* ---- per-kernel timing (bench.py's roofline leg) ---------------------------------------------
 * When enabled, every kernel launch is bracketed by cudaEvents recorded on the
 * launching stream; ctr_profile_read synchronises those events and returns the summed
 * device time and launch count of one kernel since the last reset. */
enum {
    CTR_K_PACK_IMAGE = 0, CTR_K_PACK_SINO = 1, CTR_K_FORWARD = 2, CTR_K_ADJ_EXACT = 3,
    CTR_K_ADJ_TF = 4, CTR_K_FBP_FILTER = 5, CTR_K_FBP_BP = 6, CTR_K_XCHG_SUM = 7, CTR_K_FBP_FUSED = 8, CTR_K_COUNT = 9
};
int ctr_profile_enable(int on);
int ctr_profile_reset(void);
int ctr_profile_read(int kernel_id, double* total_ms, long long* launches);
const char* ctr_kernel_name(int kernel_id);

/* Blocks until every stream of `device` has drained.  For host frameworks whose compute stream cannot be passed in
 * (the TensorFlow binding): synchronise, call with stream = NULL, synchronise. */
int ctr_device_synchronize(int device);

/* ---- host-side geometry (no GPU needed) ------------------------------------------------ */

/* forward_functions.py:29-30: ceil((sqrt(X^2+Y^2)+2)/2)*2 */
int ctr_num_proj_pix(int X, int Y);
/* forward_functions.py:32-39: frame the image is padded to and the "before" pads */
int ctr_frame(int X, int Y, int pad, int* H, int* W, int* padx, int* pady);
/* forward_functions.py:113 -> tfa.image.rotate(imgs, -theta): [A,8] float32 projective
 * transforms (angles_to_projective_transforms), float32 arithmetic, no FMA */
int ctr_make_transforms(const double* theta, int A, int H, int W, float* out_a8);
/* TF's gradient of ImageProjectiveTransformV3: float32 3x3 inverse of each transform */
int ctr_invert_transforms(const float* t_a8, int A, float* out_a8);
/* real(ifft(filter_1d)): the spatial kernel equivalent to fbp_tensorflow.py:49-50's
 * frequency-domain product for real sinograms (filt_im may be NULL) */
int ctr_filter_to_spatial(const double* filt_re, const double* filt_im, int P, double* out_p);

/* ---- projector plans --------------------------------------------------------------------- */

/* Geometry of one project_tf_fast call: the angles `theta` (radians, as passed to the
 * reference function; the minus sign of forward_functions.py:113 is applied inside),
 * the image size X x Y (rows x cols) and whether pad_phantom is applied.  Uploads the
 * transform tables and the class-sorted ray table to `device`. */
int ctr_plan_create(const double* theta, int A, int X, int Y, int pad, int device, ctr_plan** out);
int ctr_plan_destroy(ctr_plan* plan);
/* any of the out pointers may be NULL */
int ctr_plan_info(const ctr_plan* plan, int* A, int* X, int* Y, int* H, int* W, int* padx, int* pady);
/* copies the plan's [A,8] tables to host buffers (either may be NULL) */
int ctr_plan_tables(const ctr_plan* plan, float* fwd_a8, float* inv_a8);

/* one line describing the forward kernel shape a batch of B images would use (diagnostics) */
int ctr_plan_describe(const ctr_plan* plan, int B, char* buf, size_t n);

size_t ctr_forward_workspace_bytes(const ctr_plan* plan, int B);
size_t ctr_adjoint_workspace_bytes(const ctr_plan* plan, int B);

/* project_tf_fast (interp = CTR_INTERP_NEAREST, forward_functions.py:80-123) and
 * project_tf_low_mem (CTR_INTERP_BILINEAR, :49-78):
 *   sino[b,a,j] = sum_i rotate_{-theta_a}(pad(img_b))[i,j]
 * img [B,X,Y] -> sino [B,A,W], both on the plan's device. */
int ctr_radon_forward(const ctr_plan* plan, const float* img, float* sino, int B, int interp,
                      void* workspace, size_t workspace_bytes, void* stream);

/* Gradient of the above w.r.t. img given d(loss)/d(sino)  (main_ct_vae.py:471-481):
 * dsino [B,A,W] -> dimg [B,X,Y].  mode selects the exact transpose or TF's gradient. */
int ctr_radon_adjoint(const ctr_plan* plan, const float* dsino, float* dimg, int B, int interp, int mode,
                      void* workspace, size_t workspace_bytes, void* stream);

/* Same, with the result multiplied by `scale` (the upstream scalar of a summed loss). */
int ctr_radon_adjoint_scaled(const ctr_plan* plan, const float* dsino, float* dimg, int B, int interp, int mode,
                             float scale, void* workspace, size_t workspace_bytes, void* stream);

/* ---- angle subsets (training's angle minibatch) ----------------------------------------------------
 * The training loop projects a random subset of the angles every iteration (helper_functions.py:350-357:
 * tf.gather(theta, angles_i) before the projector call).  Instead of a plan per subset, ONE plan over all the angles
 * serves every subset: `sel` is a DEVICE array of n_sel int32 indices into the plan's angles; sinogram row k of the
 * call is angle sel[k].  sino / dsino are [B, n_sel, W].  Nothing is allocated, uploaded or freed on this path.
 * Workspaces: the plain ctr_*_workspace_bytes of the same B. */
int ctr_radon_forward_sel(const ctr_plan* plan, const float* img, float* sino, int B, int interp, const int* sel, int n_sel,
                          void* workspace, size_t workspace_bytes, void* stream);
int ctr_radon_adjoint_sel(const ctr_plan* plan, const float* dsino, float* dimg, int B, int interp, int mode, float scale,
                          const int* sel, int n_sel, void* workspace, size_t workspace_bytes, void* stream);
/* ctr_radon_loglik (below) on a subset: mask [B,A] and meas [B,A,W] cover ALL the plan's angles and are read at
 * column sel[k]; dproj is [B, n_sel, W] */
int ctr_radon_loglik_sel(const ctr_plan* plan, const float* img, const float* mask, const float* meas, const int* sel,
                         int n_sel, float pnm, float sqrt_reg, float* loglik, float* dproj, int B, int interp,
                         void* workspace, size_t workspace_bytes, void* stream);

/* ---- fused measurement log-likelihood ------------------------------------------
 * calculate_log_prob_M_given_R (ctvae/helper_functions.py:336-368) reduced over angles and
 * bins as find_loss_vae_unsup does (:305-311), in ONE pass of the projector:
 *   proj   = project_tf_fast(img, theta)                      (never written)
 *   pm     = proj * mask[b, angle_map[a]]
 *   lp     = Normal(pm, sqrt_reg + sqrt(pm/pnm + sqrt_reg)).log_prob(meas[b, angle_map[a], j])
 *   loglik[b] = sum_{a,j} lp          dproj[b,a,j] = d lp / d proj
 * dproj [B,A,W] is the cotangent ctr_radon_adjoint[_scaled] turns into d loglik / d img.
 * mask [B,A_all], meas [B,A_all,W], angle_map [A] device int32 (NULL = identity, the
 * reference's angles_i gather :355-357), loglik [B]; all on the plan's device. */
size_t ctr_loglik_workspace_bytes(const ctr_plan* plan, int B);
int ctr_radon_loglik(const ctr_plan* plan, const float* img, const float* mask, const float* meas,
                     const int* angle_map, int A_all, float pnm, float sqrt_reg, float* loglik, float* dproj,
                     int B, int interp, void* workspace, size_t workspace_bytes, void* stream);

/* ---- filtered back-projection ------------------------------------------------------------ */

/* iradon(sinogram, theta, x_size, y_size, filter_1d)  (fbp_tensorflow.py:14-75).
 * filter_1d is the length-P frequency-domain filter in FFT order (imaginary part
 * optional).  Returns CTR_EINVAL if A does not match at call time, like the
 * reference's ValueError (:43-45). */
int ctr_fbp_plan_create(const double* theta, int A, int P, int x_size, int y_size, const double* filt_re,
                        const double* filt_im, int device, ctr_fbp_plan** out);
int ctr_fbp_plan_destroy(ctr_fbp_plan* plan);
/* Two execution paths, same arithmetic:
 *  - two kernels (default): row filter into a packed sinogram (L2-resident), then the interpolating gather;
 *  - ONE kernel (on != 0; images of up to 128 x 128 = 8 x 2048 pixels): a thread-block cluster filters the sinogram
 *    rows in shared memory and back-projects them from distributed shared memory (ctr_fbp_fused_kernel), so the
 *    filtered rows never leave the chip.  Both paths are bit-identical.  The single kernel
 *    holds 64 accumulators per thread and runs one 512-thread CTA per SM; measured on B200 it is the slower of the two
 *    (DESIGN.md section 4), which is why it is opt-in.
 * Returns 1 if on != 0 but the plan's geometry has no single-kernel path. */
int ctr_fbp_plan_set_fused(ctr_fbp_plan* plan, int on);
size_t ctr_fbp_workspace_bytes(const ctr_fbp_plan* plan, int B);
/* sino [B,A,P] -> recon [B,x_size,y_size] (float32 on device; the Python shim widens
 * to float64 to keep the reference's return dtype) */
int ctr_fbp(const ctr_fbp_plan* plan, const float* sino, int A, float* recon, int B, void* workspace,
            size_t workspace_bytes, void* stream);

/* ---- host-buffer pipeline -------------------------------------------------------------------
 * The reference's callers hand NumPy arrays to project_tf_fast (main_ct_vae.py:523-524,
 * scripts/images_to_sinograms.py:61-68).  A pipe owns device staging for `chunk` images per
 * slot (6 slots), its workspace and three streams; a call cuts the host batch into chunks
 * and overlaps copy-in, kernels and copy-out.  Calls only ENQUEUE and return; results are
 * valid after ctr_hostpipe_wait.  Page-locked (pinned) host buffers are copied directly.  Pageable
 * memory -- a plain NumPy array, what the reference's callers pass -- is staged through pinned
 * buffers owned by the slots (worker threads fill the next chunk's buffer while the previous one
 * crosses PCIe); a pageable input may be released as soon as the call returns, a pageable result
 * is complete after ctr_hostpipe_wait.  Calls on one pipe are serialised by a mutex and
 * their chunks share the ring, so back-to-back calls overlap each other's copies.
 * ctr_hostpipe_done: 1 if everything enqueued so far has completed, 0 if not, < 0 on error. */
typedef struct ctr_hostpipe ctr_hostpipe;
int ctr_hostpipe_create(const ctr_plan* plan, int chunk, ctr_hostpipe** out);
int ctr_hostpipe_destroy(ctr_hostpipe* pipe);
/* img_host [B,X,Y] -> sino_host [B,A,W] */
int ctr_hostpipe_forward(ctr_hostpipe* pipe, const float* img_host, float* sino_host, int B, int interp);
/* dsino_host [B,A,W] -> dimg_host [B,X,Y] */
int ctr_hostpipe_adjoint(ctr_hostpipe* pipe, const float* dsino_host, float* dimg_host, int B, int interp, int mode);
int ctr_hostpipe_wait(ctr_hostpipe* pipe);
int ctr_hostpipe_done(ctr_hostpipe* pipe);
/* diagnostics: record every chunk's copy-in / kernel / copy-out interval; ctr_hostpipe_wait prints them to stderr */
int ctr_hostpipe_trace(ctr_hostpipe* pipe, int on);

/* ---- angle-sharded mode (BASELINE configs[3]) ---------------------------------------
 * The reference has no multi-device layer at all (configs/config_gpu.yaml:44 is a string for an external
 * launcher); this one is the new framework's.  One rank per GPU.  Every rank holds all B images and a contiguous
 * block of the angles; forward outputs are disjoint sinogram rows (ctr_radon_forward on the rank's own plan, no
 * exchange).  The adjoint / FBP of an angle block is a full-size PARTIAL image; the partials are summed over the
 * ranks and the result is left batch-sharded: rank r receives images [r*B/nranks, (r+1)*B/nranks).
 *
 * CTR_EXCHANGE_P2P (the product path): the back-projection kernel's epilogue stores each tile straight into the
 *   owner rank's exchange buffer over NVLink peer memory (no separate collective); a flag exchange and one
 *   fixed-order sum of the nranks slots follow in ctr_xchg_sum_kernel.  Deterministic (rank order).
 * CTR_EXCHANGE_NCCL (the baseline): the kernel writes the partial locally, ncclReduceScatter sums it.
 *
 * Set-up between processes: every rank calls ctr_comm_create, ctr_comm_export, exchanges the
 * CTR_COMM_HANDLE_BYTES blobs with the other ranks by any means (torch.distributed all_gather, MPI, a file) and
 * passes all of them, in rank order, to ctr_comm_connect.  Inside one process ctr_comm_create_all does it all
 * (the ncclCommInitAll analogue).  NCCL is optional: rank 0 calls ctr_comm_nccl_unique_id, broadcasts the 128 bytes,
 * every rank calls ctr_comm_nccl_init (collective).  libnccl.so.2 is opened at run time.
 *
 * Failure detection: a rank that waits longer than the timeout (default 10 s) for a peer's flag stops waiting and
 * raises the comm's error word; ctr_comm_check then returns CTR_ECOMM naming the missing rank.  It also forwards
 * ncclCommGetAsyncError.  Calls on one comm must be issued in the same order on every rank. */
typedef struct ctr_comm ctr_comm;
#define CTR_COMM_HANDLE_BYTES 128
#define CTR_NCCL_ID_BYTES 128
enum { CTR_EXCHANGE_P2P = 0, CTR_EXCHANGE_NCCL = 1 };
/* exchange_bytes >= B*X*Y*4 of the largest call this comm will serve */
int ctr_comm_create(int nranks, int rank, int device, size_t exchange_bytes, ctr_comm** out);
int ctr_comm_export(const ctr_comm* comm, void* handle_out);
int ctr_comm_connect(ctr_comm* comm, const void* handles_in_rank_order);
int ctr_comm_create_all(int nranks, const int* devices, size_t exchange_bytes, ctr_comm** out_nranks);
int ctr_comm_nccl_unique_id(void* id_out);
int ctr_comm_nccl_init(ctr_comm* comm, const void* id_in);
int ctr_comm_set_timeout_ms(ctr_comm* comm, int milliseconds);
int ctr_comm_info(const ctr_comm* comm, int* nranks, int* rank, int* connected, int* has_nccl, size_t* exchange_bytes);
int ctr_comm_check(ctr_comm* comm);
int ctr_comm_destroy(ctr_comm* comm);
size_t ctr_adjoint_sharded_workspace_bytes(const ctr_comm* comm, const ctr_plan* plan, int B);
/* dsino_local [B, A_local, W] (the plan's angles are this rank's block) -> dimg_shard [B/nranks, X, Y] */
int ctr_radon_adjoint_sharded(ctr_comm* comm, const ctr_plan* plan, const float* dsino_local, float* dimg_shard, int B,
                              int interp, int mode, int algo, void* workspace, size_t workspace_bytes, void* stream);
/* angle-sharded iradon: sino_local [B, A_local, P] -> recon_shard [B/nranks, x_size, y_size]; the pi/(2A) scale of
 * fbp_tensorflow.py:74 uses A_total.  Workspace: ctr_fbp_workspace_bytes. */
int ctr_fbp_sharded(ctr_comm* comm, const ctr_fbp_plan* plan, const float* sino_local, int A_local, int A_total,
                    float* recon_shard, int B, int algo, void* workspace, size_t workspace_bytes, void* stream);

/* ---- zero-copy DLPack entry points ---------------------------------------------------------
 * Same operations on borrowed DLTensors (kDLCUDA, float32, compact row-major).
 * img may be [B,X,Y] or [B,X,Y,1]; sino [B,A,W] or [B,A,W,1].  `workspace` is any
 * compact CUDA tensor with at least ctr_*_workspace_bytes() bytes. */
int ctr_radon_forward_dl(const ctr_plan* plan, const DLTensor* img, DLTensor* sino, int interp,
                         DLTensor* workspace, void* stream);
int ctr_radon_adjoint_dl(const ctr_plan* plan, const DLTensor* dsino, DLTensor* dimg, int interp, int mode,
                         DLTensor* workspace, void* stream);
int ctr_fbp_dl(const ctr_fbp_plan* plan, const DLTensor* sino, DLTensor* recon, DLTensor* workspace, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CTRADON_H_ */
