/*
 * tr_b200.h — C ABI of libtrb200.so: the CP (Kruskal) tensor-regression fit iteration
 * (forward contraction, residual-weighted MTTKRP gradients, penalty + Adam update) as
 * hand-written sm_100a CUDA kernels.
 *
 * The reference (kimerein/tensor_regression) has no FFI: its boundary for this path is the
 * Python surface of standard_tensor_regression.py / multinomial_tensor_regression.py.  Each
 * entry point below names the reference lines whose work it replaces; the Python modules in
 * tensor_regression_b200/ keep the reference's signatures and call these through ctypes
 * (INTEGRATION.md shows the binding).
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless marked host;
 *   - the caller owns every buffer; the library owns only the opaque handle + its workspace;
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); no call synchronises
 *     the device except tr_create / tr_reserve / tr_destroy (allocation);
 *   - every function returns 0 on success, non-zero on error (TR_ERR_*); the message is
 *     available from tr_last_error(); no exceptions cross the boundary;
 *   - there is no CPU fallback: without a CUDA device tr_create fails.
 *   - dtype: TR_F32 / TR_F64 is the type of X, theta, w, y (std), the predictions and `grad`.
 *     `gradsum` and `loss` are always double.
 *
 * Parameter vector layout (theta, grad, Adam state), row-major (I_m, R) blocks:
 *     theta = [ F_0 | F_1 | ... | F_{k-1} | F_C (multinomial only, (C,R)) | bias (standard only) ]
 * which is the reference's Kruskal list `Bcp` (README.md:12-13,21-22; mn:280) laid end to end.
 * Pf = R * (sum_m I_m + C) is the number of factor entries; P = Pf + (C == 0 ? 1 : 0).
 *
 * gradsum layout (unnormalised LOCAL sums over the N samples of this call — the vector one
 * NCCL all-reduce sums across GPUs when the sample axis is sharded):
 *     standard   : [ dFt (Pf) | sum_n res_n | sum_n res_n^2 ]                 (Pf + 2 doubles)
 *     multinomial: [ dFt (Pf, class factor last) | sum_n -omega[y_n] log Q[n,y_n] ]   (Pf + 1)
 * where dFt is the derivative with respect to the softplus-ed factors before the 2/N (or
 * 1/W) normalisation, res_n = yhat_n - y_n.
 */
#ifndef TR_B200_H
#define TR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TR_B200_VERSION 210

enum { TR_F32 = 0, TR_F64 = 1 };
enum { TR_OK = 0, TR_ERR_INVALID = 1, TR_ERR_CUDA = 2, TR_ERR_UNSUPPORTED = 3, TR_ERR_NOMEM = 4 };

#define TR_MAX_MODES 8      /* feature modes of X (k) */
#define TR_MAX_CLASSES 128  /* multinomial classes */

typedef struct tr_handle tr_handle;

int tr_version(void);

/* Plan for one model geometry.  dims = host array of the k feature-mode sizes (X.shape[1:]),
 * R = CP rank, C = 0 for the standard model, n_classes for the multinomial model.
 * Replaces the geometry bookkeeping of CP_linear_regression.__init__ (std:204-303) and
 * CP_logistic_regression.__init__ (mn:212-286). */
int tr_create(tr_handle** out, int dtype, int k, const int64_t* dims, int R, int C, int device);
int tr_destroy(tr_handle* h);

/* Last error message of this handle (h may be NULL: message of the last failed tr_create). */
const char* tr_last_error(tr_handle* h);

/* P (all trainable scalars, incl. bias for the standard model) and Pf (factor entries). */
int tr_param_count(tr_handle* h, int64_t* P, int64_t* Pf);

/* Number of doubles in gradsum (Pf + 2 standard, Pf + 1 multinomial). */
int tr_gradsum_count(tr_handle* h, int64_t* count);

/* Pre-allocate the workspace for calls with up to N samples (otherwise grown on demand by the
 * first call that needs it, which then synchronises once). */
int tr_reserve(tr_handle* h, int64_t N);

/* Forward only — lin_model (std:87-130): yhat[n] = <X_n, cp_to_tensor(w, softplus?(F))> + bias.
 * X: (N, I_1..I_k) row-major contiguous; yhat: (N).  nn_mask bit m = softplus on factor m
 * (non_neg_fn, std:53-85) with torch.nn.functional.softplus(beta, threshold) semantics. */
int tr_forward_std(tr_handle* h, const void* X, int64_t N, const void* theta, const void* w,
                   uint32_t nn_mask, double sp_beta, double sp_thr, void* yhat, void* stream);

/* Forward only — model (mn:148-187): P[n,:] = softmax(<X_n, cp_to_tensor(w, F incl. class factor)>).
 * P: (N, C).  pred (int64, N) may be NULL; otherwise argmax_c P[n,c] (predict, mn:526-527). */
int tr_forward_mn(tr_handle* h, const void* X, int64_t N, const void* theta, const void* w,
                  uint32_t nn_mask, double sp_beta, double sp_thr, void* P, int64_t* pred, void* stream);

/* One closure evaluation without the penalty and normalisation (std:368-373 / 459-462):
 * forward, residual, and all per-mode residual-weighted MTTKRP gradients in two streaming
 * passes over X.  y: (N) same dtype as X.  yhat may be NULL. */
int tr_fwd_grad_std(tr_handle* h, const void* X, const void* y, int64_t N, const void* theta,
                    const void* w, uint32_t nn_mask, double sp_beta, double sp_thr,
                    double* gradsum, void* yhat, void* stream);

/* Same for the multinomial model (mn:357-362 / 454-457): softmax, the reference's SECOND
 * softmax inside CrossEntropyLoss (mn:364-366), class-weighted CE, dZ, and all factor
 * gradients incl. the class factor.  y: int64 (N) in [0,C); class_w: (C) of dtype.  P may be NULL. */
int tr_fwd_grad_mn(tr_handle* h, const void* X, const int64_t* y, const void* class_w, int64_t N,
                   const void* theta, const void* w, uint32_t nn_mask, double sp_beta, double sp_thr,
                   double* gradsum, void* P, void* stream);

/* Vector-Jacobian product of lin_model for an arbitrary upstream gradient dyhat (N) — what
 * autograd's MmBackward0 + the cp_to_tensor backward compute (std:372,462).  Writes
 * gradsum = [ dFt | sum_n dyhat_n | 0 ]. */
int tr_backward_std(tr_handle* h, const void* X, const void* dyhat, int64_t N, const void* theta,
                    const void* w, uint32_t nn_mask, double sp_beta, double sp_thr,
                    double* gradsum, void* stream);

/* Vector-Jacobian product of model (mn:148-187) for an arbitrary upstream gradient dP (N,C) wrt the
 * softmax probabilities: recomputes the forward pass, dZ = P*(dP - sum_c dP*P), then all factor
 * gradients incl. the class factor.  gradsum = [ dFt | 0 ]. */
int tr_backward_mn(tr_handle* h, const void* X, const void* dP, int64_t N, const void* theta, const void* w,
                   uint32_t nn_mask, double sp_beta, double sp_thr, double* gradsum, void* stream);

/* ---- spectral_tensor_regression.py: X (T, W, D), y (T, n_out) ---------------------------------
 * The model of that file's fit / fit_Adam closures (spectral:573-586 / 727-731):
 *     yhat = lin_model(X, Bcp_n, weights[:rank_normal], ...) + stepwise_spectral_model(X, Bcp_c, ...)
 * lin_model (spectral:118-165) is the standard CP model with n_out outputs; stepwise_spectral_model
 * (spectral:339-390) contracts the window axis with a (W, rank_spectral, complex_dim) factor, takes the NORM over
 * the complex axis, then contracts D and maps to the outputs.  Loss = MSE over (T, n_out) (spectral:581).
 *
 * Parameter vector (the reference's optimizer parameter list Bcp_n + Bcp_c + [bias], laid end to end, row-major):
 *     theta = [ Fn0 (W,Rn) | Fn1 (D,Rn) | Fn2 (n_out,Rn) | Fc0 (W,Rs,CC) | Fc1 (D,Rs) | Fc2 (n_out,Rs) | bias (n_out) ]
 * P = Pf + n_out.  nn_mask bit i (i < 6) = softplus on the i-th block (the reference applies non_negative[m] to
 * block m of both lists: bits m and m + 3).  w = rank weights (Rn + Rs values; only the first Rn are used, as in the
 * reference).  gradsum = [ dFt (Pf) | nb * sum_t res[t,n] (n_out) | sum res^2 ]  (P + 1 doubles; nb = number of
 * non-empty parts, because both parts add the bias), to be summed across GPUs and passed to tr_finish_grad with
 * grad_scale = 2 / (T_total * n_out), loss_scale = 1 / (T_total * n_out).  tr_param_count, tr_gradsum_count,
 * tr_finish_grad, tr_adam_step, tr_lbfgs_*, tr_allreduce, tr_destroy work on these handles as on the others. */
int tr_spec_create(tr_handle** out, int dtype, int64_t W, int64_t D, int64_t n_out, int rank_normal,
                   int rank_spectral, int complex_dim, int device);

/* One closure evaluation without penalty and normalisation (spectral:573-586): two streaming passes over X.
 * y: (N, n_out) of dtype; yhat (N, n_out) may be NULL. */
int tr_spec_fwd_grad(tr_handle* h, const void* X, const void* y, int64_t N, const void* theta, const void* w,
                     uint32_t nn_mask, double sp_beta, double sp_thr, double* gradsum, void* yhat, void* stream);

/* Forward only; every output may be NULL:
 *   yhat     (N, n_out)  the model of the fit closures (above)
 *   yhat_lin (N, n_out)  lin_model alone, bias once                               (spectral:118-165)
 *   spec_pred (N, n_out) spectral_model (spectral:168-221): sqrt(sum_c z_c^2) + bias with z_c the complete CP
 *                        contraction of X with the c-th complex slice of Fc0 (rank weights w[Rn:] applied) — what the
 *                        reference's predict adds to lin_model (spectral:960-961); NOT the model of the fit
 *   latents  (N, Rn)     stepwise_latents_model: s_n without rank weights          (spectral:284-337, predict_latents) */
int tr_spec_forward(tr_handle* h, const void* X, int64_t N, const void* theta, const void* w, uint32_t nn_mask,
                    double sp_beta, double sp_thr, void* yhat, void* yhat_lin, void* spec_pred, void* latents,
                    void* stream);

/* gradsum (after the cross-GPU sum, if any) -> gradient wrt the RAW parameters and the losses:
 *   grad[F_m] = grad_scale * dFt_m * softplus'(F_m) + lambda_L2 * F_m / ||F_m||_F   (L2_penalty, std:180-196)
 *   grad[bias] = grad_scale * gradsum[Pf]            (standard only)
 *   loss[0] = loss_scale * gradsum[last]  (MSE or CE),  loss[1] = loss[0] + lambda_L2 * sum_m ||F_m||_F
 * grad_scale = 2/N_total, loss_scale = 1/N_total (standard); both 1/sum_n omega[y_n] (multinomial). */
int tr_finish_grad(tr_handle* h, const double* gradsum, double grad_scale, double loss_scale,
                   const void* theta, double lambda_L2, uint32_t nn_mask, double sp_beta, double sp_thr,
                   void* grad, double* loss, void* stream);

/* torch.optim.Adam step (std:453,463 / mn:447,458; torch/optim/adam.py single-tensor path) on
 * the flat parameter vector.  step is 1-based.  vmax may be NULL (amsgrad off). */
int tr_adam_step(tr_handle* h, void* theta, const void* grad, void* m, void* v, void* vmax,
                 int64_t step, double lr, double beta1, double beta2, double eps, double weight_decay,
                 void* stream);

/* Same step with one learning rate per parameter group, the groups being the factors of theta in order
 * (feature factors, class factor) plus, for the standard model, the bias: the three Adam parameter groups of
 * multinomial_tensor_regression_hierarchical.py (hier:436-440) are lr_groups = {lr, lr, lr}.  lr_groups is a
 * HOST array of n_groups doubles; n_groups must equal k + 1 (multinomial: k feature factors + class factor;
 * standard: k factors + bias). */
int tr_adam_step_groups(tr_handle* h, void* theta, const void* grad, void* m, void* v, void* vmax, int64_t step,
                        const double* lr_groups, int n_groups, double beta1, double beta2, double eps,
                        double weight_decay, void* stream);

/* Cross-GPU sum of `count` doubles in place (the packed gradsum, one call per closure evaluation):
 * ncclAllReduce(sum, float64) on `stream`.  nccl_comm is an ncclComm_t — either the host framework's (PyTorch:
 * ProcessGroupNCCL._comm_ptr()) or one made by tr_comm_create.  NCCL is resolved with dlopen at first use
 * (the copy already loaded in the process, else libnccl.so.2, else $TR_B200_NCCL_LIB); single-GPU hosts never
 * need it.  The reference has no multi-GPU path; this is the one collective of the sharded fit. */
int tr_allreduce(tr_handle* h, double* buf, int64_t count, void* nccl_comm, void* stream);

/* Communicator for hosts without a framework: rank 0 calls tr_comm_unique_id (fills 128 HOST bytes), ships them
 * to the other ranks by any means, then every rank calls tr_comm_create(&comm, id128, rank, world, device).
 * On failure the message is available from tr_last_error(NULL). */
int tr_comm_unique_id(void* id128);
int tr_comm_create(void** comm, const void* id128, int rank, int world, int device);
int tr_comm_destroy(void* comm);

/* Host array -> device memory (the reference's one `.to(device)` of X, std:339-345 / mn:255) at close to the
 * pinned-memory DMA rate even when src_host is PAGEABLE: a ring of pinned staging buffers is filled by `threads`
 * host threads (0 = min(hardware threads, 16); env TR_B200_UPLOAD_THREADS) in parallel slices while the previous
 * buffer is in flight; memory that is already pinned / registered is copied from in place.  chunk_bytes = size
 * of one staging buffer (0 = 32 MiB).  Work queued on `stream` before the call is waited for, work queued after
 * it sees the data; the call returns when src_host may be reused.  Errors: tr_host_last_error().
 * tr_upload_stats: out4 = { seconds of the last upload, of which host fill seconds, threads used, 1 if staged }. */
int tr_upload(void* dst_device, const void* src_host, size_t bytes, int device, int threads, size_t chunk_bytes,
              void* stream);
int tr_upload_stats(double* out4);
const char* tr_host_last_error(void);

/* L-BFGS building blocks (torch.optim.LBFGS as used by fit, std:366,392 / mn:355,381; algorithm of
 * torch/optim/lbfgs.py:333-536).  The update history, the two-loop recursion and every dot
 * product stay on the device; the host keeps only the strong-Wolfe control flow and reads back
 * the scalars it branches on.  Caller-owned state: S, Y (history x P, dtype), prev_g, d (P, dtype),
 * lstate (4 + history doubles, zero-initialised: H_diag, num_old, ring head, -, ro[history]).
 *
 * tr_lbfgs_direction: first != 0 -> d = -g and the history is cleared; otherwise y = g - prev_g,
 *   s = t*d are pushed when y.s > 1e-10 (H_diag = y.s / y.y) and d = two-loop(g).  Then prev_g = g
 *   and scal4 = { g.d, sum|g|, max|g|, max|d| } (device doubles).
 * tr_lbfgs_point: out = x + t*d (the trial point of a line-search evaluation / the accepted step).
 * tr_lbfgs_gtd: scal2 = { g.d, max|g| } (d may be NULL: only max|g|). */
int tr_lbfgs_direction(tr_handle* h, const void* g, void* prev_g, void* d, double t, int first, void* S, void* Y,
                       double* lstate, int history, double* scal4, void* stream);
int tr_lbfgs_point(tr_handle* h, void* out, const void* x, double t, const void* d, void* stream);
int tr_lbfgs_gtd(tr_handle* h, const void* g, const void* d, double* scal2, void* stream);

/* Optional timing of the two streaming kernels with CUDA events recorded on the launching stream
 * (what bench.py's roofline uses).  tr_profile_enable(h, 1) resets the sums; tr_profile_read waits
 * for the last recorded launch and returns out6 = { forward-pass ms total, forward launches,
 * gradient-pass ms total, gradient launches, fused single-pass ms total, fused launches }. */
int tr_profile_enable(tr_handle* h, int enable);
int tr_profile_read(tr_handle* h, double* out6);

/* Options.  "fused": -1 = auto (default; env TR_B200_FUSED overrides), 0 = always the two-pass
 * kernels, 1 = always the single-pass cluster kernel of tr_fwd_grad_std (error if the geometry
 * does not fit a cluster's shared memory).
 * "flow": 0 = never (default), 1 = run tr_fwd_grad_std / tr_fwd_grad_mn as ONE cooperative
 * dataflow kernel whose gradient warps re-read X from L2 a bounded window behind the forward warps
 * (experimental: correct, but measured slower than the two-pass kernels; error if the
 * geometry / alignment is not eligible); "flow_window_mb": size of that window in MiB (default 32).
 * "spec_single" (spectral handles): -1 = auto (default; env TR_B200_SPEC_SINGLE overrides: the single-pass
 * kernel from 4 x SMs samples when the geometry fits), 0 = never, 1 = always the single-pass kernel of
 * tr_spec_fwd_grad that keeps a ring of whole samples in shared memory and reads X once (error unless
 * 16-byte rows, D <= 32 lanes x 16 bytes, Q <= 8 channels, W <= 64 window rows and two samples fit);
 * "spec_single_ns": cap on its shared-memory stages (testing knob). */
int tr_set_option(tr_handle* h, const char* name, int64_t value);

/* How the last tr_fwd_grad_* / tr_forward_* call was executed (host ints):
 * info[0]=kernel launches, [1]=forward grid, [2]=gradient grid, [3]=tiles per sample,
 * [4]=sample groups (forward), [5]=sample groups (gradient), [6]=channels RK, [7]=vector width.
 * After a single-pass launch: [1]=grid, [2]=cluster size, [3]=stages, [4]=clusters, [5]=chunks,
 * [7]= -(vector width).  After a dataflow launch: [1]=grid, [2]=0, [3]=window in samples per group,
 * [4]=sample groups, [5]=chunks, [6]=channels, [7]= -(vector width). */
int tr_last_launch_info(tr_handle* h, int64_t* info8);

#ifdef __cplusplus
}
#endif
#endif /* TR_B200_H */
