/*
 * mogp_b200 -- C ABI of the B200-native exact multi-output GP engine.
 *
 * This is the drop-in boundary for the exact-GP hot path of GAMES-UChile/mogptk
 * (SURVEY.md section 8b).  The reference is pure Python on PyTorch and has no FFI
 * of its own; the entry points below are what a ctypes binding placed behind the
 * reference's `inference=` builder seam (mogptk/model.py:89-100,231) binds.  Each
 * entry point names the reference interface it replaces.
 *
 * Conventions
 *   - plain C, no torch types.  All matrices are fp64, row-major.
 *   - pointers named *_dev are DEVICE pointers on the handle's device, pointers
 *     named *_host are host pointers.  The library never frees or keeps caller
 *     buffers past the call; scratch lives in the handle.
 *   - rows of x / y are sorted by channel: channel c owns rows
 *     [chan_off[c], chan_off[c+1]) ; chan_off has C+1 entries and lives on the HOST.
 *     x holds the input coordinates only (N x D), not the channel-id column.
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*).
 *   - return value 0 = ok, <0 = error (text via mogp_last_error).  Cholesky
 *     failure is NOT an error return: it is reported LAPACK-style as info = k > 0
 *     ("leading minor k not positive definite") in the output block, and the
 *     host raises mogptk.gpr.model.CholeskyException (gpr/model.py:71-78,255).
 *
 * Packed constrained hyper-parameters (`params`, length mogp_num_params()):
 *   MOSM : weight[C][Q] | mean[C][Q][D] | variance[C][Q][D] | delay[C][Q][D] | phase[C][Q]
 *          (mogptk/gpr/multioutput.py:156-171)
 *   SM   : magnitude[C][Q] | mean[C][Q][D] | variance[C][Q][D]
 *          (IndependentMultiOutputKernel of SpectralMixtureKernel,
 *           gpr/multioutput.py:5-39, gpr/singleoutput.py:583-592)
 *   CONV : weight[Q][C] | variance[Q][C][D] | base_variance[Q][D]
 *          (MixtureKernel of GaussianConvolutionProcessKernel,
 *           gpr/kernel.py:264-276, gpr/multioutput.py:520-529)
 *   CSM  : amplitude[Q][C][Rq] | mean[Q][D] | variance[Q][D] | shift[Q][C][Rq]
 *          (MixtureKernel of CrossSpectralKernel, gpr/multioutput.py:397-454; what mogptk.CSM builds)
 *   SMLMC: weight[C][Q][Rq] | magnitude[Q] | mean[Q][D] | variance[Q][D]
 *          (LinearModelOfCoregionalizationKernel of SpectralKernel, gpr/multioutput.py:456-502,
 *           gpr/singleoutput.py:520-561; what mogptk.SM_LMC builds)
 *   UMOSM: weight[Q][C][C] (lower triangle used) | mean[Q][C][D] | variance[Q][C][D] | delay[Q][C][D] | phase[Q][C]
 *          (MixtureKernel of UncoupledMultiOutputSpectralKernel, gpr/multioutput.py:212-293)
 *   MOHSM: weight[Q][C] | mean[Q][C][D] | variance[Q][C][D] | lengthscale[Q][C] | center[Q][D] | delay[Q][C][D] | phase[Q][C]
 *          (MixtureKernel of MultiOutputHarmonizableSpectralKernel, gpr/multioutput.py:295-395; what mogptk.MOHSM builds with
 *           Q = P*Q of that model).  Non-stationary: every term carries a Gaussian window in the mid-point (x + x')/2, so the
 *           Gram diagonal depends on the row and K_diag needs the inputs: use mogp_kdiag_x.
 */
#ifndef MOGP_B200_H
#define MOGP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mogp_handle_s* mogp_handle_t;

enum { MOGP_KIND_MOSM = 0, MOGP_KIND_SM = 1, MOGP_KIND_CONV = 2, MOGP_KIND_CSM = 3, MOGP_KIND_SMLMC = 4, MOGP_KIND_UMOSM = 5,
       MOGP_KIND_MOHSM = 6 };
/* CSM and SM-LMC have Rq sub-components per mixture term: pass kind = MOGP_KIND_WITH_RQ(MOGP_KIND_CSM, Rq)
 * (family in the low 8 bits, Rq above; Rq = 0 means 1).  Every `kind` argument below accepts this form. */
#define MOGP_KIND_WITH_RQ(kind, Rq) ((kind) | ((Rq) << 8))
enum { MOGP_MAX_D = 8 };

/* library / ABI version (major*1000 + minor) */
int mogp_version(void);

/* Number of packed constrained kernel parameters for (kind, C, Q, D); <0 on bad args. */
int mogp_num_params(int kind, int C, int Q, int D);

/* Workspace for problems of up to max_n rows on CUDA device `device`.
 * Replaces nothing in the reference (torch's caching allocator plays this role). */
int mogp_create(int device, int64_t max_n, mogp_handle_t* out);
int mogp_destroy(mogp_handle_t h);
const char* mogp_last_error(mogp_handle_t h);

/* K(X1, X2) -- replaces MultiOutputKernel.K (gpr/kernel.py:446-481) with the Ksub of
 * MOSM (gpr/multioutput.py:178-204), SM (gpr/singleoutput.py:594-600 under
 * gpr/multioutput.py:26-34) or CONV (gpr/multioutput.py:531-547 under gpr/kernel.py:242-243).
 *   x2_dev == NULL  => Gram matrix of x1 (n2 = n1, chan_off2 ignored), full symmetric output.
 *   noise_sigma_dev (C) / data_var_dev (n1) / jitter_rel are only used for the Gram
 *   matrix: diag += sigma_c^2 + data_var_r, then diag += jitter_rel * mean(diag)
 *   (gpr/model.py:440-442,244).  Pass NULL / 0.0 for the bare kernel matrix.
 *   K_dev is n1 x n2 with leading dimension ldk (elements). */
int mogp_kbuild(mogp_handle_t h, int kind, int C, int Q, int D, const double* params_dev,
                const double* x1_dev, const int32_t* chan_off1_host,
                const double* x2_dev, const int32_t* chan_off2_host,
                const double* noise_sigma_dev, const double* data_var_dev, double jitter_rel,
                double* K_dev, int64_t ldk, void* stream);

/* K_diag(X) -- replaces MultiOutputKernel.K_diag / Ksub_diag (gpr/kernel.py:483-495,
 * gpr/multioutput.py:36-39,206-210,549-553).  out_dev has n entries. */
int mogp_kdiag(mogp_handle_t h, int kind, int C, int Q, int D, const double* params_dev,
               const int32_t* chan_off_host, double* out_dev, void* stream);
/* The same with the n x D inputs (sorted by channel): required for MOHSM (gpr/multioutput.py:389-395: the prior variance
 * depends on x), accepted for every kind. */
int mogp_kdiag_x(mogp_handle_t h, int kind, int C, int Q, int D, const double* params_dev, const double* x_dev,
                 const int32_t* chan_off_host, double* out_dev, void* stream);

/* In-place lower Cholesky A = L L^T of an n x n row-major matrix (only the lower triangle
 * is read; on return the lower triangle holds L, the strict upper triangle is unspecified).
 * Replaces torch.linalg.cholesky at gpr/model.py:246.  *info_dev (device int32) = 0 or the
 * 1-based index of the first non-positive pivot. */
int mogp_potrf(mogp_handle_t h, double* A_dev, int64_t n, int64_t lda, int32_t* info_dev, void* stream);

/* One exact-GP evaluation -- replaces gpr.Exact.log_marginal_likelihood
 * (gpr/model.py:438-453) and, with want_grad, the autograd backward of gpr.Model.loss
 * (gpr/model.py:279-292) in constrained-parameter space.
 *   y_dev: N targets with any mean function already subtracted (gpr/model.py:445-448).
 *   noise_sigma_dev: C Gaussian-likelihood scales (gpr/likelihood.py:326-330).
 *   out_dev: 2 + P + C doubles:
 *     out[0]      log marginal likelihood
 *     out[1]      info (0, or k>0: leading minor k not positive definite)
 *     out[2..2+P) d(-LML)/d params   (packed layout above)          [want_grad only]
 *     out[2+P..)  d(-LML)/d noise_sigma_c                            [want_grad only]
 * The factor (L^-1, alpha) stays in the handle for mogp_predict. */
int mogp_lml_grad(mogp_handle_t h, int kind, int C, int Q, int D, const double* params_dev,
                  const double* x_dev, const int32_t* chan_off_host, const double* y_dev,
                  const double* noise_sigma_dev, const double* data_var_dev, double jitter_rel,
                  int want_grad, double* out_dev, void* stream);

/* Same evaluation with HOST buffers: copies inputs host->device, runs, copies the output
 * block back and synchronises.  This is the call `bench.py`'s end-to-end leg times. */
int mogp_lml_grad_host(mogp_handle_t h, int kind, int C, int Q, int D, const double* params_host,
                       const double* x_host, const int32_t* chan_off_host, const double* y_host,
                       const double* noise_sigma_host, const double* data_var_host, double jitter_rel,
                       int want_grad, double* out_host);

/* Posterior of f at M test rows using the factor of the last mogp_lml_grad call --
 * replaces gpr.Exact.predict_f (gpr/model.py:455-483), without re-factorising.
 *   xs_dev: M x D test inputs sorted by channel, chan_off_s_host: C+1 offsets.
 *   mu_dev: M means.  var_dev: M variances, or (full != 0) the M x M covariance, row-major. */
int mogp_predict(mogp_handle_t h, const double* xs_dev, const int32_t* chan_off_s_host,
                 int full, double* mu_dev, double* var_dev, void* stream);

/* alpha = K~^-1 (y - mean) of the last mogp_lml_grad call, N doubles in the caller's (channel-sorted) row order.
 * d LML / d y = -alpha: lets the host chain the gradient into a trainable mean function, which the reference gets
 * from autograd through y - mean(X) (gpr/model.py:445-452). */
int mogp_alpha(mogp_handle_t h, double* alpha_dev, void* stream);

/* Early loss.  gpr.Model.loss() (gpr/model.py:279-292) returns the loss VALUE (and raises CholeskyException); its gradients are
 * consumed by the optimiser afterwards.  With on = 1 every following mogp_lml_grad / mogp_loss_grad on this handle writes
 * host_buf[0..2] = {lml, info, seq} into mapped pinned host memory as soon as the solves are done -- before K^-1, the gradient
 * reduction and the chain rule -- where seq counts the evaluations since the switch.  The caller polls host_buf[2] for
 * mogp_early_expected(h) instead of synchronising the stream; the gradient outputs are complete in stream order. */
int mogp_early_loss(mogp_handle_t h, int on, double** host_buf);
unsigned long long mogp_early_expected(mogp_handle_t h);

/* ---- constrained parameters on the device ------------------------------------------------
 * Replaces Parameter.constrained / Softplus.forward / Sigmoid.forward (gpr/parameter.py:30-96,186-201) and
 * their autograd backward for the training loop (mogptk/model.py:563-565): forward maps the raw leaves to the
 * packed constrained vector that mogp_lml_grad consumes (kernel parameters, then the C noise scales) and
 * records d constrained / d raw; backward multiplies the gradient block of mogp_lml_grad by it and writes the
 * raw-space gradients into the p.grad buffers.  Entries live in HOST memory; pointers inside are DEVICE. */
typedef struct {
    const double* raw;     /* n raw (unconstrained) values */
    double* grad;          /* n raw-space gradients to fill, or NULL */
    const double* lower;   /* lower bound(s): lower_n == 1 (broadcast) or n values; NULL for type 0 */
    const double* upper;   /* upper bound(s), type 2 only */
    int64_t n;             /* elements */
    int64_t off;           /* offset of this parameter in the packed vector */
    int32_t type;          /* 0: identity, 1: Softplus(lower, beta), 2: Sigmoid(lower, upper) */
    int32_t lower_n, upper_n, pad;
    double beta;           /* Softplus slope (reference default 0.1; -0.1 for an upper-only bound) */
} mogp_param_entry;

int mogp_params_forward(mogp_handle_t h, const mogp_param_entry* entries_host, int n_entries,
                        double* packed_dev, double* dcons_dev, void* stream);
/* gcons_dev: d(-LML)/d constrained in packed order (= out + 2 of mogp_lml_grad); lml_dev: out of mogp_lml_grad;
 * loss_out_dev (optional): receives -lml. */
int mogp_params_backward(mogp_handle_t h, const mogp_param_entry* entries_host, int n_entries,
                         const double* gcons_dev, const double* dcons_dev, const double* lml_dev,
                         double* loss_out_dev, void* stream);

/* forward + mogp_lml_grad + backward in one call -- replaces gpr.Model.loss (gpr/model.py:279-292) for a model whose
 * parameters all live in `entries_host`.  work_dev: 3 * (2 + P + C) doubles: packed constrained values | d constrained /
 * d raw | [lml, info, d(-LML)/d constrained].  loss_out_dev (optional): -lml. */
int mogp_loss_grad(mogp_handle_t h, int kind, int C, int Q, int D, const mogp_param_entry* entries_host, int n_entries,
                   const double* x_dev, const int32_t* chan_off_host, const double* y_dev, const double* data_var_dev,
                   double jitter_rel, double* work_dev, double* loss_out_dev, void* stream);

/* Device-resident Adam training -- replaces `iters` passes of the body of mogptk.Model.train's loop
 * (mogptk/model.py:563-565: loss = gpr.loss(); torch.optim.Adam.step()) without any host synchronisation.
 *   entries_host: the raw leaves, as for mogp_params_forward (kernel parameters first, the C noise scales last); the
 *                 raw values are updated IN PLACE, the raw-space gradient of the last iteration is left in .grad.
 *   work_dev    : 3 * (2 + P + C) doubles of scratch.
 *   exp_avg_dev, exp_avg_sq_dev: Adam moments, P + C doubles each in packed order (zero them for a fresh optimiser).
 *   step0       : optimiser steps already taken (bias correction uses step0 + i + 1).
 *   losses_dev  : iters doubles; losses_dev[i] = -LML at the parameters BEFORE update i (what the reference records).
 *   fail_dev    : 2 int32, zeroed by the caller: [0] = LAPACK-style info of the first failed Cholesky (or -1 for a
 *                 non-finite evaluation), [1] = its 1-based iteration.  From that iteration on the parameters are
 *                 frozen, so the host can re-evaluate there and raise CholeskyException (gpr/model.py:246-255).
 * torch.optim.Adam semantics with amsgrad = False, weight_decay = 0, maximize = False. */
int mogp_train_adam(mogp_handle_t h, int kind, int C, int Q, int D, const mogp_param_entry* entries_host, int n_entries,
                    const double* x_dev, const int32_t* chan_off_host, const double* y_dev, const double* data_var_dev,
                    double jitter_rel, double* work_dev, double* exp_avg_dev, double* exp_avg_sq_dev, long long step0,
                    int iters, double lr, double beta1, double beta2, double eps, double* losses_dev, int32_t* fail_dev,
                    void* stream);

/* ---- building blocks exposed for tests and micro-benchmarks --------------------------- */

/* C = alpha * op(A) * op(B) + beta * C on the fp64 tensor pipe (DMMA).
 * transa/transb: 0 = operand stored (rows x k) / (k x cols) "N", 1 = transposed storage.
 * M, N multiples of 64, K multiple of 16; leading dimensions even. */
int mogp_dgemm(mogp_handle_t h, int transa, int transb, int M, int N, int K, double alpha,
               const double* A_dev, int64_t lda, const double* B_dev, int64_t ldb,
               double beta, double* C_dev, int64_t ldc, void* stream);

/* Inverse of the lower Cholesky factor and (L L^T)^-1 (lower triangle), n multiple of 128,
 * all matrices n x n with leading dimension n: used by tests of the inverse path. */
int mogp_trtri_kinv(mogp_handle_t h, double* A_dev /* in: K, out: L */, double* Linv_dev,
                    double* Kinv_dev, int64_t n, int32_t* info_dev, void* stream);

/* Device micro-benchmarks: sustained DMMA (mma.sync m8n8k4 f64) and DFMA rates, TFLOP/s. */
int mogp_peak_fp64(mogp_handle_t h, double* dmma_tflops_host, double* dfma_tflops_host);

#ifdef __cplusplus
}
#endif
#endif /* MOGP_B200_H */
