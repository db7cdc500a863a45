/*
 * iwvi_b200.h -- C ABI of the B200 (sm_100a) IW-ELBO hot path.
 *
 * Drop-in boundary for the float64 importance-weighted ELBO forward/backward of
 * hughsalimbeni/DGPs_with_IWVI.  Each entry point names the reference interface it replaces
 * (paths relative to the reference repository root).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to a contiguous row-major float64 buffer unless stated;
 *   - the caller (PyTorch) allocates every input, output, context, save and workspace buffer; the
 *     library never allocates or frees device memory and keeps no state between calls;
 *   - all launches are asynchronous on the cudaStream_t passed as `stream` (void* here so that the
 *     header is plain C); no host synchronisation happens inside the library;
 *   - return value: 0 success; IWVI_ERR_* (negative) for a bad descriptor / unsupported size /
 *     launch failure.  Numerical failure of the Cholesky (non-positive pivot) is reported
 *     LAPACK-style through the device-side `info` word written by iwvi_gp_prologue_fwd
 *     (0 = ok, i>0 = leading minor of order i is not positive definite);
 *   - "points" are the T = B*K rows the reference calls [N,K] (IW, models.py:113) or [S*N] (VI,
 *     models.py:50) or [S,N] (prediction, models.py:96), flattened row-major.
 */
#ifndef IWVI_B200_H
#define IWVI_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)   /* the library is built with -fvisibility=hidden; these are its exports */
#endif

#define IWVI_VERSION 100

/* limits of this build */
#define IWVI_MAX_M 512   /* inducing points per layer                 */
#define IWVI_MAX_D 32    /* GP layer input width                       */
#define IWVI_MAX_R 8     /* latent GPs per layer (num_outputs)        */
#define IWVI_MAX_P 32    /* GP layer output width                      */
#define IWVI_MAX_ENC_LAYERS 8
#define IWVI_MAX_ENC_WIDTH 64
#define IWVI_MAX_LW 8

/* error codes */
#define IWVI_OK 0
#define IWVI_ERR_BAD_DESC   (-1)
#define IWVI_ERR_UNSUPPORTED (-2)
#define IWVI_ERR_LAUNCH     (-3)
#define IWVI_ERR_NULL       (-4)

/* kernels: gpflow.kernels.RBF / Matern12 / Matern32 / Matern52 (call sites temp_workaround.py:39,44,45) */
#define IWVI_KERN_RBF      0
#define IWVI_KERN_MATERN12 1
#define IWVI_KERN_MATERN32 2
#define IWVI_KERN_MATERN52 3

/* mean functions: gpflow.mean_functions.Zero / Identity / Linear (call site layers.py:46) */
#define IWVI_MF_ZERO     0
#define IWVI_MF_IDENTITY 1
#define IWVI_MF_LINEAR   2

/* encoder non-linearities: tf.nn.tanh (the reference default, layers.py:122) / relu / sigmoid / softplus / elu / identity */
#define IWVI_ACT_TANH     0
#define IWVI_ACT_RELU     1
#define IWVI_ACT_SIGMOID  2
#define IWVI_ACT_SOFTPLUS 3
#define IWVI_ACT_ELU      4
#define IWVI_ACT_IDENTITY 5

/* flags */
#define IWVI_FLAG_SAMPLE 1  /* eps given: produce sample = mean + eps*sqrt(var) (temp_workaround.py:89-91) */
#define IWVI_FLAG_SAVE   2  /* keep A, U, latent mean/var for the backward pass */
#define IWVI_FLAG_ACCUM  4  /* iwvi_gp_prologue_bwd adds into its outputs instead of overwriting them */
/* iwvi_gp_rows_bwd is a five-launch sequence (epilogue adjoint, tile kernel, gram adjoint, split-K reduce, finalize).
 * When any of these is set only the selected launches run, on the buffers the earlier ones have filled: used to time
 * each alone, and by the host to run EPI|TILE|GRAM (dX, Bbar, per-CTA partials) on one stream and REDUCE|FINAL (every
 * parameter gradient) on another, so that the reductions of one layer overlap with the per-point kernels of the layer
 * below.  REDUCE needs TILE's output only; FINAL needs GRAM's and REDUCE's. */
#define IWVI_FLAG_ONLY_GRAM   8
#define IWVI_FLAG_ONLY_EPI    16
#define IWVI_FLAG_ONLY_TILE   32
#define IWVI_FLAG_ONLY_REDUCE 64
#define IWVI_FLAG_ONLY_FINAL  128
#define IWVI_FLAG_ONLY_MASK   (8 | 16 | 32 | 64 | 128)
/* The reduce + finalize launches in two halves, so that iwvi_gp_prologue_bwd (needs dLm, dZ, dls, dvariance only) can
 * overlap with the second one on another stream:  PART_A: dLm, dZ, dls, dvariance, dW, dmfA, dmfb;  PART_B: dq_mu, dq_sqrt.
 * Neither flag: everything.  With the split, run iwvi_gp_prologue_bwd with IWVI_FLAG_SKIP_KL after part A and with
 * IWVI_FLAG_ONLY_KL (the whitened-KL adjoint, which adds into dq_mu / dq_sqrt) after part B.
 * ORDER the two halves: launch part B's REDUCE only after part A's REDUCE + FINAL launches have completed (an event
 * between the two streams; iwvi_gp_prologue_bwd of part A may still overlap part B).  With the two REDUCE launches in
 * flight together part B's sums were observed to come out wrong by about one point's contribution on cold buffers
 * (DESIGN.md section 8); the memory they touch is disjoint, the cause is not understood, the order costs nothing
 * measurable (part A is one short wave). */
#define IWVI_FLAG_PART_A   256
#define IWVI_FLAG_PART_B   512
#define IWVI_FLAG_SKIP_KL  1024
#define IWVI_FLAG_ONLY_KL  2048
/* iwvi_gp_prologue_fwd in two halves, for callers that update a layer's parameters in two steps (training.Trainer updates
 * Z / kernel parameters and q_mu / q_sqrt of the first GP layer at different moments of the backward pass, and prepares the
 * NEXT step's factorisation as soon as the former are final):  PRO_HYP: everything that depends on Z, lengthscales,
 * variance (scaled inducing inputs, constants, Cholesky factor and its block copies; resets info);  PRO_Q: everything that
 * depends on q_mu, q_sqrt (padded copies, blocks of tril(q_sqrt), the KL).  Neither flag: both. */
#define IWVI_FLAG_PRO_HYP 16384
#define IWVI_FLAG_PRO_Q   32768
/* iwvi_gp_rows_bwd (REDUCE / FINAL launches): the per-point half ran as two point chains (iwvi_gp_rows_bwd_range): sum the
 * per-CTA partials of both. */
#define IWVI_FLAG_TWO_CHAINS 8192
/* OPTIONAL reduced-precision fast path, NOT the float64 parity path: iwvi_gp_rows_bwd forms the parameter contractions over the
 * points (dq_sqrt, dLm, dq_mu) on the tcgen05 tensor cores -- every float64 operand split into two TF32 numbers, three
 * TF32 products per term, FP32 accumulation in tensor memory, float64 sums of the split-K partials.  Relative error about
 * 1e-6 of the largest entry of each gradient; everything else (ELBO, dX, dZ, kernel parameters) is unchanged.  Needs M
 * padded to a multiple of 128; otherwise the flag is ignored.  Pass it with the REDUCE and FINAL launches alike. */
#define IWVI_FLAG_FAST_REDUCE 65536
/* iwvi_gp_rows_bwd: leave the Kdiag term (d var / d variance = 1 per point) out of dvariance -- set by callers that
 * differentiate the prior covariance k(X, X) themselves (iwvi_gp_fullcov_bwd). */
#define IWVI_FLAG_NO_KDIAG 4096

typedef struct iwvi_gp_desc {
  int32_t T;      /* points in this call                                    */
  int32_t M;      /* inducing points (len(Z), layers.py:19)                 */
  int32_t D;      /* input width D_in                                       */
  int32_t R;      /* latent GPs = num_outputs (layers.py:16)                */
  int32_t P;      /* output width: rows of the Mok mixing W, else == R      */
  int32_t kern;   /* IWVI_KERN_*                                            */
  int32_t mix;    /* 1: SharedMixedMok mixing W [P,R] (temp_workaround.py:142-145) */
  int32_t mf;     /* IWVI_MF_*                                              */
  int32_t flags;  /* IWVI_FLAG_*                                            */
  int32_t reserved;
  double  jitter; /* gpflow settings.numerics.jitter_level (temp_workaround.py:39) */
} iwvi_gp_desc;

int iwvi_version(void);
/* Name of the CUDA runtime error behind the calling thread's most recent IWVI_ERR_LAUNCH ("cudaSuccess" if none yet). */
const char* iwvi_last_cuda_error(void);

/* derived sizes (host-side helpers, no device work) */
int32_t iwvi_gp_mp(int32_t M);                         /* M padded to a multiple of 64                 */
int32_t iwvi_gp_lda(int32_t M);                        /* leading dimension of the point-major panels  */
int64_t iwvi_gp_aux_doubles(const iwvi_gp_desc* d);    /* size of `aux`  (doubles)                     */
int64_t iwvi_gp_save_doubles(const iwvi_gp_desc* d);   /* size of `save` (doubles) for d->T points     */
int64_t iwvi_gp_bwd_ws_doubles(const iwvi_gp_desc* d); /* workspace of iwvi_gp_rows_bwd (doubles)      */
int64_t iwvi_gp_pbwd_ws_doubles(const iwvi_gp_desc* d);/* workspace of iwvi_gp_prologue_bwd (doubles)  */

/*
 * Once per GP layer per step.  Replaces temp_workaround.py:39 (Kuu + jitter), :48 (tf.cholesky) and
 * layers.py:44 -> temp_workaround.py:167-188 -> gpflow gauss_kl (whitened KL).
 *   Z [M,D], ls [D] (ARD lengthscales, constrained), variance [1], q_mu [M,R], q_sqrt [R,M,M]
 *   out: Lm [Mp,Mp] lower Cholesky factor (identity on the padding), aux (opaque: inverted diagonal
 *        blocks, padded tril(q_sqrt), scaled inducing inputs, padded q_mu, constants), kl [1], info [1] (int32).
 */
int iwvi_gp_prologue_fwd(const iwvi_gp_desc* d, const double* Z, const double* ls, const double* variance,
                         const double* q_mu, const double* q_sqrt,
                         double* Lm, double* aux, double* kl, int32_t* info, void* stream);

/*
 * Per-point stage.  Replaces independent_multisample_sample_conditional (temp_workaround.py:44-91, the
 * full_cov=False branch the Mok path forces at :125-129 and whose diagonal is all that models.py:133
 * keeps), the Mok mixing (:142-145) and the mean function add (layers.py:46-48).
 *   X [T,D]; W [P,R] or NULL; mfA [D,P], mfb [P] (Linear) or NULL; eps [T,R] or NULL
 *   out: sample [T,P] (NULL unless IWVI_FLAG_SAMPLE), mean [T,P], var [T,P]; save (IWVI_FLAG_SAVE).
 */
int iwvi_gp_rows_fwd(const iwvi_gp_desc* d, const double* Lm, const double* aux, const double* X,
                     const double* W, const double* mfA, const double* mfb, const double* eps,
                     double* sample, double* mean, double* var, double* save, void* stream);

/*
 * The same stage over the points [point_begin, point_end) only (all pointers and the descriptor are those of the whole
 * call; point_begin a multiple of iwvi_gp_tile_points(d), point_end too unless it is d->T).  Points are independent
 * through the whole layer chain, so a caller can run the chain of disjoint ranges on different streams -- the engine
 * splits a minibatch into its full waves of tiles and the remainder, which fills the SMs the last wave leaves idle.
 */
int iwvi_gp_rows_fwd_range(const iwvi_gp_desc* d, const double* Lm, const double* aux, const double* X,
                           const double* W, const double* mfA, const double* mfb, const double* eps,
                           double* sample, double* mean, double* var, double* save,
                           int64_t point_begin, int64_t point_end, void* stream);
int iwvi_gp_tile_points(const iwvi_gp_desc* d);        /* points per tile (64 or 32) the row kernels use for d */

/*
 * Adjoint of iwvi_gp_rows_fwd (the reference: tf.gradients through temp_workaround.py:44-91,142-145,
 * layers.py:46-48; formulas in DESIGN.md).  Cotangents d_sample/d_mean/d_var [T,P] may be NULL.
 *   out (overwritten): dX [T,D], dZ [M,D], dls [D], dvariance [1], dq_mu [M,R], dq_sqrt [R,M,M] (lower),
 *        dLm [Mp,Mp] (lower), dW [P,R] (if mix), dmfA [D,P], dmfb [P] (if Linear).
 */
int iwvi_gp_rows_bwd(const iwvi_gp_desc* d, const double* Lm, const double* aux, const double* save,
                     const double* X, const double* W, const double* mfA, const double* mfb, const double* eps,
                     const double* d_sample, const double* d_mean, const double* d_var,
                     double* dX, double* dZ, double* dls, double* dvariance, double* dq_mu, double* dq_sqrt,
                     double* dLm, double* dW, double* dmfA, double* dmfb, double* ws, void* stream);

/*
 * The per-point half of iwvi_gp_rows_bwd (flags: any of IWVI_FLAG_ONLY_EPI / _TILE / _GRAM) over one of TWO point
 * chains: [0, point_end) or [point_begin, T), split on a multiple of iwvi_gp_bwd_tile_points(d).  Points are independent
 * through the whole backward chain of GP layers as they are through the forward one, so the caller can run the chain of
 * the full waves of tiles and the chain of the remainder on two streams (the remainder, at most one tile per SM, fills
 * the SMs the last wave leaves idle); the parameter half (REDUCE / FINAL, which contracts over ALL points) then runs
 * once both have finished, with IWVI_FLAG_TWO_CHAINS.  All pointers and the descriptor are those of the whole call.
 */
int iwvi_gp_rows_bwd_range(const iwvi_gp_desc* d, const double* Lm, const double* aux, const double* save,
                           const double* X, const double* W, const double* mfA, const double* mfb, const double* eps,
                           const double* d_sample, const double* d_mean, const double* d_var,
                           double* dX, double* dZ, double* dls, double* dvariance, double* dq_mu, double* dq_sqrt,
                           double* dLm, double* dW, double* dmfA, double* dmfb, double* ws,
                           int64_t point_begin, int64_t point_end, void* stream);
int iwvi_gp_bwd_tile_points(const iwvi_gp_desc* d);    /* points per tile of the backward tile kernel for d */

/*
 * Adjoint of iwvi_gp_prologue_fwd: Cholesky adjoint (TF CholeskyGrad), gram adjoint of Kuu and the KL
 * adjoint.  dLm [Mp,Mp] lower, dkl [1] (device scalar cotangent of kl).
 *   out (overwritten, or added to under IWVI_FLAG_ACCUM so that the outputs of iwvi_gp_rows_bwd can be passed
 *        straight in): dZ [M,D], dls [D], dvariance [1], dq_mu [M,R], dq_sqrt [R,M,M].
 */
int iwvi_gp_prologue_bwd(const iwvi_gp_desc* d, const double* Lm, const double* aux, const double* Z,
                         const double* ls, const double* variance, const double* q_mu, const double* q_sqrt,
                         const double* dLm, const double* dkl,
                         double* dZ, double* dls, double* dvariance, double* dq_mu, double* dq_sqrt,
                         double* ws, void* stream);

/*
 * Covariance over the inner axis and the joint draw, forward, for plain-kernel layers: the full_cov=True branch of
 * independent_multisample_sample_conditional (temp_workaround.py:45, :55-57, :82-83) and the joint sampler intended at
 * :92-96 (the reference's own lines add an [S,N,R] mean to an [S,R,N,1] draw and never execute).  Runs after
 * iwvi_gp_rows_fwd with IWVI_FLAG_SAVE on the same descriptor (d->T == S*N, d->mix == 0; the draw serves N <= 256, the
 * covariance any N).  ws: iwvi_gp_fullcov_ws_doubles(d, S, N) doubles (0 for N <= 64: may be NULL):
 *   X [S*N,D]; save (A, U_r panels); mean [S*N,R] as written by iwvi_gp_rows_fwd; eps [S,R,N] (the [S,R,N,1] draw of :94)
 *   out: cov [S,R,N,N] = k(X_s,X_s) - A_s^T A_s + U_rs^T U_rs (or NULL);
 *        sample [S*N,R] = mean + chol(cov + chol_jitter I) eps (or NULL); info [1] (int32, first failing leading minor).
 */
int iwvi_gp_fullcov_fwd(const iwvi_gp_desc* d, int32_t S, int32_t N, const double* aux, const double* X,
                        const double* save, const double* mean, const double* eps, double chol_jitter,
                        double* cov, double* sample, int32_t* info, double* ws, void* stream);
int64_t iwvi_gp_fullcov_ws_doubles(const iwvi_gp_desc* d, int32_t S, int32_t N);   /* workspace of _fwd / _bwd (doubles) */

/*
 * Adjoint of iwvi_gp_fullcov_fwd (tf.gradients through temp_workaround.py:45,55-57,82-83 and the joint draw; Cholesky
 * adjoint as in TF's CholeskyGrad, symmetrised), N <= 256 (ws as for the forward call).  Cotangents d_sample [S*N,R] and d_cov [S,R,N,N] (either may
 * be NULL).  It prepares what iwvi_gp_rows_bwd needs to finish the job with its existing kernels:
 *   save2 (same size as save, zero-initialised by the caller): save2.A = A_s (sum_r H_r) / R, save2.U_r = U_rs H_r, with
 *         H_r the symmetric N x N cotangent of C_r.  Run iwvi_gp_rows_bwd(IWVI_FLAG_ONLY_EPI | ONLY_TILE | ONLY_GRAM | NO_KDIAG) on
 *         save2 with d_mean = d_mean + d_sample, d_var = ones [T,R], d_sample = NULL; then copy save.A over save2.A
 *         and run the ONLY_REDUCE | ONLY_FINAL half on save2.
 *   dX_knn [S*N,D] and part [S,40] (dls partials at 0..D-1, dvariance partial at 32): the adjoint of k(X_s, X_s); the
 *         caller adds dX_knn to dX and the column sums of part to dls / dvariance.
 */
int iwvi_gp_fullcov_bwd(const iwvi_gp_desc* d, int32_t S, int32_t N, const double* aux, const double* X,
                        const double* save, const double* eps, double chol_jitter, const double* d_sample,
                        const double* d_cov, double* save2, double* dX_knn, double* part, double* ws, void* stream);

/*
 * Whitened KL[q(u)||p(u)] on its own, for callers of the operator-level gauss_kl(q_mu, q_sqrt) (temp_workaround.py:167-188
 * with K=None -> gpflow gauss_kl): kl = 0.5 (sum q_mu^2 - M R - sum log diag(Lq)^2 + sum Lq^2), Lq = tril(q_sqrt).
 * iwvi_gp_prologue_fwd returns the same number as a by-product.  dkl [1] device scalar cotangent; outputs overwritten.
 */
int iwvi_gauss_kl_fwd(int32_t M, int32_t R, const double* q_mu, const double* q_sqrt, double* kl, void* stream);
int iwvi_gauss_kl_bwd(int32_t M, int32_t R, const double* q_mu, const double* q_sqrt, const double* dkl,
                      double* dq_mu, double* dq_sqrt, void* stream);

/* ---- LatentVariableLayer + Encoder (layers.py:72-105, :137-152) ---- */
typedef struct iwvi_lv_desc {
  int32_t Be;        /* distinct encoder rows                                              */
  int32_t Kt;        /* each row serves Kt consecutive points (point = n*Kt + k); 1 = none  */
  int32_t Df;        /* width of F                                                         */
  int32_t Dxy;       /* encoder input width (XY_dim, layers.py:55)                          */
  int32_t Lw;        /* latent_dim                                                         */
  int32_t n_layers;  /* encoder layers = len(network_dims)+1 (layers.py:122)               */
  int32_t dims[IWVI_MAX_ENC_LAYERS + 1]; /* layer_dims (layers.py:122)                     */
  int32_t sampled;   /* 1: log q(W) - log p(W) per sample (layers.py:98-100); 0: closed-form KL (:103) */
  int32_t f_bcast;   /* 1: F is [Be,Df], broadcast over Kt; 0: F is [Be*Kt,Df]             */
  int32_t prior;     /* 1: no encoder, q_mu/q_sqrt = prior_mu/prior_sigma (layers.py:73-81) */
  int32_t act;       /* IWVI_ACT_*: Encoder(activation_func=...) between the layers (layers.py:109,122,144; default tanh) */
  int32_t reserved;
  double  prior_mu, prior_sigma;
} iwvi_lv_desc;

int64_t iwvi_lv_param_doubles(const iwvi_lv_desc* d);  /* packed W0,b0,W1,b1,...                */
int64_t iwvi_lv_bwd_ws_doubles(const iwvi_lv_desc* d);

/* out: samples [T,Df+Lw] = [F, W], kl [T,Lw], mu [Be,Lw], sigma [Be,Lw] */
int iwvi_lv_fwd(const iwvi_lv_desc* d, const double* F, const double* enc_in, const double* params,
                const double* eps, double* samples, double* kl, double* mu, double* sigma, void* stream);
/* cotangents d_samples [T,Df+Lw], d_kl [T,Lw] (either may be NULL); d_mu/d_sigma [Be,Lw] extra cotangents or NULL.
 * out (overwritten): d_params (packed), dF ([Be,Df] if f_bcast else [T,Df]; may be NULL). */
int iwvi_lv_bwd(const iwvi_lv_desc* d, const double* F, const double* enc_in, const double* params,
                const double* eps, const double* mu, const double* sigma,
                const double* d_samples, const double* d_kl, const double* d_mu, const double* d_sigma,
                double* d_params, double* dF, double* ws, void* stream);

/* ---- likelihood + K-way logsumexp (models.py:133-150; VI: models.py:66-86) ---- */
typedef struct iwvi_elbo_desc {
  int32_t B;          /* minibatch rows                                                     */
  int32_t K;          /* importance samples / VI samples                                    */
  int32_t Dy;         /* output columns                                                     */
  int32_t Lw;         /* total local-regulariser columns (0: none)                          */
  int32_t iw;         /* 1: logsumexp_K - log K (models.py:148); 0: mean over K (models.py:84) */
  int32_t data_major; /* 1: point = n*K + k (models.py:113); 0: point = k*B + n (models.py:50) */
  double  scale;      /* num_data / B_global (models.py:144-145)                            */
} iwvi_elbo_desc;

int64_t iwvi_elbo_ws_doubles(const iwvi_elbo_desc* d);

/* out: elbo_data [1] = scale * sum_n logp_n, logp [B], w [B,K] (softmax over K, or 1/K) */
int iwvi_iwelbo_fwd(const iwvi_elbo_desc* d, const double* fmean, const double* fvar, const double* Y,
                    const double* lik_var, const double* kl_local,
                    double* elbo_data, double* logp, double* w, double* ws, void* stream);
/* d_elbo [1] device scalar. out: dmean, dvar [T,Dy], dkl_local [T,Lw] (or NULL), dlik [1] */
int iwvi_iwelbo_bwd(const iwvi_elbo_desc* d, const double* fmean, const double* fvar, const double* Y,
                    const double* lik_var, const double* w, const double* d_elbo,
                    double* dmean, double* dvar, double* dkl_local, double* dlik, double* ws, void* stream);

/* ---- counter-based N(0,1) noise, identical for any number of GPUs ----
 * Replaces tf.random_normal at layers.py:86 and temp_workaround.py:89 (index order [n,k,c] row-major).
 * element (point p, column c) of a [*, C] tensor uses Philox4x32-10 with key = (seed_lo, seed_hi) and
 * counter = ((first_point + p) * C + c) >> 1, lane (.. & 1) of a Box-Muller pair. */
int iwvi_normal_fill(double* out, int64_t n_points, int32_t C, int64_t first_point, uint64_t seed, void* stream);

/* The same with the step taken from DEVICE memory, so that a captured CUDA graph of the whole training step draws fresh
 * noise on every replay: state[0] = optimiser steps completed so far (advanced by iwvi_adam_step_counter);
 * seed = seed_base * 0x9E3779B97F4A7C15 + (state[0] + 1 + step_add) * 0xBF58476D1CE4E5B9 + (layer + 1) * 0x94D049BB133111EB. */
int iwvi_normal_fill_counter(double* out, int64_t n_points, int32_t C, int64_t first_point, uint64_t seed_base,
                             int32_t layer, int64_t step_add, const int64_t* state, void* stream);

/* ---- minibatch assembly (gpflow.params.Minibatch feeding X / Y, models.py:25-26; the [X, Y] concatenation of
 * models.py:53,116) ----
 * Rows idx[b] (int64 device indices; NULL: row b) of the resident data X [N,Dx], Y [N,Dy] into Xb [B,Dx], Yb [B,Dy] and
 * XYb [B,Dx+Dy] (any of the three outputs may be NULL) in one launch. */
int iwvi_batch_gather(const double* X, const double* Y, const int64_t* idx, int32_t B, int32_t Dx, int32_t Dy,
                      double* Xb, double* Yb, double* XYb, void* stream);

/* ---- exchange step of data-parallel training (the reference is single-device: its optimiser consumes
 * tf.gradients of the whole minibatch, build_models.py:293-295) as a one-shot all-reduce over NVLink peer memory ----
 * Every rank owns a receive buffer recv [2][world][bucket_len] doubles and a flag array [n_segments][world] of 64-bit
 * words, both ZERO-INITIALISED and mapped into every peer (e.g. torch.distributed._symmetric_memory); recv_ptrs /
 * flag_ptrs are HOST arrays [world] of this rank's device addresses of every rank's buffers (own one included).
 * epoch [n_segments] (64-bit, zero-initialised) and counters [2 * n_segments] (32-bit, zero-initialised) are plain
 * device memory of the calling rank.  A segment = n entries g[index[0 .. n-1]] (index: int64 device indices into the
 * gradient bucket g), stored at seg_off inside the packed bucket of bucket_len entries.
 *   iwvi_dp_push    stores this rank's segment into every peer's slot and publishes the segment's next epoch;
 *   iwvi_dp_reduce  (same stream, after the push) waits for all ranks' epochs, sums the slots in rank order (bit-identical
 *                   on every rank) into g[index[.]] and advances the epoch.
 * All ranks must issue the same sequence of (push, reduce) pairs per segment.  Both launches replay inside a CUDA graph. */
#define IWVI_DP_MAX_RANKS 16
int iwvi_dp_push(const double* g, const int64_t* index, int64_t n, int64_t seg_off, int64_t bucket_len, int32_t seg,
                 int32_t rank, int32_t world, const uint64_t* recv_ptrs, const uint64_t* flag_ptrs, const uint64_t* epoch,
                 uint32_t* counters, void* stream);
int iwvi_dp_reduce(double* g, const int64_t* index, int64_t n, int64_t seg_off, int64_t bucket_len, int32_t seg,
                   int32_t world, const double* recv_local, const uint64_t* flags_local, uint64_t* epoch,
                   uint32_t* counters, void* stream);

/* ---- optimiser step either side of the path (experiments/build_models.py:284-295) ----
 * gpflow.transforms.positive (Log1pe): theta = softplus(x) + 1e-6 for the first n entries. */
int iwvi_positive_fwd(const double* x, double* theta, int64_t n, void* stream);
/* tf.train.AdamOptimizer(lr) defaults on the flat unconstrained parameter buffer x [n], maximising the ELBO:
 * grad_elbo [n] holds d ELBO / d (constrained value); the first n_pos entries are `positive`-transformed (the chain
 * rule factor sigmoid(x) is applied here and theta_pos [n_pos] is refreshed after the update).  mask [n] (0/1) or
 * NULL freezes entries (set_trainable(False), build_models.py:209,213,225-227).  t = 1, 2, ... is the step count. */
int iwvi_adam_step(double* x, const double* grad_elbo, double* m, double* v, const double* mask, double* theta_pos,
                   int64_t n, int64_t n_pos, double lr, double beta1, double beta2, double eps, int64_t t,
                   void* stream);

/* iwvi_adam_step with the step count t = state[0] + 1 and the learning rate lr[0] read from DEVICE memory; increments
 * state[0] afterwards (second, one-thread launch).  For CUDA-graph replay of the training step. */
int iwvi_adam_step_counter(double* x, const double* grad_elbo, double* m, double* v, const double* mask,
                           double* theta_pos, int64_t n, int64_t n_pos, const double* lr, double beta1, double beta2,
                           double eps, int64_t* state, void* stream);

/* ---- measurement utility (no reference counterpart) ----
 * Register-only mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4) issue loop: blocks x warps warps, 8 independent accumulator pairs
 * each, `iters` iterations -> 2*256*8*iters*warps*blocks flops.  bench.py times it with CUDA events to state the FP64
 * tensor-pipe roofline denominator from the run itself.  out: blocks*warps*32 doubles (written, never read). */
int iwvi_probe_dmma(double* out, int32_t blocks, int32_t warps, int32_t iters, void* stream);
/* slots[idx] = %globaltimer (ns) when `stream` reaches this point; a one-thread launch, capturable (tools/graph_timeline.py) */
int iwvi_debug_stamp(unsigned long long* slots, int32_t idx, void* stream);

/* iwvi_adam_step_counter restricted to the entries with mask != 0 WITHOUT touching the others (their value, moments and
 * constrained copies stay as they are), so that a step can update the parameters segment by segment as their gradients
 * become final; advance != 0: increment state[0] afterwards (the last segment of a step). */
int iwvi_adam_step_counter_part(double* x, const double* grad_elbo, double* m, double* v, const double* mask,
                                double* theta_pos, int64_t n, int64_t n_pos, const double* lr, double beta1, double beta2,
                                double eps, int64_t* state, int32_t advance, void* stream);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* IWVI_B200_H */
