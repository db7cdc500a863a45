// gp_fullcov.cu -- covariance over the inner axis and the joint draw (forward), for plain-kernel GP layers.
//
// Replaces the full_cov=True branch of independent_multisample_sample_conditional (reference temp_workaround.py:45,
// :55-57, :82-83) and the joint sampler the reference intends at :92-96 (its own version adds an [S,N,R] mean to an
// [S,R,N,1] draw and never executes).  For every group s of N consecutive points
// (Xnew is [S, N, D]; DGP_IWVI passes S = minibatch rows, N = importance samples, models.py:118-123):
//
//   C_r   = k(X_s, X_s) - A_s^T A_s + U_{r,s}^T U_{r,s}          cov [S, R, N, N]
//   L_r   = chol(C_r + chol_jitter I)                             (:95 has no jitter; gpflow's 2-D _sample_mvn adds it)
//   smp   = mean[:, r] + L_r z_r                                  sample [S, N, R], z = eps [S, R, N]  (:93-94 draws in
//                                                                 the [S, R, N, 1] order)
//
// A = Lm^-1 Kuf and U_r = tril(q_sqrt_r)^T A are the block-major panels iwvi_gp_rows_fwd saved (IWVI_FLAG_SAVE); `mean`
// is its output (mean function included, layers.py:46-48).  One CTA owns a group: the N x N Gram products run on the
// FP64 tensor pipe (DMMA, one 8-row tile of C per warp), operands staged one 64-wide m-block at a time; covariances
// of any N are produced in 64 x 64 blocks; the joint draw factorises C_r by a right-looking Cholesky -- in shared memory
// inside the same kernel for N <= 64, and for 64 < N <= 256 (BASELINE config c4 has K = 256) by a batched BLOCKED
// Cholesky over the covariances in global memory (fc_chol_kernel: 64-wide block columns, diagonal block and block
// column panel resident in shared memory, one CTA per matrix), which also applies the draw.  The adjoint follows the
// same split (gp_fullcov_bwd_kernel / gp_fullcov_bwd_large_kernel).
#include "common.cuh"

namespace {

#define FC_THREADS 256
#define FC_LDC 65      // leading dimension of the N x N matrices in shared memory (odd: column walks are conflict-free)

#define FC_MAXN 256    // largest inner axis served by the joint draw and the covariance adjoint

struct FullCovParams {
  iwvi_gp_desc d;
  int S, N;              // groups of this launch, points per group
  int64_t s0;            // absolute index of the launch's first group (points, eps, mean are indexed absolutely)
  int64_t cov_s_base;    // cov[0] is the covariance of group cov_s_base (0: the caller's full array; s0: a workspace chunk)
  const double *aux, *X, *save, *mean, *eps;
  double chol_jitter;
  double *cov, *sample;
  int* info;
};

// acc[j][c] (+)= sum_m Pi(8 w + g, m) Pj(8 j + 2 t + c, m) over the 64 columns of the staged slabs Pi, Pj [64][68]
__device__ __forceinline__ void slab_syrk(double (&acc)[8][2], const double* __restrict__ slab_i,
                                          const double* __restrict__ slab_j, int warp, int lane, int ntn) {
  const int g = lane >> 2, t = lane & 3;
  const double* ap = slab_i + (warp * 8 + g) * IWVI_LDS + t;
  const double* bp = slab_j + g * IWVI_LDS + t;
#pragma unroll 4
  for (int k0 = 0; k0 < IWVI_BLK; k0 += 4) {
    const double a = ap[k0];
#pragma unroll
    for (int j = 0; j < 8; j++)
      if (j < ntn) dmma884(acc[j], a, bp[j * 8 * IWVI_LDS + k0]);
  }
}

// one m-block of a saved block-major panel, rows pt0 .. pt0 + n - 1 (rows >= n zero), into a [64][68] slab
__device__ __forceinline__ void load_slab(double* slab, const double* __restrict__ panel, int64_t pt0, int n, int mb, int NB,
                                          int tid) {
  for (int idx = tid; idx < IWVI_BLK * IWVI_BLK; idx += FC_THREADS) {
    const int r = idx >> 6, c = idx & 63;
    slab[r * IWVI_LDS + c] = (r < n) ? panel[iwvi_blk_off((int)(pt0 + r), mb * IWVI_BLK + c, NB)] : 0.0;
  }
}

// Work item = (group s, 64-point row block bi, 64-point column block bj) of the N x N covariances of the group; the
// joint draw needs the whole matrix in one CTA and is built for N <= 64 (one block).
__global__ void __launch_bounds__(FC_THREADS) gp_fullcov_fwd_kernel(const FullCovParams p) {
  extern __shared__ __align__(16) double smem[];
  const iwvi_gp_desc& d = p.d;
  const AuxLayout al = iwvi_aux_layout(d.M, d.D, d.R);
  const SaveLayout sv = iwvi_save_layout(d.T, d.M, d.R);
  const int N = p.N, R = d.R, D = d.D, NB = al.NB;
  const int nblk = (N + IWVI_BLK - 1) / IWVI_BLK;
  double* slab_i = smem;                             // [64][68] one m-block of A_s or U_{r,s}, row points of the item
  double* slab_jb = slab_i + IWVI_STAGE_DOUBLES;     // the same for the column points (unused on diagonal items)
  double* C0 = slab_jb + IWVI_STAGE_DOUBLES;         // [64][65] k(X_i, X_j) - A_i^T A_j
  double* Cr = C0 + IWVI_BLK * FC_LDC;               // [64][65] C_r, then its Cholesky factor
  double* xs_i = Cr + IWVI_BLK * FC_LDC;             // [64][D] length-scaled inputs of the row / column points
  double* xs_j = xs_i + IWVI_BLK * IWVI_MAX_D;
  double* xn_i = xs_j + IWVI_BLK * IWVI_MAX_D;       // [64] squared norms
  double* xn_j = xn_i + IWVI_BLK;
  __shared__ int bad;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const double* consts = p.aux + al.off_consts;
  const double variance = consts[IWVI_C_VARIANCE];
  const double* Apanel = p.save + sv.off_a;

  const int64_t n_items = (int64_t)p.S * nblk * nblk;
  for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int64_t s = p.s0 + item / (nblk * nblk);
    const int bi = (int)(item % (nblk * nblk)) / nblk, bj = (int)(item % (nblk * nblk)) % nblk;
    const int i0 = bi * IWVI_BLK, j0 = bj * IWVI_BLK;
    const int ni = min(IWVI_BLK, N - i0), nj = min(IWVI_BLK, N - j0);
    const int64_t pt_i = (int64_t)s * N + i0, pt_j = (int64_t)s * N + j0;
    const bool diag = bi == bj;
    const double* slab_j = diag ? slab_i : slab_jb;
    const int ntn = iwvi_round_up(nj, 8) / 8;
    const bool active = warp * 8 < iwvi_round_up(ni, 8);   // this warp owns rows 8 warp .. 8 warp + 7 of the block
    __syncthreads();
    // ---- k(X_i, X_j) with gpflow's expanded squared distance
    for (int idx = tid; idx < ni * D; idx += FC_THREADS) {
      const int n = idx / D, k = idx - n * D;
      xs_i[n * IWVI_MAX_D + k] = p.X[(pt_i + n) * D + k] * consts[IWVI_C_INVLS + k];
    }
    for (int idx = tid; idx < nj * D; idx += FC_THREADS) {
      const int n = idx / D, k = idx - n * D;
      xs_j[n * IWVI_MAX_D + k] = p.X[(pt_j + n) * D + k] * consts[IWVI_C_INVLS + k];
    }
    __syncthreads();
    if (tid < ni) {
      double q = 0.0;
      for (int k = 0; k < D; k++) { const double v = xs_i[tid * IWVI_MAX_D + k]; q += v * v; }
      xn_i[tid] = q;
    } else if (tid >= IWVI_BLK && tid - IWVI_BLK < nj) {
      const int n = tid - IWVI_BLK;
      double q = 0.0;
      for (int k = 0; k < D; k++) { const double v = xs_j[n * IWVI_MAX_D + k]; q += v * v; }
      xn_j[n] = q;
    }
    __syncthreads();
    for (int idx = tid; idx < IWVI_BLK * IWVI_BLK; idx += FC_THREADS) {
      const int i = idx >> 6, j = idx & 63;
      double v = 0.0;
      if (i < ni && j < nj) {
        double dot = 0.0;
        for (int k = 0; k < D; k++) dot += xs_i[i * IWVI_MAX_D + k] * xs_j[j * IWVI_MAX_D + k];
        v = kern_k(d.kern, xn_i[i] + xn_j[j] - 2.0 * dot, variance);
      }
      C0[i * FC_LDC + j] = v;
    }
    // ---- C0 -= A_i^T A_j
    double acc[8][2];
#pragma unroll
    for (int j = 0; j < 8; j++) { acc[j][0] = 0.0; acc[j][1] = 0.0; }
    for (int mb = 0; mb < NB; mb++) {
      __syncthreads();
      load_slab(slab_i, Apanel, pt_i, ni, mb, NB, tid);
      if (!diag) load_slab(slab_jb, Apanel, pt_j, nj, mb, NB, tid);
      __syncthreads();
      if (active) slab_syrk(acc, slab_i, slab_j, warp, lane, ntn);
    }
    if (active) {
#pragma unroll
      for (int j = 0; j < 8; j++)
        if (j < ntn) {
          C0[(warp * 8 + g) * FC_LDC + j * 8 + 2 * t] -= acc[j][0];
          C0[(warp * 8 + g) * FC_LDC + j * 8 + 2 * t + 1] -= acc[j][1];
        }
    }
    for (int r = 0; r < R; r++) {
      // ---- C_r = C0 + U_ri^T U_rj (accumulators start from this thread's own entries of C0)
      const double* Upanel = p.save + sv.off_u + (int64_t)r * sv.u_stride;
      __syncthreads();
      if (active) {
#pragma unroll
        for (int j = 0; j < 8; j++)
          if (j < ntn) {
            acc[j][0] = C0[(warp * 8 + g) * FC_LDC + j * 8 + 2 * t];
            acc[j][1] = C0[(warp * 8 + g) * FC_LDC + j * 8 + 2 * t + 1];
          }
      }
      for (int mb = 0; mb < NB; mb++) {
        __syncthreads();
        load_slab(slab_i, Upanel, pt_i, ni, mb, NB, tid);
        if (!diag) load_slab(slab_jb, Upanel, pt_j, nj, mb, NB, tid);
        __syncthreads();
        if (active) slab_syrk(acc, slab_i, slab_j, warp, lane, ntn);
      }
      if (active) {
#pragma unroll
        for (int j = 0; j < 8; j++)
          if (j < ntn) {
            Cr[(warp * 8 + g) * FC_LDC + j * 8 + 2 * t] = acc[j][0];
            Cr[(warp * 8 + g) * FC_LDC + j * 8 + 2 * t + 1] = acc[j][1];
          }
      }
      if (tid == 0) bad = 0;
      __syncthreads();
      if (p.cov) {
        double* dst = p.cov + ((s - p.cov_s_base) * R + r) * N * N + (int64_t)i0 * N + j0;
        for (int idx = tid; idx < ni * nj; idx += FC_THREADS) {
          const int i = idx / nj, j = idx - i * nj;
          dst[(int64_t)i * N + j] = Cr[i * FC_LDC + j];
        }
      }
      if (!p.sample) continue;
      // ---- right-looking Cholesky of C_r + chol_jitter I, in place (lower); N <= 64: the item is the whole matrix
      for (int j = 0; j < N; j++) {
        __syncthreads();
        const double piv = Cr[j * FC_LDC + j] + p.chol_jitter;
        if (!(piv > 0.0) && tid == 0 && !bad) { bad = j + 1; }
        const double dj = sqrt(piv);
        const double inv = 1.0 / dj;
        __syncthreads();
        if (tid == 0) Cr[j * FC_LDC + j] = dj;
        for (int i = j + 1 + tid; i < N; i += FC_THREADS) Cr[i * FC_LDC + j] *= inv;
        __syncthreads();
        const int n_tr = N - j - 1;
        for (int idx = tid; idx < n_tr * n_tr; idx += FC_THREADS) {
          const int a = idx / n_tr, b = idx - a * n_tr;
          if (b <= a) Cr[(j + 1 + a) * FC_LDC + j + 1 + b] -= Cr[(j + 1 + a) * FC_LDC + j] * Cr[(j + 1 + b) * FC_LDC + j];
        }
      }
      __syncthreads();
      if (tid == 0 && bad && p.info) atomicCAS(p.info, 0, bad);   // LAPACK-style: order of the first bad leading minor
      // ---- joint draw over the group
      if (tid < N) {
        const double* z = p.eps + (s * R + r) * N;
        double v = p.mean[(pt_i + tid) * R + r];
        for (int j = 0; j <= tid; j++) v += Cr[tid * FC_LDC + j] * z[j];
        p.sample[(pt_i + tid) * R + r] = v;
      }
    }
  }
}


// ------------------------------------------------------------------------------------------------------------------
// Adjoint of the stage above, in a form the existing per-point backward kernels can finish.
//
// With H_r = sym(dC_r) + 1/2 L_r^-T (P + P^T) L_r^-1, P = Phi(L_r^T Lbar_r), Lbar_r = tril(dsmp_r z_r^T) (the TF / torch
// Cholesky adjoint, symmetrised; here L^T Lbar is rank one below the diagonal: P_ij = q_i z_j with q = L^T dsmp_r), the
// cotangent of A is
//     Abar / 2 = q_mu gmean_bar^T / 2 - A_s (sum_r H_r) + sum_r tril(Lq_r) (U_rs H_r),
// which is gp_tile_bwd_kernel's formula with the per-point scalings diag(gvar_bar_r) replaced by the N x N matrices H_r.
// This kernel therefore writes, in the block-major layout of the saved panels,
//     save2.A  = A_s (sum_r H_r) / R          save2.U_r = U_rs H_r
// so that iwvi_gp_rows_bwd run on save2 with d_var = 1 (gvar_bar = 1, sum over r = R; IWVI_FLAG_NO_KDIAG) yields Bbar,
// dX, dZ, dls, dvariance of the Kuf path; its parameter reductions (dLm = -tril(Bbar A^T), dq_mu = A gmean_bar,
// dLq_r = 2 tril(A (U_r H_r)^T)) then need the ORIGINAL A back in save2.A.  The adjoint of k(X_s, X_s) (cotangent
// sum_r H_r) is formed here: dX_knn [T, D] and per-group partials part [S, 40] (dls at 0..D-1, dvariance at 32).
struct FullCovBwdParams {
  iwvi_gp_desc d;
  int S, N;
  const double *aux, *X, *save, *eps, *d_sample, *d_cov;
  double chol_jitter;
  double *save2, *dXk, *part;
};

// acc[j][c] = sum_i Hm(8 w + g, i) slab(i, 8 j + 2 t + c): Hm symmetric [64][65], slab [64][68] (rows = points)
template <bool ZERO = true>
__device__ __forceinline__ void h_times_slab(double (&acc)[8][2], const double* __restrict__ Hm, const double* __restrict__ slab,
                                             int warp, int lane, int kmax) {
  const int g = lane >> 2, t = lane & 3;
  const double* ap = Hm + (warp * 8 + g) * FC_LDC + t;
  const double* bp = slab + t * IWVI_LDS + g;
  if (ZERO) {
#pragma unroll
    for (int j = 0; j < 8; j++) { acc[j][0] = 0.0; acc[j][1] = 0.0; }
  }
  for (int k0 = 0; k0 < kmax; k0 += 4) {
    const double a = ap[k0];
#pragma unroll
    for (int j = 0; j < 8; j++) dmma884(acc[j], a, bp[k0 * IWVI_LDS + j * 8]);
  }
}

__device__ __forceinline__ void store_rows(double* __restrict__ panel, const double (&acc)[8][2], int64_t pt0, int n, int mb,
                                           int NB, int warp, int lane) {
  const int g = lane >> 2, t = lane & 3, row = warp * 8 + g;
  if (row >= n) return;
#pragma unroll
  for (int j = 0; j < 8; j++) {
    double* dst = panel + iwvi_blk_off((int)(pt0 + row), mb * IWVI_BLK + j * 8 + 2 * t, NB);
    dst[0] = acc[j][0];
    dst[1] = acc[j][1];
  }
}

__global__ void __launch_bounds__(FC_THREADS) gp_fullcov_bwd_kernel(const FullCovBwdParams p) {
  extern __shared__ __align__(16) double smem[];
  const iwvi_gp_desc& d = p.d;
  const AuxLayout al = iwvi_aux_layout(d.M, d.D, d.R);
  const SaveLayout sv = iwvi_save_layout(d.T, d.M, d.R);
  const int N = p.N, R = d.R, D = d.D, NB = al.NB;
  const int Np = iwvi_round_up(N, 8), ntn = Np / 8;
  double* slab = smem;                               // [64][68]
  double* C0 = slab + IWVI_STAGE_DOUBLES;            // [64][65] k(X_s, X_s) - A_s^T A_s
  double* Lr = C0 + IWVI_BLK * FC_LDC;               // C_r, then its Cholesky factor (lower)
  double* Hr = Lr + IWVI_BLK * FC_LDC;               // H_r (symmetric, zero outside N x N)
  double* Hs = Hr + IWVI_BLK * FC_LDC;               // sum_r H_r
  double* Sc = Hs + IWVI_BLK * FC_LDC;               // scratch: Q -> L^-T Q -> L^-T Q L^-1 ; later Hs * dk/dr2
  double* xs = Sc + IWVI_BLK * FC_LDC;               // [64][32] length-scaled inputs
  double* xn = xs + IWVI_BLK * IWVI_MAX_D;           // [64]
  double* dv = xn + IWVI_BLK;                        // [64] cotangent of the draw, output r
  double* zv = dv + IWVI_BLK;                        // [64] its noise
  double* qv = zv + IWVI_BLK;                        // [64] L^T dv
  __shared__ double red[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const double* consts = p.aux + al.off_consts;
  const double variance = consts[IWVI_C_VARIANCE];
  const double* Apanel = p.save + sv.off_a;
  const bool active = warp * 8 < Np;

  for (int s = blockIdx.x; s < p.S; s += gridDim.x) {
    const int64_t pt0 = (int64_t)s * N;
    __syncthreads();
    for (int idx = tid; idx < N * D; idx += FC_THREADS) {
      const int n = idx / D, k = idx - n * D;
      xs[n * IWVI_MAX_D + k] = p.X[(pt0 + n) * D + k] * consts[IWVI_C_INVLS + k];
    }
    for (int idx = tid; idx < IWVI_BLK * FC_LDC; idx += FC_THREADS) Hs[idx] = 0.0;
    __syncthreads();
    if (tid < N) {
      double q = 0.0;
      for (int k = 0; k < D; k++) { const double v = xs[tid * IWVI_MAX_D + k]; q += v * v; }
      xn[tid] = q;
    }
    __syncthreads();
    for (int idx = tid; idx < IWVI_BLK * IWVI_BLK; idx += FC_THREADS) {
      const int i = idx >> 6, j = idx & 63;
      double v = 0.0;
      if (i < N && j < N) {
        double dot = 0.0;
        for (int k = 0; k < D; k++) dot += xs[i * IWVI_MAX_D + k] * xs[j * IWVI_MAX_D + k];
        v = kern_k(d.kern, xn[i] + xn[j] - 2.0 * dot, variance);
      }
      C0[i * FC_LDC + j] = v;
    }
    double acc[8][2];
#pragma unroll
    for (int j = 0; j < 8; j++) { acc[j][0] = 0.0; acc[j][1] = 0.0; }
    for (int mb = 0; mb < NB; mb++) {
      __syncthreads();
      load_slab(slab, Apanel, pt0, N, mb, NB, tid);
      __syncthreads();
      if (active) slab_syrk(acc, slab, slab, warp, lane, ntn);
    }
    if (active) {
#pragma unroll
      for (int j = 0; j < 8; j++)
        if (j < ntn) {
          C0[(warp * 8 + g) * FC_LDC + j * 8 + 2 * t] -= acc[j][0];
          C0[(warp * 8 + g) * FC_LDC + j * 8 + 2 * t + 1] -= acc[j][1];
        }
    }
    for (int r = 0; r < R; r++) {
      const double* Upanel = p.save + sv.off_u + (int64_t)r * sv.u_stride;
      double* Vpanel = p.save2 + sv.off_u + (int64_t)r * sv.u_stride;
      const double* dc = p.d_cov ? p.d_cov + ((int64_t)s * R + r) * N * N : nullptr;
      __syncthreads();
      // ---- H_r starts as sym(dC_r)
      for (int idx = tid; idx < IWVI_BLK * IWVI_BLK; idx += FC_THREADS) {
        const int i = idx >> 6, j = idx & 63;
        Hr[i * FC_LDC + j] = (dc && i < N && j < N) ? 0.5 * (dc[i * N + j] + dc[j * N + i]) : 0.0;
      }
      if (p.d_sample) {
        // ---- C_r and its Cholesky factor, as in the forward kernel
        if (active) {
#pragma unroll
          for (int j = 0; j < 8; j++)
            if (j < ntn) {
              acc[j][0] = C0[(warp * 8 + g) * FC_LDC + j * 8 + 2 * t];
              acc[j][1] = C0[(warp * 8 + g) * FC_LDC + j * 8 + 2 * t + 1];
            }
        }
        for (int mb = 0; mb < NB; mb++) {
          __syncthreads();
          load_slab(slab, Upanel, pt0, N, mb, NB, tid);
          __syncthreads();
          if (active) slab_syrk(acc, slab, slab, warp, lane, ntn);
        }
        if (active) {
#pragma unroll
          for (int j = 0; j < 8; j++)
            if (j < ntn) {
              Lr[(warp * 8 + g) * FC_LDC + j * 8 + 2 * t] = acc[j][0];
              Lr[(warp * 8 + g) * FC_LDC + j * 8 + 2 * t + 1] = acc[j][1];
            }
        }
        for (int j = 0; j < N; j++) {
          __syncthreads();
          const double dj = sqrt(Lr[j * FC_LDC + j] + p.chol_jitter);
          const double inv = 1.0 / dj;
          __syncthreads();
          if (tid == 0) Lr[j * FC_LDC + j] = dj;
          for (int i = j + 1 + tid; i < N; i += FC_THREADS) Lr[i * FC_LDC + j] *= inv;
          __syncthreads();
          const int n_tr = N - j - 1;
          for (int idx = tid; idx < n_tr * n_tr; idx += FC_THREADS) {
            const int a = idx / n_tr, b = idx - a * n_tr;
            if (b <= a) Lr[(j + 1 + a) * FC_LDC + j + 1 + b] -= Lr[(j + 1 + a) * FC_LDC + j] * Lr[(j + 1 + b) * FC_LDC + j];
          }
        }
        if (tid < N) {
          dv[tid] = p.d_sample[(pt0 + tid) * R + r];
          zv[tid] = p.eps[((int64_t)s * R + r) * N + tid];
        }
        __syncthreads();
        if (tid < N) {        // q = L^T dv
          double q = 0.0;
          for (int k = tid; k < N; k++) q += Lr[k * FC_LDC + tid] * dv[k];
          qv[tid] = q;
        }
        __syncthreads();
        // Q = P + P^T, P = Phi(L^T Lbar): Q_ij = q_i z_j (i > j), q_j z_i (i < j), q_i z_i (i == j)
        for (int idx = tid; idx < N * N; idx += FC_THREADS) {
          const int i = idx / N, j = idx - i * N;
          Sc[i * FC_LDC + j] = i >= j ? qv[i] * zv[j] : qv[j] * zv[i];
        }
        __syncthreads();
        if (tid < N) {        // Y = L^-T Q, one column per thread (back substitution)
          const int c = tid;
          for (int i = N - 1; i >= 0; i--) {
            double v = Sc[i * FC_LDC + c];
            for (int k = i + 1; k < N; k++) v -= Lr[k * FC_LDC + i] * Sc[k * FC_LDC + c];
            Sc[i * FC_LDC + c] = v / Lr[i * FC_LDC + i];
          }
        }
        __syncthreads();
        if (tid < N) {        // S1 = Y L^-1: row c of S1 solves L^T w = (row c of Y)^T
          const int c = tid;
          for (int i = N - 1; i >= 0; i--) {
            double v = Sc[c * FC_LDC + i];
            for (int k = i + 1; k < N; k++) v -= Lr[k * FC_LDC + i] * Sc[c * FC_LDC + k];
            Sc[c * FC_LDC + i] = v / Lr[i * FC_LDC + i];
          }
        }
        __syncthreads();
        for (int idx = tid; idx < N * N; idx += FC_THREADS) {
          const int i = idx / N, j = idx - i * N;
          Hr[i * FC_LDC + j] += 0.25 * (Sc[i * FC_LDC + j] + Sc[j * FC_LDC + i]);
        }
      }
      __syncthreads();
      for (int idx = tid; idx < IWVI_BLK * FC_LDC; idx += FC_THREADS) Hs[idx] += Hr[idx];
      // ---- save2.U_r = U_rs H_r
      for (int mb = 0; mb < NB; mb++) {
        __syncthreads();
        load_slab(slab, Upanel, pt0, N, mb, NB, tid);
        __syncthreads();
        if (active) {
          h_times_slab(acc, Hr, slab, warp, lane, Np);
          store_rows(Vpanel, acc, pt0, N, mb, NB, warp, lane);
        }
      }
    }
    // ---- save2.A = A_s (sum_r H_r) / R
    __syncthreads();
    const double invR = 1.0 / (double)R;
    for (int idx = tid; idx < IWVI_BLK * FC_LDC; idx += FC_THREADS) Hr[idx] = Hs[idx] * invR;
    for (int mb = 0; mb < NB; mb++) {
      __syncthreads();
      load_slab(slab, Apanel, pt0, N, mb, NB, tid);
      __syncthreads();
      if (active) {
        h_times_slab(acc, Hr, slab, warp, lane, Np);
        store_rows(p.save2 + sv.off_a, acc, pt0, N, mb, NB, warp, lane);
      }
    }
    // ---- adjoint of k(X_s, X_s) with cotangent Hs: Sc = Hs * dk/dr2, dvariance partial = sum Hs * k / variance
    __syncthreads();
    double vpart = 0.0;
    for (int idx = tid; idx < N * N; idx += FC_THREADS) {
      const int i = idx / N, j = idx - i * N;
      double dot = 0.0;
      for (int k = 0; k < D; k++) dot += xs[i * IWVI_MAX_D + k] * xs[j * IWVI_MAX_D + k];
      double K, dK;
      kern_k_dk(d.kern, xn[i] + xn[j] - 2.0 * dot, variance, K, dK);
      const double h = Hs[i * FC_LDC + j];
      Sc[i * FC_LDC + j] = h * dK;
      vpart += h * K;
    }
    vpart = block_sum(vpart, red);      // total in thread 0
    if (tid == 0) qv[0] = vpart;
    __syncthreads();
    // dx~_id = 4 sum_j (Hs dk)_ij (x~_id - x~_jd); C0 is free now: C0[i][d] = -dx~_id x~_id / ls_d (dls summands)
    for (int idx = tid; idx < N * D; idx += FC_THREADS) {
      const int i = idx / D, k = idx - i * D;
      const double xi = xs[i * IWVI_MAX_D + k];
      double sum = 0.0;
      for (int j = 0; j < N; j++) sum += Sc[i * FC_LDC + j] * (xi - xs[j * IWVI_MAX_D + k]);
      const double dx = 4.0 * sum * consts[IWVI_C_INVLS + k];
      p.dXk[(pt0 + i) * D + k] = dx;
      C0[i * FC_LDC + k] = -dx * xi;
    }
    __syncthreads();
    double* part = p.part + (int64_t)s * 40;
    if (tid < 40) {
      double v = 0.0;
      if (tid < D) for (int i = 0; i < N; i++) v += C0[i * FC_LDC + tid];
      else if (tid == 32) v = qv[0] / variance;
      part[tid] = v;
    }
  }
}


// ------------------------------------------------------------------------------------------------------------------
// Groups of 64 < N <= FC_MAXN points: the N x N matrices live in global memory (the caller's workspace).
// ------------------------------------------------------------------------------------------------------------------
// Batched blocked right-looking Cholesky, in place (lower triangle and diagonal; the strict upper triangle is left as
// it is), one CTA per matrix: per 64-wide block column the diagonal block is factorised in shared memory, the block
// column below it is solved against it row by row (one thread per row) and kept in shared memory for the trailing update
// C(i, j) -= sum_c L(i, c) L(j, c), which runs over the lower triangle of what is left.  Optionally applies the joint
// draw sample = mean + L z (temp_workaround.py:92-96 as intended) while the block columns are on chip.
struct FcCholParams {
  double* mats;           // [n_mats][N][N]; matrix i belongs to group s0 + i / R, output r = i % R
  int n_mats, N, R;
  int64_t s0;
  double jitter;
  int* info;
  const double *mean, *eps;
  double* sample;         // NULL: factorise only
};

__global__ void __launch_bounds__(FC_THREADS) fc_chol_kernel(const FcCholParams p) {
  extern __shared__ __align__(16) double smem[];
  double* Dg = smem;                                   // [64][65] diagonal block
  double* Pn = Dg + IWVI_BLK * FC_LDC;                 // [FC_MAXN - 64][65] block column below it
  double* zv = Pn + (FC_MAXN - IWVI_BLK) * FC_LDC;     // [FC_MAXN] noise of this matrix
  double* vv = zv + FC_MAXN;                           // [FC_MAXN] L z so far
  __shared__ int bad;
  const int N = p.N, R = p.R, tid = threadIdx.x;
  for (int mi = blockIdx.x; mi < p.n_mats; mi += gridDim.x) {
    double* Cm = p.mats + (int64_t)mi * N * N;
    const int r = mi % R;
    const int64_t s = p.s0 + mi / R;
    __syncthreads();
    if (tid == 0) bad = 0;
    if (p.sample)
      for (int i = tid; i < N; i += FC_THREADS) { zv[i] = p.eps[(s * R + r) * N + i]; vv[i] = 0.0; }
    for (int k0 = 0; k0 < N; k0 += IWVI_BLK) {
      const int nk = min(IWVI_BLK, N - k0), nrem = N - k0 - nk;
      __syncthreads();
      for (int idx = tid; idx < nk * nk; idx += FC_THREADS) {
        const int i = idx / nk, j = idx - i * nk;
        if (j <= i) Dg[i * FC_LDC + j] = Cm[(int64_t)(k0 + i) * N + k0 + j] + (i == j ? p.jitter : 0.0);
      }
      for (int idx = tid; idx < nrem * nk; idx += FC_THREADS) {
        const int i = idx / nk, j = idx - i * nk;
        Pn[i * FC_LDC + j] = Cm[(int64_t)(k0 + nk + i) * N + k0 + j];
      }
      for (int j = 0; j < nk; j++) {
        __syncthreads();
        const double piv = Dg[j * FC_LDC + j];
        if (!(piv > 0.0) && tid == 0 && !bad) bad = k0 + j + 1;
        const double dj = sqrt(piv);
        const double inv = 1.0 / dj;
        __syncthreads();
        if (tid == 0) Dg[j * FC_LDC + j] = dj;
        for (int i = j + 1 + tid; i < nk; i += FC_THREADS) Dg[i * FC_LDC + j] *= inv;
        __syncthreads();
        const int n_tr = nk - j - 1;
        for (int idx = tid; idx < n_tr * n_tr; idx += FC_THREADS) {
          const int a = idx / n_tr, b = idx - a * n_tr;
          if (b <= a) Dg[(j + 1 + a) * FC_LDC + j + 1 + b] -= Dg[(j + 1 + a) * FC_LDC + j] * Dg[(j + 1 + b) * FC_LDC + j];
        }
      }
      __syncthreads();
      if (tid < nrem) {            // row tid of the block column: x L_kk^T = a (forward substitution along the row)
        double* row = Pn + tid * FC_LDC;
        for (int c = 0; c < nk; c++) {
          double v = row[c];
          for (int q = 0; q < c; q++) v -= row[q] * Dg[c * FC_LDC + q];
          row[c] = v / Dg[c * FC_LDC + c];
        }
      }
      __syncthreads();
      for (int idx = tid; idx < nk * nk; idx += FC_THREADS) {
        const int i = idx / nk, j = idx - i * nk;
        if (j <= i) Cm[(int64_t)(k0 + i) * N + k0 + j] = Dg[i * FC_LDC + j];
      }
      for (int idx = tid; idx < nrem * nk; idx += FC_THREADS) {
        const int i = idx / nk, j = idx - i * nk;
        Cm[(int64_t)(k0 + nk + i) * N + k0 + j] = Pn[i * FC_LDC + j];
      }
      if (p.sample)
        for (int i = tid; i < nk + nrem; i += FC_THREADS) {
          const double* row = i < nk ? Dg + i * FC_LDC : Pn + (i - nk) * FC_LDC;
          const int lim = i < nk ? i + 1 : nk;
          double a = 0.0;
          for (int c = 0; c < lim; c++) a += row[c] * zv[k0 + c];
          vv[k0 + i] += a;
        }
      for (int idx = tid; idx < nrem * nrem; idx += FC_THREADS) {
        const int a = idx / nrem, b = idx - a * nrem;
        if (b <= a) {
          double v = 0.0;
          for (int c = 0; c < nk; c++) v += Pn[a * FC_LDC + c] * Pn[b * FC_LDC + c];
          Cm[(int64_t)(k0 + nk + a) * N + k0 + nk + b] -= v;
        }
      }
    }
    __syncthreads();
    if (tid == 0 && bad && p.info) atomicCAS(p.info, 0, bad);
    if (p.sample)
      for (int i = tid; i < N; i += FC_THREADS) p.sample[(s * N + i) * R + r] = p.mean[(s * N + i) * R + r] + vv[i];
  }
}

// Adjoint for 64 < N <= FC_MAXN: the same algebra as gp_fullcov_bwd_kernel with the N x N matrices H_r, sum_r H_r and
// the two scratch matrices of the Cholesky adjoint in a per-CTA slice of the workspace (4 N^2 doubles), the Cholesky
// factors L_r read from global memory (formed beforehand by the covariance kernel + fc_chol_kernel), and the products
// U_rs H_r / A_s (sum_r H_r) assembled from 64 x 64 blocks of H.
struct FullCovBwdLargeParams {
  iwvi_gp_desc d;
  int S, N;
  int64_t s0;
  const double *aux, *X, *save, *eps, *d_sample, *d_cov, *Lfac;   // Lfac [S][R][N][N] (this launch's groups) or NULL
  double *save2, *dXk, *part, *ws;
};

// X := L^-T X for an N x N matrix X in global memory, one column per thread (N <= FC_THREADS)
__device__ __forceinline__ void solve_lt_columns(const double* __restrict__ L, double* Xm, int N, int tid) {
  if (tid < N) {
    const int c = tid;
    for (int i = N - 1; i >= 0; i--) {
      double v = Xm[(int64_t)i * N + c];
      for (int k = i + 1; k < N; k++) v -= L[(int64_t)k * N + i] * Xm[(int64_t)k * N + c];
      Xm[(int64_t)i * N + c] = v / L[(int64_t)i * N + i];
    }
  }
}

__global__ void __launch_bounds__(FC_THREADS) gp_fullcov_bwd_large_kernel(const FullCovBwdLargeParams p) {
  extern __shared__ __align__(16) double smem[];
  const iwvi_gp_desc& d = p.d;
  const AuxLayout al = iwvi_aux_layout(d.M, d.D, d.R);
  const SaveLayout sv = iwvi_save_layout(d.T, d.M, d.R);
  const int N = p.N, R = d.R, D = d.D, NB = al.NB;
  const int nblk = (N + IWVI_BLK - 1) / IWVI_BLK;
  double* slab = smem;                                  // [64][68]
  double* Hb = slab + IWVI_STAGE_DOUBLES;               // [64][65] one block of H
  double* xs = Hb + IWVI_BLK * FC_LDC;                  // [FC_MAXN][32] length-scaled inputs
  double* tmp = xs + FC_MAXN * IWVI_MAX_D;              // [FC_MAXN][32] dls summands
  double* xn = tmp + FC_MAXN * IWVI_MAX_D;              // [FC_MAXN]
  double* dv = xn + FC_MAXN;
  double* zv = dv + FC_MAXN;
  double* qv = zv + FC_MAXN;
  __shared__ double red[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double* consts = p.aux + al.off_consts;
  const double variance = consts[IWVI_C_VARIANCE];
  const double* Apanel = p.save + sv.off_a;
  const int64_t NN = (int64_t)N * N;
  double* Hr = p.ws + (int64_t)blockIdx.x * 4 * NN;
  double* Hs = Hr + NN;
  double* Sc = Hs + NN;
  double* Sc2 = Sc + NN;

  // out panel rows of point block bi, m-block mb:  sum_bj H(bi, bj) src(bj, mb)
  auto h_times_panel = [&](const double* Hm, double scale, const double* src, double* dst, int64_t pt0) {
    for (int mb = 0; mb < NB; mb++)
      for (int bi = 0; bi < nblk; bi++) {
        const int nbi = min(IWVI_BLK, N - bi * IWVI_BLK);
        const bool active = warp * 8 < iwvi_round_up(nbi, 8);
        double acc[8][2];
#pragma unroll
        for (int j = 0; j < 8; j++) { acc[j][0] = 0.0; acc[j][1] = 0.0; }
        for (int bj = 0; bj < nblk; bj++) {
          const int nbj = min(IWVI_BLK, N - bj * IWVI_BLK);
          __syncthreads();
          load_slab(slab, src, pt0 + bj * IWVI_BLK, nbj, mb, NB, tid);
          for (int idx = tid; idx < IWVI_BLK * IWVI_BLK; idx += FC_THREADS) {
            const int i = idx >> 6, j = idx & 63;
            Hb[i * FC_LDC + j] = (i < nbi && j < nbj) ? scale * Hm[(int64_t)(bi * IWVI_BLK + i) * N + bj * IWVI_BLK + j] : 0.0;
          }
          __syncthreads();
          if (active) h_times_slab<false>(acc, Hb, slab, warp, lane, IWVI_BLK);
        }
        if (active) store_rows(dst, acc, pt0 + bi * IWVI_BLK, nbi, mb, NB, warp, lane);
      }
  };

  for (int sl = blockIdx.x; sl < p.S; sl += gridDim.x) {
    const int64_t s = p.s0 + sl;
    const int64_t pt0 = s * N;
    __syncthreads();
    for (int idx = tid; idx < N * D; idx += FC_THREADS) {
      const int n = idx / D, k = idx - n * D;
      xs[n * IWVI_MAX_D + k] = p.X[(pt0 + n) * D + k] * consts[IWVI_C_INVLS + k];
    }
    for (int64_t idx = tid; idx < NN; idx += FC_THREADS) Hs[idx] = 0.0;
    __syncthreads();
    for (int n = tid; n < N; n += FC_THREADS) {
      double q = 0.0;
      for (int k = 0; k < D; k++) { const double v = xs[n * IWVI_MAX_D + k]; q += v * v; }
      xn[n] = q;
    }
    for (int r = 0; r < R; r++) {
      const double* Upanel = p.save + sv.off_u + (int64_t)r * sv.u_stride;
      double* Vpanel = p.save2 + sv.off_u + (int64_t)r * sv.u_stride;
      const double* dc = p.d_cov ? p.d_cov + (s * R + r) * NN : nullptr;
      __syncthreads();
      for (int64_t idx = tid; idx < NN; idx += FC_THREADS) {
        const int i = (int)(idx / N), j = (int)(idx - (int64_t)i * N);
        Hr[idx] = dc ? 0.5 * (dc[idx] + dc[(int64_t)j * N + i]) : 0.0;
      }
      if (p.d_sample) {
        const double* L = p.Lfac + ((int64_t)sl * R + r) * NN;
        for (int i = tid; i < N; i += FC_THREADS) {
          dv[i] = p.d_sample[(pt0 + i) * R + r];
          zv[i] = p.eps[(s * R + r) * N + i];
        }
        __syncthreads();
        for (int i = tid; i < N; i += FC_THREADS) {       // q = L^T dv
          double q = 0.0;
          for (int k = i; k < N; k++) q += L[(int64_t)k * N + i] * dv[k];
          qv[i] = q;
        }
        __syncthreads();
        for (int64_t idx = tid; idx < NN; idx += FC_THREADS) {
          const int i = (int)(idx / N), j = (int)(idx - (int64_t)i * N);
          Sc[idx] = i >= j ? qv[i] * zv[j] : qv[j] * zv[i];
        }
        __syncthreads();
        solve_lt_columns(L, Sc, N, tid);                  // Y = L^-T Q
        __syncthreads();
        for (int64_t idx = tid; idx < NN; idx += FC_THREADS) {
          const int i = (int)(idx / N), j = (int)(idx - (int64_t)i * N);
          Sc2[idx] = Sc[(int64_t)j * N + i];              // Y^T
        }
        __syncthreads();
        solve_lt_columns(L, Sc2, N, tid);                 // S1^T = L^-T Y^T
        __syncthreads();
        for (int64_t idx = tid; idx < NN; idx += FC_THREADS) {
          const int i = (int)(idx / N), j = (int)(idx - (int64_t)i * N);
          Hr[idx] += 0.25 * (Sc2[idx] + Sc2[(int64_t)j * N + i]);
        }
      }
      __syncthreads();
      for (int64_t idx = tid; idx < NN; idx += FC_THREADS) Hs[idx] += Hr[idx];
      h_times_panel(Hr, 1.0, Upanel, Vpanel, pt0);        // save2.U_r = U_rs H_r
    }
    __syncthreads();
    h_times_panel(Hs, 1.0 / (double)R, Apanel, p.save2 + sv.off_a, pt0);   // save2.A = A_s (sum_r H_r) / R
    // ---- adjoint of k(X_s, X_s) with cotangent Hs
    __syncthreads();
    double vpart = 0.0;
    for (int64_t idx = tid; idx < NN; idx += FC_THREADS) {
      const int i = (int)(idx / N), j = (int)(idx - (int64_t)i * N);
      double dot = 0.0;
      for (int k = 0; k < D; k++) dot += xs[i * IWVI_MAX_D + k] * xs[j * IWVI_MAX_D + k];
      double K, dK;
      kern_k_dk(d.kern, xn[i] + xn[j] - 2.0 * dot, variance, K, dK);
      const double h = Hs[idx];
      Sc[idx] = h * dK;
      vpart += h * K;
    }
    vpart = block_sum(vpart, red);
    if (tid == 0) qv[0] = vpart;
    __syncthreads();
    for (int idx = tid; idx < N * D; idx += FC_THREADS) {
      const int i = idx / D, k = idx - i * D;
      const double xi = xs[i * IWVI_MAX_D + k];
      double sum = 0.0;
      for (int j = 0; j < N; j++) sum += Sc[(int64_t)i * N + j] * (xi - xs[j * IWVI_MAX_D + k]);
      const double dx = 4.0 * sum * consts[IWVI_C_INVLS + k];
      p.dXk[(pt0 + i) * D + k] = dx;
      tmp[i * IWVI_MAX_D + k] = -dx * xi;
    }
    __syncthreads();
    double* part = p.part + s * 40;
    if (tid < 40) {
      double v = 0.0;
      if (tid < D) for (int i = 0; i < N; i++) v += tmp[i * IWVI_MAX_D + tid];
      else if (tid == 32) v = qv[0] / variance;
      part[tid] = v;
    }
  }
}

}  // namespace

#define FC_CHUNK 256   // groups per pass of the N > 64 paths (bounds the workspace: FC_CHUNK * R * N^2 doubles)

static int fc_device(int* nsm) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return IWVI_ERR_LAUNCH;
  *nsm = 148;
  cudaDeviceGetAttribute(nsm, cudaDevAttrMultiProcessorCount, dev);
  return IWVI_OK;
}

static int fc_check(const iwvi_gp_desc* d, int32_t S, int32_t N) {
  if (!d) return IWVI_ERR_NULL;
  if (d->T < 0 || d->M < 1 || d->D < 1 || d->R < 1) return IWVI_ERR_BAD_DESC;
  if (d->M > IWVI_MAX_M || d->D > IWVI_MAX_D || d->R > IWVI_MAX_R) return IWVI_ERR_UNSUPPORTED;
  if (d->mix || d->P != d->R) return IWVI_ERR_BAD_DESC;            // the Mok branch forces full_cov=False (:125-129)
  if (S < 0 || N < 1 || (int64_t)S * N != d->T) return IWVI_ERR_BAD_DESC;
  return IWVI_OK;
}

extern "C" int64_t iwvi_gp_fullcov_ws_doubles(const iwvi_gp_desc* d, int32_t S, int32_t N) {
  if (fc_check(d, S, N) != IWVI_OK) return -1;
  if (N <= IWVI_BLK) return 0;
  int nsm = 148;
  fc_device(&nsm);
  const int64_t chunk = S < FC_CHUNK ? S : FC_CHUNK;
  return chunk * d->R * (int64_t)N * N + (int64_t)nsm * 4 * N * N;
}

// covariance kernel over the groups [s0, s0 + sc)
static int fc_launch_cov(const iwvi_gp_desc* d, int sc, int64_t s0, int N, const double* aux, const double* X,
                         const double* save, const double* mean, const double* eps, double chol_jitter, double* cov,
                         int64_t cov_s_base, double* sample, int32_t* info, int nsm, cudaStream_t st) {
  const int smem_bytes = (2 * IWVI_STAGE_DOUBLES + 2 * IWVI_BLK * FC_LDC + 2 * IWVI_BLK * IWVI_MAX_D + 2 * IWVI_BLK) * 8;
  if (cudaFuncSetAttribute(gp_fullcov_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes) != cudaSuccess)
    return IWVI_ERR_LAUNCH;
  FullCovParams p;
  p.d = *d; p.S = sc; p.N = N; p.s0 = s0; p.cov_s_base = cov_s_base; p.aux = aux; p.X = X; p.save = save; p.mean = mean;
  p.eps = eps; p.chol_jitter = chol_jitter; p.cov = cov; p.sample = sample; p.info = info;
  const int nblk = (N + IWVI_BLK - 1) / IWVI_BLK;
  const int64_t items = (int64_t)sc * nblk * nblk;
  const int grid = (int)(items < 4 * nsm ? items : 4 * nsm);
  gp_fullcov_fwd_kernel<<<grid, FC_THREADS, smem_bytes, st>>>(p);
  IWVI_CHECK_LAUNCH();
  return IWVI_OK;
}

static int fc_launch_chol(double* mats, int n_mats, int N, int R, int64_t s0, double jitter, int32_t* info,
                          const double* mean, const double* eps, double* sample, int nsm, cudaStream_t st) {
  const int smem_bytes = (FC_MAXN * FC_LDC + 2 * FC_MAXN) * 8;
  if (cudaFuncSetAttribute(fc_chol_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes) != cudaSuccess)
    return IWVI_ERR_LAUNCH;
  FcCholParams p;
  p.mats = mats; p.n_mats = n_mats; p.N = N; p.R = R; p.s0 = s0; p.jitter = jitter; p.info = info; p.mean = mean;
  p.eps = eps; p.sample = sample;
  const int grid = n_mats < nsm ? n_mats : nsm;
  fc_chol_kernel<<<grid, FC_THREADS, smem_bytes, st>>>(p);
  IWVI_CHECK_LAUNCH();
  return IWVI_OK;
}

extern "C" int iwvi_gp_fullcov_fwd(const iwvi_gp_desc* d, int32_t S, int32_t N, const double* aux, const double* X,
                                   const double* save, const double* mean, const double* eps, double chol_jitter,
                                   double* cov, double* sample, int32_t* info, double* ws, void* stream) {
  int rc = fc_check(d, S, N);
  if (rc != IWVI_OK) return rc;
  if (sample && N > FC_MAXN) return IWVI_ERR_UNSUPPORTED;
  if (!aux || !X || !save) return IWVI_ERR_NULL;
  if (sample && (!eps || !mean)) return IWVI_ERR_NULL;
  if (!cov && !sample) return IWVI_ERR_NULL;
  if (S == 0) return IWVI_OK;
  int nsm = 148;
  if ((rc = fc_device(&nsm)) != IWVI_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (!sample || N <= IWVI_BLK)   // covariance in 64 x 64 blocks (any N); the draw of a one-block group in the same kernel
    return fc_launch_cov(d, S, 0, N, aux, X, save, mean, eps, chol_jitter, cov, 0, sample, info, nsm, st);
  // 64 < N <= FC_MAXN: covariances of a chunk of groups into the workspace, blocked Cholesky + draw over them
  if (!ws) return IWVI_ERR_NULL;
  const int64_t per_group = (int64_t)d->R * N * N;
  for (int64_t s0 = 0; s0 < S; s0 += FC_CHUNK) {
    const int sc = (int)(S - s0 < FC_CHUNK ? S - s0 : FC_CHUNK);
    rc = fc_launch_cov(d, sc, s0, N, aux, X, save, mean, eps, chol_jitter, cov ? cov : ws, cov ? 0 : s0, nullptr, nullptr,
                       nsm, st);
    if (rc != IWVI_OK) return rc;
    if (cov && cudaMemcpyAsync(ws, cov + s0 * per_group, (size_t)sc * per_group * 8, cudaMemcpyDeviceToDevice, st) != cudaSuccess)
      return IWVI_ERR_LAUNCH;
    rc = fc_launch_chol(ws, sc * d->R, N, d->R, s0, chol_jitter, info, mean, eps, sample, nsm, st);
    if (rc != IWVI_OK) return rc;
  }
  return IWVI_OK;
}

extern "C" int iwvi_gp_fullcov_bwd(const iwvi_gp_desc* d, int32_t S, int32_t N, const double* aux, const double* X,
                                   const double* save, const double* eps, double chol_jitter, const double* d_sample,
                                   const double* d_cov, double* save2, double* dX_knn, double* part, double* ws,
                                   void* stream) {
  int rc = fc_check(d, S, N);
  if (rc != IWVI_OK) return rc;
  if (N > FC_MAXN) return IWVI_ERR_UNSUPPORTED;
  if (!aux || !X || !save || !save2 || !dX_knn || !part) return IWVI_ERR_NULL;
  if (d_sample && !eps) return IWVI_ERR_NULL;
  if (S == 0) return IWVI_OK;
  int nsm = 148;
  if ((rc = fc_device(&nsm)) != IWVI_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (N <= IWVI_BLK) {
    const int smem_bytes = (IWVI_STAGE_DOUBLES + 5 * IWVI_BLK * FC_LDC + IWVI_BLK * IWVI_MAX_D + 4 * IWVI_BLK) * 8;
    if (cudaFuncSetAttribute(gp_fullcov_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes) != cudaSuccess)
      return IWVI_ERR_LAUNCH;
    FullCovBwdParams p;
    p.d = *d; p.S = S; p.N = N; p.aux = aux; p.X = X; p.save = save; p.eps = eps; p.d_sample = d_sample; p.d_cov = d_cov;
    p.chol_jitter = chol_jitter; p.save2 = save2; p.dXk = dX_knn; p.part = part;
    const int grid = S < nsm ? S : nsm;
    gp_fullcov_bwd_kernel<<<grid, FC_THREADS, smem_bytes, st>>>(p);
    IWVI_CHECK_LAUNCH();
    return IWVI_OK;
  }
  if (!ws) return IWVI_ERR_NULL;
  const int64_t chunk = S < FC_CHUNK ? S : FC_CHUNK;
  double* Lfac = ws;
  double* per_cta = ws + chunk * d->R * (int64_t)N * N;
  const int smem_bytes = (IWVI_STAGE_DOUBLES + IWVI_BLK * FC_LDC + 2 * FC_MAXN * IWVI_MAX_D + 4 * FC_MAXN) * 8;
  if (cudaFuncSetAttribute(gp_fullcov_bwd_large_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes) != cudaSuccess)
    return IWVI_ERR_LAUNCH;
  for (int64_t s0 = 0; s0 < S; s0 += FC_CHUNK) {
    const int sc = (int)(S - s0 < FC_CHUNK ? S - s0 : FC_CHUNK);
    if (d_sample) {   // the Cholesky factors of the chunk, as the forward pass formed them
      rc = fc_launch_cov(d, sc, s0, N, aux, X, save, nullptr, nullptr, chol_jitter, Lfac, s0, nullptr, nullptr, nsm, st);
      if (rc != IWVI_OK) return rc;
      rc = fc_launch_chol(Lfac, sc * d->R, N, d->R, s0, chol_jitter, nullptr, nullptr, nullptr, nullptr, nsm, st);
      if (rc != IWVI_OK) return rc;
    }
    FullCovBwdLargeParams p;
    p.d = *d; p.S = sc; p.N = N; p.s0 = s0; p.aux = aux; p.X = X; p.save = save; p.eps = eps; p.d_sample = d_sample;
    p.d_cov = d_cov; p.Lfac = d_sample ? Lfac : nullptr; p.save2 = save2; p.dXk = dX_knn; p.part = part; p.ws = per_cta;
    const int grid = sc < nsm ? sc : nsm;
    gp_fullcov_bwd_large_kernel<<<grid, FC_THREADS, smem_bytes, st>>>(p);
    IWVI_CHECK_LAUNCH();
  }
  return IWVI_OK;
}
