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
// FP64 tensor pipe (DMMA, one 8-row tile of C per warp), operands staged one 64-wide m-block at a time; the N x N
// Cholesky (N <= 64) is a right-looking factorisation in shared memory, one barrier pair per column.
#include "common.cuh"

namespace {

#define FC_THREADS 256
#define FC_LDC 65      // leading dimension of the N x N matrices in shared memory (odd: column walks are conflict-free)

struct FullCovParams {
  iwvi_gp_desc d;
  int S, N;
  const double *aux, *X, *save, *mean, *eps;
  double chol_jitter;
  double *cov, *sample;
  int* info;
};

// acc[j][c] (+)= sum_m P(8 w + g, m) P(8 j + 2 t + c, m) over the 64 columns of one staged slab P [Np][68]
__device__ __forceinline__ void slab_syrk(double (&acc)[8][2], const double* __restrict__ slab, int warp, int lane, int ntn) {
  const int g = lane >> 2, t = lane & 3;
  const double* ap = slab + (warp * 8 + g) * IWVI_LDS + t;
  const double* bp = slab + g * IWVI_LDS + t;
#pragma unroll 4
  for (int k0 = 0; k0 < IWVI_BLK; k0 += 4) {
    const double a = ap[k0];
#pragma unroll
    for (int j = 0; j < 8; j++)
      if (j < ntn) dmma884(acc[j], a, bp[j * 8 * IWVI_LDS + k0]);
  }
}

__global__ void __launch_bounds__(FC_THREADS) gp_fullcov_fwd_kernel(const FullCovParams p) {
  extern __shared__ __align__(16) double smem[];
  const iwvi_gp_desc& d = p.d;
  const AuxLayout al = iwvi_aux_layout(d.M, d.D, d.R);
  const SaveLayout sv = iwvi_save_layout(d.T, d.M, d.R);
  const int N = p.N, R = d.R, D = d.D, NB = al.NB;
  const int Np = iwvi_round_up(N, 8), ntn = Np / 8;
  double* slab = smem;                               // [64][68] one m-block of A_s or U_{r,s}, rows >= N zero
  double* C0 = slab + IWVI_STAGE_DOUBLES;            // [64][65] k(X_s, X_s) - A_s^T A_s
  double* Cr = C0 + IWVI_BLK * FC_LDC;               // [64][65] C_r, then its Cholesky factor
  double* xs = Cr + IWVI_BLK * FC_LDC;               // [64][D] length-scaled inputs of the group
  double* xn = xs + IWVI_BLK * IWVI_MAX_D;           // [64]
  __shared__ int bad;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const double* consts = p.aux + al.off_consts;
  const double variance = consts[IWVI_C_VARIANCE];
  const bool active = warp * 8 < Np;                 // this warp owns rows 8 warp .. 8 warp + 7 of C

  for (int s = blockIdx.x; s < p.S; s += gridDim.x) {
    const int64_t pt0 = (int64_t)s * N;
    __syncthreads();
    // ---- k(X_s, X_s) with gpflow's expanded squared distance
    for (int idx = tid; idx < N * D; idx += FC_THREADS) {
      const int n = idx / D, k = idx - n * D;
      xs[n * IWVI_MAX_D + k] = p.X[(pt0 + n) * D + k] * consts[IWVI_C_INVLS + k];
    }
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
    // ---- C0 -= A_s^T A_s
    double acc[8][2];
#pragma unroll
    for (int j = 0; j < 8; j++) { acc[j][0] = 0.0; acc[j][1] = 0.0; }
    for (int mb = 0; mb < NB; mb++) {
      __syncthreads();
      for (int idx = tid; idx < IWVI_BLK * IWVI_BLK; idx += FC_THREADS) {
        const int n = idx >> 6, c = idx & 63;
        slab[n * IWVI_LDS + c] = (n < N) ? p.save[sv.off_a + iwvi_blk_off((int)(pt0 + n), mb * IWVI_BLK + c, NB)] : 0.0;
      }
      __syncthreads();
      if (active) slab_syrk(acc, slab, warp, lane, ntn);
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
      // ---- C_r = C0 + U_r^T U_r (accumulators start from this thread's own entries of C0)
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
        for (int idx = tid; idx < IWVI_BLK * IWVI_BLK; idx += FC_THREADS) {
          const int n = idx >> 6, c = idx & 63;
          slab[n * IWVI_LDS + c] =
              (n < N) ? p.save[sv.off_u + (int64_t)r * sv.u_stride + iwvi_blk_off((int)(pt0 + n), mb * IWVI_BLK + c, NB)] : 0.0;
        }
        __syncthreads();
        if (active) slab_syrk(acc, slab, warp, lane, ntn);
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
        double* dst = p.cov + ((int64_t)s * R + r) * N * N;
        for (int idx = tid; idx < N * N; idx += FC_THREADS) {
          const int i = idx / N, j = idx - i * N;
          dst[idx] = Cr[i * FC_LDC + j];
        }
      }
      if (!p.sample) continue;
      // ---- right-looking Cholesky of C_r + chol_jitter I, in place (lower)
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
        const double* z = p.eps + ((int64_t)s * R + r) * N;
        double v = p.mean[(pt0 + tid) * R + r];
        for (int j = 0; j <= tid; j++) v += Cr[tid * FC_LDC + j] * z[j];
        p.sample[(pt0 + tid) * R + r] = v;
      }
    }
  }
}

}  // namespace

extern "C" int iwvi_gp_fullcov_fwd(const iwvi_gp_desc* d, int32_t S, int32_t N, const double* aux, const double* X,
                                   const double* save, const double* mean, const double* eps, double chol_jitter,
                                   double* cov, double* sample, int32_t* info, void* stream) {
  if (!d) return IWVI_ERR_NULL;
  if (d->T < 0 || d->M < 1 || d->D < 1 || d->R < 1) return IWVI_ERR_BAD_DESC;
  if (d->M > IWVI_MAX_M || d->D > IWVI_MAX_D || d->R > IWVI_MAX_R) return IWVI_ERR_UNSUPPORTED;
  if (d->mix || d->P != d->R) return IWVI_ERR_BAD_DESC;            // the Mok branch forces full_cov=False (:125-129)
  if (S < 0 || N < 1 || (int64_t)S * N != d->T) return IWVI_ERR_BAD_DESC;
  if (N > IWVI_BLK) return IWVI_ERR_UNSUPPORTED;
  if (!aux || !X || !save) return IWVI_ERR_NULL;
  if (sample && (!eps || !mean)) return IWVI_ERR_NULL;
  if (!cov && !sample) return IWVI_ERR_NULL;
  if (S == 0) return IWVI_OK;
  int dev = 0, nsm = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
  const int smem_bytes = (IWVI_STAGE_DOUBLES + 2 * IWVI_BLK * FC_LDC + IWVI_BLK * IWVI_MAX_D + IWVI_BLK) * 8;
  if (cudaFuncSetAttribute(gp_fullcov_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes) != cudaSuccess)
    return IWVI_ERR_LAUNCH;
  FullCovParams p;
  p.d = *d; p.S = S; p.N = N; p.aux = aux; p.X = X; p.save = save; p.mean = mean; p.eps = eps;
  p.chol_jitter = chol_jitter; p.cov = cov; p.sample = sample; p.info = info;
  const int grid = S < 2 * nsm ? S : 2 * nsm;
  gp_fullcov_fwd_kernel<<<grid, FC_THREADS, smem_bytes, (cudaStream_t)stream>>>(p);
  IWVI_CHECK_LAUNCH();
  return IWVI_OK;
}
