// gp_prologue.cu -- once-per-layer-per-step stage of a GPLayer and its adjoint.
//
// Forward (reference temp_workaround.py:39 Kuu + jitter, :48 tf.cholesky, layers.py:44 -> gauss_kl whitened):
//   gp_pack_kernel   length-scaled inducing inputs Zt and their norms, padded q_mu, constants (8 CTAs); one CTA per lower
//                    block of every tril(q_sqrt_r) (block-major padded copy); per-CTA partial sums of the whitened KL.
//   gp_chol_kernel   blocked Cholesky of Kuu + jitter*I with 64x64 blocks as a DATAFLOW over NB CTAs, one per block row,
//                    handing finished blocks to the rows below through global memory (acquire/release counters).  The
//                    gram blocks are produced on the fly (-2 Z Z^T GEMM + norm epilogue + kernel function, DMMA), the
//                    block updates and the panel scaling S(i,k) Dinv_k^T run on the FP64 tensor pipe, the 64x64 diagonal
//                    blocks are factorised in shared memory with 8-wide panels and inverted by recursive doubling.
//                    The last row also forms the final KL sum.
// Backward (the reference: tf.gradients; TF CholeskyGrad = Phi-form below):
//   pbwd_phi_kernel    P = Phi(Lm^T tril(dLm))                     one 64x64 block per CTA, DMMA
//   pbwd_solve_kernel  Out = In Lm^-1 (blocked back substitution on 32-row panels, same pipeline as the row stage);
//                      run twice: S1^T = P^T Lm^-1, then Kbar' = S1 Lm^-1
//   pbwd_gram_kernel   Kbar = sym(Kbar'), gram adjoint of Kuu (dZ, per-row partials of dls, dvariance)
//   pbwd_kl_kernel     adjoint of the whitened KL (adds into dq_mu, dq_sqrt)
//   pbwd_final_kernel  fixed-order sums of the per-row partials
#include "common.cuh"

namespace {

struct ProParams {
  iwvi_gp_desc d;
  const double *Z, *ls, *variance, *q_mu, *q_sqrt;
  double *Lm, *aux, *kl;
  int32_t* info;
  int mode;   // 0: everything; 1 (IWVI_FLAG_PRO_HYP): what depends on Z / kernel parameters; 2 (IWVI_FLAG_PRO_Q): on q_mu / q_sqrt
};

// ------------------------------------------------------------------------------------------------
// pack
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gp_pack_kernel(const ProParams p) {
  __shared__ double red[32];
  const iwvi_gp_desc& d = p.d;
  const AuxLayout al = iwvi_aux_layout(d.M, d.D, d.R);
  const int M = d.M, D = d.D, R = d.R, Mp = al.Mp, ldz = al.ldz;
  double* aux = p.aux;
  double klp = 0.0;
  if (blockIdx.x >= IWVI_PACK_SMALL) {
    // one CTA per lower 64x64 block of tril(q_sqrt_r): block-major padded [64][68] copy; trace and log-det terms of the KL
    const int blk = blockIdx.x - IWVI_PACK_SMALL;
    const int r = blk / al.npairs, pr = blk - r * al.npairs;
    int bi = 0;
    while ((bi + 1) * (bi + 2) / 2 <= pr) bi++;
    const int bj = pr - bi * (bi + 1) / 2;
    double* dst = aux + al.off_lqb + (size_t)blk * IWVI_STAGE_DOUBLES;
    for (int e = threadIdx.x; e < IWVI_STAGE_DOUBLES; e += 256) {
      const int row = e / IWVI_LDS, col = e - row * IWVI_LDS;
      const int a = bi * IWVI_BLK + row, b = bj * IWVI_BLK + col;
      double v = 0.0;
      if (col < IWVI_BLK && a < M && b <= a) {
        v = p.q_sqrt[((size_t)r * M + a) * M + b];
        klp += v * v;
        if (a == b) klp -= log(v * v);
      }
      dst[e] = v;
    }
  } else {
    const int gtid = blockIdx.x * 256 + threadIdx.x, gsz = IWVI_PACK_SMALL * 256;
    if (p.mode != 2) {
      // Zt = Z / ls
      for (int e = gtid; e < Mp * ldz; e += gsz) {
        const int m = e / ldz, k = e - m * ldz;
        aux[al.off_zt + e] = (m < M && k < D) ? p.Z[(size_t)m * D + k] / p.ls[k] : 0.0;
      }
      for (int m = gtid; m < Mp; m += gsz) {
        double s = 0.0;
        if (m < M)
          for (int k = 0; k < D; k++) { const double v = p.Z[(size_t)m * D + k] / p.ls[k]; s += v * v; }
        aux[al.off_zn + m] = s;
      }
      if (gtid < 64) {
        double v = 0.0;
        if (gtid == IWVI_C_VARIANCE) v = p.variance[0];
        else if (gtid >= IWVI_C_INVLS && gtid < IWVI_C_INVLS + D) v = 1.0 / p.ls[gtid - IWVI_C_INVLS];
        aux[al.off_consts + gtid] = v;
      }
      if (gtid < 16) reinterpret_cast<int*>(aux + al.off_prog)[gtid] = 0;   // Cholesky hand-off counters
      if (gtid == 0) p.info[0] = 0;
    }
    if (p.mode == 1) return;      // the q_mu / q_sqrt dependent half (and its KL partials) belongs to the other call
    // q_mu padded to [Mp, 8]; mahalanobis term of the KL
    for (int e = gtid; e < Mp * IWVI_MAX_R; e += gsz) {
      const int m = e / IWVI_MAX_R, r = e - m * IWVI_MAX_R;
      double v = 0.0;
      if (m < M && r < R) { v = p.q_mu[(size_t)m * R + r]; klp += v * v; }
      aux[al.off_qmu + e] = v;
    }
  }
  const double tot = block_sum(klp, red);
  if (threadIdx.x == 0) aux[al.off_scratch + blockIdx.x] = tot;
}

// final KL sum in the fixed order of the pack launch's CTAs (what the last Cholesky row does when both halves run together)
__global__ void gp_kl_sum_kernel(const ProParams p) {
  const iwvi_gp_desc& d = p.d;
  const AuxLayout al = iwvi_aux_layout(d.M, d.D, d.R);
  if (threadIdx.x == 0) {
    double s = 0.0;
    const int nparts = IWVI_PACK_SMALL + d.R * al.npairs;
    for (int q = 0; q < nparts; q++) s += p.aux[al.off_scratch + q];
    p.kl[0] = 0.5 * (s - (double)d.M * (double)d.R);
  }
}

// ------------------------------------------------------------------------------------------------
// Cholesky
// ------------------------------------------------------------------------------------------------
// copy a 64x64 block (row-major, leading dimension ld) from global into a padded shared stage
__device__ __forceinline__ void load_block(double* dst, const double* src, int ld, int tid) {
  // all 16 loads of a thread are issued before the first store: one memory latency per block, not sixteen
  double v[16];
#pragma unroll
  for (int q = 0; q < 16; q++) v[q] = src[(size_t)((tid >> 6) + 4 * q) * ld + (tid & 63)];
#pragma unroll
  for (int q = 0; q < 16; q++) dst[((tid >> 6) + 4 * q) * IWVI_LDS + (tid & 63)] = v[q];
}
// the same through L2 (data another CTA of the running grid has just published)
__device__ __forceinline__ void load_block_cg(double* dst, const double* src, int ld, int tid) {
  double v[16];
#pragma unroll
  for (int q = 0; q < 16; q++) v[q] = __ldcg(src + (size_t)((tid >> 6) + 4 * q) * ld + (tid & 63));
#pragma unroll
  for (int q = 0; q < 16; q++) dst[((tid >> 6) + 4 * q) * IWVI_LDS + (tid & 63)] = v[q];
}

template <int KIND>
__global__ void __launch_bounds__(256, 1) gp_chol_kernel(const ProParams p) {
  extern __shared__ __align__(16) double smem[];
  const iwvi_gp_desc& d = p.d;
  const AuxLayout al = iwvi_aux_layout(d.M, d.D, d.R);
  const int M = d.M, Mp = al.Mp, NB = al.NB, ldz = al.ldz;
  const int Dk = iwvi_round_up(d.D, 4);
  double* bufA = smem;                              // [64][68]
  double* bufB = bufA + IWVI_STAGE_DOUBLES;         // [64][68]
  double* S = bufB + IWVI_STAGE_DOUBLES;            // [64][68] block being formed (row-major) / X of the inversion
  double* St = S + IWVI_STAGE_DOUBLES;              // [64][68] column-major copy used by the diagonal factorisation
  double* Dv = St + IWVI_STAGE_DOUBLES;             // [64][68] inverted diagonal block of the current column
  double* zi = Dv + IWVI_STAGE_DOUBLES;             // [64][ldz]
  double* zk = zi + IWVI_BLK * 36;                  // [64][ldz]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int wm0 = (warp & 1) * 32, wn0 = (warp >> 1) * 16;
  const double* aux = p.aux;
  const double* Zt = aux + al.off_zt;
  const double* zn = aux + al.off_zn;
  const double variance = p.variance[0];
  const double jitter = d.jitter;
  double* Lm = p.Lm;
  double* Lmb = p.aux + al.off_lmb;   // block-major padded copies: L(i,k) for i > k, inverted diagonal blocks at (k,k)
  // Dataflow schedule: CTA i owns block row i and forms its blocks (i,0) .. (i,i) left to right.  prog[r] counts the
  // blocks row r has published; block (i,k) needs L(k,j), j < k (prog[k] >= j + 1) and, off the diagonal, the inverted
  // diagonal block of row k (prog[k] >= k + 1).  Row i only ever waits for rows above it, and block rows are handed out by
  // an atomic TICKET (prog[15], zeroed by the pack kernel) rather than by blockIdx: the CTA that owns row k has started
  // before any CTA that could wait on it, whatever order the hardware dispatches blocks in and however few of the NB <= 8
  // CTAs are resident at once (each needs ~211 KB of shared memory, so it never shares an SM), so the waits cannot deadlock.
  // The critical path is the chain of diagonal blocks, not a serial sweep.
  int* prog = reinterpret_cast<int*>(p.aux + al.off_prog);
  __shared__ int s_row;
  if (threadIdx.x == 0) s_row = atomicAdd(prog + 15, 1);
  __syncthreads();
  const int i = s_row;
  auto wait_row = [&](int row, int need) {
    if (tid == 0) while (ld_acquire(prog + row) < need) {}
    __syncthreads();
  };
  for (int idx = tid; idx < IWVI_BLK * ldz; idx += 256) zi[idx] = Zt[(size_t)i * IWVI_BLK * ldz + idx];

  PHASE_DECL;
  // gram block (i, kk) of Kuu + jitter I (identity on the padding) into `out`; the kernel function is evaluated in lock step
  // over the 4 entries of an accumulator row (see kern_n)
  auto gram_block = [&](int kk, double (&out)[4][2][2]) {
    __syncthreads();
    for (int idx = tid; idx < IWVI_BLK * ldz; idx += 256) zk[idx] = Zt[(size_t)kk * IWVI_BLK * ldz + idx];
    __syncthreads();
    acc_zero<4, 2>(out);
    warp_gemm<4, 2, 0, 0>(out, zi + wm0 * ldz, ldz, zk + wn0 * ldz, ldz, Dk, lane);
#pragma unroll
    for (int a = 0; a < 4; a++) {
      const int mg = i * IWVI_BLK + wm0 + a * 8 + g;
      double kv[4], unused[4];
#pragma unroll
      for (int b = 0; b < 2; b++)
#pragma unroll
        for (int c = 0; c < 2; c++) {
          const int ng = kk * IWVI_BLK + wn0 + b * 8 + 2 * t + c;
          kv[b * 2 + c] = zn[mg] + zn[ng] - 2.0 * out[a][b][c];
        }
      kern_n<KIND, 4, false>(kv, unused, variance);
#pragma unroll
      for (int b = 0; b < 2; b++)
#pragma unroll
        for (int c = 0; c < 2; c++) {
          const int ng = kk * IWVI_BLK + wn0 + b * 8 + 2 * t + c;
          double v = (mg == ng) ? 1.0 : 0.0;
          if (mg < M && ng < M) v = kv[b * 2 + c] + ((mg == ng) ? jitter : 0.0);
          out[a][b][c] = v;
        }
    }
  };
  // The row's own diagonal block is on the critical path of every row below it: its gram block is formed up front and each
  // L(i,k) is folded into it (acc_diag -= L(i,k) L(i,k)^T) right after it has been published, while the row would otherwise
  // be waiting -- so when column i is reached the block only has to be factorised.
  double acc[4][2][2], acc_next[4][2][2], acc_diag[4][2][2];
  gram_block(i, acc_diag);
  if (i > 0) gram_block(0, acc_next);
  for (int k = 0; k <= i; k++) {
    {
#pragma unroll
      for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 2; b++)
#pragma unroll
          for (int c = 0; c < 2; c++) acc[a][b][c] = (k == i) ? acc_diag[a][b][c] : acc_next[a][b][c];
      PHASE_MARK(0);
      // left-looking update of an off-diagonal block: -= sum_{j<k} L(i,j) L(k,j)^T
      for (int j = 0; j < k && k < i; j++) {
        __syncthreads();
        wait_row(k, j + 1);                 // L(k,j) is another row's block
        load_block(bufA, Lm + (size_t)(i * IWVI_BLK) * Mp + j * IWVI_BLK, Mp, tid);
        load_block_cg(bufB, Lm + (size_t)(k * IWVI_BLK) * Mp + j * IWVI_BLK, Mp, tid);
        __syncthreads();
        double upd[4][2][2];
        acc_zero<4, 2>(upd);
        warp_gemm<4, 2, 0, 0>(upd, bufA + wm0 * IWVI_LDS, IWVI_LDS, bufB + wn0 * IWVI_LDS, IWVI_LDS, IWVI_BLK, lane);
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
          for (int b = 0; b < 2; b++)
#pragma unroll
            for (int c = 0; c < 2; c++) acc[a][b][c] -= upd[a][b][c];
      }
      // the next column's gram block does not depend on other rows: form it before waiting for row k's diagonal block
      if (k + 1 < i) gram_block(k + 1, acc_next);
      PHASE_MARK(1);
      if (i == k) {
        // ---- diagonal block: right-looking factorisation in shared memory with 8-wide panels -- (1) one warp factors the
        //      8x8 pivot block (the only strictly serial part: 8 dependent reciprocal square roots), (2) one thread per
        //      row below solves its 8 entries against it, (3) the trailing matrix gets its rank-8 update as 8x8x4 DMMA
        //      tiles -- then inversion by recursive doubling (8 -> 16 -> 32 -> 64) with DMMA tile products.
        __syncthreads();
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
          for (int b = 0; b < 2; b++)
#pragma unroll
            for (int c = 0; c < 2; c++)
              S[(wm0 + a * 8 + g) * IWVI_LDS + wn0 + b * 8 + 2 * t + c] = acc[a][b][c];
        __syncthreads();
        double* rdiag = St;                               // [64] reciprocals of the diagonal of L(k,k)
        for (int p0 = 0; p0 < IWVI_BLK; p0 += 8) {
          double* P = S + p0 * IWVI_LDS + p0;             // pivot block corner
          if (warp == 0) {
            const int r = lane & 7, cq = lane >> 3;        // lane handles (r, cq) and (r, cq + 4) of the 8x8 block
            for (int j = 0; j < 8; j++) {
              const double dj = P[j * IWVI_LDS + j];
              if (lane == 0 && !(dj > 0.0)) atomicCAS(p.info, 0, k * IWVI_BLK + p0 + j + 1);
              const double rs = rsqrt(dj);
              const double lrj = P[r * IWVI_LDS + j] * rs;               // L(r, j) for r >= j (r == j: sqrt(dj))
              const double lc0 = P[cq * IWVI_LDS + j] * rs, lc1 = P[(cq + 4) * IWVI_LDS + j] * rs;
              __syncwarp();
              if (cq == (j & 3) && r >= j) P[r * IWVI_LDS + j] = lrj;    // column j final (one lane per row)
              if (lane == 0) rdiag[p0 + j] = rs;                         // 1 / L(j,j): divisions below become products
              if (cq > j && r >= cq) P[r * IWVI_LDS + cq] -= lrj * lc0;
              if (cq + 4 > j && r >= cq + 4) P[r * IWVI_LDS + cq + 4] -= lrj * lc1;
              __syncwarp();
            }
          }
          __syncthreads();
          const int nb = IWVI_BLK - p0 - 8;                // rows below the pivot block
          if (tid < nb) {                                  // (2) row i of the panel: x L11^T = s
            double* rowp = S + (p0 + 8 + tid) * IWVI_LDS + p0;
            double x[8];
#pragma unroll
            for (int c = 0; c < 8; c++) {
              double s_ = rowp[c];
#pragma unroll
              for (int q = 0; q < 8; q++)
                if (q < c) s_ -= x[q] * P[c * IWVI_LDS + q];
              x[c] = s_ * rdiag[p0 + c];
            }
#pragma unroll
            for (int c = 0; c < 8; c++) rowp[c] = x[c];
          }
          __syncthreads();
          // (3) trailing update, lower 8x8 tiles only: S22 -= X X^T  (X = the panel just solved, K = 8: two DMMAs per tile)
          const int nt = nb >> 3;
          for (int tl = warp; tl < nt * (nt + 1) / 2; tl += 8) {
            int ti = 0;
            while ((ti + 1) * (ti + 2) / 2 <= tl) ti++;
            const int tj = tl - ti * (ti + 1) / 2;
            const double* Xi = S + (p0 + 8 + ti * 8) * IWVI_LDS + p0;
            const double* Xj = S + (p0 + 8 + tj * 8) * IWVI_LDS + p0;
            double* Cc = S + (p0 + 8 + ti * 8 + g) * IWVI_LDS + p0 + 8 + tj * 8 + 2 * t;
            double c2[2] = {-Cc[0], -Cc[1]};
            dmma884(c2, Xi[g * IWVI_LDS + t], Xj[g * IWVI_LDS + t]);
            dmma884(c2, Xi[g * IWVI_LDS + 4 + t], Xj[g * IWVI_LDS + 4 + t]);
            Cc[0] = -c2[0]; Cc[1] = -c2[1];
          }
          __syncthreads();
        }
        PHASE_MARK(2);
        {
          // L(k,k): zero the upper part of S, copy to global; Dv := 0
          for (int idx = tid; idx < IWVI_BLK * IWVI_BLK; idx += 256) {
            const int row = idx >> 6, c = idx & 63;
            const double v = (c <= row) ? S[row * IWVI_LDS + c] : 0.0;
            S[row * IWVI_LDS + c] = v;
            Lm[(size_t)(k * IWVI_BLK + row) * Mp + k * IWVI_BLK + c] = v;
          }
          for (int idx = tid; idx < IWVI_STAGE_DOUBLES; idx += 256) Dv[idx] = 0.0;
        }
        __syncthreads();
        // level 0: the eight 8x8 diagonal sub-blocks, one thread per (sub-block, column)
        if (tid < IWVI_BLK) {
          const int b0 = (tid >> 3) * 8, c = tid & 7;
          double x[8];
#pragma unroll
          for (int i2 = 0; i2 < 8; i2++) {
            double s_ = (i2 == c) ? 1.0 : 0.0;
#pragma unroll
            for (int j = 0; j < 8; j++)
              if (j < i2) s_ -= S[(b0 + i2) * IWVI_LDS + b0 + j] * x[j];
            x[i2] = (i2 < c) ? 0.0 : s_ * rdiag[b0 + i2];
          }
#pragma unroll
          for (int i2 = 0; i2 < 8; i2++) Dv[(b0 + i2) * IWVI_LDS + b0 + c] = x[i2];
        }
        __syncthreads();
        // levels 1..3: inv([L11 0; L21 L22]) = [inv11 0; -inv22 L21 inv11, inv22]
        for (int sz = 8; sz < IWVI_BLK; sz *= 2) {
          const int tpp = (sz / 8) * (sz / 8);            // 8x8 output tiles per pair
          const int ntile = (IWVI_BLK / (2 * sz)) * tpp;
          for (int pass = 0; pass < 2; pass++) {
            // pass 0: bufA(lower-left) = L21 inv11 ; pass 1: Dv(lower-left) = -inv22 bufA(lower-left)
            for (int tl = warp; tl < ntile; tl += 8) {
              const int pr = tl / tpp, rem = tl - pr * tpp;
              const int ti = rem / (sz / 8), tj = rem - ti * (sz / 8);
              const int r0 = pr * 2 * sz;
              const int orow = r0 + sz + ti * 8, ocol = r0 + tj * 8;
              const double* Am = pass == 0 ? S : Dv;       // rows orow.., k-columns: pass 0: r0.., pass 1: r0+sz..
              const double* Bm = pass == 0 ? Dv : bufA;    // k-rows: pass 0: r0.., pass 1: r0+sz.. ; columns ocol..
              const int ak0 = pass == 0 ? r0 : r0 + sz;
              double c2[2] = {0.0, 0.0};
              for (int k0 = 0; k0 < sz; k0 += 4)
                dmma884(c2, Am[(orow + g) * IWVI_LDS + ak0 + k0 + t], Bm[(ak0 + k0 + t) * IWVI_LDS + ocol + g]);
              double* Cm = pass == 0 ? bufA : Dv;
              const double sg = pass == 0 ? 1.0 : -1.0;
              Cm[(orow + g) * IWVI_LDS + ocol + 2 * t] = sg * c2[0];
              Cm[(orow + g) * IWVI_LDS + ocol + 2 * t + 1] = sg * c2[1];
            }
            __syncthreads();
          }
        }
        for (int idx = tid; idx < IWVI_STAGE_DOUBLES; idx += 256) {
          const int c = idx % IWVI_LDS;
          Lmb[(size_t)iwvi_pair(k, k) * IWVI_STAGE_DOUBLES + idx] = (c < IWVI_BLK) ? Dv[idx] : 0.0;
        }
        PHASE_MARK(3);
        // zero the blocks to the right of the diagonal
        for (int jb = k + 1; jb < NB; jb++)
          for (int idx = tid; idx < IWVI_BLK * IWVI_BLK; idx += 256) {
            const int r = idx >> 6, c = idx & 63;
            Lm[(size_t)(k * IWVI_BLK + r) * Mp + jb * IWVI_BLK + c] = 0.0;
          }
      } else {
        // ---- L(i,k) = S(i,k) Dinv_k^T, with row k's inverted diagonal block fetched once it is published
        wait_row(k, k + 1);
        PHASE_MARK(4);
        {
          const double* src = Lmb + (size_t)iwvi_pair(k, k) * IWVI_STAGE_DOUBLES;
          double v[IWVI_STAGE_DOUBLES / 256];                        // 17 loads in flight per thread
#pragma unroll
          for (int q = 0; q < IWVI_STAGE_DOUBLES / 256; q++) v[q] = __ldcg(src + tid + 256 * q);
#pragma unroll
          for (int q = 0; q < IWVI_STAGE_DOUBLES / 256; q++) Dv[tid + 256 * q] = v[q];
        }
        __syncthreads();
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
          for (int b = 0; b < 2; b++)
#pragma unroll
            for (int c = 0; c < 2; c++)
              S[(wm0 + a * 8 + g) * IWVI_LDS + wn0 + b * 8 + 2 * t + c] = acc[a][b][c];
        __syncthreads();
        double out[4][2][2];
        acc_zero<4, 2>(out);
        warp_gemm<4, 2, 0, 0>(out, S + wm0 * IWVI_LDS, IWVI_LDS, Dv + wn0 * IWVI_LDS, IWVI_LDS, IWVI_BLK, lane);
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
          for (int b = 0; b < 2; b++)
#pragma unroll
            for (int c = 0; c < 2; c++) {
              const int rr = wm0 + a * 8 + g, cc = wn0 + b * 8 + 2 * t + c;
              Lm[(size_t)(i * IWVI_BLK + rr) * Mp + k * IWVI_BLK + cc] = out[a][b][c];
              Lmb[(size_t)iwvi_pair(i, k) * IWVI_STAGE_DOUBLES + rr * IWVI_LDS + cc] = out[a][b][c];
              bufA[rr * IWVI_LDS + cc] = out[a][b][c];     // kept for the update of this row's diagonal block below
            }
      }
    }
    PHASE_MARK(5);
    // publish block (i,k)
    __syncthreads();
    if (tid == 0) { __threadfence(); st_release(prog + i, k + 1); }
    if (k < i) {   // fold L(i,k) (in bufA) into the row's diagonal block
      double upd[4][2][2];
      acc_zero<4, 2>(upd);
      warp_gemm<4, 2, 0, 0>(upd, bufA + wm0 * IWVI_LDS, IWVI_LDS, bufA + wn0 * IWVI_LDS, IWVI_LDS, IWVI_BLK, lane);
#pragma unroll
      for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 2; b++)
#pragma unroll
          for (int c = 0; c < 2; c++) acc_diag[a][b][c] -= upd[a][b][c];
    }
  }
  if (i == NB - 1) PHASE_FLUSH(2);
  if (i == NB - 1 && tid == 0 && p.mode == 0) {   // the last row depends on every other row: it finishes last
    double s = 0.0;
    const int nparts = IWVI_PACK_SMALL + d.R * al.npairs;   // CTAs of this layer's pack launch
    for (int q = 0; q < nparts; q++) s += aux[al.off_scratch + q];
    p.kl[0] = 0.5 * (s - (double)M * (double)d.R);
  }
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
struct PbwdWs { int64_t off_p, off_s1, off_dls, off_dvar, total; };
__host__ __device__ inline PbwdWs pbwd_ws_layout(int Mp) {
  PbwdWs w; int64_t o = 0;
  w.off_p = o;   o += (int64_t)Mp * Mp;     // P, later Kbar'
  w.off_s1 = o;  o += (int64_t)Mp * Mp;     // S1^T
  w.off_dls = o; o += (int64_t)Mp * 32;     // per-row partials of dls
  w.off_dvar = o; o += Mp;                  // per-row partials of dvariance
  w.total = o;
  return w;
}

struct PbwdParams {
  iwvi_gp_desc d;
  const double *Lm, *aux, *Z, *ls, *variance, *q_mu, *q_sqrt, *dLm, *dkl;
  double *dZ, *dls, *dvariance, *dq_mu, *dq_sqrt, *ws;
  int accumulate;
};

__global__ void __launch_bounds__(256, 1) pbwd_phi_kernel(const PbwdParams p) {
  extern __shared__ __align__(16) double smem[];
  double* bufA = smem;
  double* bufB = smem + IWVI_STAGE_DOUBLES;
  const AuxLayout al = iwvi_aux_layout(p.d.M, p.d.D, p.d.R);
  const int Mp = al.Mp, NB = al.NB;
  const PbwdWs wl = pbwd_ws_layout(Mp);
  double* P = p.ws + wl.off_p;
  const int bi = blockIdx.x / NB, bj = blockIdx.x % NB;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int wm0 = (warp & 1) * 32, wn0 = (warp >> 1) * 16;
  double acc[4][2][2];
  acc_zero<4, 2>(acc);
  if (bi >= bj) {
    for (int l = bi; l < NB; l++) {
      __syncthreads();
      load_block(bufA, p.Lm + (size_t)(l * IWVI_BLK) * Mp + bi * IWVI_BLK, Mp, tid);
      load_block(bufB, p.dLm + (size_t)(l * IWVI_BLK) * Mp + bj * IWVI_BLK, Mp, tid);
      __syncthreads();
      warp_gemm<4, 2, 1, 1>(acc, bufA + wm0, IWVI_LDS, bufB + wn0, IWVI_LDS, IWVI_BLK, lane);
    }
  }
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int b = 0; b < 2; b++)
#pragma unroll
      for (int c = 0; c < 2; c++) {
        const int m = bi * IWVI_BLK + wm0 + a * 8 + g;
        const int n = bj * IWVI_BLK + wn0 + b * 8 + 2 * t + c;
        double v = acc[a][b][c];
        if (m < n) v = 0.0;
        else if (m == n) v *= 0.5;
        P[(size_t)m * Mp + n] = v;
      }
}

struct SolveSeq {   // blocks of the back substitution: for i = NB-1..0: Lm(j,i) j = i+1..NB-1, then Dinv_i
  int NB;
  const double* Lmb;
  int i, j;
  bool fin;
  __device__ __forceinline__ void init() { i = NB - 1; j = NB; fin = false; }
  __device__ __forceinline__ bool done() const { return fin; }
  __device__ __forceinline__ BlockSrc get() const {
    BlockSrc b;
    b.bytes = IWVI_STAGE_DOUBLES * 8;
    b.src = Lmb + (size_t)(j < NB ? iwvi_pair(j, i) : iwvi_pair(i, i)) * IWVI_STAGE_DOUBLES;
    return b;
  }
  __device__ __forceinline__ void advance() {
    if (j < NB) ++j;
    else { --i; j = i + 1; if (i < 0) fin = true; }
  }
};

#define PS_TP 32
// Out[n][:] = (Lm^-T In_n), In_n = row n of In (TRANS == 0) or column n of In (TRANS == 1); i.e. Out = In' Lm^-1
template <int TRANS>
__global__ void __launch_bounds__(256, 1) pbwd_solve_kernel(const PbwdParams p, const double* In, double* Out) {
  extern __shared__ __align__(16) double smem[];
  const AuxLayout al = iwvi_aux_layout(p.d.M, p.d.D, p.d.R);
  const int Mp = al.Mp, NB = al.NB, ldA = Mp + 4;
  double* panel = smem;                                       // [PS_TP][ldA]
  double* stages = panel + PS_TP * ldA;
  uint64_t* bars = reinterpret_cast<uint64_t*>(stages + IWVI_NST * IWVI_STAGE_DOUBLES);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  // 8 warps: 4 along the 64 block rows (16 each), 2 along the 32 panel rows (16 each)
  const int wm0 = (warp & 3) * 16, wn0 = (warp >> 2) * 16;
  const int n0 = blockIdx.x * PS_TP;

  StagePipe pipe;
  pipe.setup(bars, stages);
  SolveSeq seq;
  seq.NB = NB; seq.Lmb = p.aux + al.off_lmb;
  seq.init();
  pipe.prime(seq, warp, lane);

  if (TRANS) {
    for (int idx = tid; idx < PS_TP * Mp; idx += 256) {
      const int m = idx / PS_TP, n = idx - m * PS_TP;
      panel[n * ldA + m] = In[(size_t)m * Mp + n0 + n];
    }
  } else {
    for (int idx = tid; idx < PS_TP * Mp; idx += 256) {
      const int n = idx / Mp, m = idx - n * Mp;
      panel[n * ldA + m] = In[(size_t)(n0 + n) * Mp + m];
    }
  }
  __syncthreads();

  for (int i = NB - 1; i >= 0; i--) {
    double acc[2][2][2];
    acc_zero<2, 2>(acc);
    for (int j = i + 1; j < NB; j++) {
      const double* st = pipe.wait();
      warp_gemm<2, 2, 1, 0>(acc, st + wm0, IWVI_LDS, panel + wn0 * ldA + j * IWVI_BLK, ldA, IWVI_BLK, lane);
      pipe.release(seq, warp, lane);
    }
    if (i < NB - 1) {
#pragma unroll
      for (int a = 0; a < 2; a++)
#pragma unroll
        for (int b = 0; b < 2; b++)
#pragma unroll
          for (int c = 0; c < 2; c++)
            panel[(wn0 + b * 8 + 2 * t + c) * ldA + i * IWVI_BLK + wm0 + a * 8 + g] -= acc[a][b][c];
      __syncthreads();
    }
    const double* st = pipe.wait();
    acc_zero<2, 2>(acc);
    warp_gemm<2, 2, 1, 0>(acc, st + wm0, IWVI_LDS, panel + wn0 * ldA + i * IWVI_BLK, ldA, IWVI_BLK, lane);
    pipe.release(seq, warp, lane);
#pragma unroll
    for (int a = 0; a < 2; a++)
#pragma unroll
      for (int b = 0; b < 2; b++)
#pragma unroll
        for (int c = 0; c < 2; c++)
          panel[(wn0 + b * 8 + 2 * t + c) * ldA + i * IWVI_BLK + wm0 + a * 8 + g] = acc[a][b][c];
    __syncthreads();
  }
  for (int idx = tid; idx < PS_TP * Mp; idx += 256) {
    const int n = idx / Mp, m = idx - n * Mp;
    Out[(size_t)(n0 + n) * Mp + m] = panel[n * ldA + m];
  }
}

// one warp per inducing point i: gram adjoint of Kuu; grid-stride KL adjoint
__global__ void __launch_bounds__(256) pbwd_gram_kernel(const PbwdParams p) {
  const iwvi_gp_desc& d = p.d;
  const AuxLayout al = iwvi_aux_layout(d.M, d.D, d.R);
  const int M = d.M, D = d.D, R = d.R, Mp = al.Mp, ldz = al.ldz;
  const PbwdWs wl = pbwd_ws_layout(Mp);
  const double* Kp = p.ws + wl.off_p;
  const double* Zt = p.aux + al.off_zt;
  const double* zn = p.aux + al.off_zn;
  const double* consts = p.aux + al.off_consts;
  const double variance = consts[IWVI_C_VARIANCE];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i = blockIdx.x * 8 + warp;
  // the scaled inducing inputs are staged once per CTA (row stride ldz + 1: conflict-free for lanes = consecutive rows)
  extern __shared__ __align__(16) double zs[];
  const int lds = ldz + 1;
  for (int idx = threadIdx.x; idx < Mp * ldz; idx += 256) zs[(idx / ldz) * lds + idx % ldz] = Zt[idx];
  __syncthreads();
  if (i < Mp) {
    double dz[IWVI_MAX_D], dl[IWVI_MAX_D];
#pragma unroll
    for (int k = 0; k < IWVI_MAX_D; k++) { dz[k] = 0.0; dl[k] = 0.0; }
    double dv = 0.0;
    if (i < M) {
      const double* zi = zs + (size_t)i * lds;
      // symmetrised cotangent of this row; the next element's two loads are issued before the current one is used
      auto load_kb = [&](int j) { return (j < M) ? 0.5 * (Kp[(size_t)i * Mp + j] + Kp[(size_t)j * Mp + i]) : 0.0; };
      const double zni = zn[i];
      double kb_next = load_kb(lane);
      for (int j = lane; j < M; j += 32) {
        const double kb = kb_next;
        kb_next = load_kb(j + 32);
        const double* zj = zs + (size_t)j * lds;
        double dot = 0.0;
        for (int k = 0; k < D; k++) dot += zi[k] * zj[k];
        double K, dK;
        kern_k_dk(d.kern, zni + zn[j] - 2.0 * dot, variance, K, dK);
        dv += kb * K;
        if (j != i) {
          const double G = kb * dK;
#pragma unroll
          for (int k = 0; k < IWVI_MAX_D; k++) {
            if (k < D) {
              const double df = zi[k] - zj[k];
              dz[k] += G * df;
              dl[k] += G * df * df;
            }
          }
        }
      }
    }
    dv = warp_sum(dv);
#pragma unroll
    for (int k = 0; k < IWVI_MAX_D; k++) {
      if (k < D) { dz[k] = warp_sum(dz[k]); dl[k] = warp_sum(dl[k]); }
    }
    if (lane == 0) {
      p.ws[wl.off_dvar + i] = dv;
#pragma unroll
      for (int k = 0; k < IWVI_MAX_D; k++) {
        if (k < D) {
          const double il = consts[IWVI_C_INVLS + k];
          p.ws[wl.off_dls + (size_t)i * 32 + k] = -2.0 * il * dl[k];
          if (i < M) {
            const double v = 4.0 * il * dz[k];
            if (p.accumulate) p.dZ[(size_t)i * D + k] += v; else p.dZ[(size_t)i * D + k] = v;
          }
        }
      }
    }
  }
}

// KL adjoint (whitened gauss_kl): dq_mu = dkl q_mu, dLq = dkl (Lq - diag(1/diag Lq)) on the lower triangle
__global__ void __launch_bounds__(256) pbwd_kl_kernel(const PbwdParams p) {
  const int M = p.d.M, R = p.d.R;
  const double dkl = p.dkl[0];
  const int n_sq = R * M * M, n_all = n_sq + M * R;      // R M^2 <= 8 * 512^2 fits an int
  for (int e = blockIdx.x * 256 + threadIdx.x; e < n_all; e += gridDim.x * 256) {
    if (e < n_sq) {
      const int rem = e % (M * M);
      const int a = rem / M, b = rem - a * M;
      double v = 0.0;
      if (b <= a) {
        const double q = p.q_sqrt[e];
        v = dkl * (a == b ? q - 1.0 / q : q);
      }
      if (p.accumulate) p.dq_sqrt[e] += v; else p.dq_sqrt[e] = v;
    } else {
      const int e2 = e - n_sq;
      const double v = dkl * p.q_mu[e2];
      if (p.accumulate) p.dq_mu[e2] += v; else p.dq_mu[e2] = v;
    }
  }
}

// fixed-order sums of the per-row partials of dls and dvariance, one warp per output
__global__ void __launch_bounds__(32) pbwd_final_kernel(const PbwdParams p) {
  const AuxLayout al = iwvi_aux_layout(p.d.M, p.d.D, p.d.R);
  const PbwdWs wl = pbwd_ws_layout(al.Mp);
  const int k = blockIdx.x, lane = threadIdx.x;
  double s = 0.0;
  if (k < p.d.D) {
    for (int i = lane; i < p.d.M; i += 32) s += p.ws[wl.off_dls + (size_t)i * 32 + k];
    s = warp_sum(s);
    if (lane == 0) { if (p.accumulate) p.dls[k] += s; else p.dls[k] = s; }
  } else {
    for (int i = lane; i < p.d.M; i += 32) s += p.ws[wl.off_dvar + i];
    s = warp_sum(s) / p.aux[al.off_consts + IWVI_C_VARIANCE];
    if (lane == 0) { if (p.accumulate) p.dvariance[0] += s; else p.dvariance[0] = s; }
  }
}

// standalone whitened KL (reference temp_workaround.py:167-188 -> gpflow gauss_kl with K=None) for operator-level callers;
// the training path gets the same number from gp_pack_kernel/gp_chol_kernel.  One CTA, fixed-order reduction.
__global__ void __launch_bounds__(1024) gauss_kl_fwd_kernel(int M, int R, const double* q_mu, const double* q_sqrt, double* kl) {
  __shared__ double red[32];
  double s = 0.0;
  for (int64_t e = threadIdx.x; e < (int64_t)R * M * M; e += blockDim.x) {
    const int64_t rem = e % ((int64_t)M * M);
    const int a = (int)(rem / M), b = (int)(rem - (int64_t)a * M);
    if (b <= a) {
      const double v = q_sqrt[e];
      s += v * v;
      if (a == b) s -= log(v * v);
    }
  }
  for (int64_t e = threadIdx.x; e < (int64_t)M * R; e += blockDim.x) s += q_mu[e] * q_mu[e];
  const double tot = block_sum(s, red);
  if (threadIdx.x == 0) kl[0] = 0.5 * (tot - (double)M * (double)R);
}
__global__ void gauss_kl_bwd_kernel(int M, int R, const double* q_mu, const double* q_sqrt, const double* dkl,
                                    double* dq_mu, double* dq_sqrt) {
  const double g = dkl[0];
  const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, gsz = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = gtid; e < (int64_t)R * M * M; e += gsz) {
    const int64_t rem = e % ((int64_t)M * M);
    const int a = (int)(rem / M), b = (int)(rem - (int64_t)a * M);
    double v = 0.0;
    if (b <= a) { const double q = q_sqrt[e]; v = g * (a == b ? q - 1.0 / q : q); }
    dq_sqrt[e] = v;
  }
  for (int64_t e = gtid; e < (int64_t)M * R; e += gsz) dq_mu[e] = g * q_mu[e];
}

template <int KIND>
int launch_chol(const ProParams& p, int smem_bytes, cudaStream_t st) {
  if (cudaFuncSetAttribute(gp_chol_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes) != cudaSuccess)
    return IWVI_ERR_LAUNCH;
  gp_chol_kernel<KIND><<<iwvi_round_up(p.d.M, IWVI_BLK) / IWVI_BLK, 256, smem_bytes, st>>>(p);
  IWVI_CHECK_LAUNCH();
  return IWVI_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" int iwvi_version(void) { return IWVI_VERSION; }
static thread_local cudaError_t g_last_cuda_error = cudaSuccess;
void iwvi_note_cuda_error(cudaError_t e) { g_last_cuda_error = e; }
extern "C" const char* iwvi_last_cuda_error(void) { return cudaGetErrorName(g_last_cuda_error); }
extern "C" int32_t iwvi_gp_mp(int32_t M) { return iwvi_round_up(M, IWVI_BLK); }
extern "C" int32_t iwvi_gp_lda(int32_t M) { return iwvi_round_up(M, IWVI_BLK) + 4; }
extern "C" int64_t iwvi_gp_aux_doubles(const iwvi_gp_desc* d) {
  if (iwvi_check_gp_desc(d) != IWVI_OK) return -1;
  return iwvi_aux_layout(d->M, d->D, d->R).total;
}
extern "C" int64_t iwvi_gp_save_doubles(const iwvi_gp_desc* d) {
  if (iwvi_check_gp_desc(d) != IWVI_OK) return -1;
  return iwvi_save_layout(d->T, d->M, d->R).total;
}
extern "C" int64_t iwvi_gp_pbwd_ws_doubles(const iwvi_gp_desc* d) {
  if (iwvi_check_gp_desc(d) != IWVI_OK) return -1;
  return pbwd_ws_layout(iwvi_round_up(d->M, IWVI_BLK)).total;
}

extern "C" int iwvi_gp_prologue_fwd(const iwvi_gp_desc* d, const double* Z, const double* ls, const double* variance,
                                    const double* q_mu, const double* q_sqrt, double* Lm, double* aux, double* kl,
                                    int32_t* info, void* stream) {
  int rc = iwvi_check_gp_desc(d);
  if (rc != IWVI_OK) return rc;
  if (!Z || !ls || !variance || !q_mu || !q_sqrt || !Lm || !aux || !kl || !info) return IWVI_ERR_NULL;
  ProParams p;
  p.d = *d; p.Z = Z; p.ls = ls; p.variance = variance; p.q_mu = q_mu; p.q_sqrt = q_sqrt;
  p.Lm = Lm; p.aux = aux; p.kl = kl; p.info = info;
  const bool hyp = (d->flags & IWVI_FLAG_PRO_HYP) != 0, qq = (d->flags & IWVI_FLAG_PRO_Q) != 0;
  if (hyp && qq) return IWVI_ERR_BAD_DESC;
  p.mode = hyp ? 1 : (qq ? 2 : 0);
  cudaStream_t st = (cudaStream_t)stream;
  const int pack_grid = hyp ? IWVI_PACK_SMALL : IWVI_PACK_SMALL + d->R * iwvi_aux_layout(d->M, d->D, d->R).npairs;
  gp_pack_kernel<<<pack_grid, 256, 0, st>>>(p);
  IWVI_CHECK_LAUNCH();
  if (qq) {
    gp_kl_sum_kernel<<<1, 32, 0, st>>>(p);
    IWVI_CHECK_LAUNCH();
    return IWVI_OK;
  }
  const int smem_bytes = (5 * IWVI_STAGE_DOUBLES + 2 * IWVI_BLK * 36) * 8;
  switch (d->kern) {
    case IWVI_KERN_RBF: return launch_chol<IWVI_KERN_RBF>(p, smem_bytes, st);
    case IWVI_KERN_MATERN12: return launch_chol<IWVI_KERN_MATERN12>(p, smem_bytes, st);
    case IWVI_KERN_MATERN32: return launch_chol<IWVI_KERN_MATERN32>(p, smem_bytes, st);
    default: return launch_chol<IWVI_KERN_MATERN52>(p, smem_bytes, st);
  }
}

extern "C" int iwvi_gp_prologue_bwd(const iwvi_gp_desc* d, const double* Lm, const double* aux, const double* Z,
                                    const double* ls, const double* variance, const double* q_mu,
                                    const double* q_sqrt, const double* dLm, const double* dkl, double* dZ,
                                    double* dls, double* dvariance, double* dq_mu, double* dq_sqrt, double* ws,
                                    void* stream) {
  int rc = iwvi_check_gp_desc(d);
  if (rc != IWVI_OK) return rc;
  if (!Lm || !aux || !Z || !ls || !variance || !q_mu || !q_sqrt || !dLm || !dkl || !dZ || !dls || !dvariance ||
      !dq_mu || !dq_sqrt || !ws)
    return IWVI_ERR_NULL;
  const AuxLayout al = iwvi_aux_layout(d->M, d->D, d->R);
  const PbwdWs wl = pbwd_ws_layout(al.Mp);
  PbwdParams p;
  p.d = *d; p.Lm = Lm; p.aux = aux; p.Z = Z; p.ls = ls; p.variance = variance; p.q_mu = q_mu; p.q_sqrt = q_sqrt;
  p.dLm = dLm; p.dkl = dkl; p.dZ = dZ; p.dls = dls; p.dvariance = dvariance; p.dq_mu = dq_mu; p.dq_sqrt = dq_sqrt;
  p.ws = ws; p.accumulate = (d->flags & IWVI_FLAG_ACCUM) ? 1 : 0;
  cudaStream_t st = (cudaStream_t)stream;
  const bool only_kl = (d->flags & IWVI_FLAG_ONLY_KL) != 0, skip_kl = (d->flags & IWVI_FLAG_SKIP_KL) != 0;
  if (!only_kl) {
  const int phi_smem = 2 * IWVI_STAGE_DOUBLES * 8;
  if (cudaFuncSetAttribute(pbwd_phi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, phi_smem) != cudaSuccess)
    return IWVI_ERR_LAUNCH;
  pbwd_phi_kernel<<<al.NB * al.NB, 256, phi_smem, st>>>(p);
  IWVI_CHECK_LAUNCH();
  const int smem_bytes = (PS_TP * (al.Mp + 4) + IWVI_NST * IWVI_STAGE_DOUBLES + IWVI_NST) * 8;
  if (cudaFuncSetAttribute(pbwd_solve_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes) != cudaSuccess)
    return IWVI_ERR_LAUNCH;
  // S1^T = P^T Lm^-1 (panel rows = columns of P), then Kbar' = S1 Lm^-1 (panel rows = rows of S1 = columns of S1^T)
  pbwd_solve_kernel<1><<<al.Mp / PS_TP, 256, smem_bytes, st>>>(p, ws + wl.off_p, ws + wl.off_s1);
  IWVI_CHECK_LAUNCH();
  pbwd_solve_kernel<1><<<al.Mp / PS_TP, 256, smem_bytes, st>>>(p, ws + wl.off_s1, ws + wl.off_p);
  IWVI_CHECK_LAUNCH();
  {
    const int gram_smem = al.Mp * (al.ldz + 1) * 8;
    if (cudaFuncSetAttribute(pbwd_gram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, gram_smem) != cudaSuccess)
      return IWVI_ERR_LAUNCH;
    pbwd_gram_kernel<<<(al.Mp + 7) / 8, 256, gram_smem, st>>>(p);
  }
  IWVI_CHECK_LAUNCH();
  }
  if (!skip_kl) {
    const int n_all = d->R * d->M * d->M + d->M * d->R;
    int grid = (n_all + 1023) / 1024;
    if (grid > 592) grid = 592;
    pbwd_kl_kernel<<<grid, 256, 0, st>>>(p);
    IWVI_CHECK_LAUNCH();
  }
  if (!only_kl) {
    pbwd_final_kernel<<<d->D + 1, 32, 0, st>>>(p);
    IWVI_CHECK_LAUNCH();
  }
  return IWVI_OK;
}

extern "C" int iwvi_gauss_kl_fwd(int32_t M, int32_t R, const double* q_mu, const double* q_sqrt, double* kl, void* stream) {
  if (M < 1 || R < 1) return IWVI_ERR_BAD_DESC;
  if (!q_mu || !q_sqrt || !kl) return IWVI_ERR_NULL;
  gauss_kl_fwd_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(M, R, q_mu, q_sqrt, kl);
  IWVI_CHECK_LAUNCH();
  return IWVI_OK;
}

extern "C" int iwvi_gauss_kl_bwd(int32_t M, int32_t R, const double* q_mu, const double* q_sqrt, const double* dkl,
                                 double* dq_mu, double* dq_sqrt, void* stream) {
  if (M < 1 || R < 1) return IWVI_ERR_BAD_DESC;
  if (!q_mu || !q_sqrt || !dkl || !dq_mu || !dq_sqrt) return IWVI_ERR_NULL;
  gauss_kl_bwd_kernel<<<148, 256, 0, (cudaStream_t)stream>>>(M, R, q_mu, q_sqrt, dkl, dq_mu, dq_sqrt);
  IWVI_CHECK_LAUNCH();
  return IWVI_OK;
}
