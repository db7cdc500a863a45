// gp_rows_bwd.cu -- adjoint of the per-point stage (gp_rows_fwd.cu).  The reference obtains it from tf.gradients
// through temp_workaround.py:44-91,142-145 and layers.py:46-48; the formulas are derived in DESIGN.md and checked
// against autograd in tests/test_staged.py.
//
// Five launches (the host may run 1-3 and 4-5 on different streams, see IWVI_FLAG_ONLY_* / IWVI_FLAG_PART_*):
//   1 gp_epi_bwd_kernel   32 points per CTA: un-mix the cotangents (gmean_bar, gvar_bar), mean-function part of dX
//                         (skinny DMMA), partial sums of dW / dmfA / dmfb / dvariance.
//   2 gp_tile_bwd_kernel  persistent, one tile of TP points per CTA iteration, shared-memory resident panel
//                         (block-major), producer warp + TMA ring as in the forward kernel:
//        Abar/2 = q_mu gmean_bar^T / 2 - A (sum_r gvar_bar_r) + sum_r tril(Lq_r) (U_r * gvar_bar_r)
//        Bbar/2 = Lm^-T Abar/2                           (blocked back substitution, in place; stored for 3 and 4; the
//                                                         exact factor 2 is restored where Bbar is consumed)
//   3 gp_gram_bwd_kernel  (gp_gram_bwd.cu) strips of 32 points, two CTAs per SM:
//        G      = Bbar * dK/dr2 ; dX += 2/ls (x~ colsum(G) - G^T z~) ; dZ, dls, dvariance partials per CTA
//   4 gp_reduce_bwd_kernel  contractions over the T points, split-K, one 64x64 output block per CTA (diagonal blocks:
//                         lower triangle only):
//        dLq_r = 2 tril(A diag(gvar_bar_r) U_r^T),  dLm = -tril(Bbar A^T),  dq_mu = A gmean_bar
//   5 gp_finalize_bwd_kernel  fixed-order sums of all partials (deterministic; the only atomics are fire-and-forget
//                         adds to addresses owned by a single thread).
#include <stdlib.h>
#include "common.cuh"
#include "gp_bwd.cuh"

namespace {

template <int TP> struct TileCfg {
  static constexpr int NW = 8;
  static constexpr int WNG = TP >= 64 ? 4 : 2;
  static constexpr int WMG = NW / WNG;
  static constexpr int WM = IWVI_BLK / WMG;
  static constexpr int WN = TP / WNG;
  static constexpr int TM = WM / 8;
  static constexpr int TN = WN / 8;
};

#define RED_NST 6          // ring depth of the reduce kernel: two stages per 64-point chunk step, three steps in flight
#define TILE_THREADS IWVI_WS_THREADS    // 8 consumer warps + the producer warpgroup (1 active warp)
#define BAR_ALL 1
#define BAR_COL 2


// ------------------------------------------------------------------------------------------------
// 1. per-point epilogue adjoint
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gp_epi_bwd_kernel(const BwdParams p) {
  // One CTA per EPI_PTS = 32 points.  The cotangent tiles are loaded coalesced into shared memory; thread (point, r)
  // un-mixes them, thread (point, k) forms the mean-function part of dX, and (only for trainable W / mean function)
  // each thread sums one entry of dW / dmfA / dmfb over the CTA's points.
  __shared__ double ds[EPI_PTS][IWVI_MAX_P + 1], dm[EPI_PTS][IWVI_MAX_P + 1], dv[EPI_PTS][IWVI_MAX_P + 1];
  __shared__ double red[32];
  const iwvi_gp_desc& d = p.d;
  const int T = d.T, R = d.R, P = d.P, D = d.D;
  const SaveLayout sv = iwvi_save_layout(T, d.M, R);
  const bool sampled = (d.flags & IWVI_FLAG_SAMPLE) != 0;
  const double* gvar = p.save + sv.off_gvar;
  const double* gmean = p.save + sv.off_gmean;
  double* gmb = p.ws + p.wl.off_gmb;
  double* gvb = p.ws + p.wl.off_gvb;
  const int tid = threadIdx.x;
  const int epi_cta = blockIdx.x + p.epi0;
  const int p0 = epi_cta * EPI_PTS;
  const int npts = max(0, min(EPI_PTS, T - p0));     // real points of this CTA (the rest are zero padding)
  for (int idx = tid; idx < EPI_PTS * P; idx += 256) {
    const int n = idx / P, q = idx - n * P;
    const bool ok = n < npts;
    const size_t g_ = (size_t)p0 * P + idx;
    ds[n][q] = (ok && p.d_sample) ? p.d_sample[g_] : 0.0;
    dm[n][q] = (ok && p.d_mean) ? p.d_mean[g_] : 0.0;
    dv[n][q] = (ok && p.d_var) ? p.d_var[g_] : 0.0;
  }
  __syncthreads();
  double gv_sum = 0.0;
  {
    const int n = tid >> 3, r = tid & 7;
    const size_t pt = (size_t)p0 + n;
    double gm_ = 0.0, gv_ = 0.0;
    if (n < npts && r < R) {
      double gsb = 0.0;
      if (d.mix) {
        for (int q = 0; q < P; q++) {
          const double w = p.W[q * R + r];
          gsb += ds[n][q] * w; gm_ += dm[n][q] * w; gv_ += dv[n][q] * w * w;
        }
      } else {
        gsb = ds[n][r]; gm_ = dm[n][r]; gv_ = dv[n][r];
      }
      gm_ += gsb;
      if (sampled) gv_ += gsb * p.eps[pt * R + r] / (2.0 * sqrt(gvar[pt * R + r]));
      gv_sum = gv_;
    }
    if (pt < (size_t)p.wl.Tp) {
      gmb[pt * IWVI_MAX_R + r] = gm_;
      gvb[pt * IWVI_MAX_R + r] = gv_;
      p.ws[p.wl.off_gvt + (size_t)r * p.wl.Tp + pt] = 2.0 * gv_;
    }
  }
  // mean-function part of dX (the gram part is added by the tile kernel): Identity copies, Linear is the skinny product
  // (d_sample + d_mean) [32, P] x mfA^T [P, D] on the tensor pipe, one 8-point row tile per warp
  if (d.mf == IWVI_MF_LINEAR) {
    if (tid < 128) {
      const int lane = tid & 31, g = lane >> 2, t = lane & 3, n = (tid >> 5) * 8 + g;
      double acc[4][2] = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}};
      for (int q0 = 0; q0 < P; q0 += 4) {
        const int q = q0 + t;
        const double a = (q < P) ? ds[n][q] + dm[n][q] : 0.0;
#pragma unroll
        for (int b = 0; b < 4; b++) {
          const int k = b * 8 + g;
          dmma884(acc[b], a, (k < D && q < P) ? p.mfA[k * P + q] : 0.0);
        }
      }
      if (n < npts) {
#pragma unroll
        for (int b = 0; b < 4; b++)
#pragma unroll
          for (int c = 0; c < 2; c++)
            if (b * 8 + 2 * t + c < D) p.dX[((size_t)p0 + n) * D + b * 8 + 2 * t + c] = acc[b][c];
      }
    }
  } else {
    for (int idx = tid; idx < npts * D; idx += 256) {
      const int n = idx / D, k = idx - n * D;
      p.dX[(size_t)p0 * D + idx] = (d.mf == IWVI_MF_IDENTITY) ? ds[n][k] + dm[n][k] : 0.0;
    }
  }
  double* part = p.ws + p.wl.off_epi + (size_t)epi_cta * EPI_STRIDE;
  // (the Kdiag term of dvariance: d fvar / d variance = 1 per point and output; absent when the caller handles the prior
  //  covariance k(X, X) itself, IWVI_FLAG_NO_KDIAG -- the full-covariance adjoint in gp_fullcov.cu)
  const double tot = block_sum(gv_sum, red);
  if (tid == 0) part[P * R + D * P + P] = (d.flags & IWVI_FLAG_NO_KDIAG) ? 0.0 : tot;
  // partial sums over this CTA's points of dW, dmfA, dmfb
  const int nW = (d.mix && p.dW) ? P * R : 0;
  const int nA = (d.mf == IWVI_MF_LINEAR && p.dmfA) ? D * P : 0;
  const int nb = (d.mf == IWVI_MF_LINEAR && p.dmfb) ? P : 0;
  for (int e = tid; e < P * R + D * P + P; e += 256) {
    double s = 0.0;
    if (e < P * R) {
      if (nW) {
        const int q = e / R, r = e - q * R;
        const double w = p.W[q * R + r];
        for (int n = 0; n < npts; n++) {
          const size_t pt = (size_t)p0 + n;
          const double gvv = gvar[pt * R + r], gmm = gmean[pt * R + r];
          const double gss = sampled ? gmm + p.eps[pt * R + r] * sqrt(gvv) : 0.0;
          s += ds[n][q] * gss + dm[n][q] * gmm + 2.0 * dv[n][q] * w * gvv;
        }
      }
    } else if (e < P * R + D * P) {
      if (nA) {
        const int e2 = e - P * R;
        const int k = e2 / P, q = e2 - k * P;
        for (int n = 0; n < npts; n++) s += p.X[((size_t)p0 + n) * D + k] * (ds[n][q] + dm[n][q]);
      }
    } else if (nb) {
      const int q = e - P * R - D * P;
      for (int n = 0; n < npts; n++) s += ds[n][q] + dm[n][q];
    }
    part[e] = s;
  }
}

// ------------------------------------------------------------------------------------------------
// 2. tile kernel
// ------------------------------------------------------------------------------------------------
struct BwdSeq {
  int NB, R, npairs, tp_bytes;
  const double *Lmb, *Lqb, *A_T, *U_T;
  int64_t u_stride;
  int64_t tile_off;   // offset of this tile's rows inside block 0 of its 64-point chunk (block-major saved arrays)
  int ph, r, i, j;
  bool a_turn;        // first block column of r == 0 only: the saved A block i goes ahead of tril(q_sqrt_0) block (i, 0)
  __device__ __forceinline__ void init(int n0) {
    ph = 0; r = 0; i = -1; j = 0; a_turn = false;
    tile_off = (int64_t)(n0 >> 6) * NB * IWVI_STAGE_DOUBLES + (int64_t)(n0 & 63) * IWVI_LDS;
  }
  __device__ __forceinline__ bool done() const { return ph == 2; }
  __device__ __forceinline__ BlockSrc get() const {
    BlockSrc b;
    b.bytes = IWVI_STAGE_DOUBLES * 8;
    if (ph == 0) {
      if (i < j) {           // saved U_r, m-block j, rows of this tile
        b.src = U_T + (int64_t)r * u_stride + tile_off + (int64_t)j * IWVI_STAGE_DOUBLES; b.bytes = (uint32_t)tp_bytes;
      } else if (a_turn) {   // saved A, m-block i, rows of this tile (contiguous)
        b.src = A_T + tile_off + (int64_t)i * IWVI_STAGE_DOUBLES; b.bytes = (uint32_t)tp_bytes;
      } else {               // tril(q_sqrt_r) block (row block i, col block j), i >= j
        b.src = Lqb + ((size_t)r * npairs + iwvi_pair(i, j)) * IWVI_STAGE_DOUBLES;
      }
    } else {                 // Lm block (row block j, col block i), j > i; j == NB: inverted diagonal block i
      b.src = Lmb + (size_t)(j < NB ? iwvi_pair(j, i) : iwvi_pair(i, i)) * IWVI_STAGE_DOUBLES;
    }
    return b;
  }
  __device__ __forceinline__ void advance() {
    if (ph == 0) {
      if (a_turn) { a_turn = false; return; }
      if (++i == NB) { ++j; i = j - 1; if (j == NB) { ++r; j = 0; i = -1; if (r == R) { ph = 1; i = NB - 1; j = NB; } } }
      a_turn = (ph == 0 && r == 0 && j == 0 && i >= 0);
    } else {
      if (j < NB) ++j;
      else { --i; j = i + 1; if (i < 0) ph = 2; }
    }
  }
};

// acc += Lq_block * V for one streamed 64x64 block of tril(q_sqrt_r) (row-major, A operand) and the register-resident
// V fragments of this warp's 16 points.  DIAG: the block is lower triangular; row tile ti = WMI + WMG * a only needs
// k < 8 (ti + 1), and because WMI is a template parameter the skipped DMMAs vanish at compile time.
// The A fragments are fetched two k-step groups ahead of the DMMAs that consume them, with a warp barrier per group as
// a scheduling fence: left to itself ptxas sinks every LDS next to its two DMMAs and recycles ONE register pair for
// all 64 fragments (SASS: LDS R132 / DMMA / DMMA / LDS R132 ...), which exposes the shared-memory latency once per
// two DMMAs (short-scoreboard stalls were a third of this loop's samples, profiles/ncu_c3_r01e.md).
template <int TM, int WMG, int WMI, bool DIAG>
__device__ __forceinline__ void lq_v_product(double (&acc)[TM][2][2], const double* __restrict__ ap,
                                             const double (&vb)[16][2]) {
  constexpr int G = 2;               // k-steps per group
  constexpr int NG = 16 / G;
  double a[3][G][TM];
  auto need = [](int ks, int a_) { return !DIAG || ks < 2 * (WMI + WMG * a_ + 1); };
  auto fetch = [&](auto buf_c, auto grp_c) {
    constexpr int buf = decltype(buf_c)::value, grp = decltype(grp_c)::value;
#pragma unroll
    for (int kk = 0; kk < G; kk++)
#pragma unroll
      for (int a_ = 0; a_ < TM; a_++)
        if (need(grp * G + kk, a_)) a[buf][kk][a_] = ap[a_ * 8 * WMG * IWVI_LDS + (grp * G + kk) * 4];
  };
  auto step = [&](auto grp_c) {
    constexpr int grp = decltype(grp_c)::value, buf = grp % 3;
    __syncwarp();
    if constexpr (grp + 2 < NG) fetch(std::integral_constant<int, (grp + 2) % 3>(), std::integral_constant<int, grp + 2>());
#pragma unroll
    for (int kk = 0; kk < G; kk++)
#pragma unroll
      for (int a_ = 0; a_ < TM; a_++)
        if (need(grp * G + kk, a_)) {
          dmma884(acc[a_][0], a[buf][kk][a_], vb[grp * G + kk][0]);
          dmma884(acc[a_][1], a[buf][kk][a_], vb[grp * G + kk][1]);
        }
  };
  fetch(std::integral_constant<int, 0>(), std::integral_constant<int, 0>());
  fetch(std::integral_constant<int, 1>(), std::integral_constant<int, 1>());
  step(std::integral_constant<int, 0>()); step(std::integral_constant<int, 1>());
  step(std::integral_constant<int, 2>()); step(std::integral_constant<int, 3>());
  step(std::integral_constant<int, 4>()); step(std::integral_constant<int, 5>());
  step(std::integral_constant<int, 6>()); step(std::integral_constant<int, 7>());
  static_assert(NG == 8, "lq_v_product is written out for eight groups of two k-steps");
}

struct TileSmem { int panel, stages, gmb, gvb, bars, total_doubles; };
__host__ __device__ inline TileSmem tile_smem_layout(int TP, int Mp) {
  TileSmem s; int o = 0;
  s.panel = o;  o += (Mp / IWVI_BLK) * TP * IWVI_LDS;   // block-major [m-block][point][68], see gp_rows_fwd.cu
  s.stages = o; o += IWVI_NST * IWVI_STAGE_DOUBLES;
  s.gmb = o;    o += IWVI_MAX_R * TP;
  s.gvb = o;    o += IWVI_MAX_R * TP;
  s.bars = o;   o += 2 * IWVI_NST;
  s.total_doubles = o;
  return s;
}

template <int TP, int KIND>   // KIND: compile-time kernel family, see gp_rows_fwd.cu
__global__ void __launch_bounds__(TILE_THREADS, 1) gp_tile_bwd_kernel(const BwdParams p) {
  using C = TileCfg<TP>;
  static_assert(C::TN == 2, "register-resident V fragments assume two n-tiles per warp");
  extern __shared__ __align__(16) double smem[];
  const iwvi_gp_desc& d = p.d;
  const AuxLayout al = iwvi_aux_layout(d.M, d.D, d.R);
  const int Mp = al.Mp, NB = al.NB, R = d.R, T = d.T, M = d.M;
  constexpr int PSTR = TP * IWVI_LDS;   // element (point n, m) of m-block b: panel[b * PSTR + n * IWVI_LDS + m]
  const TileSmem sl = tile_smem_layout(TP, Mp);
  double* panel = smem + sl.panel;
  double* gmb_s = smem + sl.gmb;
  double* gvb_s = smem + sl.gvb;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double* aux = p.aux;

  // the 4 pad columns of every panel row travel with the bulk stores: keep them zero (nothing else writes them)
  for (int idx = threadIdx.x; idx < NB * TP * 4; idx += TILE_THREADS) panel[(idx >> 2) * IWVI_LDS + IWVI_BLK + (idx & 3)] = 0.0;
  RingT<IWVI_NST> pipe;
  pipe.setup(reinterpret_cast<uint64_t*>(smem + sl.bars), smem + sl.stages, C::NW);
  if (warp >= C::NW) {
    reg_dealloc<IWVI_PRODUCER_REGS>();
    if (warp > C::NW) return;      // padding of the producer warpgroup
    // producer warp: the block sequence is the same for every tile
    BwdSeq seq;
    seq.NB = NB; seq.R = R; seq.npairs = al.npairs;
    seq.Lmb = aux + al.off_lmb; seq.Lqb = aux + al.off_lqb;
    const SaveLayout svp = iwvi_save_layout(T, M, R);
    seq.A_T = p.save + svp.off_a; seq.U_T = p.save + svp.off_u; seq.u_stride = svp.u_stride;
    seq.tp_bytes = TP * IWVI_LDS * 8;
    for (int tile = p.tile0 + blockIdx.x; tile < p.tile1; tile += gridDim.x) {
      seq.init(tile * TP);
      while (!seq.done()) { pipe.produce(seq.get(), lane); seq.advance(); }
    }
    return;
  }

  reg_alloc<IWVI_CONSUMER_REGS>();
  const int g = lane >> 2, t = lane & 3;
  // row tiles of a 64-row block dealt round-robin to the warps of a column group, order flipped in warps 4-7 (see
  // gp_rows_fwd.cu): balances the work skipped in triangular diagonal blocks across warps and SM sub-partitions
  const int wmi = ((warp >> 2) & 1) ? (C::WMG - 1 - warp % C::WMG) : (warp % C::WMG);
  const int wr0 = wmi * 8;
  constexpr int MR = 8 * C::WMG;
  const int wn0 = (warp / C::WMG) * C::WN;
  const int colbar = BAR_COL + warp / C::WMG;

  const double* qmu = aux + al.off_qmu;
  const SaveLayout sv = iwvi_save_layout(T, M, R);
  const double* A_T = p.save + sv.off_a;
  const double* U_T = p.save + sv.off_u;
  const double* gmb = p.ws + p.wl.off_gmb;
  const double* gvb = p.ws + p.wl.off_gvb;
  double* bbar_T = p.ws + p.wl.off_bbar;

  // Everything a warp touches outside the ring belongs to its COLUMN GROUP (the C::WMG warps that share WN points): panel
  // columns, per-point cotangents, the Bbar stores.  So the tile boundary needs no block-wide barrier: each group waits
  // for its own stores, swaps in its own cotangents (fetched into registers before the previous tile's back
  // substitution, so that their latency is not exposed here) and goes on.
  const bool group_lead = (warp % C::WMG == 0) && lane == 0;
  const int gt = (warp % C::WMG) * 32 + lane;                 // thread index inside the column group
  constexpr int GTH = C::WMG * 32;
  constexpr int NPRE = IWVI_MAX_R * C::WN / GTH;              // cotangent entries per thread and array
  static_assert(NPRE * GTH == IWVI_MAX_R * C::WN, "column-group cotangent loader");
  double pre_m[NPRE], pre_v[NPRE];
  auto prefetch_cot = [&](int tile_) {                        // entries (point wn0 + e / 8, r = e % 8): contiguous
    const size_t base = ((size_t)tile_ * TP + wn0) * IWVI_MAX_R;
#pragma unroll
    for (int q = 0; q < NPRE; q++) {
      asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(pre_m[q]) : "l"(gmb + base + gt + q * GTH));
      asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(pre_v[q]) : "l"(gvb + base + gt + q * GTH));
    }
  };
  if (p.tile0 + (int)blockIdx.x < p.tile1) prefetch_cot(p.tile0 + blockIdx.x);

  PHASE_DECL;
  for (int tile = p.tile0 + blockIdx.x; tile < p.tile1; tile += gridDim.x) {
    const int n0 = tile * TP;
    if (group_lead) bulk_wait_read();     // this group's Bbar stores of the previous tile have read the panel
#pragma unroll
    for (int q = 0; q < NPRE; q++) {
      const int e = gt + q * GTH, n = e / IWVI_MAX_R, rr = e % IWVI_MAX_R;
      gmb_s[rr * TP + wn0 + n] = pre_m[q];
      gvb_s[rr * TP + wn0 + n] = pre_v[q];
    }
    named_bar_sync(colbar, GTH);
    PHASE_MARK(0);

    // ---- Abar / 2 = (q_mu gmean_bar^T) / 2 - A gsum  +  sum_r tril(Lq_r) V_r,  V_r = U_r * gvar_bar_r held as register
    //      B-fragments per k-block j (the saved U_r block comes through the ring, ahead of the tril(q_sqrt) blocks that
    //      multiply it).  The factor 2 of the second term is exact in binary floating point, so the panel carries Abar / 2
    //      until it is restored where Bbar is consumed.  The first term is never written on its own: every thread forms
    //      it for the accumulator entries it owns when block row i is first touched (r == 0, j == 0), from the saved A block
    //      that the producer warp slots into the ring right before tril(q_sqrt_0) block (i, 0) -- its latency hides behind
    //      the previous product instead of being exposed in a streaming pass of its own, and the panel is neither
    //      written nor re-read for it.
    for (int r = 0; r < R; r++) {
      for (int j = 0; j < NB; j++) {
        double vb[16][2];
        {
          const double* su = pipe.wait();
#pragma unroll
          for (int b = 0; b < 2; b++) {
            const int n = wn0 + b * 8 + g;
            const double sc = gvb_s[r * TP + n];
            const double* up = su + n * IWVI_LDS + t;
#pragma unroll
            for (int ks = 0; ks < 16; ks++) vb[ks][b] = up[ks * 4] * sc;
          }
          pipe.release(lane);
        }
        for (int i = j; i < NB; i++) {
          // the accumulators start from this thread's own panel entries (exclusive owner): no read-modify-write pass
          double acc[C::TM][2][2];
          if (r == 0 && j == 0) {
            // (q_mu gmean_bar^T) first, r outermost so that only the accumulators stay live (q_mu in aux is padded to
            // [Mp, 8], gmean_bar to 8 columns, zeros beyond R / M); then the saved A block, which by now has arrived
            acc_zero<C::TM, 2>(acc);
#pragma unroll
            for (int rr = 0; rr < IWVI_MAX_R; rr++) {
              double qa[C::TM], gb[2][2];
#pragma unroll
              for (int a_ = 0; a_ < C::TM; a_++)
                qa[a_] = __ldg(qmu + (size_t)(i * IWVI_BLK + wr0 + a_ * MR + g) * IWVI_MAX_R + rr);
#pragma unroll
              for (int b = 0; b < 2; b++)
#pragma unroll
                for (int c = 0; c < 2; c++) gb[b][c] = gmb_s[rr * TP + wn0 + b * 8 + 2 * t + c];
#pragma unroll
              for (int a_ = 0; a_ < C::TM; a_++)
#pragma unroll
                for (int b = 0; b < 2; b++)
#pragma unroll
                  for (int c = 0; c < 2; c++) acc[a_][b][c] += qa[a_] * gb[b][c];
            }
            const double* sa = pipe.wait();   // saved A, m-block i: sa[n * IWVI_LDS + m]
#pragma unroll
            for (int b = 0; b < 2; b++)
#pragma unroll
              for (int c = 0; c < 2; c++) {
                const int n = wn0 + b * 8 + 2 * t + c;
                double gs_n = 0.0;                      // sum_r gvar_bar_r (columns beyond R hold zeros)
#pragma unroll
                for (int rr = 0; rr < IWVI_MAX_R; rr++) gs_n += gvb_s[rr * TP + n];
#pragma unroll
                for (int a_ = 0; a_ < C::TM; a_++)
                  acc[a_][b][c] = 0.5 * acc[a_][b][c] - sa[n * IWVI_LDS + wr0 + a_ * MR + g] * gs_n;   // Abar / 2, first term
              }
            pipe.release(lane);
          } else {
#pragma unroll
            for (int a_ = 0; a_ < C::TM; a_++)
#pragma unroll
              for (int b = 0; b < 2; b++)
#pragma unroll
                for (int c = 0; c < 2; c++)
                  acc[a_][b][c] = panel[i * PSTR + (wn0 + b * 8 + 2 * t + c) * IWVI_LDS + wr0 + a_ * MR + g];
          }
          const double* st = pipe.wait();
          const double* ap = st + (wr0 + g) * IWVI_LDS + t;
          if (i == j) {   // diagonal block of tril(q_sqrt_r): the structural zeros are skipped (compile-time pattern per wmi)
            if (C::WMG == 2) {
              if (wmi == 0) lq_v_product<C::TM, C::WMG, 0, true>(acc, ap, vb);
              else lq_v_product<C::TM, C::WMG, 1, true>(acc, ap, vb);
            } else {
              if (wmi == 0) lq_v_product<C::TM, C::WMG, 0, true>(acc, ap, vb);
              else if (wmi == 1) lq_v_product<C::TM, C::WMG, 1, true>(acc, ap, vb);
              else if (wmi == 2) lq_v_product<C::TM, C::WMG, 2, true>(acc, ap, vb);
              else lq_v_product<C::TM, C::WMG, 3, true>(acc, ap, vb);
            }
          } else {
            lq_v_product<C::TM, C::WMG, 0, false>(acc, ap, vb);
          }
          pipe.release(lane);
#pragma unroll
          for (int a_ = 0; a_ < C::TM; a_++)
#pragma unroll
            for (int b = 0; b < 2; b++)
#pragma unroll
              for (int c = 0; c < 2; c++) {
                const int n = wn0 + b * 8 + 2 * t + c;
                panel[i * PSTR + n * IWVI_LDS + wr0 + a_ * MR + g] = acc[a_][b][c];
              }
        }
      }
    }
    named_bar_sync(colbar, GTH);
    PHASE_MARK(2);
    if (tile + (int)gridDim.x < p.tile1) prefetch_cot(tile + gridDim.x);   // (the cotangents in shared memory are dead from here on)

    // ---- Bbar / 2 = Lm^-T (Abar / 2), blocked back substitution in place.  The factor 2 is restored where Bbar is
    //      consumed: the gram adjoint below and the reduce kernel's dLm scale.
    for (int i = NB - 1; i >= 0; i--) {
      double acc[C::TM][C::TN][2];
      if (i < NB - 1) {
        // acc = -(rhs_i) + sum_{j>i} L(j,i)^T Bbar_j, then rhs_i := -acc
#pragma unroll
        for (int a = 0; a < C::TM; a++)
#pragma unroll
          for (int b = 0; b < C::TN; b++)
#pragma unroll
            for (int c = 0; c < 2; c++)
              acc[a][b][c] = -panel[i * PSTR + (wn0 + b * 8 + 2 * t + c) * IWVI_LDS + wr0 + a * MR + g];
        for (int j = i + 1; j < NB; j++) {
          const double* st = pipe.wait();
          warp_gemm_pf<C::TM, C::TN, 1, 0, C::WMG>(acc, st + wr0, IWVI_LDS, panel + j * PSTR + wn0 * IWVI_LDS, IWVI_LDS, lane);
          pipe.release(lane);
        }
#pragma unroll
        for (int a = 0; a < C::TM; a++)
#pragma unroll
          for (int b = 0; b < C::TN; b++)
#pragma unroll
            for (int c = 0; c < 2; c++)
              panel[i * PSTR + (wn0 + b * 8 + 2 * t + c) * IWVI_LDS + wr0 + a * MR + g] = -acc[a][b][c];
        named_bar_sync(colbar, C::WMG * 32);
      }
      const double* st = pipe.wait();   // inverted diagonal block i, used transposed
      acc_zero<C::TM, C::TN>(acc);
      warp_gemm_tri<C::TM, C::TN, 1, 0, C::WMG, 0>(acc, st, IWVI_LDS, panel + i * PSTR + wn0 * IWVI_LDS, IWVI_LDS, wmi, lane);
      pipe.release(lane);
      named_bar_sync(colbar, C::WMG * 32);   // every warp of the column group has read the right-hand side
#pragma unroll
      for (int a = 0; a < C::TM; a++)
#pragma unroll
        for (int b = 0; b < C::TN; b++)
#pragma unroll
          for (int c = 0; c < 2; c++)
            panel[i * PSTR + (wn0 + b * 8 + 2 * t + c) * IWVI_LDS + wr0 + a * MR + g] = acc[a][b][c];
      // Block row i of Bbar / 2 is final: this column group's WN points of it go out now (needed by the gram-adjoint
      // and reduce kernels), as one asynchronous TMA store straight from the panel (block-major panel == block-major
      // destination; rows of invalid points are zero by construction) that overlaps with the rows still to be solved.
      fence_async_smem();
      named_bar_sync(colbar, GTH);
      if (group_lead) {
        double* dst = bbar_T + ((int64_t)(n0 >> 6) * NB + i) * IWVI_STAGE_DOUBLES + (int64_t)((n0 & 63) + wn0) * IWVI_LDS;
        bulk_s2g(dst, panel + i * PSTR + wn0 * IWVI_LDS, C::WN * IWVI_LDS * 8);
        bulk_commit();
      }
    }
    PHASE_MARK(3);
  }
  if (group_lead) bulk_wait_read();       // the stores read this CTA's shared memory: it must outlive them
  PHASE_FLUSH(1);
}

// ------------------------------------------------------------------------------------------------
// 3. contractions over the points (split-K)
// ------------------------------------------------------------------------------------------------
struct RedSeq {
  const double *Aop, *Bop;   // block-major [chunk][m-block][64][68], already offset to their m-block
  int NB, c, c1, which;
  __device__ __forceinline__ bool done() const { return c >= c1; }
  __device__ __forceinline__ BlockSrc get() const {
    BlockSrc b;
    b.src = (which == 0 ? Aop : Bop) + (size_t)c * NB * IWVI_STAGE_DOUBLES;
    b.bytes = IWVI_STAGE_DOUBLES * 8;
    return b;
  }
  __device__ __forceinline__ void advance() { if (which == 0) which = 1; else { which = 0; ++c; } }
};

#define RED_SC_SLOTS (RED_NST / 2)   // one scale vector per chunk step in flight
struct ReduceLoopArgs { const double *gvb, *gmb, *sc_s; int q; bool is_lm; int c0, c1, wm0, wn0, lane; };

// The per-point factors 2 gvar_bar_r[k] of dLq_r = 2 tril(A diag(gvar_bar_r) U_r^T) arrive in shared memory with the
// chunk itself (gp_epi_bwd_kernel leaves them transposed, [r][Tp], so that a chunk's 64 factors are ONE 512-byte bulk
// copy signalled on the B block's barrier): no per-lane gather from global memory and no 2 x 16 registers of prefetch
// buffer in the product loop (which had cost it its software pipelining: 0.410 -> 0.383 ms per launch at c3 with the
// factors out of the way).  dLm = -tril(Bbar A^T) (Bbar / 2 is stored) needs the constant -2 only: applied once to the
// finished sums (exact).
template <bool QMU>
__device__ __forceinline__ void reduce_loop(RingT<RED_NST>& pipe, const ReduceLoopArgs& la, double (&acc)[4][2][2],
                                            double (&accq)[4][2]) {
  const int g = la.lane >> 2, t = la.lane & 3;
  for (int c = la.c0; c < la.c1; c++) {
    const double* scs = la.sc_s + ((pipe.it / 2) % RED_SC_SLOTS) * IWVI_BLK + t;
    const double* sa = pipe.wait(0);   // [k = point][m]  -> A operand, k-major
    const double* sb = pipe.wait(1);   // [k = point][n]  -> B operand, k-major (+ the chunk's scale vector)
    const double* ap = sa + t * IWVI_LDS + la.wm0 + g;
    const double* bp = sb + t * IWVI_LDS + la.wn0 + g;
    // The B fragments of HALF a chunk are loaded and scaled in one burst (16 LDS + 16 DMUL, kept in registers), a warp
    // barrier keeps ptxas from sinking the DMULs back between the DMMAs (where they stall the shared FP64 pipe far beyond
    // their own issue time), and the product loop is pure LDS (A fragments) + DMMA.
#pragma unroll
    for (int h = 0; h < 2; h++) {
      double b[8][2];
#pragma unroll
      for (int kk = 0; kk < 8; kk++)
#pragma unroll
        for (int j = 0; j < 2; j++) b[kk][j] = bp[4 * (8 * h + kk) * IWVI_LDS + j * 8];
      if (!la.is_lm) {
#pragma unroll
        for (int kk = 0; kk < 8; kk++) {
          const double sck = scs[4 * (8 * h + kk)];
          b[kk][0] *= sck; b[kk][1] *= sck;
        }
      }
      __syncwarp();
#pragma unroll
      for (int kk = 0; kk < 8; kk++) {
        const int k0 = 4 * (8 * h + kk);
        double a[4];
#pragma unroll
        for (int i = 0; i < 4; i++) a[i] = ap[k0 * IWVI_LDS + i * 8];
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
          for (int j = 0; j < 2; j++) dmma884(acc[i][j], a[i], b[kk][j]);
        if (QMU) {
          const double bq = __ldg(la.gmb + ((size_t)c * IWVI_BLK + t + k0) * IWVI_MAX_R + g);
#pragma unroll
          for (int i = 0; i < 4; i++) dmma884(accq[i], a[i], bq);
        }
      }
      __syncwarp();
    }
    pipe.release(la.lane, 2);
  }
  if (la.is_lm) {
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 2; j++) { acc[i][j][0] *= -2.0; acc[i][j][1] *= -2.0; }
  }
}

// Diagonal output blocks (bi == bj): only the lower triangle of the 64x64 block is needed (dLq_r and dLm are tril'd),
// i.e. 36 of the 64 8x8 tiles.  Warp w (and w + 4) owns row tiles {w, 7 - w} of the tile grid = 9 tiles, and the two
// warps split every 64-point chunk in halves (k split) whose partial sums are combined through shared memory at the
// end: 72 instead of 128 DMMAs per warp per chunk, evenly spread over the four SM sub-partitions.  The per-point scale
// multiplies the two A fragments (2 DMULs per k-step) instead of the up-to-8 B fragments.
template <int WQ, bool QMU>
__device__ __forceinline__ void reduce_diag(RingT<RED_NST>& pipe, const ReduceLoopArgs& la, int warp, double* scratch,
                                            double* out, double* oq) {
  const int g = la.lane >> 2, t = la.lane & 3;
  const int half = warp >> 2;
  constexpr int R1 = WQ, R2 = 7 - WQ, N1 = WQ + 1, N2 = 8 - WQ;
  double acc1[N1][2], acc2[N2][2], accq[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
  for (int j = 0; j < N1; j++) { acc1[j][0] = 0.0; acc1[j][1] = 0.0; }
#pragma unroll
  for (int j = 0; j < N2; j++) { acc2[j][0] = 0.0; acc2[j][1] = 0.0; }
  for (int c = la.c0; c < la.c1; c++) {
    const double* scs = la.sc_s + ((pipe.it / 2) % RED_SC_SLOTS) * IWVI_BLK + 32 * half + t;   // (see reduce_loop)
    const double* sa = pipe.wait(0);
    const double* sb = pipe.wait(1);
    const double* ap = sa + (32 * half + t) * IWVI_LDS + g;
    const double* bp = sb + (32 * half + t) * IWVI_LDS + g;
#pragma unroll
    for (int ks = 0; ks < 8; ks++) {
      const int k0 = 4 * ks;
      const double a1u = ap[k0 * IWVI_LDS + R1 * 8], a2u = ap[k0 * IWVI_LDS + R2 * 8];
      const double sck = la.is_lm ? 1.0 : scs[k0];
      const double a1 = a1u * sck, a2 = a2u * sck;
#pragma unroll
      for (int j = 0; j < N2; j++) {
        const double b = bp[k0 * IWVI_LDS + j * 8];
        if (j < N1) dmma884(acc1[j], a1, b);
        dmma884(acc2[j], a2, b);
      }
      if (QMU) {   // compile-time: a DMMA under a runtime predicate would occupy the pipe even when off
        const double bq = __ldg(la.gmb + ((size_t)c * IWVI_BLK + 32 * half + t + k0) * IWVI_MAX_R + g);
        dmma884(accq[0], a1u, bq);
        dmma884(accq[1], a2u, bq);
      }
    }
    pipe.release(la.lane, 2);
  }
  const double fin = la.is_lm ? -2.0 : 1.0;     // dLm: the constant factor, once (exact)
  // combine the two k halves through the ring memory, once EVERY consumer warp has finished reading its last stages
  named_bar_sync(1, 256);
  double* scr = scratch + (size_t)((warp & 3) * 32 + la.lane) * 24;
  if (half == 1) {
#pragma unroll
    for (int j = 0; j < N1; j++) { scr[2 * j] = acc1[j][0]; scr[2 * j + 1] = acc1[j][1]; }
#pragma unroll
    for (int j = 0; j < N2; j++) { scr[2 * (N1 + j)] = acc2[j][0]; scr[2 * (N1 + j) + 1] = acc2[j][1]; }
    scr[18] = accq[0][0]; scr[19] = accq[0][1]; scr[20] = accq[1][0]; scr[21] = accq[1][1];
  }
  named_bar_sync(1, 256);
  if (half == 0) {
#pragma unroll
    for (int j = 0; j < N1; j++)
#pragma unroll
      for (int c = 0; c < 2; c++) out[(R1 * 8 + g) * IWVI_BLK + j * 8 + 2 * t + c] = fin * (acc1[j][c] + scr[2 * j + c]);
#pragma unroll
    for (int j = 0; j < N2; j++)
#pragma unroll
      for (int c = 0; c < 2; c++) out[(R2 * 8 + g) * IWVI_BLK + j * 8 + 2 * t + c] = fin * (acc2[j][c] + scr[2 * (N1 + j) + c]);
    if (QMU) {
#pragma unroll
      for (int c = 0; c < 2; c++) {
        oq[(R1 * 8 + g) * IWVI_MAX_R + 2 * t + c] = accq[0][c] + scr[18 + c];
        oq[(R2 * 8 + g) * IWVI_MAX_R + 2 * t + c] = accq[1][c] + scr[20 + c];
      }
    }
  }
}

#define RED_THREADS 288   // 8 consumer warps + 1 producer warp
__global__ void __launch_bounds__(RED_THREADS, 1) gp_reduce_bwd_kernel(const BwdParams p) {
  extern __shared__ __align__(16) double smem[];
  const iwvi_gp_desc& d = p.d;
  const AuxLayout al = iwvi_aux_layout(d.M, d.D, d.R);
  const int NB = al.NB, R = d.R;
  const BwdWs& wl = p.wl;
  double* stages = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + RED_NST * IWVI_STAGE_DOUBLES);
  double* sc_s = smem + RED_NST * IWVI_STAGE_DOUBLES + 2 * RED_NST;      // [RED_SC_SLOTS][64] scale vectors

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  // decode the work item.  Rasterisation: the block pair runs fastest, then the matrix q, the point range s slowest, so
  // that the CTAs resident at any time (blockIdx order) all stream the SAME range of points: a 64-point chunk of A, U_r
  // or Bbar is then fetched from DRAM once and served to the other (up to 2 NB (R+1)) readers of it from L2.  With q
  // slowest (round 1) a wave held the ranges of 2-3 matrices only and A was streamed from DRAM once per wave
  // (652 MB per launch against 390 MB of operands at c3, ncu).
  int item = blockIdx.x;
  int pair, q, s;
  if (p.qmu_only) {                                              // grid = S x NB: block column 0 of matrix q_lo
    pair = iwvi_pair(item % NB, 0); q = p.q_lo; s = item / NB;
  } else {
    pair = item % wl.npairs; item /= wl.npairs;
    q = p.q_lo + item % p.q_n;                                   // q < R: dLq_q ; q == R: dLm
    s = item / p.q_n;
  }
  int bi = 0, acc_pairs = 0;
  while (acc_pairs + bi + 1 <= pair) { acc_pairs += bi + 1; bi++; }
  const int bj = pair - acc_pairs;                               // bj <= bi
  const int c0 = s * wl.chunks_per_split;
  const int c1 = min(c0 + wl.chunks_per_split, wl.Tp / IWVI_BLK);

  const SaveLayout sv = iwvi_save_layout(d.T, d.M, R);
  const double* A_T = p.save + sv.off_a;
  const double* U_T = p.save + sv.off_u;
  const double* bbar_T = p.ws + wl.off_bbar;
  const double* gmb = p.ws + wl.off_gmb;
  const double* gvb = p.ws + wl.off_gvb;
  const bool is_lm = (q == R);
  const bool do_qmu = (q == 0 && bj == 0);

  RingT<RED_NST> pipe;
  pipe.setup(bars, stages, 8);
  if (warp == 8) {
    RedSeq seq;
    seq.Aop = (is_lm ? bbar_T : A_T) + (size_t)bi * IWVI_STAGE_DOUBLES;
    seq.Bop = (is_lm ? A_T : U_T + (size_t)q * sv.u_stride) + (size_t)bj * IWVI_STAGE_DOUBLES;
    seq.NB = NB; seq.c = c0; seq.c1 = c1; seq.which = 0;
    const double* gvt = p.ws + wl.off_gvt + (size_t)(is_lm ? 0 : q) * wl.Tp;
    while (!seq.done()) {
      // the B block of a chunk travels with the chunk's 64 per-point factors (dLq_r only)
      if (seq.which == 1 && !is_lm)
        pipe.produce2(seq.get(), sc_s + ((pipe.it / 2) % RED_SC_SLOTS) * IWVI_BLK, gvt + (size_t)seq.c * IWVI_BLK,
                      IWVI_BLK * 8, lane);
      else
        pipe.produce(seq.get(), lane);
      seq.advance();
    }
    return;
  }
  double* out = p.ws + wl.off_red + (((size_t)q * wl.S + s) * wl.npairs + pair) * IWVI_BLK * IWVI_BLK;
  double* oq = p.ws + wl.off_qred + ((size_t)s * NB + bi) * IWVI_BLK * IWVI_MAX_R;
  if (bi == bj) {
    const ReduceLoopArgs la = {gvb, gmb, sc_s, q, is_lm, c0, c1, 0, 0, lane};
    if (do_qmu) {
      switch (warp & 3) {
        case 0: reduce_diag<0, true>(pipe, la, warp, stages, out, oq); break;
        case 1: reduce_diag<1, true>(pipe, la, warp, stages, out, oq); break;
        case 2: reduce_diag<2, true>(pipe, la, warp, stages, out, oq); break;
        default: reduce_diag<3, true>(pipe, la, warp, stages, out, oq); break;
      }
    } else {
      switch (warp & 3) {
        case 0: reduce_diag<0, false>(pipe, la, warp, stages, out, oq); break;
        case 1: reduce_diag<1, false>(pipe, la, warp, stages, out, oq); break;
        case 2: reduce_diag<2, false>(pipe, la, warp, stages, out, oq); break;
        default: reduce_diag<3, false>(pipe, la, warp, stages, out, oq); break;
      }
    }
    return;
  }
  const int g = lane >> 2, t = lane & 3;
  const int wm0 = (warp & 1) * 32, wn0 = (warp >> 1) * 16;     // warp tile 32 x 16 of the 64 x 64 output block

  double acc[4][2][2];
  acc_zero<4, 2>(acc);
  double accq[4][2];
#pragma unroll
  for (int a = 0; a < 4; a++) { accq[a][0] = 0.0; accq[a][1] = 0.0; }

  // The two warps of the few CTAs that also form dq_mu = A gmean_bar take a separate copy of the loop: a DMMA that is
  // merely predicated off still occupies the FP64 pipe for its full 16 cycles (measured: 24 instead of 16 cycles per
  // useful DMMA when the ride-along sat under a predicate in the common loop).
  const ReduceLoopArgs la = {gvb, gmb, sc_s, q, is_lm, c0, c1, wm0, wn0, lane};
  if (do_qmu && wn0 == 0) reduce_loop<true>(pipe, la, acc, accq);
  else reduce_loop<false>(pipe, la, acc, accq);

#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 2; j++)
#pragma unroll
      for (int c = 0; c < 2; c++)
        out[(wm0 + i * 8 + g) * IWVI_BLK + wn0 + j * 8 + 2 * t + c] = acc[i][j][c];
  if (do_qmu && wn0 == 0) {
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int c = 0; c < 2; c++) oq[(wm0 + i * 8 + g) * IWVI_MAX_R + 2 * t + c] = accq[i][c];
  }
}

#include "gp_reduce_fast.cuh"

// ------------------------------------------------------------------------------------------------
// 4. fixed-order sums of the partials
// ------------------------------------------------------------------------------------------------
// fixed-order sum of `count` partials `src[i * stride]` by one warp: lane l adds partials l, l+32, ... in order, then a
// fixed shuffle tree combines the lanes
__device__ __forceinline__ double warp_partial_sum(const double* src, int count, size_t stride, int lane) {
  double s = 0.0;
  for (int i = lane; i < count; i += 32) s += src[(size_t)i * stride];
  return warp_sum(s);
}

struct FinalLayout { int64_t n_blk, n_up, n_qmu, n_elem; int n_z, n_small, n_warp_out; int grid_elem, grid_warp; };
__host__ __device__ inline FinalLayout final_layout(const iwvi_gp_desc& d, const BwdWs& wl) {
  FinalLayout f;
  const AuxLayout al = iwvi_aux_layout(d.M, d.D, d.R);
  f.n_blk = (int64_t)(d.R + 1) * wl.npairs * IWVI_BLK * IWVI_BLK;                   // lower blocks of dLq_r, dLm
  f.n_up = (int64_t)(d.R + 1) * (al.NB * (al.NB - 1) / 2) * IWVI_BLK * IWVI_BLK;    // blocks above the diagonal: zeros
  f.n_qmu = (int64_t)d.M * d.R;
  f.n_elem = f.n_blk + f.n_up + f.n_qmu;
  f.n_z = d.M * d.D;
  f.n_small = d.D + 1 + d.P * d.R + d.D * d.P + d.P;
  f.n_warp_out = f.n_z + f.n_small;
  f.grid_elem = (int)((f.n_elem + 255) / 256);
  f.grid_warp = (f.n_warp_out + 7) / 8;
  return f;
}

__global__ void __launch_bounds__(256) gp_finalize_bwd_kernel(const BwdParams p) {
  const iwvi_gp_desc& d = p.d;
  const AuxLayout al = iwvi_aux_layout(d.M, d.D, d.R);
  const BwdWs& wl = p.wl;
  const int M = d.M, Mp = al.Mp, R = d.R, D = d.D, P = d.P, NB = al.NB, ldz = al.ldz;
  const FinalLayout f = final_layout(d, wl);
  const double* red = p.ws + wl.off_red;
  const double* qred = p.ws + wl.off_qred;
  const double* tile = p.ws + wl.off_tile;
  const double* epi = p.ws + wl.off_epi;
  const int BB = IWVI_BLK * IWVI_BLK;
  if ((int)blockIdx.x < f.grid_elem) {
    // ---- one thread per output element: split-K partials of the reduce kernel
    int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (e < f.n_blk) {
      const int off = (int)(e % BB);
      const int64_t e2 = e / BB;
      const int pair = (int)(e2 % wl.npairs), q = (int)(e2 / wl.npairs);
      if (p.fin_part && (p.fin_part == 1) != (q == R)) return;   // part A owns dLm, part B the dLq_r
      int bi = 0;
      while ((bi + 1) * (bi + 2) / 2 <= pair) bi++;
      const int bj = pair - bi * (bi + 1) / 2;
      const int a = bi * IWVI_BLK + (off >> 6), b = bj * IWVI_BLK + (off & 63);
      double s = 0.0;
      if (a >= b)
        for (int s_ = 0; s_ < wl.S; s_++) s += red[(((size_t)q * wl.S + s_) * wl.npairs + pair) * BB + off];
      if (q < R) { if (a < M && b < M) p.dq_sqrt[((size_t)q * M + a) * M + b] = s; }
      else p.dLm[(size_t)a * Mp + b] = s;
      return;
    }
    e -= f.n_blk;
    if (e < f.n_up) {
      const int off = (int)(e % BB);
      const int64_t e2 = e / BB;
      const int nup = NB * (NB - 1) / 2;
      const int up = (int)(e2 % nup), q = (int)(e2 / nup);
      if (p.fin_part && (p.fin_part == 1) != (q == R)) return;
      int bj = 1;                                  // block column bj > block row bi; enumerate (bi, bj) with bi < bj
      while (bj * (bj + 1) / 2 <= up) bj++;
      const int bi = up - bj * (bj - 1) / 2;
      const int a = bi * IWVI_BLK + (off >> 6), b = bj * IWVI_BLK + (off & 63);
      if (q < R) { if (a < M && b < M) p.dq_sqrt[((size_t)q * M + a) * M + b] = 0.0; }
      else p.dLm[(size_t)a * Mp + b] = 0.0;
      return;
    }
    e -= f.n_up;
    if (e < f.n_qmu && p.fin_part != 1) {
      const int m = (int)(e / R), r = (int)(e - (int64_t)m * R);
      const int bi = m / IWVI_BLK;
      double s = 0.0;
      for (int s_ = 0; s_ < wl.S; s_++)
        s += qred[((size_t)s_ * NB + bi) * IWVI_BLK * IWVI_MAX_R + (m - bi * IWVI_BLK) * IWVI_MAX_R + r];
      p.dq_mu[e] = s;
    }
    return;
  }
  // ---- one warp per output: per-CTA partials of the tile kernel and of the epilogue kernel (part A)
  if (p.fin_part == 2) return;
  const int lane = threadIdx.x & 31;
  const int o = ((int)blockIdx.x - f.grid_elem) * 8 + (threadIdx.x >> 5);
  if (o >= f.n_warp_out) return;
  if (o < f.n_z) {
    const int m = o / D, dd = o - m * D;
    const double s = warp_partial_sum(tile + (size_t)m * ldz + dd, p.grid_tile, wl.tile_stride, lane);
    if (lane == 0) p.dZ[o] = s;
    return;
  }
  const int k = o - f.n_z;
  if (k < D) {
    const double s = warp_partial_sum(tile + (size_t)Mp * ldz + k, p.grid_tile, wl.tile_stride, lane);
    if (lane == 0) p.dls[k] = s;
  } else if (k == D) {
    const double s = warp_partial_sum(tile + (size_t)Mp * ldz + 32, p.grid_tile, wl.tile_stride, lane) +
                     warp_partial_sum(epi + P * R + D * P + P, wl.n_epi, EPI_STRIDE, lane);
    if (lane == 0) p.dvariance[0] = s;
  } else {
    const int j = k - D - 1;   // index into [dW | dmfA | dmfb]
    double* dst = j < P * R ? (p.dW ? p.dW + j : nullptr)
                : j < P * R + D * P ? (p.dmfA ? p.dmfA + (j - P * R) : nullptr)
                                    : (p.dmfb ? p.dmfb + (j - P * R - D * P) : nullptr);
    if (!dst) return;          // frozen parameter: its partials were not formed
    const double s = warp_partial_sum(epi + j, wl.n_epi, EPI_STRIDE, lane);
    if (lane == 0) *dst = s;
  }
}

template <int TP, int KIND>
static int launch_tile_k(const BwdParams& p, int smem_bytes, cudaStream_t st) {
  static const bool pool_ok = iwvi_ws_pool_ok(gp_tile_bwd_kernel<TP, KIND>);
  if (!pool_ok) return IWVI_ERR_LAUNCH;
  if (cudaFuncSetAttribute(gp_tile_bwd_kernel<TP, KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes) != cudaSuccess)
    return IWVI_ERR_LAUNCH;
  const int grid = (p.tile1 - p.tile0) < p.grid_tile ? (p.tile1 - p.tile0) : p.grid_tile;
  gp_tile_bwd_kernel<TP, KIND><<<grid, TILE_THREADS, smem_bytes, st>>>(p);
  IWVI_CHECK_LAUNCH();
  return IWVI_OK;
}
template <int TP>
static int launch_tile(const BwdParams& p, int smem_bytes, cudaStream_t st) {
  switch (p.d.kern) {
    case IWVI_KERN_RBF: return launch_tile_k<TP, IWVI_KERN_RBF>(p, smem_bytes, st);
    case IWVI_KERN_MATERN12: return launch_tile_k<TP, IWVI_KERN_MATERN12>(p, smem_bytes, st);
    case IWVI_KERN_MATERN32: return launch_tile_k<TP, IWVI_KERN_MATERN32>(p, smem_bytes, st);
    default: return launch_tile_k<TP, IWVI_KERN_MATERN52>(p, smem_bytes, st);
  }
}

// tile width of the tile kernel: 64 points when that still gives every SM at least two tiles (or 32 does not fit),
// else 32 -- the rule of the forward kernel (iwvi_pick_tp), c2: 0.529 -> 0.507 ms/step.  With few points (c2: 160 tiles of 64 on 148 SMs) the
// wider tile left most of the second wave empty.
int pick_bwd_tp(int Tp, int Mp, int nsm, int max_smem, int* smem_bytes) {
  const int b64 = tile_smem_layout(64, Mp).total_doubles * 8, b32 = tile_smem_layout(32, Mp).total_doubles * 8;
  const bool fits64 = b64 <= max_smem, fits32 = b32 <= max_smem;
  // (a single, partly filled wave of 64-point tiles is kept: at c1's size the narrower tiles measured slower)
  if (fits64 && (Tp / 64 >= 2 * nsm || Tp / 64 <= nsm || !fits32)) { *smem_bytes = b64; return 64; }
  if (fits32) { *smem_bytes = b32; return 32; }
  return -1;
}

}  // namespace

static int device_info(int* nsm, int* max_smem) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return IWVI_ERR_LAUNCH;
  cudaDeviceGetAttribute(nsm, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  return IWVI_OK;
}

extern "C" int64_t iwvi_gp_bwd_ws_doubles(const iwvi_gp_desc* d) {
  if (iwvi_check_gp_desc(d) != IWVI_OK) return -1;
  int nsm = 148, max_smem = 0;
  if (device_info(&nsm, &max_smem) != IWVI_OK) nsm = 148;
  const int64_t a = bwd_ws_layout(*d, nsm, false).total, b = bwd_ws_layout(*d, nsm, true).total;
  return a > b ? a : b;   // either reduce variant (IWVI_FLAG_FAST_REDUCE) may run on it
}

extern "C" int iwvi_gp_bwd_tile_points(const iwvi_gp_desc* d) {
  if (iwvi_check_gp_desc(d) != IWVI_OK) return IWVI_ERR_BAD_DESC;
  int nsm = 148, max_smem = 0;
  if (device_info(&nsm, &max_smem) != IWVI_OK) return IWVI_ERR_LAUNCH;
  const AuxLayout al = iwvi_aux_layout(d->M, d->D, d->R);
  int smem_bytes = 0;
  const int TP = pick_bwd_tp(iwvi_round_up(d->T, 128), al.Mp, nsm, max_smem, &smem_bytes);
  return TP < 0 ? IWVI_ERR_UNSUPPORTED : TP;
}

static int rows_bwd_impl(const iwvi_gp_desc* d, const double* Lm, const double* aux, const double* save,
                         const double* X, const double* W, const double* mfA, const double* mfb,
                         const double* eps, const double* d_sample, const double* d_mean, const double* d_var,
                         double* dX, double* dZ, double* dls, double* dvariance, double* dq_mu,
                         double* dq_sqrt, double* dLm, double* dW, double* dmfA, double* dmfb, double* ws,
                         int64_t point_begin, int64_t point_end, bool ranged, void* stream) {
  int rc = iwvi_check_gp_desc(d);
  if (rc != IWVI_OK) return rc;
  if (!Lm || !aux || !save || !X || !dX || !dZ || !dls || !dvariance || !dq_mu || !dq_sqrt || !dLm || !ws)
    return IWVI_ERR_NULL;
  if (d->mix && !W) return IWVI_ERR_NULL;
  if (d->mf == IWVI_MF_LINEAR && !mfA) return IWVI_ERR_NULL;
  if ((d->flags & IWVI_FLAG_SAMPLE) && !eps) return IWVI_ERR_NULL;
  if (d->T == 0) return IWVI_ERR_BAD_DESC;
  int nsm = 148, max_smem = 0;
  rc = device_info(&nsm, &max_smem);
  if (rc != IWVI_OK) return rc;
  const AuxLayout al = iwvi_aux_layout(d->M, d->D, d->R);
  BwdParams p;
  p.d = *d; p.Lm = Lm; p.aux = aux; p.save = save; p.X = X; p.W = W; p.mfA = mfA; p.mfb = mfb; p.eps = eps;
  p.d_sample = d_sample; p.d_mean = d_mean; p.d_var = d_var;
  p.dX = dX; p.dZ = dZ; p.dls = dls; p.dvariance = dvariance; p.dq_mu = dq_mu; p.dq_sqrt = dq_sqrt; p.dLm = dLm;
  p.dW = dW; p.dmfA = dmfA; p.dmfb = dmfb; p.ws = ws;
  const bool fast = fast_reduce_ok(*d);
  p.wl = bwd_ws_layout(*d, nsm, fast);
  int smem_bytes = 0;
  const int TP = pick_bwd_tp(p.wl.Tp, al.Mp, nsm, max_smem, &smem_bytes);
  if (TP < 0) return IWVI_ERR_UNSUPPORTED;
  p.ntiles = p.wl.Tp / TP;   // covers the zero-padded rows too, so every row of Bbar is written
  p.grid_tile = p.ntiles < nsm ? p.ntiles : nsm;
  cudaStream_t st = (cudaStream_t)stream;
  const int only = d->flags & IWVI_FLAG_ONLY_MASK;
  p.q_lo = 0; p.q_n = d->R + 1; p.fin_part = 0; p.qmu_only = 0;
  p.tile0 = 0; p.tile1 = p.ntiles; p.slot0 = 0; p.n_slots = 0; p.epi0 = 0;
  p.strip0 = 0; p.strip1 = p.wl.Tp / GRAM_PTS;
  int n_epi = p.wl.n_epi;
  if (ranged) {
    // one of the two point chains of the per-point half: [0, point_end) or [point_begin, T)
    if (!only || (only & (IWVI_FLAG_ONLY_REDUCE | IWVI_FLAG_ONLY_FINAL))) return IWVI_ERR_BAD_DESC;
    if (point_begin < 0 || point_end > d->T || point_begin >= point_end) return IWVI_ERR_BAD_DESC;
    if (point_begin != 0 && point_end != d->T) return IWVI_ERR_BAD_DESC;
    if (point_begin % TP || (point_end != d->T && point_end % TP)) return IWVI_ERR_BAD_DESC;
    p.tile0 = (int)(point_begin / TP);
    p.tile1 = point_end == d->T ? p.ntiles : (int)(point_end / TP);
    p.strip0 = p.tile0 * (TP / GRAM_PTS); p.strip1 = p.tile1 * (TP / GRAM_PTS);
    p.epi0 = (int)(point_begin / EPI_PTS);
    n_epi = (point_end == d->T ? p.wl.n_epi : (int)(point_end / EPI_PTS)) - p.epi0;
    p.n_slots = p.wl.gram_slots;
    if (point_begin != 0) {
      if (p.tile1 - p.tile0 > nsm) return IWVI_ERR_UNSUPPORTED;   // the second chain runs one tile per CTA
      p.slot0 = p.wl.gram_slots;
    }
  }

  if (!only || (only & IWVI_FLAG_ONLY_EPI)) {
    gp_epi_bwd_kernel<<<n_epi, 256, 0, st>>>(p);
    IWVI_CHECK_LAUNCH();
  }

  if (!only || (only & IWVI_FLAG_ONLY_TILE)) {
    rc = TP == 64 ? launch_tile<64>(p, smem_bytes, st) : launch_tile<32>(p, smem_bytes, st);
    if (rc != IWVI_OK) return rc;
  }

  if (!only || (only & IWVI_FLAG_ONLY_GRAM)) {
    rc = iwvi_launch_gram_bwd(p, nsm, max_smem, st);
    if (rc != IWVI_OK) return rc;
  }

  // IWVI_FLAG_PART_A / _B restrict the reduce + finalize launches to one half of the parameter gradients, so that the
  // caller can start iwvi_gp_prologue_bwd (needs dLm, dZ, dls, dvariance only) while the other half is still running:
  //   A: reduce (dLm items), finalize (dLm, dZ, dls, dvariance, dW, dmfA, dmfb)      B: reduce (dLq_r items, dq_mu), finalize (dq_sqrt, dq_mu)
  const int part = d->flags & (IWVI_FLAG_PART_A | IWVI_FLAG_PART_B);
  const bool do_a = part != IWVI_FLAG_PART_B, do_b = part != IWVI_FLAG_PART_A;
  if (!only || (only & IWVI_FLAG_ONLY_REDUCE)) {
    const int red_smem = (RED_NST * IWVI_STAGE_DOUBLES + 2 * RED_NST + RED_SC_SLOTS * IWVI_BLK) * 8;
    if (cudaFuncSetAttribute(gp_reduce_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, red_smem) != cudaSuccess)
      return IWVI_ERR_LAUNCH;
    const int per_q = p.wl.S * p.wl.npairs;
    p.q_lo = (do_a && !do_b) ? d->R : 0;
    const int nq = (do_a && do_b) ? d->R + 1 : (do_a ? 1 : d->R);
    p.q_n = nq;
    if (fast) {
      // tcgen05 variant for dLq_r and dq_mu; dLm stays on the float64 kernel: its error would be amplified by the
      // Cholesky adjoint behind it (measured: 1e-4 on dZ / kernel parameters instead of 1e-6)
      const int q_hi = p.q_lo + nq;                        // one past the last matrix of this launch
      const int nq_fast = (q_hi < d->R ? q_hi : d->R) - p.q_lo;
      if (nq_fast > 0) {
        const int fsmem = FR_STAGES * FR_STAGE_FLOATS * 4 + FR_MAX_RANGE * 4 + 1024;
        if (p.wl.chunks_per_split * IWVI_BLK > FR_MAX_RANGE) return IWVI_ERR_UNSUPPORTED;
        if (cudaFuncSetAttribute(gp_reduce_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, fsmem) != cudaSuccess)
          return IWVI_ERR_LAUNCH;
        p.q_n = nq_fast;
        gp_reduce_fast_kernel<<<nq_fast * p.wl.S * fast_reduce_tiles(al.Mp), FR_THREADS, fsmem, st>>>(p);
        IWVI_CHECK_LAUNCH();
        if (p.q_lo == 0) {
          // the launch that owns matrix 0 owns dq_mu = A gmean_bar: the float64 kernel's (q = 0, block column 0) items,
          // which form it as a ride-along (their dLq_0 blocks overwrite the tensor-core ones with float64 values)
          p.qmu_only = 1; p.q_n = 1;
          gp_reduce_bwd_kernel<<<p.wl.S * al.NB, RED_THREADS, red_smem, st>>>(p);
          IWVI_CHECK_LAUNCH();
          p.qmu_only = 0;
        }
      }
      if (q_hi > d->R) {
        p.q_lo = d->R; p.q_n = 1;
        gp_reduce_bwd_kernel<<<per_q, RED_THREADS, red_smem, st>>>(p);
      }
    } else {
      gp_reduce_bwd_kernel<<<nq * per_q, RED_THREADS, red_smem, st>>>(p);
    }
    IWVI_CHECK_LAUNCH();
  }
  if (!only || (only & IWVI_FLAG_ONLY_FINAL)) {
    const FinalLayout fl = final_layout(*d, p.wl);
    p.fin_part = (do_a && do_b) ? 0 : (do_a ? 1 : 2);
    // the per-CTA partials of the gram-adjoint kernel: the slots of its one launch over all strips, or
    // (IWVI_FLAG_TWO_CHAINS) both chains' slots
    p.strip0 = 0; p.strip1 = p.wl.Tp / GRAM_PTS;
    p.grid_tile = (d->flags & IWVI_FLAG_TWO_CHAINS) ? 2 * p.wl.gram_slots : iwvi_gram_bwd_grid(p);
    gp_finalize_bwd_kernel<<<fl.grid_elem + fl.grid_warp, 256, 0, st>>>(p);
    IWVI_CHECK_LAUNCH();
  }
  return IWVI_OK;
}

extern "C" int iwvi_gp_rows_bwd(const iwvi_gp_desc* d, const double* Lm, const double* aux, const double* save,
                                const double* X, const double* W, const double* mfA, const double* mfb,
                                const double* eps, const double* d_sample, const double* d_mean, const double* d_var,
                                double* dX, double* dZ, double* dls, double* dvariance, double* dq_mu,
                                double* dq_sqrt, double* dLm, double* dW, double* dmfA, double* dmfb, double* ws,
                                void* stream) {
  return rows_bwd_impl(d, Lm, aux, save, X, W, mfA, mfb, eps, d_sample, d_mean, d_var, dX, dZ, dls, dvariance, dq_mu,
                       dq_sqrt, dLm, dW, dmfA, dmfb, ws, 0, d ? d->T : 0, false, stream);
}

extern "C" int iwvi_gp_rows_bwd_range(const iwvi_gp_desc* d, const double* Lm, const double* aux, const double* save,
                                      const double* X, const double* W, const double* mfA, const double* mfb,
                                      const double* eps, const double* d_sample, const double* d_mean,
                                      const double* d_var, double* dX, double* dZ, double* dls, double* dvariance,
                                      double* dq_mu, double* dq_sqrt, double* dLm, double* dW, double* dmfA, double* dmfb,
                                      double* ws, int64_t point_begin, int64_t point_end, void* stream) {
  return rows_bwd_impl(d, Lm, aux, save, X, W, mfA, mfb, eps, d_sample, d_mean, d_var, dX, dZ, dls, dvariance, dq_mu,
                       dq_sqrt, dLm, dW, dmfA, dmfb, ws, point_begin, point_end, true, stream);
}
