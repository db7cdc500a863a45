// gp_rows_fwd.cu -- per-point stage of the sparse-variational conditional, forward.
//
// Replaces independent_multisample_sample_conditional (reference temp_workaround.py:44-91, diag branch),
// the SharedMixedMok mixing (:142-145) and the mean-function add (layers.py:46-48) with ONE persistent kernel:
// a CTA owns a tile of TP points, keeps the M x TP panel (Kuf -> A = Lm^-1 Kuf) resident in shared memory
// (block-major [m-block][point][68], the layout of the saved arrays), has a dedicated producer warp stream the 64x64
// blocks of Lm / inverted diagonal blocks / tril(q_sqrt) from L2 through a bulk-TMA + mbarrier ring, and does every
// contraction on the FP64 tensor pipe (DMMA); triangular diagonal blocks only issue the 8x8x4 tiles that touch the triangle:
//
//   G  Kuf_i   = k(|z|^2 + |x|^2 - 2 z.x)                (gram: -2XZ^T GEMM + norm epilogue + kernel function)
//   T  A_i     = Dinv_i (Kuf_i - sum_{j<i} Lm_ij A_j)    (blocked forward substitution, inverted diagonal blocks)
//   S  fvar0   = variance - sum_m A^2 ; gmean = A^T q_mu
//   U  U_r,i   = sum_{j>=i} Lq_r[j,i]^T A_j ; gvar_r = fvar0 + sum_m U_r^2     (never materialised unless saved)
//   E  sample  = gmean + eps*sqrt(gvar) ; Mok mixing by W, W^2 ; + mean function
//
// Nothing of size M x T touches HBM unless IWVI_FLAG_SAVE asks for A and U (kept for the backward pass: on B200 an
// HBM round trip costs less than recomputing them at the 37 TFLOP/s fp64 rate).
#include <type_traits>
#include <stdlib.h>
#include "common.cuh"

namespace {

struct FwdSeq {  // order in which blocks are consumed by one tile
  int NB, R, npairs, ldz;
  const double *Zt, *Lmb, *Lqb;
  int ph, r, i, j;
  // (the scaled inducing inputs are read straight from L1/L2 by the gram phase: streaming those small blocks through
  //  the ring exposed one TMA latency per block and kept the first Lm blocks from being prefetched)
  __device__ __forceinline__ void init() { ph = 1; r = 0; i = 0; j = 0; }
  __device__ __forceinline__ bool done() const { return ph == 3; }
  __device__ __forceinline__ BlockSrc get() const {
    BlockSrc b;
    if (ph == 0) {          // (unused) scaled inducing inputs of block i
      b.src = Zt + (size_t)i * IWVI_BLK * ldz; b.bytes = (uint32_t)(IWVI_BLK * ldz * 8);
    } else if (ph == 1) {   // Lm(i,j), j < i; then the inverted diagonal block (slot (i,i))
      b.src = Lmb + (size_t)iwvi_pair(i, j) * IWVI_STAGE_DOUBLES; b.bytes = IWVI_STAGE_DOUBLES * 8;
    } else {                // tril(q_sqrt_r) block (j, i), j >= i
      b.src = Lqb + ((size_t)r * npairs + iwvi_pair(j, i)) * IWVI_STAGE_DOUBLES; b.bytes = IWVI_STAGE_DOUBLES * 8;
    }
    return b;
  }
  __device__ __forceinline__ void advance() {
    if (ph == 0) {
      if (++i == NB) { ph = 1; i = 0; j = 0; }
    } else if (ph == 1) {
      if (j < i) ++j;
      else { ++i; j = 0; if (i == NB) { ph = (R > 0) ? 2 : 3; r = 0; i = 0; j = 0; } }
    } else {
      if (++j == NB) { ++i; j = i; if (i == NB) { ++r; i = 0; j = 0; if (r == R) ph = 3; } }
    }
  }
};

template <int TP> struct TileCfg {
  static constexpr int NW = 8;
  static constexpr int WNG = TP >= 64 ? 4 : 2;  // warps along the point axis (TP is 64 or 32)
  static constexpr int WMG = NW / WNG;          // warps along the 64 rows of a block
  static constexpr int WM = IWVI_BLK / WMG;
  static constexpr int WN = TP / WNG;
  static constexpr int TM = WM / 8;
  static constexpr int TN = WN / 8;
};

struct FwdParams {
  iwvi_gp_desc d;
  const double *Lm, *aux, *X, *W, *mfA, *mfb, *eps;
  double *sample, *mean, *var, *save;
  int tile0, ntiles;   // this launch covers tiles [tile0, ntiles)
};

// dynamic shared memory carve-up (doubles), shared with the host-side size computation
struct FwdSmem {
  int panel, stages, xs, xn, fv0, usq, gm, bars, total_doubles;
};
__host__ __device__ inline FwdSmem fwd_smem_layout(int TP, int Mp, int ldx) {
  FwdSmem s; int o = 0;
  // the panel is block-major like the saved arrays: [m-block][point][68], so that a whole m-block of the tile leaves
  // for HBM as ONE bulk-TMA store (per-row 512-byte stores were bound by the TMA engine's per-operation cost)
  s.panel = o;  o += (Mp / IWVI_BLK) * TP * IWVI_LDS;
  s.stages = o; o += IWVI_NST * IWVI_STAGE_DOUBLES;
  const int n_xs = TP * ldx, n_usq = IWVI_MAX_R * TP * (TP >= 64 ? 2 : 4);   // usq: one slot per warp row group
  s.xs = o;     o += n_xs;
  if (n_usq <= n_xs) s.usq = s.xs;            // x tile is dead after the gram phase; usq lives from U to E
  else { s.usq = o; o += n_usq; }
  s.xn = o;     o += TP;
  s.fv0 = o;    o += TP;
  s.gm = o;     o += IWVI_MAX_R * TP;
  s.bars = o;   o += 2 * IWVI_NST;
  s.total_doubles = o;
  return s;
}

#define FWD_THREADS IWVI_WS_THREADS   // 8 consumer warps + the producer warpgroup (1 active warp)
#define BAR_ALL 1         // named barrier of the 256 consumer threads
#define BAR_COL 2         // + column-group index: the WMG warps that share a set of points

// KIND (the stationary kernel, IWVI_KERN_*) is a template parameter: with a run-time switch the sixteen inlined
// copies of the kernel function per thread carried all four families and pushed the code past the instruction cache.
template <int TP, int KIND>
__global__ void __launch_bounds__(FWD_THREADS, 1) gp_rows_fwd_kernel(const FwdParams p) {
  using C = TileCfg<TP>;
  extern __shared__ __align__(16) double smem[];
  const iwvi_gp_desc& d = p.d;
  const AuxLayout al = iwvi_aux_layout(d.M, d.D, d.R);
  const int Mp = al.Mp, NB = al.NB, ldz = al.ldz, R = d.R, D = d.D, T = d.T;
  constexpr int PSTR = TP * IWVI_LDS;        // doubles per m-block of the panel: element (point n, m) of block b is
                                             // panel[b * PSTR + n * IWVI_LDS + m]
  const int Dk = iwvi_round_up(D, 4);
  const FwdSmem sl = fwd_smem_layout(TP, Mp, ldz);
  double* panel = smem + sl.panel;
  double* xs = smem + sl.xs;
  double* xn = smem + sl.xn;
  double* fv0 = smem + sl.fv0;
  double* usq = smem + sl.usq;
  double* gms = smem + sl.gm;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double* aux = p.aux;

  // the 4 pad columns of every panel row travel with the bulk stores: keep them zero (nothing else writes them)
  for (int idx = threadIdx.x; idx < NB * TP * 4; idx += FWD_THREADS) panel[(idx >> 2) * IWVI_LDS + IWVI_BLK + (idx & 3)] = 0.0;
  RingT<IWVI_NST> ring;
  ring.setup(reinterpret_cast<uint64_t*>(smem + sl.bars), smem + sl.stages, C::NW);

  if (warp >= C::NW) {
    reg_dealloc<IWVI_PRODUCER_REGS>();
    if (warp > C::NW) return;      // padding of the producer warpgroup (setmaxnreg is warpgroup-aligned, common.cuh)
    // ---- producer warp: streams the tile-independent block sequence once per tile, running ahead of the consumers
    FwdSeq seq;
    seq.NB = NB; seq.R = R; seq.npairs = al.npairs; seq.ldz = ldz;
    seq.Zt = aux + al.off_zt; seq.Lmb = aux + al.off_lmb; seq.Lqb = aux + al.off_lqb;
    for (int tile = p.tile0 + blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
      seq.init();
      while (!seq.done()) { ring.produce(seq.get(), lane); seq.advance(); }
    }
    return;
  }

  reg_alloc<IWVI_CONSUMER_REGS>();
  const int g = lane >> 2, t = lane & 3;
  // Row tiles (8 rows) of a 64-row block are dealt round-robin to the WMG warps of a column group: warp `wmi` owns
  // tiles wmi, wmi + WMG, ...  With the triangular diagonal blocks this balances the skipped work, and flipping the
  // order in warps 4-7 balances it across the four SM sub-partitions too (warps w and w + 4 share one).
  const int wmi = ((warp >> 2) & 1) ? (C::WMG - 1 - warp % C::WMG) : (warp % C::WMG);
  const int wr0 = wmi * 8;                   // first row of this warp's first tile; tile a starts at wr0 + a * MR
  constexpr int MR = 8 * C::WMG;             // row distance between consecutive tiles of one warp
  const int wn0 = (warp / C::WMG) * C::WN;   // first point of this warp
  const int colbar = BAR_COL + warp / C::WMG;
  const double* zn = aux + al.off_zn;
  const double* qmu = aux + al.off_qmu;
  const double* consts = aux + al.off_consts;
  const double variance = consts[IWVI_C_VARIANCE];
  const SaveLayout sv = iwvi_save_layout(T, d.M, R);
  const bool do_save = (d.flags & IWVI_FLAG_SAVE) != 0;
  const bool do_sample = (d.flags & IWVI_FLAG_SAMPLE) != 0;

  PHASE_DECL;
  for (int tile = p.tile0 + blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
    const int n0 = tile * TP;
    if (warp == 0) bulk_wait_read();   // the previous tile's bulk stores no longer read the panel
    named_bar_sync(BAR_ALL, 256);      // previous tile fully done with every shared buffer

    // ---- x tile: xs[n][k] = X[n0+n][k] / ls[k] (zero padded), xn[n] = |xs[n]|^2
    for (int idx = tid; idx < TP * ldz; idx += 256) {
      const int n = idx / ldz, k = idx - n * ldz;
      double v = 0.0;
      if (k < D && n0 + n < T) v = p.X[(size_t)(n0 + n) * D + k] * consts[IWVI_C_INVLS + k];
      xs[idx] = v;
    }
    named_bar_sync(BAR_ALL, 256);
    if (tid < TP) {
      double s = 0.0;
      for (int k = 0; k < Dk; k++) { const double v = xs[tid * ldz + k]; s += v * v; }
      xn[tid] = s;
    }
    named_bar_sync(BAR_ALL, 256);

    PHASE_MARK(0);
    // ---- G: gram blocks -> panel (Kuf).  A fragments (Z / ls, [Mp][ldz]) and |z|^2 come straight from L1/L2; they are
    //      fetched for block i + 1 while block i's kernel function is evaluated.
    //      The k range is the zero-padded row length of Z / ls (5 or 8 k-steps of 4, a compile-time constant per
    //      branch): a run-time trip count would put the DMMAs under predicates, which occupy the pipe even when off.
    auto gram_phase = [&](auto nk_c) {
      constexpr int NK = decltype(nk_c)::value;
      const double* ztw = aux + al.off_zt + (size_t)(wr0 + g) * ldz + t;
      double za[C::TM][NK], znr[C::TM];
      auto fetch = [&](int i) {
#pragma unroll
        for (int a = 0; a < C::TM; a++) {
          const double* zp = ztw + (size_t)(i * IWVI_BLK + a * MR) * ldz;
#pragma unroll
          for (int ks = 0; ks < NK; ks++) za[a][ks] = __ldg(zp + 4 * ks);
          znr[a] = __ldg(zn + i * IWVI_BLK + wr0 + a * MR + g);
        }
      };
      fetch(0);
      for (int i = 0; i < NB; i++) {
        double acc[C::TM][C::TN][2];
        acc_zero<C::TM, C::TN>(acc);
        const double* bp = xs + (wn0 + g) * ldz + t;
#pragma unroll
        for (int ks = 0; ks < NK; ks++) {
          double b[C::TN];
#pragma unroll
          for (int j = 0; j < C::TN; j++) b[j] = bp[j * 8 * ldz + 4 * ks];
#pragma unroll
          for (int a = 0; a < C::TM; a++)
#pragma unroll
            for (int j = 0; j < C::TN; j++) dmma884(acc[a][j], za[a][ks], b[j]);
        }
        double zc[C::TM];
#pragma unroll
        for (int a = 0; a < C::TM; a++) zc[a] = znr[a];
        if (i + 1 < NB) fetch(i + 1);
#pragma unroll
        for (int a = 0; a < C::TM; a++) {
          const int mg = i * IWVI_BLK + wr0 + a * MR + g;
          double kv[C::TN * 2], unused[C::TN * 2];
#pragma unroll
          for (int b = 0; b < C::TN; b++)
#pragma unroll
            for (int c = 0; c < 2; c++) kv[b * 2 + c] = zc[a] + xn[wn0 + b * 8 + 2 * t + c] - 2.0 * acc[a][b][c];
          kern_n<KIND, C::TN * 2, false>(kv, unused, variance);     // the row's kernel values in lock step
#pragma unroll
          for (int b = 0; b < C::TN; b++)
#pragma unroll
            for (int c = 0; c < 2; c++)
              panel[i * PSTR + (wn0 + b * 8 + 2 * t + c) * IWVI_LDS + wr0 + a * MR + g] = (mg < d.M) ? kv[b * 2 + c] : 0.0;
        }
      }
    };
    if (ldz == 20) gram_phase(std::integral_constant<int, 5>());
    else gram_phase(std::integral_constant<int, 8>());
    named_bar_sync(colbar, C::WMG * 32);

    PHASE_MARK(1);
    // ---- T: blocked forward substitution, in place.  Data only flows between the WMG warps of a column group.
    for (int i = 0; i < NB; i++) {
      double acc[C::TM][C::TN][2];
      if (i > 0) {
        // acc = -(Kuf_i) + sum_{j<i} Lm(i,j) A_j, then rhs_i := -acc (each thread owns its entries: no read-modify-write)
#pragma unroll
        for (int a = 0; a < C::TM; a++)
#pragma unroll
          for (int b = 0; b < C::TN; b++)
#pragma unroll
            for (int c = 0; c < 2; c++)
              acc[a][b][c] = -panel[i * PSTR + (wn0 + b * 8 + 2 * t + c) * IWVI_LDS + wr0 + a * MR + g];
        for (int j = 0; j < i; j++) {
          const double* st = ring.wait();
          warp_gemm_pf<C::TM, C::TN, 0, 0, C::WMG>(acc, st + wr0 * IWVI_LDS, IWVI_LDS, panel + j * PSTR + wn0 * IWVI_LDS,
                                                   IWVI_LDS, lane);
          ring.release(lane);
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
      const double* st = ring.wait();   // inverted diagonal block
      acc_zero<C::TM, C::TN>(acc);
      warp_gemm_tri<C::TM, C::TN, 0, 0, C::WMG, 1>(acc, st, IWVI_LDS, panel + i * PSTR + wn0 * IWVI_LDS, IWVI_LDS, wmi, lane);
      ring.release(lane);
      named_bar_sync(colbar, C::WMG * 32);   // every warp of the group has read the right-hand side
#pragma unroll
      for (int a = 0; a < C::TM; a++)
#pragma unroll
        for (int b = 0; b < C::TN; b++)
#pragma unroll
          for (int c = 0; c < 2; c++) {
            const int n = wn0 + b * 8 + 2 * t + c;
            panel[i * PSTR + n * IWVI_LDS + wr0 + a * MR + g] = acc[a][b][c];
          }
      named_bar_sync(colbar, C::WMG * 32);
    }
    fence_async_smem();   // the panel (A) is read by bulk stores below
    named_bar_sync(BAR_ALL, 256);

    PHASE_MARK(2);
    // ---- S: fvar0 = sum_m A^2 and the latent means gmean = A^T q_mu, as one skinny DMMA product per 8 points
    //      (A fragments from the panel, q_mu [Mp, 8] fragments straight from L1/L2); two accumulators break the
    //      dependency chain.  The squares ride along on the A fragments.
    for (int mt = warp; mt < TP / 8; mt += C::NW) {
      const double* ap = panel + (mt * 8 + g) * IWVI_LDS + t;
      const double* bp = qmu + (size_t)t * IWVI_MAX_R + g;
      double c0[2] = {0.0, 0.0}, c1[2] = {0.0, 0.0}, sq0 = 0.0, sq1 = 0.0;
      for (int mb = 0; mb < NB; mb++) {
#pragma unroll 4
        for (int k0 = 0; k0 < IWVI_BLK; k0 += 8) {
          const double a0 = ap[mb * PSTR + k0], a1 = ap[mb * PSTR + k0 + 4];
          const double b0 = __ldg(bp + (size_t)(mb * IWVI_BLK + k0) * IWVI_MAX_R);
          const double b1 = __ldg(bp + (size_t)(mb * IWVI_BLK + k0 + 4) * IWVI_MAX_R);
          dmma884(c0, a0, b0);
          dmma884(c1, a1, b1);
          sq0 += a0 * a0; sq1 += a1 * a1;
        }
      }
      double sq = sq0 + sq1;
      sq += __shfl_xor_sync(0xffffffffu, sq, 1);
      sq += __shfl_xor_sync(0xffffffffu, sq, 2);
      if (t == 0) fv0[mt * 8 + g] = sq;
      gms[(2 * t) * TP + mt * 8 + g] = c0[0] + c1[0];
      gms[(2 * t + 1) * TP + mt * 8 + g] = c0[1] + c1[1];
    }
    if (do_save) {
      const int nvalid = min(TP, T - n0);       // real points of this tile (pad points are saved as zeros)
      double* dst = p.save + sv.off_a + (int64_t)(n0 >> 6) * NB * IWVI_STAGE_DOUBLES + (int64_t)(n0 & 63) * IWVI_LDS;
      if (nvalid == TP) {
        // asynchronous TMA stores straight from the panel: one bulk operation per m-block and 64-point chunk
        if (warp == 0 && lane < NB) {
          bulk_s2g(dst + (int64_t)lane * IWVI_STAGE_DOUBLES, panel + lane * PSTR, TP * IWVI_LDS * 8);
          bulk_commit();
        }
      } else {
        const int mm = tid & 63;
        for (int mb = 0; mb < NB; mb++)
          for (int n = tid >> 6; n < TP; n += 4)
            dst[(int64_t)mb * IWVI_STAGE_DOUBLES + n * IWVI_LDS + mm] = (n < nvalid) ? panel[mb * PSTR + n * IWVI_LDS + mm] : 0.0;
      }
    }

    // ---- U: triangular products with tril(q_sqrt_r)^T, column sums of squares (no inter-warp data flow)
    // this thread's corner of every saved U block of the tile (block-major [point][68]); pad points are saved as zeros
    double* ubase = p.save + sv.off_u + (int64_t)(n0 >> 6) * NB * IWVI_STAGE_DOUBLES +
                    (int64_t)((n0 & 63) + wn0 + 2 * t) * IWVI_LDS + wr0 + g;
    unsigned pad_mask = 0;
#pragma unroll
    for (int b = 0; b < C::TN; b++)
#pragma unroll
      for (int c = 0; c < 2; c++)
        if (n0 + wn0 + b * 8 + 2 * t + c >= T) pad_mask |= 1u << (b * 2 + c);
    for (int r = 0; r < R; r++) {
      double csq[C::TN][2];
#pragma unroll
      for (int b = 0; b < C::TN; b++) { csq[b][0] = 0.0; csq[b][1] = 0.0; }
      for (int i = 0; i < NB; i++) {
        double acc[C::TM][C::TN][2];
        acc_zero<C::TM, C::TN>(acc);
        double* ub = ubase + (int64_t)r * sv.u_stride + (int64_t)i * IWVI_STAGE_DOUBLES;
        for (int j = i; j < NB; j++) {
          const double* st = ring.wait();
          if (j == i)   // diagonal block of tril(q_sqrt_r), used transposed: upper triangular in (m, k)
            warp_gemm_tri<C::TM, C::TN, 1, 0, C::WMG, 0>(acc, st, IWVI_LDS, panel + j * PSTR + wn0 * IWVI_LDS, IWVI_LDS, wmi, lane);
          else
            warp_gemm_pf<C::TM, C::TN, 1, 0, C::WMG>(acc, st + wr0, IWVI_LDS, panel + j * PSTR + wn0 * IWVI_LDS, IWVI_LDS, lane);
          ring.release(lane);
        }
#pragma unroll
        for (int a = 0; a < C::TM; a++)
#pragma unroll
          for (int b = 0; b < C::TN; b++)
#pragma unroll
            for (int c = 0; c < 2; c++) {
              const double u = acc[a][b][c];
              csq[b][c] += u * u;
              if (do_save) ub[(b * 8 + c) * IWVI_LDS + a * MR] = ((pad_mask >> (b * 2 + c)) & 1u) ? 0.0 : u;
            }
      }
#pragma unroll
      for (int b = 0; b < C::TN; b++)
#pragma unroll
        for (int c = 0; c < 2; c++) {
          double v = csq[b][c];
          v += __shfl_xor_sync(0xffffffffu, v, 4);
          v += __shfl_xor_sync(0xffffffffu, v, 8);
          v += __shfl_xor_sync(0xffffffffu, v, 16);
          if (g == 0) usq[(r * C::WMG + warp % C::WMG) * TP + wn0 + b * 8 + 2 * t + c] = v;
        }
    }
    named_bar_sync(BAR_ALL, 256);

    PHASE_MARK(4);
    // ---- E: per-point epilogue, 256/TP threads per point sharing its output columns
    {
      constexpr int QT = 256 / TP;
      const int n = tid / QT, q0 = tid % QT;
      if (n0 + n < T) {
        const size_t pt = (size_t)(n0 + n);
        double gm[IWVI_MAX_R], gv[IWVI_MAX_R], gs[IWVI_MAX_R];
        for (int r = 0; r < R; r++) {
          gm[r] = gms[r * TP + n];
          double us = 0.0;
#pragma unroll
          for (int wg = 0; wg < C::WMG; wg++) us += usq[(r * C::WMG + wg) * TP + n];
          gv[r] = variance - fv0[n] + us;
          // (training: the per-point outputs are formed from the saved latent moments by gp_epi_fwd_kernel, at full
          //  occupancy -- here, with one CTA per SM, their dependent global loads were 5 % of the kernel)
          gs[r] = (do_sample && !do_save) ? gm[r] + p.eps[pt * R + r] * sqrt(gv[r]) : 0.0;
          if (do_save && q0 == 0) {
            p.save[sv.off_gvar + pt * R + r] = gv[r];
            p.save[sv.off_gmean + pt * R + r] = gm[r];
          }
        }
        const int P = do_save ? 0 : d.P;
        for (int q = q0; q < P; q += QT) {
          double mf = 0.0;
          if (d.mf == IWVI_MF_IDENTITY) mf = p.X[pt * D + q];
          else if (d.mf == IWVI_MF_LINEAR) {
            mf = p.mfb[q];
            for (int k = 0; k < D; k++) mf += p.X[pt * D + k] * p.mfA[k * P + q];
          }
          double om, ov, os;
          if (d.mix) {
            om = 0.0; ov = 0.0; os = 0.0;
            for (int r = 0; r < R; r++) {
              const double w = p.W[q * R + r];
              om += gm[r] * w; ov += gv[r] * w * w; os += gs[r] * w;
            }
          } else { om = gm[q]; ov = gv[q]; os = gs[q]; }
          p.mean[pt * P + q] = om + mf;
          p.var[pt * P + q] = ov;
          if (do_sample) p.sample[pt * P + q] = os + mf;
        }
      }
    }
    PHASE_MARK(5);
  }
  PHASE_FLUSH(0);
  if (warp == 0) bulk_wait_read();   // shared memory must outlive the last tile's bulk stores
}

// Per-point epilogue of a SAVING call (training), from the latent moments the row kernel left in `save`: sample / mean /
// var = mixing by W (temp_workaround.py:142-145) + mean function (layers.py:46-48).  One CTA per 32 points: the points'
// inputs and moments are staged in shared memory by coalesced loads (one global round trip per CTA), then thread
// (point, output column) forms its outputs from shared memory and the stores are coalesced.  Same operation order as the
// in-kernel epilogue above (the non-saving path), so both give identical bits.
#define EPIF_PTS 32
__global__ void __launch_bounds__(256) gp_epi_fwd_kernel(const FwdParams p, int pt0, int pt1) {
  __shared__ double Ws[IWVI_MAX_P * IWVI_MAX_R], As[IWVI_MAX_D * IWVI_MAX_P], bs[IWVI_MAX_P];
  __shared__ double Xs[EPIF_PTS * IWVI_MAX_D], gms[EPIF_PTS * IWVI_MAX_R], gvs[EPIF_PTS * IWVI_MAX_R], gss[EPIF_PTS * IWVI_MAX_R];
  const iwvi_gp_desc& d = p.d;
  const int R = d.R, P = d.P, D = d.D;
  const SaveLayout sv = iwvi_save_layout(d.T, d.M, R);
  const bool do_sample = (d.flags & IWVI_FLAG_SAMPLE) != 0;
  const int tid = threadIdx.x;
  const int64_t p0 = (int64_t)pt0 + (int64_t)blockIdx.x * EPIF_PTS;
  const int npts = (int)((pt1 - p0) < EPIF_PTS ? (pt1 - p0) : EPIF_PTS);
  if (d.mix) for (int i = tid; i < P * R; i += 256) Ws[i] = p.W[i];
  if (d.mf == IWVI_MF_LINEAR) {
    for (int i = tid; i < D * P; i += 256) As[i] = p.mfA[i];
    if (tid < P) bs[tid] = p.mfb[tid];
  }
  if (d.mf != IWVI_MF_ZERO)
    for (int i = tid; i < npts * D; i += 256) Xs[i] = p.X[p0 * D + i];
  for (int i = tid; i < npts * R; i += 256) {
    const double gm = p.save[sv.off_gmean + p0 * R + i], gv = p.save[sv.off_gvar + p0 * R + i];
    gms[i] = gm; gvs[i] = gv;
    gss[i] = do_sample ? gm + p.eps[p0 * R + i] * sqrt(gv) : 0.0;
  }
  __syncthreads();
  for (int idx = tid; idx < npts * P; idx += 256) {
    const int n = idx / P, q = idx - n * P;
    const double *gm = gms + n * R, *gv = gvs + n * R, *gs = gss + n * R;
    double mf = 0.0;
    if (d.mf == IWVI_MF_IDENTITY) mf = Xs[n * D + q];
    else if (d.mf == IWVI_MF_LINEAR) {
      mf = bs[q];
      for (int k = 0; k < D; k++) mf += Xs[n * D + k] * As[k * P + q];
    }
    double om, ov, os;
    if (d.mix) {
      om = 0.0; ov = 0.0; os = 0.0;
      for (int r = 0; r < R; r++) {
        const double w = Ws[q * R + r];
        om += gm[r] * w; ov += gv[r] * w * w; os += gs[r] * w;
      }
    } else { om = gm[q]; ov = gv[q]; os = gs[q]; }
    p.mean[p0 * P + idx] = om + mf;
    p.var[p0 * P + idx] = ov;
    if (do_sample) p.sample[p0 * P + idx] = os + mf;
  }
}

template <int TP, int KIND>
int launch_fwd_k(const FwdParams& p, int smem_bytes, int grid, cudaStream_t stream) {
  static const bool pool_ok = iwvi_ws_pool_ok(gp_rows_fwd_kernel<TP, KIND>);
  if (!pool_ok) return IWVI_ERR_LAUNCH;
  if (cudaFuncSetAttribute(gp_rows_fwd_kernel<TP, KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes) != cudaSuccess)
    return IWVI_ERR_LAUNCH;
  gp_rows_fwd_kernel<TP, KIND><<<grid, FWD_THREADS, smem_bytes, stream>>>(p);
  IWVI_CHECK_LAUNCH();
  return IWVI_OK;
}

template <int TP>
int launch_fwd(const FwdParams& p, int smem_bytes, int grid, cudaStream_t stream) {
  switch (p.d.kern) {
    case IWVI_KERN_RBF: return launch_fwd_k<TP, IWVI_KERN_RBF>(p, smem_bytes, grid, stream);
    case IWVI_KERN_MATERN12: return launch_fwd_k<TP, IWVI_KERN_MATERN12>(p, smem_bytes, grid, stream);
    case IWVI_KERN_MATERN32: return launch_fwd_k<TP, IWVI_KERN_MATERN32>(p, smem_bytes, grid, stream);
    default: return launch_fwd_k<TP, IWVI_KERN_MATERN52>(p, smem_bytes, grid, stream);
  }
}

}  // namespace

int iwvi_check_gp_desc(const iwvi_gp_desc* d) {
  if (!d) return IWVI_ERR_NULL;
  if (d->T < 0 || d->M < 1 || d->D < 1 || d->R < 1 || d->P < 1) return IWVI_ERR_BAD_DESC;
  if (d->M > IWVI_MAX_M || d->D > IWVI_MAX_D || d->R > IWVI_MAX_R || d->P > IWVI_MAX_P) return IWVI_ERR_UNSUPPORTED;
  if (d->kern < 0 || d->kern > 3 || d->mf < 0 || d->mf > 2) return IWVI_ERR_BAD_DESC;
  if (!d->mix && d->P != d->R) return IWVI_ERR_BAD_DESC;
  if (d->mf == IWVI_MF_IDENTITY && d->P != d->D) return IWVI_ERR_BAD_DESC;
  return IWVI_OK;
}

// pick the tile width: the largest TP whose panel fits and that still yields >= 2 tiles per SM, else smaller
int iwvi_pick_tp(int T, int Mp, int ldz, int nsm, int max_smem, int* smem_bytes) {
  // (a tile never straddles a 64-point chunk of the block-major saved arrays: TP divides 64)
  const int cands[2] = {64, 32};
  int best = -1, best_bytes = 0;
  for (int c = 0; c < 2; c++) {
    const int TP = cands[c];
    const int bytes = fwd_smem_layout(TP, Mp, ldz).total_doubles * 8;
    if (bytes > max_smem) continue;
    if (best < 0) { best = TP; best_bytes = bytes; }
    const int tiles = (T + TP - 1) / TP;
    if (tiles >= 2 * nsm || TP == 32) { best = TP; best_bytes = bytes; break; }
    best = TP; best_bytes = bytes;
  }
  *smem_bytes = best_bytes;
  return best;
}

extern "C" int iwvi_gp_tile_points(const iwvi_gp_desc* d) {
  int rc = iwvi_check_gp_desc(d);
  if (rc != IWVI_OK) return rc;
  int dev = 0, nsm = 148, max_smem = 0, smem_bytes = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  const AuxLayout al = iwvi_aux_layout(d->M, d->D, d->R);
  const int TP = iwvi_pick_tp(d->T, al.Mp, al.ldz, nsm, max_smem, &smem_bytes);
  return TP < 0 ? IWVI_ERR_UNSUPPORTED : TP;
}

extern "C" int iwvi_gp_rows_fwd_range(const iwvi_gp_desc* d, const double* Lm, const double* aux, const double* X,
                                      const double* W, const double* mfA, const double* mfb, const double* eps,
                                      double* sample, double* mean, double* var, double* save,
                                      int64_t point_begin, int64_t point_end, void* stream) {
  int rc = iwvi_check_gp_desc(d);
  if (rc != IWVI_OK) return rc;
  if (!Lm || !aux || !X || !mean || !var) return IWVI_ERR_NULL;
  if (d->mix && !W) return IWVI_ERR_NULL;
  if (d->mf == IWVI_MF_LINEAR && (!mfA || !mfb)) return IWVI_ERR_NULL;
  if ((d->flags & IWVI_FLAG_SAMPLE) && (!eps || !sample)) return IWVI_ERR_NULL;
  if ((d->flags & IWVI_FLAG_SAVE) && !save) return IWVI_ERR_NULL;
  if (point_begin < 0 || point_end > d->T || point_begin > point_end) return IWVI_ERR_BAD_DESC;
  if (d->T == 0 || point_begin == point_end) return IWVI_OK;
  int dev = 0, nsm = 148, max_smem = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  const AuxLayout al = iwvi_aux_layout(d->M, d->D, d->R);
  int smem_bytes = 0;
  const int TP = iwvi_pick_tp(d->T, al.Mp, al.ldz, nsm, max_smem, &smem_bytes);
  if (TP < 0) return IWVI_ERR_UNSUPPORTED;
  if (point_begin % TP) return IWVI_ERR_BAD_DESC;                       // ranges start on a tile boundary ...
  if (point_end != d->T && point_end % TP) return IWVI_ERR_BAD_DESC;    // ... and end on one, or at the last point
  FwdParams p;
  p.d = *d; p.Lm = Lm; p.aux = aux; p.X = X; p.W = W; p.mfA = mfA; p.mfb = mfb; p.eps = eps;
  p.sample = sample; p.mean = mean; p.var = var; p.save = save;
  p.tile0 = (int)(point_begin / TP);
  // when saving, the range that ends at the last point also covers the zero-padded rows of the saved arrays (at most
  // 128/TP - 1 extra, all-zero tiles)
  if (point_end == d->T)
    p.ntiles = (d->flags & IWVI_FLAG_SAVE) ? iwvi_save_layout(d->T, d->M, d->R).Tp / TP : (d->T + TP - 1) / TP;
  else
    p.ntiles = (int)(point_end / TP);
  const int count = p.ntiles - p.tile0;
  // A partial range is one of several chains running side by side on different streams: one tile per CTA lets the
  // block scheduler interleave the chains at tile granularity (measured: c3 3.38 -> 3.35 ms/step); a whole-range call
  // has the GPU to itself and keeps the persistent grid (the producer warp prefetches across tiles).
  const bool partial = point_begin != 0 || point_end != d->T;
  const int grid = (partial || count < nsm) ? count : nsm;
  cudaStream_t st = (cudaStream_t)stream;
  rc = TP == 64 ? launch_fwd<64>(p, smem_bytes, grid, st) : launch_fwd<32>(p, smem_bytes, grid, st);
  if (rc != IWVI_OK) return rc;
  if (d->flags & IWVI_FLAG_SAVE) {
    const int64_t npts = point_end - point_begin;
    const int egrid = (int)((npts + EPIF_PTS - 1) / EPIF_PTS);
    gp_epi_fwd_kernel<<<egrid, 256, 0, st>>>(p, (int)point_begin, (int)point_end);
    IWVI_CHECK_LAUNCH();
  }
  return IWVI_OK;
}

extern "C" int iwvi_gp_rows_fwd(const iwvi_gp_desc* d, const double* Lm, const double* aux, const double* X,
                                const double* W, const double* mfA, const double* mfb, const double* eps,
                                double* sample, double* mean, double* var, double* save, void* stream) {
  if (!d) return IWVI_ERR_NULL;
  return iwvi_gp_rows_fwd_range(d, Lm, aux, X, W, mfA, mfb, eps, sample, mean, var, save, 0, d->T, stream);
}
