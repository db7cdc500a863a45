// gp_gram_bwd.cu -- gram adjoint of the per-point stage: the cotangent Bbar of Kuf (left by gp_tile_bwd_kernel) pulled
// back through the kernel function to the layer's inputs, the inducing inputs and the kernel parameters.  The reference
// obtains it from tf.gradients through Kuf / the stationary kernels (temp_workaround.py:36, GPflow kernels.py
// `scaled_square_dist`, restated in SURVEY.md A.1); formulas in DESIGN.md section 4 and oracle/staged_np.py.
//
//   G      = Bbar * dK/dr2          (elementwise, [M, T];  r2 = |z~|^2 + |x~|^2 - 2 z~ x~^T on length-scaled inputs)
//   dX    += 2/ls (x~ colsum(G) - G^T z~)
//   dZ     = -2/ls (G x~ - z~ rowsum(G))            per-CTA partials, summed in fixed order by gp_finalize_bwd_kernel
//   dls    = -2/ls sum_mn G (x~ - z~)^2             (from the by-products of dX / dZ: the expanded square)
//   dvar   = sum_mn Bbar K / variance
//
// Why a kernel of its own.  These phases used to close every tile of gp_tile_bwd_kernel (20 % of its time for 6 % of
// its flops): short dependent chains -- a 20-deep gram product, the kernel function, skinny DMMAs, shuffles -- on a
// kernel that keeps ONE CTA of eight warps per SM because its panel fills the shared memory, so nothing hid their
// latencies.  Here the same arithmetic runs with two CTAs (sixteen warps) per SM, no block-wide barrier inside the
// work loop and no shared-memory round trip for G:
//   * a CTA walks over strips of GRAM_PTS = 16 points (32 would need ~160 registers per thread and spill at the 128
//     that two CTAs per SM allow); per strip and 64-row block of inducing points, warp w owns rows 8w..8w+7: r2 as one
//     8 x 16 DMMA tile row, K and dK/dr2 in registers, G in the accumulator layout;
//   * G x~ uses the accumulator registers directly as the DMMA A operand (the k index of an m8n8k4 product is free to
//     permute: k-step (b, c) pairs G[m][8b + 2t + c] with x~[8b + 2t + c][.]);
//   * G^T z~ needs G as a B operand: two shuffles per fragment move it there, no shared memory;
//   * when the input width is 8 k + 1 (the data dimensions plus ONE latent dimension: every shipped configuration) the
//     last column is handled by scalar FMAs instead of a whole DMMA column tile that would be 7/8 padding;
//   * dZ accumulates in shared memory by plain read-modify-write (each (m, d) has one owner thread in the CTA) and is
//     written once per CTA; dX has one owner per (point, d) in the whole grid.  No atomics: bit-reproducible.
#include "gp_bwd.cuh"

namespace {

#define GRAM_SLOT_LD (GRAM_PTS + 8)   // slot row stride (24 or 40): 128-bit stores of (2t, 2t+1) pairs are conflict-free
#define GRAM_NSLOT 4         // warps 4-7 hand their strip sums to warps 0-3 through these

struct GramSmem { int dz, zt, zn, xs, xn, slots, dlx, red, total_doubles, ldd, slot_rows; };
__host__ __device__ inline GramSmem gram_smem_layout(int Mp, int ldz, int D, int ndx, int xcol, int pts, bool zsmem) {
  GramSmem s; int o = 0;
  s.ldd = D | 1;                                  // odd row stride of the dZ accumulator
  s.slot_rows = ndx + xcol + 1;                   // dX^T rows of the DMMA tiles, the extra column, colsum(G)
  s.dz = o;    o += Mp * s.ldd;
  o = (o + 1) & ~1;
  // length-scaled inducing inputs and their squared norms: a shared-memory copy when that still leaves room for two
  // CTAs per SM (ZSMEM), else read through L1
  s.zt = o;    o += zsmem ? Mp * ldz : 0;
  s.zn = o;    o += zsmem ? Mp : 0;
  s.xs = o;    o += 2 * pts * ldz;                // length-scaled inputs of the strip, double buffered
  s.xn = o;    o += 2 * pts;
  s.slots = o; o += GRAM_NSLOT * s.slot_rows * GRAM_SLOT_LD;
  s.dlx = o;   o += 32 * pts;                     // x part of the lengthscale adjoint per (d, point of a strip)
  s.red = o;   o += 9 * 32;
  s.total_doubles = o;
  return s;
}

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
// A load that stays where it is written: ptxas sinks ordinary loads down to their first use to save registers, which
// turns a prefetch back into an exposed HBM round trip (volatile asm keeps its order relative to the DMMAs, also volatile).
__device__ __forceinline__ double ldg_here(const double* p) {
  double v;
  asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}

template <int KIND, int ND8, bool XCOL, int NBT, bool ZSMEM>
__global__ void __launch_bounds__(GRAM_THREADS, GRAM_CTAS_PER_SM) gp_gram_bwd_kernel(const BwdParams p) {
  extern __shared__ __align__(16) double smem[];
  const iwvi_gp_desc& d = p.d;
  const AuxLayout al = iwvi_aux_layout(d.M, d.D, d.R);
  const int Mp = al.Mp, NB = al.NB, ldz = al.ldz, D = d.D, T = d.T, M = d.M;
  constexpr int NDX = ND8 * 8;                    // input columns covered by DMMA column tiles
  constexpr int PTS = 8 * NBT;                    // points per strip
  constexpr int KS = XCOL ? ND8 * 2 : 0;          // k-steps of the gram product (XCOL: exactly the tiled columns)
  const int ks_n = XCOL ? KS : iwvi_round_up(D, 4) / 4;
  const GramSmem sl = gram_smem_layout(Mp, ldz, D, NDX, XCOL ? 1 : 0, PTS, ZSMEM);
  const int ldd = sl.ldd;
  double* dz_s = smem + sl.dz;
  double* xs_b = smem + sl.xs;
  double* xn_b = smem + sl.xn;
  double* slots = smem + sl.slots;
  double* dlx_s = smem + sl.dlx;
  double* red = smem + sl.red;
  constexpr int ROW_E = NDX, ROW_CS = NDX + (XCOL ? 1 : 0);
  const int slot_doubles = sl.slot_rows * GRAM_SLOT_LD;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const double* aux = p.aux;
  const double* zt = ZSMEM ? smem + sl.zt : aux + al.off_zt;
  const double* zn = ZSMEM ? smem + sl.zn : aux + al.off_zn;
  if (ZSMEM) {
    for (int idx = tid; idx < Mp * ldz; idx += GRAM_THREADS) smem[sl.zt + idx] = __ldg(aux + al.off_zt + idx);
    for (int idx = tid; idx < Mp; idx += GRAM_THREADS) smem[sl.zn + idx] = __ldg(aux + al.off_zn + idx);
  }
  const double* consts = aux + al.off_consts;
  const double variance = consts[IWVI_C_VARIANCE];
  const double* bbar = p.ws + p.wl.off_bbar;
  double* mypart = p.ws + p.wl.off_tile + (size_t)(p.slot0 + blockIdx.x) * p.wl.tile_stride;

  for (int idx = tid; idx < Mp * ldd; idx += GRAM_THREADS) dz_s[idx] = 0.0;
  for (int idx = tid; idx < 32 * PTS; idx += GRAM_THREADS) dlx_s[idx] = 0.0;
  // one of two point chains (iwvi_gp_rows_bwd_range): the finalize kernel sums ALL slots of both chains, and the slots of
  // this chain beyond this launch's grid may hold partials of an earlier call -- clear them
  if (p.n_slots > 0)
    for (int s2 = blockIdx.x + gridDim.x; s2 < p.n_slots; s2 += gridDim.x) {
      double* other = p.ws + p.wl.off_tile + (size_t)(p.slot0 + s2) * p.wl.tile_stride;
      for (int idx = tid; idx < p.wl.tile_stride; idx += GRAM_THREADS) other[idx] = 0.0;
    }

  double dvar_acc = 0.0, dlze = 0.0;
  double dlz[ND8][2];                            // z part of the lengthscale adjoint, columns d = 8 dt + 2 t + c
#pragma unroll
  for (int dt = 0; dt < ND8; dt++) { dlz[dt][0] = 0.0; dlz[dt][1] = 0.0; }

  // loader of a strip's inputs: 8 threads per point, thread j takes columns j, j + 8, j + 16, j + 24
  const int xl_n = tid >> 3, xl_j = tid & 7;
  auto load_x = [&](int strip, double (&xv)[4]) {
    const size_t pt = (size_t)strip * PTS + xl_n;
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const int k = xl_j + 8 * q;
      xv[q] = (xl_n < PTS && k < D && pt < (size_t)T) ? ldg_here(p.X + pt * D + k) : 0.0;
    }
  };
  auto store_x = [&](int buf, const double (&xv)[4]) {
    double* xs = xs_b + buf * PTS * ldz;
    double ss = 0.0;
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const int k = xl_j + 8 * q;
      const double v = (k < D) ? xv[q] * consts[IWVI_C_INVLS + k] : 0.0;
      if (xl_n < PTS && k < ldz) xs[xl_n * ldz + k] = v;
      ss += v * v;
    }
    if (xl_n < PTS && ldz > 32 && xl_j < ldz - 32) xs[xl_n * ldz + 32 + xl_j] = 0.0;
    ss += __shfl_xor_sync(0xffffffffu, ss, 1);
    ss += __shfl_xor_sync(0xffffffffu, ss, 2);
    ss += __shfl_xor_sync(0xffffffffu, ss, 4);
    if (xl_n < PTS && xl_j == 0) xn_b[buf * PTS + xl_n] = ss;
  };

  auto bb_of = [&](int n0) {
    return bbar + ((int64_t)(n0 >> 6) * NB * IWVI_BLK + (n0 & 63) + 2 * t) * IWVI_LDS + warp * 8 + g;
  };
  int s = p.strip0 + blockIdx.x, buf = 0;
  double bb_pf[NBT][2];
  const double* bb_base = bb_of(s * PTS);
  if (s < p.strip1) {
    double xv[4];
    load_x(s, xv);
#pragma unroll
    for (int b = 0; b < NBT; b++)
#pragma unroll
      for (int c = 0; c < 2; c++) bb_pf[b][c] = ldg_here(bb_base + (8 * b + c) * IWVI_LDS);
    store_x(0, xv);
  }
  __syncthreads();

  for (; s < p.strip1; s += gridDim.x, buf ^= 1, bb_base = bb_of(s * PTS)) {
    const int n0 = s * PTS;
    const double* xs = xs_b + buf * PTS * ldz;
    const double* xn = xn_b + buf * PTS;
    const bool has_next = s + (int)gridDim.x < p.strip1;
    double xv_next[4];
    if (has_next) load_x(s + gridDim.x, xv_next);     // consumed after the block loop: the latency hides behind it

    double accx[ND8][NBT][2];                           // dX^T tiles: row d = 8 dt + g, column n = 8 b + 2 t + c
    double accxe[NBT][2], cs[NBT][2], xnr[NBT][2], xe[NBT][2];
#pragma unroll
    for (int b = 0; b < NBT; b++)
#pragma unroll
      for (int c = 0; c < 2; c++) {
        const int n = 8 * b + 2 * t + c;
#pragma unroll
        for (int dt = 0; dt < ND8; dt++) accx[dt][b][c] = 0.0;
        accxe[b][c] = 0.0; cs[b][c] = 0.0;
        xnr[b][c] = xn[n];
        xe[b][c] = XCOL ? xs[n * ldz + NDX] : 0.0;
      }
    // Bbar / 2 of (point n0 + 8 b + 2 t + c, row 8 warp + g of block i), block-major [chunk][m-block][64][68]: fetched
    // one pass ahead (it comes from HBM / L2; everything else a pass reads is L1- or shared-memory resident)
    const int n0_next = (s + (int)gridDim.x) * PTS;
    const double* bb_next_strip = bb_of(n0_next);

    for (int i = 0; i < NB; i++) {
      const int mg = i * IWVI_BLK + warp * 8 + g;
      const double* zrow = zt + (size_t)mg * ldz;
      double bb[NBT][2];
#pragma unroll
      for (int b = 0; b < NBT; b++)
#pragma unroll
        for (int c = 0; c < 2; c++) bb[b][c] = bb_pf[b][c];
      if (i + 1 < NB || has_next) {
        const double* src = (i + 1 < NB) ? bb_base + (int64_t)(i + 1) * IWVI_STAGE_DOUBLES : bb_next_strip;
#pragma unroll
        for (int b = 0; b < NBT; b++)
#pragma unroll
          for (int c = 0; c < 2; c++) bb_pf[b][c] = ldg_here(src + (8 * b + c) * IWVI_LDS);
      }
      const double znr = zn[mg];
      const double ze = XCOL ? zrow[NDX] : 0.0;

      // ---- r2 tile: 8 rows x PTS points
      double acc[NBT][2];
#pragma unroll
      for (int b = 0; b < NBT; b++) { acc[b][0] = 0.0; acc[b][1] = 0.0; }
      if (XCOL) {
#pragma unroll
        for (int ks = 0; ks < KS; ks++) {
          const double a = zrow[4 * ks + t];
#pragma unroll
          for (int b = 0; b < NBT; b++) dmma884(acc[b], a, xs[(8 * b + g) * ldz + 4 * ks + t]);
        }
      } else {
        for (int ks = 0; ks < ks_n; ks++) {
          const double a = zrow[4 * ks + t];
#pragma unroll
          for (int b = 0; b < NBT; b++) dmma884(acc[b], a, xs[(8 * b + g) * ldz + 4 * ks + t]);
        }
      }
      double kv[2 * NBT], dkv[2 * NBT];
#pragma unroll
      for (int b = 0; b < NBT; b++)
#pragma unroll
        for (int c = 0; c < 2; c++) {
          double sxy = acc[b][c];
          if (XCOL) sxy = fma(ze, xe[b][c], sxy);
          kv[b * 2 + c] = znr + xnr[b][c] - 2.0 * sxy;
        }
      // RBF: dK/dr2 = -K/2, so G = 2 (Bbar/2) dK/dr2 = -(Bbar/2) K and sum Bbar K = -2 sum G: neither dK/dr2 nor the
      // variance adjoint needs instructions of its own (the latter falls out of the column sums, once per strip)
      kern_n<KIND, 2 * NBT, KIND != IWVI_KERN_RBF>(kv, dkv, variance);
      double rs = 0.0;
      const bool mvalid = mg < M;
#pragma unroll
      for (int b = 0; b < NBT; b++)
#pragma unroll
        for (int c = 0; c < 2; c++) {
          const bool valid = mvalid && (n0 + 8 * b + 2 * t + c < T);
          double G;
          if (KIND == IWVI_KERN_RBF) {
            G = valid ? -(bb[b][c] * kv[b * 2 + c]) : 0.0;
          } else {
            const double bbv = 2.0 * bb[b][c];
            G = valid ? bbv * dkv[b * 2 + c] : 0.0;
            if (valid) dvar_acc = fma(bbv, kv[b * 2 + c], dvar_acc);
          }
          acc[b][c] = G;
          cs[b][c] += G;
          rs += G;
        }
      rs += __shfl_xor_sync(0xffffffffu, rs, 1);
      rs += __shfl_xor_sync(0xffffffffu, rs, 2);

      // ---- dZ rows of this warp: (G x~)[m][d], G straight from the accumulator registers as the A operand
      double accz[ND8 > 0 ? ND8 : 1][2];
#pragma unroll
      for (int dt = 0; dt < ND8; dt++) { accz[dt][0] = 0.0; accz[dt][1] = 0.0; }
#pragma unroll
      for (int b = 0; b < NBT; b++)
#pragma unroll
        for (int c = 0; c < 2; c++) {
          const double* xr = xs + (8 * b + 2 * t + c) * ldz + g;
#pragma unroll
          for (int dt = 0; dt < ND8; dt++) {
            const double bv = (8 * dt + 8 <= 20 || 8 * dt + g < ldz) ? xr[8 * dt] : 0.0;
            dmma884(accz[dt], acc[b][c], bv);
          }
        }
#pragma unroll
      for (int dt = 0; dt < ND8; dt++)
#pragma unroll
        for (int c = 0; c < 2; c++) {
          const int dcol = 8 * dt + 2 * t + c;
          if (8 * dt + 8 <= 20 || dcol < ldz) {
            const double zv = zrow[dcol];
            if (dcol < D) dz_s[mg * ldd + dcol] += accz[dt][c] - zv * rs;
            dlz[dt][c] = fma(zv, rs * zv - 2.0 * accz[dt][c], dlz[dt][c]);
          }
        }
      if (XCOL) {
        double ez = 0.0;
#pragma unroll
        for (int b = 0; b < NBT; b++)
#pragma unroll
          for (int c = 0; c < 2; c++) ez = fma(acc[b][c], xe[b][c], ez);
        ez += __shfl_xor_sync(0xffffffffu, ez, 1);
        ez += __shfl_xor_sync(0xffffffffu, ez, 2);
        if (t == 0) {
          dz_s[mg * ldd + NDX] += ez - ze * rs;
          dlze = fma(ze, rs * ze - 2.0 * ez, dlze);
        }
      }

      // ---- dX^T tiles: (z~^T G)[d][n]; the B fragment G[4 h + t][8 b + g] lives in lane (4 h + t, g / 2), entry g % 2
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const int src = (4 * h + t) * 4 + (g >> 1);
        const double* zr2 = zt + (size_t)(i * IWVI_BLK + warp * 8 + 4 * h + t) * ldz + g;
        double za[ND8 > 0 ? ND8 : 1];
#pragma unroll
        for (int dt = 0; dt < ND8; dt++) za[dt] = (8 * dt + 8 <= 20 || 8 * dt + g < ldz) ? zr2[8 * dt] : 0.0;
#pragma unroll
        for (int b = 0; b < NBT; b++) {
          const double v0 = shfl_d(acc[b][0], src), v1 = shfl_d(acc[b][1], src);
          const double bf = (g & 1) ? v1 : v0;
#pragma unroll
          for (int dt = 0; dt < ND8; dt++) dmma884(accx[dt][b], za[dt], bf);
        }
      }
      if (XCOL) {
#pragma unroll
        for (int b = 0; b < NBT; b++)
#pragma unroll
          for (int c = 0; c < 2; c++) accxe[b][c] = fma(acc[b][c], ze, accxe[b][c]);
      }
    }

    // ---- strip sums: over the 8 rows of a warp (lanes g), then over the warps through the slots
#pragma unroll
    for (int b = 0; b < NBT; b++)
#pragma unroll
      for (int c = 0; c < 2; c++) {
        double v = cs[b][c];
        if (KIND == IWVI_KERN_RBF) dvar_acc = fma(-2.0, v, dvar_acc);     // sum Bbar K over this thread's entries of the strip
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        cs[b][c] = v;
        if (XCOL) {
          double e = accxe[b][c];
          e += __shfl_xor_sync(0xffffffffu, e, 4);
          e += __shfl_xor_sync(0xffffffffu, e, 8);
          e += __shfl_xor_sync(0xffffffffu, e, 16);
          accxe[b][c] = e;
        }
      }
    {
      double* slot = slots + (warp & (GRAM_NSLOT - 1)) * slot_doubles;
      const bool upper = warp >= GRAM_NSLOT;
      auto exchange = [&](bool add) {
#pragma unroll
        for (int b = 0; b < NBT; b++) {
#pragma unroll
          for (int dt = 0; dt < ND8; dt++) {
            double2* q = reinterpret_cast<double2*>(slot + (8 * dt + g) * GRAM_SLOT_LD + 8 * b + 2 * t);
            double2 v = make_double2(accx[dt][b][0], accx[dt][b][1]);
            if (add) { const double2 o = *q; v.x += o.x; v.y += o.y; }
            *q = v;
          }
          if (g == 0) {
            double2* q = reinterpret_cast<double2*>(slot + ROW_CS * GRAM_SLOT_LD + 8 * b + 2 * t);
            double2 v = make_double2(cs[b][0], cs[b][1]);
            if (add) { const double2 o = *q; v.x += o.x; v.y += o.y; }
            *q = v;
            if (XCOL) {
              double2* qe = reinterpret_cast<double2*>(slot + ROW_E * GRAM_SLOT_LD + 8 * b + 2 * t);
              double2 ve = make_double2(accxe[b][0], accxe[b][1]);
              if (add) { const double2 o = *qe; ve.x += o.x; ve.y += o.y; }
              *qe = ve;
            }
          }
        }
      };
      if (upper) exchange(false);
      __syncthreads();
      if (!upper) exchange(true);
      __syncthreads();
    }
    // ---- dX += 2/ls (x~ colsum(G) - G^T z~): one owner thread per (point, d)
    {
      const int n = tid & (PTS - 1);
      const size_t pt = (size_t)n0 + n;
      double gsv = 0.0;
#pragma unroll
      for (int k = 0; k < GRAM_NSLOT; k++) gsv += slots[k * slot_doubles + ROW_CS * GRAM_SLOT_LD + n];
      for (int dcol = tid / PTS; dcol < D; dcol += GRAM_THREADS / PTS) {
        double sum = 0.0;
#pragma unroll
        for (int k = 0; k < GRAM_NSLOT; k++) sum += slots[k * slot_doubles + dcol * GRAM_SLOT_LD + n];
        if (pt < (size_t)T) {
          const double xv = xs[n * ldz + dcol];
          // (one adder per address in this launch, on top of what gp_epi_bwd_kernel stored: order is fixed)
          red_add(&p.dX[pt * D + dcol], 2.0 * consts[IWVI_C_INVLS + dcol] * (xv * gsv - sum));
          dlx_s[dcol * PTS + n] += gsv * xv * xv;
        }
      }
    }
    if (has_next) store_x(buf ^ 1, xv_next);
    __syncthreads();   // slots and this strip's inputs are free; the next strip's inputs are in place
  }

  // ---- per-CTA partials: dZ [Mp, ldz], dls [32], dvariance
  {
    const double v = warp_sum(dvar_acc);
    const double e = warp_sum(dlze);
    if (lane == 0) { red[8 * 32 + warp] = v; red[8 * 32 + 8 + warp] = e; }
  }
#pragma unroll
  for (int dt = 0; dt < ND8; dt++)
#pragma unroll
    for (int c = 0; c < 2; c++) {
      double v = dlz[dt][c];
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      dlz[dt][c] = v;
    }
  if (g == 0) {
#pragma unroll
    for (int dt = 0; dt < 4; dt++)
#pragma unroll
      for (int c = 0; c < 2; c++) red[warp * 32 + 8 * dt + 2 * t + c] = dt < ND8 ? dlz[dt < ND8 ? dt : 0][c] : 0.0;
  }
  __syncthreads();
  for (int k = lane; k < ldz; k += 32) {
    const double sc = (k < D) ? -2.0 * consts[IWVI_C_INVLS + k] : 0.0;
    for (int m = warp; m < Mp; m += GRAM_THREADS / 32) mypart[m * ldz + k] = (k < D) ? sc * dz_s[m * ldd + k] : 0.0;
  }
  if (tid < 32) {
    double sacc = 0.0;
    for (int w = 0; w < 8; w++) sacc += red[w * 32 + tid];
    if (XCOL && tid == NDX)
      for (int w = 0; w < 8; w++) sacc += red[8 * 32 + 8 + w];
    if (tid < D)
      for (int n = 0; n < PTS; n++) sacc += dlx_s[tid * PTS + n];
    mypart[(size_t)Mp * ldz + tid] = (tid < D) ? -2.0 * consts[IWVI_C_INVLS + tid] * sacc : 0.0;
  }
  if (tid == 32) {
    double tot = 0.0;
    for (int w = 0; w < 8; w++) tot += red[8 * 32 + w];
    mypart[(size_t)Mp * ldz + 32] = tot / variance;
  }
}

template <int KIND, int ND8, bool XCOL, bool ZSMEM>
int launch_gram_z(const BwdParams& p, int nsm, int max_smem, cudaStream_t st) {
  constexpr int NBT = GRAM_PTS / 8;
  const AuxLayout al = iwvi_aux_layout(p.d.M, p.d.D, p.d.R);
  const GramSmem sl = gram_smem_layout(al.Mp, al.ldz, p.d.D, ND8 * 8, XCOL ? 1 : 0, GRAM_PTS, ZSMEM);
  const int smem_bytes = sl.total_doubles * 8;
  if (smem_bytes > max_smem) return IWVI_ERR_UNSUPPORTED;
  if (cudaFuncSetAttribute(gp_gram_bwd_kernel<KIND, ND8, XCOL, NBT, ZSMEM>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes) !=
      cudaSuccess)
    return IWVI_ERR_LAUNCH;
  const int ns = p.strip1 - p.strip0;
  if (ns <= 0) return IWVI_OK;
  // equal shares: every CTA takes ceil(ns / slots) strips (the last ones may take one fewer)
  const int per = (ns + p.wl.gram_slots - 1) / p.wl.gram_slots;
  const int grid = (ns + per - 1) / per;
  (void)nsm;
  gp_gram_bwd_kernel<KIND, ND8, XCOL, NBT, ZSMEM><<<grid, GRAM_THREADS, smem_bytes, st>>>(p);
  IWVI_CHECK_LAUNCH();
  return IWVI_OK;
}

template <int KIND, int ND8, bool XCOL>
int launch_gram_v(const BwdParams& p, int nsm, int max_smem, cudaStream_t st) {
  const AuxLayout al = iwvi_aux_layout(p.d.M, p.d.D, p.d.R);
  const int with_z = gram_smem_layout(al.Mp, al.ldz, p.d.D, ND8 * 8, XCOL ? 1 : 0, GRAM_PTS, true).total_doubles * 8;
  const int two_ctas = (228 * 1024) / GRAM_CTAS_PER_SM - 1024;   // shared memory per SM, 1 KB reserved per CTA
  if (with_z <= two_ctas && with_z <= max_smem) return launch_gram_z<KIND, ND8, XCOL, true>(p, nsm, max_smem, st);
  return launch_gram_z<KIND, ND8, XCOL, false>(p, nsm, max_smem, st);
}

template <int KIND>
int launch_gram_k(const BwdParams& p, int nsm, int max_smem, cudaStream_t st) {
  const int D = p.d.D;
  if (D > 1 && D % 8 == 1) {
    switch (D / 8) {
      case 1: return launch_gram_v<KIND, 1, true>(p, nsm, max_smem, st);
      case 2: return launch_gram_v<KIND, 2, true>(p, nsm, max_smem, st);
      default: return launch_gram_v<KIND, 3, true>(p, nsm, max_smem, st);
    }
  }
  switch ((D + 7) / 8) {
    case 1: return launch_gram_v<KIND, 1, false>(p, nsm, max_smem, st);
    case 2: return launch_gram_v<KIND, 2, false>(p, nsm, max_smem, st);
    case 3: return launch_gram_v<KIND, 3, false>(p, nsm, max_smem, st);
    default: return launch_gram_v<KIND, 4, false>(p, nsm, max_smem, st);
  }
}

}  // namespace

int iwvi_gram_bwd_grid(const BwdParams& p) {
  const int ns = p.strip1 - p.strip0;
  if (ns <= 0) return 0;
  const int per = (ns + p.wl.gram_slots - 1) / p.wl.gram_slots;
  return (ns + per - 1) / per;
}

int iwvi_launch_gram_bwd(const BwdParams& p, int nsm, int max_smem, cudaStream_t st) {
  switch (p.d.kern) {
    case IWVI_KERN_RBF: return launch_gram_k<IWVI_KERN_RBF>(p, nsm, max_smem, st);
    case IWVI_KERN_MATERN12: return launch_gram_k<IWVI_KERN_MATERN12>(p, nsm, max_smem, st);
    case IWVI_KERN_MATERN32: return launch_gram_k<IWVI_KERN_MATERN32>(p, nsm, max_smem, st);
    default: return launch_gram_k<IWVI_KERN_MATERN52>(p, nsm, max_smem, st);
  }
}
