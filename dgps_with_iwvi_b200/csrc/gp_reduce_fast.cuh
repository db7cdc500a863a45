// gp_reduce_fast.cuh -- OPTIONAL reduced-precision variant of gp_reduce_bwd_kernel (IWVI_FLAG_FAST_REDUCE), included by
// gp_rows_bwd.cu inside its namespace.  NOT the parity path: results carry FP32-grade error (about 1e-6 relative to the
// largest entry of each gradient, tests/test_gpu_fastpath.py) and are reported separately (bench.py "fast_path").
//
// The parameter half of the backward pass is three families of contractions over the T points,
//     dLq_r = 2 tril(A diag(gvar_bar_r) U_r^T),   dLm = -tril(Bbar A^T) (x 2: the panel holds Bbar / 2),   dq_mu = A gmean_bar,
// i.e. plain GEMMs with K = T -- the one part of the path with no triangular solve, no cancellation-prone variance and no
// elementwise stage in between, whose outputs are gradients (an SGD step tolerates 1e-6).  tcgen05.mma has no f64 kind,
// so every float64 operand x is split on the fly into two TF32 numbers hi = tf32(x), lo = tf32(x - hi) and each product is
// formed as hi*hi + hi*lo + lo*hi ("3xTF32") with FP32 accumulation in tensor memory; split-K partials are summed in
// float64 by the unchanged finalize kernel.
//
// One CTA = one 128 x 256 output tile of one matrix q over one range of points:
//   * 8 converter warps stream the saved panels from global memory (coalesced 256-byte row segments, several loads in
//     flight per lane), split them, apply the per-point scale to the B operand, and write both operands TRANSPOSED into
//     shared memory in the K-major no-swizzle canonical layout [row/8][k/4][row%8][k%4] (one 16-byte chunk per store);
//   * one elected thread of warp 8 issues the tcgen05.mma instructions (M = 128, N = 256, K = 8 each, accumulators in
//     TMEM) for a 32-point stage and hands the stage back through tcgen05.commit -> mbarrier;
//   * two 100 KB stages double-buffer conversion against the tensor core;
//   * four epilogue warps read the accumulators with tcgen05.ld and write float64 partials in the block layout the
//     finalize kernel expects.
#define FR_M 128            // output rows per CTA (TMEM lanes)
#define FR_N 256            // output columns per CTA (TMEM columns)
#define FR_KB 32            // points per stage
#define FR_CONV_WARPS 8
#define FR_THREADS ((FR_CONV_WARPS + 1) * 32)
#define FR_STAGES 2
#define FR_A_FLOATS (FR_M * FR_KB)          // per hi / lo copy
#define FR_B_FLOATS (FR_N * FR_KB)
#define FR_Q_FLOATS (16 * FR_KB)
#define FR_STAGE_FLOATS (2 * FR_A_FLOATS + 2 * FR_B_FLOATS + 2 * FR_Q_FLOATS)

__device__ __forceinline__ float fr_tf32(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}
__device__ __forceinline__ uint64_t fr_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ void fr_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void fr_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// 32 rows x 4 points of a saved block-major panel -> one 16-byte chunk per lane of the hi and lo operand copies.
// `src` points at (point n0, column m_base + lane) of the panel block; consecutive points are IWVI_LDS doubles apart.
__device__ __forceinline__ void fr_convert4(const double (&x)[4], const double (&sc)[4], float* hi, float* lo, int off) {
  float4 h, l;
  const double v0 = x[0] * sc[0], v1 = x[1] * sc[1], v2 = x[2] * sc[2], v3 = x[3] * sc[3];
  h.x = fr_tf32((float)v0); l.x = fr_tf32((float)(v0 - (double)h.x));
  h.y = fr_tf32((float)v1); l.y = fr_tf32((float)(v1 - (double)h.y));
  h.z = fr_tf32((float)v2); l.z = fr_tf32((float)(v2 - (double)h.z));
  h.w = fr_tf32((float)v3); l.w = fr_tf32((float)(v3 - (double)h.w));
  *reinterpret_cast<float4*>(hi + off) = h;
  *reinterpret_cast<float4*>(lo + off) = l;
}

struct FastItem { int q, m0, n0, s; };

__global__ void __launch_bounds__(FR_THREADS, 1) gp_reduce_fast_kernel(const BwdParams p) {
  extern __shared__ __align__(1024) unsigned char fr_smem_raw[];
  float* stage_base = reinterpret_cast<float*>(fr_smem_raw);
  __shared__ uint64_t full_bar[FR_STAGES], empty_bar[FR_STAGES], done_bar;
  __shared__ uint32_t tmem_base_s;
  const iwvi_gp_desc& d = p.d;
  const AuxLayout al = iwvi_aux_layout(d.M, d.D, d.R);
  const int NB = al.NB, R = d.R, Mp = al.Mp;
  const BwdWs& wl = p.wl;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  // ---- decode the work item: tile fastest, then matrix, point range slowest (all resident CTAs stream the same points)
  const int mts = Mp / FR_M;
  int n_tiles = 0;                                  // tiles (mt, nt) with n0 < m0 + FR_M (lower block triangle)
  for (int mt = 0; mt < mts; mt++) n_tiles += (mt * FR_M + FR_M + FR_N - 1) / FR_N;
  int item = blockIdx.x;
  int tile = item % n_tiles; item /= n_tiles;
  const int q = p.q_lo + item % p.q_n;
  const int s = item / p.q_n;
  int m0 = 0, n0 = 0;
  for (int mt = 0; mt < mts; mt++) {
    const int nn = (mt * FR_M + FR_M + FR_N - 1) / FR_N;
    if (tile < nn) { m0 = mt * FR_M; n0 = tile * FR_N; break; }
    tile -= nn;
  }
  const int ncols = min(FR_N, Mp - n0);             // real B rows of this tile (Mp = 128: 128)
  const int c0 = s * wl.chunks_per_split;
  const int c1 = min(c0 + wl.chunks_per_split, wl.Tp / IWVI_BLK);
  const int nkb = (c1 - c0) * (IWVI_BLK / FR_KB);   // stages of this CTA
  const bool is_lm = (q == R);
  const bool do_qmu = (q == 0 && n0 == 0);

  const SaveLayout sv = iwvi_save_layout(d.T, d.M, R);
  const double* A_T = p.save + sv.off_a;
  const double* P_op = is_lm ? p.ws + wl.off_bbar : A_T;                              // rows of the output
  const double* Q_op = is_lm ? A_T : p.save + sv.off_u + (size_t)q * sv.u_stride;   // columns of the output
  const double* gmb = p.ws + wl.off_gmb;
  const double* gvb = p.ws + wl.off_gvb;

  if (tid == 0) {
    for (int i = 0; i < FR_STAGES; i++) { mbar_init(&full_bar[i], FR_CONV_WARPS); mbar_init(&empty_bar[i], 1); }
    mbar_init(&done_bar, 1);
    mbar_fence_init();
  }
  if (warp == FR_CONV_WARPS) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;

  if (warp < FR_CONV_WARPS) {
    // ================= converters =================
    // unit of work: 32 columns (lanes) x 4 points -> one 16-byte chunk per lane.  A tile: 4 column groups x 8 point
    // groups = 32 units, B tile: (ncols / 32) x 8 units; units are dealt round-robin to the 8 warps.
    const int a_units = (FR_M / 32) * (FR_KB / 4), b_units = (ncols / 32) * (FR_KB / 4);
    for (int kb = 0; kb < nkb; kb++) {
      const int st = kb % FR_STAGES;
      if (kb >= FR_STAGES) mbar_wait(&empty_bar[st], ((kb / FR_STAGES) & 1u) ^ 1u);
      float* sb = stage_base + (size_t)st * FR_STAGE_FLOATS;
      float *a_hi = sb, *a_lo = sb + FR_A_FLOATS, *b_hi = sb + 2 * FR_A_FLOATS, *b_lo = b_hi + FR_B_FLOATS;
      float *q_hi = b_lo + FR_B_FLOATS, *q_lo = q_hi + FR_Q_FLOATS;
      const int chunk = c0 + kb / (IWVI_BLK / FR_KB);
      const int pr0 = (kb % (IWVI_BLK / FR_KB)) * FR_KB;                 // first point of the stage inside its chunk
      const size_t pt0 = (size_t)chunk * IWVI_BLK + pr0;
      const double one[4] = {1.0, 1.0, 1.0, 1.0};
      // all global loads of this warp's units are issued before the first conversion: ~100 KB in flight per SM per stage
      constexpr int MAXU = ((FR_M + FR_N) / 32) * (FR_KB / 4) / FR_CONV_WARPS;     // 12
      double x[MAXU][4];
#pragma unroll
      for (int i = 0; i < MAXU; i++) {
        const int u = warp + i * FR_CONV_WARPS;
        if (u < a_units + b_units) {
          const bool isA = u < a_units;
          const int uu = isA ? u : u - a_units;
          const int cg = uu / (FR_KB / 4), pg = uu % (FR_KB / 4);         // column group of 32, point group of 4
          const int col = (isA ? m0 : n0) + cg * 32 + lane;               // inducing index of this lane
          const double* src = (isA ? P_op : Q_op) + ((size_t)chunk * NB + (col >> 6)) * IWVI_STAGE_DOUBLES +
                              (size_t)(pr0 + pg * 4) * IWVI_LDS + (col & 63);
#pragma unroll
          for (int j = 0; j < 4; j++) x[i][j] = __ldg(src + j * IWVI_LDS);
        }
      }
#pragma unroll
      for (int i = 0; i < MAXU; i++) {
        const int u = warp + i * FR_CONV_WARPS;
        if (u < a_units + b_units) {
          const bool isA = u < a_units;
          const int uu = isA ? u : u - a_units;
          const int cg = uu / (FR_KB / 4), pg = uu % (FR_KB / 4);
          const int row = cg * 32 + lane;                                  // operand row inside the tile
          const int off = ((row >> 3) * (FR_KB / 4) + pg) * 32 + (row & 7) * 4;
          if (isA) {
            fr_convert4(x[i], one, a_hi, a_lo, off);
          } else {
            double sc[4];
#pragma unroll
            for (int j = 0; j < 4; j++)
              sc[j] = is_lm ? -2.0 : 2.0 * __ldg(gvb + (pt0 + pg * 4 + j) * IWVI_MAX_R + q);
            fr_convert4(x[i], sc, b_hi, b_lo, off);
          }
        }
      }
      if (do_qmu && warp == 0) {
        // gmean_bar of the stage's 32 points as 16 extra output columns (8 real): row r, point n
        const int r = lane & 15, half = lane >> 4;
#pragma unroll
        for (int pg2 = 0; pg2 < (FR_KB / 4) / 2; pg2++) {
          const int pg = pg2 * 2 + half;
          double x[4];
#pragma unroll
          for (int j = 0; j < 4; j++) x[j] = r < IWVI_MAX_R ? __ldg(gmb + (pt0 + pg * 4 + j) * IWVI_MAX_R + r) : 0.0;
          const int off = ((r >> 3) * (FR_KB / 4) + pg) * 32 + (r & 7) * 4;
          fr_convert4(x, one, q_hi, q_lo, off);
        }
      }
      // this thread's generic-proxy writes -> visible to the async proxy (tensor core), then one arrival per warp
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&full_bar[st]);
    }
  } else {
   if (lane == 0) {
    // ================= MMA issuer (one thread) =================
    const uint32_t idesc_n = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(ncols >> 3) << 17) | ((uint32_t)(FR_M >> 4) << 24);
    const uint32_t idesc_q = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(FR_M >> 4) << 24);
    const uint32_t lbo = 128, sbo = (FR_KB / 4) * 128;
    for (int kb = 0; kb < nkb; kb++) {
      const int st = kb % FR_STAGES;
      mbar_wait(&full_bar[st], (kb / FR_STAGES) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t sa = smem_u32(stage_base + (size_t)st * FR_STAGE_FLOATS);
      const uint32_t a_hi = sa, a_lo = sa + FR_A_FLOATS * 4, b_hi = sa + 2 * FR_A_FLOATS * 4, b_lo = b_hi + FR_B_FLOATS * 4;
      const uint32_t q_hi = b_lo + FR_B_FLOATS * 4, q_lo = q_hi + FR_Q_FLOATS * 4;
#pragma unroll
      for (int k0 = 0; k0 < FR_KB; k0 += 8) {
        const uint32_t ko = (k0 / 4) * 128;
        const uint32_t acc = (kb > 0 || k0 > 0) ? 1u : 0u;
        fr_mma(tmem_base, fr_desc(a_hi + ko, lbo, sbo), fr_desc(b_hi + ko, lbo, sbo), idesc_n, acc);
        fr_mma(tmem_base, fr_desc(a_hi + ko, lbo, sbo), fr_desc(b_lo + ko, lbo, sbo), idesc_n, 1u);
        fr_mma(tmem_base, fr_desc(a_lo + ko, lbo, sbo), fr_desc(b_hi + ko, lbo, sbo), idesc_n, 1u);
        if (do_qmu) {
          fr_mma(tmem_base + FR_N, fr_desc(a_hi + ko, lbo, sbo), fr_desc(q_hi + ko, lbo, sbo), idesc_q, acc);
          fr_mma(tmem_base + FR_N, fr_desc(a_hi + ko, lbo, sbo), fr_desc(q_lo + ko, lbo, sbo), idesc_q, 1u);
          fr_mma(tmem_base + FR_N, fr_desc(a_lo + ko, lbo, sbo), fr_desc(q_hi + ko, lbo, sbo), idesc_q, 1u);
        }
      }
      fr_commit(&empty_bar[st]);          // the stage may be overwritten once these MMAs have read it
    }
    fr_commit(&done_bar);
   }
   __syncwarp();    // the issuing warp reconverges before the block-wide barrier below
  }

  // ================= epilogue: warps 0-3 own TMEM lanes 32 w .. 32 w + 31 = output rows m0 + 32 w + lane =================
  if (warp < 4) {
    mbar_wait(&done_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int row = warp * 32 + lane;                  // row inside the tile
    const int m = m0 + row, bi = m >> 6, rb = m & 63;
    double* red = p.ws + wl.off_red;
    for (int cb = 0; cb < ncols; cb += 8) {
      uint32_t r[8];
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)cb;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                   : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      const int n = n0 + cb, bj = n >> 6;
      if (bj <= bi) {                                   // lower block triangle only: the finalize kernel reads nothing else
        double* out = red + (((size_t)q * wl.S + s) * wl.npairs + iwvi_pair(bi, bj)) * IWVI_BLK * IWVI_BLK +
                      (size_t)rb * IWVI_BLK + (n & 63);
#pragma unroll
        for (int j = 0; j < 8; j++) out[j] = (double)__uint_as_float(r[j]);
      }
    }
    if (do_qmu) {
      uint32_t r[8];
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)FR_N;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                   : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      double* oq = p.ws + wl.off_qred + ((size_t)s * NB + bi) * IWVI_BLK * IWVI_MAX_R + (size_t)rb * IWVI_MAX_R;
#pragma unroll
      for (int j = 0; j < 8; j++) oq[j] = (double)__uint_as_float(r[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == FR_CONV_WARPS)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
}

// number of CTAs of a fast reduce launch over `nq` matrices
inline int fast_reduce_tiles(int Mp) {
  int n = 0;
  for (int mt = 0; mt < Mp / FR_M; mt++) n += (mt * FR_M + FR_M + FR_N - 1) / FR_N;
  return n;
}
