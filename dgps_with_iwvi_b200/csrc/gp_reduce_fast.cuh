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
// One CTA = one 256 x 256 output tile (two 128-row halves = two accumulators in the 512 columns of tensor memory, sharing the
// B operand) of one matrix q over one range of points:
//   * 16 converter warps read the saved float64 panels straight from global memory (coalesced 256-byte row segments; the
//     loads of stage k + 1 are in flight -- 64 KB per SM -- while stage k is converted: two register buffers of 4 units),
//     split every value into TF32 hi / lo, apply the per-point scale to the B operand, and write both operands TRANSPOSED
//     into shared memory in the K-major no-swizzle canonical layout [row/8][k/4][row%8][k%4] (one 16-byte chunk per
//     store) -- a 3-deep ring of 64 KB operand stages of 16 points.  (Staging the float64 data in shared memory first, by
//     bulk TMA or cp.async, was measured slower: 4-9 KB bulk copies are bound by the TMA engine's per-operation cost,
//     0.3-0.4 us each, and a second ring costs a block-wide barrier per stage.)
//   * one elected thread issues the tcgen05.mma instructions (M = 128, N <= 256, K = 8: two k-steps per stage, 3 products x
//     2 halves) and hands the stage back through tcgen05.commit -> mbarrier;
//   * four epilogue warps read the accumulators with tcgen05.ld and write float64 partials in the block layout the
//     finalize kernel expects.
// dq_mu = A gmean_bar (8 columns) stays on the float64 kernel (a launch of its (q = 0, block column 0) items only): the
// 256 x 256 tile leaves no tensor memory for a second accumulator.
#define FR_M 256            // output rows per CTA: two halves of 128 TMEM lanes
#define FR_N 256            // output columns per CTA (TMEM columns per half)
#define FR_KB 16            // points per stage = two TF32 k-steps
#define FR_CONV_WARPS 16
#define FR_THREADS ((FR_CONV_WARPS + 1) * 32)   // + the MMA-issuing warp
#define FR_STAGES 3          // TF32 operand stages
#define FR_A_FLOATS (FR_M * FR_KB)          // per hi / lo copy
#define FR_B_FLOATS (FR_N * FR_KB)
#define FR_STAGE_FLOATS (2 * FR_A_FLOATS + 2 * FR_B_FLOATS)
#define FR_MAX_RANGE 4096    // points per CTA at most (scale table in shared memory)

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
// Split of 4 scaled float64 values into TF32 hi / lo pairs, one 16-byte chunk each.  One f64 -> f32 conversion per value;
// the rest is FP32 / integer work: hi = the float with its low 13 mantissa bits cleared (exactly what the tensor core keeps
// of an FP32 operand), lo = xf - hi (exact in FP32; the tensor core keeps its top 11 bits), so hi + lo carries 22 bits.
__device__ __forceinline__ void fr_convert4(const double (&x)[4], const float (&sc)[4], float* hi, float* lo, int off) {
  float h[4], l[4];
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const float xf = (float)x[j] * sc[j];
    h[j] = __uint_as_float(__float_as_uint(xf) & 0xFFFFE000u);
    l[j] = xf - h[j];
  }
  *reinterpret_cast<float4*>(hi + off) = make_float4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<float4*>(lo + off) = make_float4(l[0], l[1], l[2], l[3]);
}

// tiles (m0, n0) of the lower block triangle: m0 = 256 mt, n0 = 256 nt <= m0
__host__ __device__ inline int fast_reduce_tiles(int Mp) {
  const int mts = (Mp + FR_M - 1) / FR_M;
  return mts * (mts + 1) / 2;
}

__global__ void __launch_bounds__(FR_THREADS, 1) gp_reduce_fast_kernel(const BwdParams p) {
  extern __shared__ __align__(1024) unsigned char fr_smem_raw[];
  float* stage_base = reinterpret_cast<float*>(fr_smem_raw);                                   // TF32 operand ring
  float* sc_tab = stage_base + FR_STAGES * FR_STAGE_FLOATS;                                    // [FR_MAX_RANGE] per-point scales
  __shared__ uint64_t full_bar[FR_STAGES], empty_bar[FR_STAGES], done_bar;
  __shared__ uint32_t tmem_base_s;
  const iwvi_gp_desc& d = p.d;
  const AuxLayout al = iwvi_aux_layout(d.M, d.D, d.R);
  const int NB = al.NB, R = d.R, Mp = al.Mp;
  const BwdWs& wl = p.wl;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  // ---- decode the work item: tile fastest, then matrix, point range slowest (all resident CTAs stream the same points)
  const int n_tiles = fast_reduce_tiles(Mp);
  int item = blockIdx.x;
  int tile = item % n_tiles; item /= n_tiles;
  const int q = p.q_lo + item % p.q_n;
  const int s = item / p.q_n;
  int mt = 0;
  while ((mt + 1) * (mt + 2) / 2 <= tile) mt++;
  const int m0 = mt * FR_M, n0 = (tile - mt * (mt + 1) / 2) * FR_N;
  const int nrows = min(FR_M, Mp - m0);             // real A rows of this tile: 128 or 256
  const int ncols = min(FR_N, Mp - n0);             // real B rows of this tile: 128 or 256
  const int halves = nrows / 128;
  const int c0 = s * wl.chunks_per_split;
  const int c1 = min(c0 + wl.chunks_per_split, wl.Tp / IWVI_BLK);
  const int nkb = (c1 - c0) * (IWVI_BLK / FR_KB);   // stages of this CTA
  const bool is_lm = (q == R);

  const SaveLayout sv = iwvi_save_layout(d.T, d.M, R);
  const double* A_T = p.save + sv.off_a;
  const double* P_op = is_lm ? p.ws + wl.off_bbar : A_T;                              // rows of the output
  const double* Q_op = is_lm ? A_T : p.save + sv.off_u + (size_t)q * sv.u_stride;   // columns of the output
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

  // per-point scales of the B operand for this CTA's range (2 gvar_bar_q, or -2 for dLm), once
  for (int i = tid; i < (c1 - c0) * IWVI_BLK; i += FR_THREADS)
    sc_tab[i] = is_lm ? -2.f : 2.f * (float)__ldg(gvb + ((size_t)c0 * IWVI_BLK + i) * IWVI_MAX_R + q);
  __syncthreads();

  if (warp < FR_CONV_WARPS) {
    // ================= converters =================
    // unit of work: 32 columns (lanes) x 4 points -> one 16-byte chunk per lane of the hi and of the lo operand copy.
    // A tile: (nrows / 32) x 4 units, B tile: (ncols / 32) x 4; dealt round-robin to the 16 warps (4 each at 256 x 256).
    const int a_units = (nrows / 32) * (FR_KB / 4), n_units = a_units + (ncols / 32) * (FR_KB / 4);
    constexpr int MAXU = ((FR_M + FR_N) / 32) * (FR_KB / 4) / FR_CONV_WARPS;       // 4
    auto load_stage = [&](int kb, double (&x)[MAXU][4]) {
      const int chunk = c0 + kb / (IWVI_BLK / FR_KB);
      const int pr0 = (kb % (IWVI_BLK / FR_KB)) * FR_KB;                 // first point of the stage inside its chunk
#pragma unroll
      for (int i = 0; i < MAXU; i++) {
        const int u = warp + i * FR_CONV_WARPS;
        if (u < n_units) {
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
    };
    auto convert_stage = [&](int kb, const double (&x)[MAXU][4]) {
      const int st = kb % FR_STAGES;
      if (kb >= FR_STAGES) mbar_wait(&empty_bar[st], ((kb / FR_STAGES) & 1u) ^ 1u);       // operand stage is free
      float* sb = stage_base + (size_t)st * FR_STAGE_FLOATS;
      float *a_hi = sb, *a_lo = sb + FR_A_FLOATS, *b_hi = sb + 2 * FR_A_FLOATS, *b_lo = b_hi + FR_B_FLOATS;
      const float* sct = sc_tab + kb * FR_KB;
      const float one[4] = {1.f, 1.f, 1.f, 1.f};
#pragma unroll
      for (int i = 0; i < MAXU; i++) {
        const int u = warp + i * FR_CONV_WARPS;
        if (u < n_units) {
          const bool isA = u < a_units;
          const int uu = isA ? u : u - a_units;
          const int cg = uu / (FR_KB / 4), pg = uu % (FR_KB / 4);
          const int lc = cg * 32 + lane;                                   // row inside the operand tile
          const int off = ((lc >> 3) * (FR_KB / 4) + pg) * 32 + (lc & 7) * 4;
          if (isA) {
            fr_convert4(x[i], one, a_hi, a_lo, off);
          } else {
            const float4 s4 = *reinterpret_cast<const float4*>(sct + pg * 4);
            const float sc[4] = {s4.x, s4.y, s4.z, s4.w};
            fr_convert4(x[i], sc, b_hi, b_lo, off);
          }
        }
      }
      // this thread's generic-proxy writes -> visible to the async proxy (tensor core); one arrival per warp
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&full_bar[st]);
    };
    double xa[MAXU][4], xb[MAXU][4];
    if (nkb > 0) load_stage(0, xa);
    for (int kb = 0; kb < nkb; kb += 2) {
      if (kb + 1 < nkb) load_stage(kb + 1, xb);
      convert_stage(kb, xa);
      if (kb + 1 < nkb) {
        if (kb + 2 < nkb) load_stage(kb + 2, xa);
        convert_stage(kb + 1, xb);
      }
    }
  } else {
   if (lane == 0) {
    // ================= MMA issuer (one thread) =================
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(ncols >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t lbo = 128, sbo = (FR_KB / 4) * 128;
    const uint32_t half_bytes = 128 * FR_KB * 4;       // operand rows 128 .. 255 of the A copies
    for (int kb = 0; kb < nkb; kb++) {
      const int st = kb % FR_STAGES;
      mbar_wait(&full_bar[st], (kb / FR_STAGES) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t sa = smem_u32(stage_base + (size_t)st * FR_STAGE_FLOATS);
      const uint32_t a_hi = sa, a_lo = sa + FR_A_FLOATS * 4, b_hi = sa + 2 * FR_A_FLOATS * 4, b_lo = b_hi + FR_B_FLOATS * 4;
#pragma unroll
      for (int k0 = 0; k0 < FR_KB; k0 += 8) {
        const uint32_t ko = (k0 / 4) * 128;
        const uint32_t acc = (kb > 0 || k0 > 0) ? 1u : 0u;
        for (int h = 0; h < halves; h++) {
          const uint32_t td = tmem_base + (uint32_t)h * FR_N, ao = (uint32_t)h * half_bytes + ko;
          fr_mma(td, fr_desc(a_hi + ao, lbo, sbo), fr_desc(b_hi + ko, lbo, sbo), idesc, acc);
          fr_mma(td, fr_desc(a_hi + ao, lbo, sbo), fr_desc(b_lo + ko, lbo, sbo), idesc, 1u);
          fr_mma(td, fr_desc(a_lo + ao, lbo, sbo), fr_desc(b_hi + ko, lbo, sbo), idesc, 1u);
        }
      }
      fr_commit(&empty_bar[st]);          // the stage may be overwritten once these MMAs have read it
    }
    fr_commit(&done_bar);
   }
   __syncwarp();    // the issuing warp reconverges before the block-wide barrier below
  }

  // ================= epilogue: warps 0-3 own TMEM lanes 32 w .. 32 w + 31 = output rows m0 + 128 h + 32 w + lane =================
  if (warp < 4) {
    mbar_wait(&done_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    double* red = p.ws + wl.off_red;
    for (int h = 0; h < halves; h++) {
      const int m = m0 + h * 128 + warp * 32 + lane, bi = m >> 6, rb = m & 63;
      for (int cb = 0; cb < ncols; cb += 8) {
        uint32_t r[8];
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(h * FR_N + cb);
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const int n = n0 + cb, bj = n >> 6;
        if (bj <= bi) {                                 // lower block triangle only: the finalize kernel reads nothing else
          double* out = red + (((size_t)q * wl.S + s) * wl.npairs + iwvi_pair(bi, bj)) * IWVI_BLK * IWVI_BLK +
                        (size_t)rb * IWVI_BLK + (n & 63);
#pragma unroll
          for (int j = 0; j < 8; j++) out[j] = (double)__uint_as_float(r[j]);
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == FR_CONV_WARPS)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
}

