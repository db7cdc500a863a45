// common.cuh -- device building blocks shared by the sm_100a IW-ELBO kernels.
//
// * DMMA: on sm_100a every f64 mma.sync shape lowers to DMMA.8x8x4 (checked with cuobjdump), so m8n8k4 is used
//   directly; tcgen05/TMEM has no f64 kind.  Measured pipe peak on B200: 37.05 TFLOP/s (profiles/FP64_PEAK_r01.md).
// * staging: 1-D bulk TMA (cp.async.bulk, SASS UBLKCP) global->shared with mbarrier transaction counting.
// * shared-memory operand layouts use leading dimensions == 4 (mod 16) doubles, which makes every DMMA fragment
//   load (8 rows x 4 k, one double per lane) conflict-free per half-warp.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include "../../include/iwvi_b200.h"

#define IWVI_BLK 64            // block size of the blocked triangular algorithms
#define IWVI_LDS 68            // leading dimension of a staged 64x64 block in shared memory (== 4 mod 16)
#define IWVI_STAGE_DOUBLES (IWVI_BLK * IWVI_LDS)
#define IWVI_PACK_SMALL 8      // CTAs of the prologue pack kernel that prepare Z / ls, |z|^2, q_mu, constants
#define IWVI_PACK_GRID (IWVI_PACK_SMALL + IWVI_MAX_R * 36)   // + one CTA per lower block of every tril(q_sqrt_r): KL partial slots

__host__ __device__ inline int iwvi_round_up(int x, int m) { return (x + m - 1) / m * m; }
// leading dimension of length-scaled inputs (Zt rows in aux, x tile in smem): == 4 (mod 16), >= round_up(D,4)
__host__ __device__ inline int iwvi_ldz(int D) { return D <= 20 ? 20 : 36; }

// ---------------------------------------------------------------------------------------------
// aux layout (doubles), written by the prologue, read by the row kernels
// ---------------------------------------------------------------------------------------------
struct AuxLayout {
  int Mp, NB, ldz, R, npairs;
  int64_t off_lmb, off_lqb, off_zt, off_zn, off_qmu, off_consts, off_scratch, off_prog, total;
};
// index of the lower block (i, j), i >= j, in the block-major arrays
__host__ __device__ inline int iwvi_pair(int i, int j) { return i * (i + 1) / 2 + j; }
__host__ __device__ inline AuxLayout iwvi_aux_layout(int M, int D, int R) {
  AuxLayout a;
  a.Mp = iwvi_round_up(M, IWVI_BLK);
  a.NB = a.Mp / IWVI_BLK;
  a.ldz = iwvi_ldz(D);
  a.R = R;
  a.npairs = a.NB * (a.NB + 1) / 2;
  int64_t o = 0;
  // block-major, padded [64][68] copies (one bulk-TMA transaction each): strictly-lower blocks of Lm, with the
  // INVERTED diagonal blocks of Lm in the diagonal slots; lower blocks of tril(q_sqrt_r)
  a.off_lmb = o;    o += (int64_t)a.npairs * IWVI_STAGE_DOUBLES;
  a.off_lqb = o;    o += (int64_t)R * a.npairs * IWVI_STAGE_DOUBLES;
  a.off_zt = o;     o += (int64_t)a.Mp * a.ldz;                 // Z / ls, zero padded
  a.off_zn = o;     o += a.Mp;                                  // |Z/ls|^2
  a.off_qmu = o;    o += (int64_t)a.Mp * IWVI_MAX_R;            // q_mu padded to [Mp, 8]
  a.off_consts = o; o += 64;                                    // [0]=variance, 1/ls[d] at [8+d]
  a.off_scratch = o; o += IWVI_PACK_GRID;                       // per-CTA partial sums of the KL (fixed-order final sum)
  a.off_prog = o;    o += 8;                                    // 16 ints: per block-row progress counters of the Cholesky
  a.total = o;
  return a;
}
#define IWVI_C_VARIANCE 0
#define IWVI_C_INVLS 8   // consts[8 + d] = 1 / ls[d], d < 32

// save layout (doubles): A and U_r are kept block-major, [chunk of 64 points][m-block][64 points][68] (so that every
// 64x64 operand block of the backward contractions is ONE contiguous bulk-TMA transaction), then gvar [T, R] and
// gmean [T, R].  Tp = T rounded up to 128; points T..Tp-1 are written as zeros by the forward kernel.
struct SaveLayout { int64_t off_a, off_u, off_gvar, off_gmean, total, u_stride; int NB; int Tp; };
__host__ __device__ inline SaveLayout iwvi_save_layout(int T, int M, int R) {
  SaveLayout s;
  s.NB = iwvi_round_up(M, IWVI_BLK) / IWVI_BLK;
  s.Tp = iwvi_round_up(T, 128);
  s.u_stride = (int64_t)(s.Tp / IWVI_BLK) * s.NB * IWVI_STAGE_DOUBLES;
  int64_t o = 0;
  s.off_a = o;     o += s.u_stride;
  s.off_u = o;     o += (int64_t)R * s.u_stride;
  s.off_gvar = o;  o += (int64_t)T * R;
  s.off_gmean = o; o += (int64_t)T * R;
  s.total = o;
  return s;
}
// offset of element (point n, inducing index m) inside a block-major [Tp x Mp] array
__host__ __device__ inline int64_t iwvi_blk_off(int n, int m, int NB) {
  return ((int64_t)((n >> 6) * NB + (m >> 6)) * IWVI_BLK + (n & 63)) * IWVI_LDS + (m & 63);
}

// ---------------------------------------------------------------------------------------------
// DMMA m8n8k4 and the warp-level tile product
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

// acc[TM][TN][2] += A(m,k) * B(k,n),  n in [0, 8*TN), k in [0, KC), KC % 4 == 0; the warp's TM row tiles of 8 rows are
// MS tiles apart: tile i covers rows 8*MS*i .. 8*MS*i + 7 relative to As (MS == 1: TM*8 consecutive rows).
//   ALAY == 0: A(m,k) = As[m*lda + k]      ALAY == 1: A(m,k) = As[k*lda + m]
//   BLAY == 0: B(k,n) = Bs[n*ldb + k]      BLAY == 1: B(k,n) = Bs[k*ldb + n]
// accumulator element acc[i][j][c] is C(8*MS*i + lane/4, 8*j + 2*(lane%4) + c).
template <int TM, int TN, int ALAY, int BLAY, int MS = 1>
__device__ __forceinline__ void warp_gemm(double (&acc)[TM][TN][2], const double* __restrict__ As, int lda,
                                          const double* __restrict__ Bs, int ldb, int KC, int lane) {
  const int g = lane >> 2, t = lane & 3;
  const double* ap = ALAY == 0 ? As + g * lda + t : As + t * lda + g;
  const double* bp = BLAY == 0 ? Bs + g * ldb + t : Bs + t * ldb + g;
#pragma unroll 2
  for (int k0 = 0; k0 < KC; k0 += 4) {
    double a[TM], b[TN];
#pragma unroll
    for (int i = 0; i < TM; i++) a[i] = ALAY == 0 ? ap[i * 8 * MS * lda + k0] : ap[k0 * lda + i * 8 * MS];
#pragma unroll
    for (int j = 0; j < TN; j++) b[j] = BLAY == 0 ? bp[j * 8 * ldb + k0] : bp[k0 * ldb + j * 8];
#pragma unroll
    for (int i = 0; i < TM; i++)
#pragma unroll
      for (int j = 0; j < TN; j++) dmma884(acc[i][j], a[i], b[j]);
  }
}

// warp_gemm over a full 64-deep block with the operand fragments of k-step s + 1 fetched while the DMMAs of k-step s
// issue (two register buffers).  In warp_gemm every iteration starts with its 12 fragment loads and then waits out the
// shared-memory latency before its first DMMA; the two warps of an SM sub-partition leave a ring wait together, run
// those phases in lock step, and the FP64 pipe idles once per iteration.  The warp barrier is a scheduling fence: it
// keeps ptxas from sinking the prefetch loads back next to their consumers (needs the consumer register budget of
// reg_alloc below).
template <int TM, int TN, int ALAY, int BLAY, int MS = 1>
__device__ __forceinline__ void warp_gemm_pf(double (&acc)[TM][TN][2], const double* __restrict__ As, int lda,
                                             const double* __restrict__ Bs, int ldb, int lane) {
  const int g = lane >> 2, t = lane & 3;
  const double* ap = ALAY == 0 ? As + g * lda + t : As + t * lda + g;
  const double* bp = BLAY == 0 ? Bs + g * ldb + t : Bs + t * ldb + g;
  double a0[TM], b0[TN], a1[TM], b1[TN];
  auto fetch = [&](double (&a)[TM], double (&b)[TN], int k0) {
#pragma unroll
    for (int i = 0; i < TM; i++) a[i] = ALAY == 0 ? ap[i * 8 * MS * lda + k0] : ap[k0 * lda + i * 8 * MS];
#pragma unroll
    for (int j = 0; j < TN; j++) b[j] = BLAY == 0 ? bp[j * 8 * ldb + k0] : bp[k0 * ldb + j * 8];
  };
  auto mma = [&](const double (&a)[TM], const double (&b)[TN]) {
#pragma unroll
    for (int i = 0; i < TM; i++)
#pragma unroll
      for (int j = 0; j < TN; j++) dmma884(acc[i][j], a[i], b[j]);
  };
  fetch(a0, b0, 0);
#pragma unroll 2
  for (int k0 = 0; k0 < IWVI_BLK; k0 += 8) {
    __syncwarp();
    fetch(a1, b1, k0 + 4);
    mma(a0, b0);
    __syncwarp();
    if (k0 + 8 < IWVI_BLK) fetch(a0, b0, k0 + 8);
    mma(a1, b1);
  }
}

// The same product for a TRIANGULAR 64x64 A block whose structural zeros are skipped at the granularity of the 8x8x4
// DMMA: Ablk is the block's (0,0) corner, the warp's row tiles are mt0, mt0 + MS, ... (units of 8 rows).  The k range
// is cut into segments inside which the set of active row tiles is a compile-time constant, so no DMMA is ever issued
// under a predicate (a predicated-off DMMA still occupies the FP64 pipe).
//   LOWER == 1: A(m,k) != 0 only for k <= m  -> row tile ti needs k in [0, 8*(ti+1))
//   LOWER == 0: A(m,k) != 0 only for k >= m  -> row tile ti needs k in [8*ti, 64)
template <int TM, int TN, int ALAY, int BLAY, int MS, int LOWER>
__device__ __forceinline__ void warp_gemm_tri(double (&acc)[TM][TN][2], const double* __restrict__ Ablk, int lda,
                                              const double* __restrict__ Bs, int ldb, int mt0, int lane) {
  const int g = lane >> 2, t = lane & 3;
  const double* ap = ALAY == 0 ? Ablk + (mt0 * 8 + g) * lda + t : Ablk + t * lda + mt0 * 8 + g;
  const double* bp = BLAY == 0 ? Bs + g * ldb + t : Bs + t * ldb + g;
#pragma unroll
  for (int s = 0; s < TM; s++) {
    const int k_lo = LOWER ? (s == 0 ? 0 : 8 * (mt0 + MS * (s - 1) + 1)) : 8 * (mt0 + MS * s);
    const int k_hi = LOWER ? 8 * (mt0 + MS * s + 1) : (s == TM - 1 ? IWVI_BLK : 8 * (mt0 + MS * (s + 1)));
#pragma unroll 2
    for (int k0 = k_lo; k0 < k_hi; k0 += 4) {
      double a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; i++)
        if (LOWER ? i >= s : i <= s) a[i] = ALAY == 0 ? ap[i * 8 * MS * lda + k0] : ap[k0 * lda + i * 8 * MS];
#pragma unroll
      for (int j = 0; j < TN; j++) b[j] = BLAY == 0 ? bp[j * 8 * ldb + k0] : bp[k0 * ldb + j * 8];
#pragma unroll
      for (int i = 0; i < TM; i++)
        if (LOWER ? i >= s : i <= s) {
#pragma unroll
          for (int j = 0; j < TN; j++) dmma884(acc[i][j], a[i], b[j]);
        }
    }
  }
}

template <int TM, int TN>
__device__ __forceinline__ void acc_zero(double (&acc)[TM][TN][2]) {
#pragma unroll
  for (int i = 0; i < TM; i++)
#pragma unroll
    for (int j = 0; j < TN; j++) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
}

// ---------------------------------------------------------------------------------------------
// mbarrier + bulk TMA
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  const uint32_t addr = smem_u32(bar);
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(addr), "r"(parity) : "memory");
  }
}
// 1-D bulk copy global -> shared, completion signalled on `bar` (bytes % 16 == 0, both addresses 16B aligned)
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// 1-D bulk copy shared -> global (asynchronous TMA store, tracked by bulk async-groups of the issuing thread)
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
               ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// the issuing thread's bulk stores have finished READING shared memory (their source may be overwritten)
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// make this thread's generic-proxy shared-memory writes visible to the async proxy (before a barrier + bulk store)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// A 2-deep (NST) ring of 64-row blocks filled by warp 0 with bulk TMA; arrival through mbarriers,
// release through the block-wide barrier that the algorithms need anyway between dependent steps.
#define IWVI_NST 2
struct BlockSrc { const double* src; uint32_t bytes; };   // one contiguous block, copied verbatim into a stage

template <int NST>
struct StagePipeT {
  uint64_t* bars;   // [NST]
  double* stages;   // [NST][IWVI_STAGE_DOUBLES]
  uint32_t it_c;    // blocks consumed so far (uniform across the CTA)
  uint32_t it_p;    // blocks produced so far (meaningful in warp 0)

  __device__ __forceinline__ void setup(uint64_t* b, double* s) {
    bars = b; stages = s; it_c = 0; it_p = 0;
    if (threadIdx.x == 0) {
      for (int i = 0; i < NST; i++) mbar_init(&bars[i], 1);
      mbar_fence_init();
    }
    __syncthreads();
  }
  // warp 0 only: issue one block into the next stage
  __device__ __forceinline__ void produce(const BlockSrc& b, int lane) {
    const uint32_t s = it_p % NST;
    if (lane == 0) {
      mbar_arrive_expect_tx(&bars[s], b.bytes);
      bulk_g2s(stages + (size_t)s * IWVI_STAGE_DOUBLES, b.src, b.bytes, &bars[s]);
    }
    it_p++;
  }
  // all threads: wait for the oldest outstanding block, return its stage
  __device__ __forceinline__ const double* wait() {
    const uint32_t s = it_c % NST;
    mbar_wait(&bars[s], (it_c / NST) & 1u);
    return stages + (size_t)s * IWVI_STAGE_DOUBLES;
  }
  // all threads: the block returned by the last wait() is no longer needed.  Contains a __syncthreads().
  // `seq` is the producer's block sequence (only warp 0's copy is used/advanced).
  template <class Seq>
  __device__ __forceinline__ void release(Seq& seq, int warp, int lane) {
    __syncthreads();
    it_c++;
    if (warp == 0 && !seq.done()) {
      produce(seq.get(), lane);
      seq.advance();
    }
  }
  // all threads: wait for the block `k` positions after the oldest outstanding one (k < NST)
  __device__ __forceinline__ const double* wait_ahead(uint32_t k) {
    const uint32_t it = it_c + k;
    const uint32_t s = it % NST;
    mbar_wait(&bars[s], (it / NST) & 1u);
    return stages + (size_t)s * IWVI_STAGE_DOUBLES;
  }
  // all threads: the n oldest blocks are no longer needed (one __syncthreads())
  template <class Seq>
  __device__ __forceinline__ void release_n(Seq& seq, int n, int warp, int lane) {
    __syncthreads();
    it_c += n;
    if (warp == 0) {
      for (int i = 0; i < n && !seq.done(); i++) {
        produce(seq.get(), lane);
        seq.advance();
      }
    }
  }
  // all threads, at the start of a tile (after a __syncthreads): prime the ring from a fresh sequence
  template <class Seq>
  __device__ __forceinline__ void prime(Seq& seq, int warp, int lane) {
    if (warp == 0) {
      for (int i = 0; i < NST && !seq.done(); i++) {
        produce(seq.get(), lane);
        seq.advance();
      }
    }
  }
};
typedef StagePipeT<IWVI_NST> StagePipe;

// ---------------------------------------------------------------------------------------------
// Warp-specialised ring: a dedicated producer warp streams 64-row blocks through an NST-deep ring with bulk TMA;
// consumer warps wait on full[s] (transaction count) and hand a stage back by arriving on empty[s] (one arrival per
// consumer warp).  No block-wide barrier is involved, so consumers only synchronise where data really flows
// between warps (named barriers, below), and the producer's issue cost is off the consumers' critical path.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Register reallocation between the roles of a warp-specialised CTA (setmaxnreg, sm_90+).  With a ninth warp one SM
// sub-partition hosts three warps and ptxas must fit 3 x 32 x regs into its 16 K registers: 168 per thread for EVERY
// warp.  Launching three full warpgroups (384 threads: 8 consumer warps, the producer warp and 3 idle warps that exit)
// lets the producer warpgroup hand its registers back (dec) and the two consumer warpgroups grow (inc) -- the
// instruction is warpgroup-aligned, which is why the producer side is padded to four warps.
template <int N> __device__ __forceinline__ void reg_dealloc() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void reg_alloc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
#define IWVI_WS_THREADS 384         // 2 consumer warpgroups + 1 producer warpgroup (one active warp)
#define IWVI_PRODUCER_REGS 24
#define IWVI_CONSUMER_REGS 240      // 168 + (168 - 24) * 128 / 256

// Host-side guard: setmaxnreg.inc waits until the CTA's register pool (registers per thread at launch x threads) can
// serve the request, so a build whose launch allocation could not cover consumers + producers would hang, not fail.
template <class Kernel>
inline bool iwvi_ws_pool_ok(Kernel kernel) {
  cudaFuncAttributes fa;
  if (cudaFuncGetAttributes(&fa, kernel) != cudaSuccess) return false;
  return (long)fa.numRegs * IWVI_WS_THREADS >= (long)IWVI_CONSUMER_REGS * 256 + (long)IWVI_PRODUCER_REGS * 128;
}

template <int NST>
struct RingT {
  uint64_t* full;    // [NST]
  uint64_t* empty;   // [NST]
  double* stages;    // [NST][IWVI_STAGE_DOUBLES]
  uint32_t it;       // blocks consumed (consumer copy) or produced (producer copy) so far

  // all threads of the CTA, once, before the roles split
  __device__ __forceinline__ void setup(uint64_t* bars, double* s, int n_consumer_warps) {
    full = bars; empty = bars + NST; stages = s; it = 0;
    if (threadIdx.x == 0) {
      for (int i = 0; i < NST; i++) { mbar_init(&full[i], 1); mbar_init(&empty[i], n_consumer_warps); }
      mbar_fence_init();
    }
    __syncthreads();
  }
  // producer warp: wait until the stage is free, then issue the block (one bulk-TMA transaction)
  __device__ __forceinline__ void produce(const BlockSrc& b, int lane) {
    const uint32_t s = it % NST;
    if (lane == 0) {
      mbar_wait(&empty[s], ((it / NST) & 1u) ^ 1u);
      mbar_arrive_expect_tx(&full[s], b.bytes);
      bulk_g2s(stages + (size_t)s * IWVI_STAGE_DOUBLES, b.src, b.bytes, &full[s]);
    }
    it++;
  }
  // producer warp: the block plus a small second transaction (`xbytes` from `xsrc` to `xdst`) signalled on the same barrier
  __device__ __forceinline__ void produce2(const BlockSrc& b, void* xdst, const void* xsrc, uint32_t xbytes, int lane) {
    const uint32_t s = it % NST;
    if (lane == 0) {
      mbar_wait(&empty[s], ((it / NST) & 1u) ^ 1u);
      mbar_arrive_expect_tx(&full[s], b.bytes + xbytes);
      bulk_g2s(stages + (size_t)s * IWVI_STAGE_DOUBLES, b.src, b.bytes, &full[s]);
      if (xbytes) bulk_g2s(xdst, xsrc, xbytes, &full[s]);
    }
    it++;
  }
  // consumer warps: oldest outstanding block (offset k blocks ahead, k < NST)
  __device__ __forceinline__ const double* wait(uint32_t k = 0) {
    const uint32_t i2 = it + k, s = i2 % NST;
    mbar_wait(&full[s], (i2 / NST) & 1u);
    return stages + (size_t)s * IWVI_STAGE_DOUBLES;
  }
  // consumer warps: this warp is done with the n oldest blocks
  __device__ __forceinline__ void release(int lane, int n = 1) {
    __syncwarp();
    if (lane == 0)
      for (int k = 0; k < n; k++) mbar_arrive(&empty[(it + k) % NST]);
    it += n;
  }
};

// ---------------------------------------------------------------------------------------------
// stationary kernels (GPflow 1.x formulas): K(r2) and dK/dr2
// ---------------------------------------------------------------------------------------------
// Branch-free exp(x) for the kernel functions (x <= ~700; in practice x <= 0).  libdevice's exp() carries a rarely
// taken slow-path branch, which splits the sixteen independent evaluations each thread makes per 64x64 gram block into
// separate basic blocks that the scheduler cannot interleave: measured ~450 cycles per element.  Straight-line code
// overlaps them.  Method: x = k ln2 + r, |r| <= ln2/2 (Cody-Waite, two-part ln2); exp(r) = (exp(r/8))^8 with a
// degree-8 Taylor polynomial on |r/8| <= 0.0434 (truncation 3e-17 relative) and three squarings; scale by 2^k through
// the exponent field.  Relative error a few ulp (<= ~2e-15 against numpy over [-700, 1]); the Lm / Kuf stage parity
// tests (rtol 1e-8) exercise it for every kernel family.
// Arguments below -708 are clamped (result 3.3e-308 instead of a denormal or 0); NaN propagates.
__device__ __forceinline__ double exp_bf(double x) {
  x = (x < -708.0) ? -708.0 : x;
  const double MAGIC = 6755399441055744.0;   // 1.5 * 2^52: adding it rounds to the nearest integer, kept in the low word
  const double t = fma(x, 1.4426950408889634074, MAGIC);
  const int k = __double2loint(t);
  const double kd = t - MAGIC;
  double r = fma(kd, -6.93147180369123816490e-01, x);
  r = fma(kd, -1.90821492927058770002e-10, r);
  const double s = 0.125 * r;
  double p = 2.48015873015873015873e-05;             // 1/8!
  p = fma(p, s, 1.98412698412698412698e-04);         // 1/7!
  p = fma(p, s, 1.38888888888888888889e-03);         // 1/6!
  p = fma(p, s, 8.33333333333333333333e-03);         // 1/5!
  p = fma(p, s, 4.16666666666666666667e-02);         // 1/4!
  p = fma(p, s, 1.66666666666666666667e-01);         // 1/3!
  p = fma(p, s, 0.5);
  p = fma(p, s, 1.0);
  double e = fma(p, s, 1.0);                         // exp(r/8)
  e *= e; e *= e; e *= e;                            // exp(r)
  return e * __hiloint2double((k + 1023) << 20, 0);  // * 2^k  (k >= -1022 thanks to the clamp)
}

// N independent evaluations in lock step (x -> exp(x) in place).  Written statement by statement over the N values
// because ptxas otherwise emits the N serial dependency chains one after the other (observed in the SASS), which
// leaves the FP64 pipe idle for most of each chain's latency.
template <int N>
__device__ __forceinline__ void exp_bf_n(double (&x)[N]) {
  const double MAGIC = 6755399441055744.0;
  double t[N], r[N], p[N];
#pragma unroll
  for (int i = 0; i < N; i++) x[i] = (x[i] < -708.0) ? -708.0 : x[i];
#pragma unroll
  for (int i = 0; i < N; i++) t[i] = fma(x[i], 1.4426950408889634074, MAGIC);
#pragma unroll
  for (int i = 0; i < N; i++) r[i] = t[i] - MAGIC;
#pragma unroll
  for (int i = 0; i < N; i++) x[i] = fma(r[i], -6.93147180369123816490e-01, x[i]);
#pragma unroll
  for (int i = 0; i < N; i++) r[i] = 0.125 * fma(r[i], -1.90821492927058770002e-10, x[i]);
#pragma unroll
  for (int i = 0; i < N; i++) p[i] = fma(2.48015873015873015873e-05, r[i], 1.98412698412698412698e-04);
#pragma unroll
  for (int i = 0; i < N; i++) p[i] = fma(p[i], r[i], 1.38888888888888888889e-03);
#pragma unroll
  for (int i = 0; i < N; i++) p[i] = fma(p[i], r[i], 8.33333333333333333333e-03);
#pragma unroll
  for (int i = 0; i < N; i++) p[i] = fma(p[i], r[i], 4.16666666666666666667e-02);
#pragma unroll
  for (int i = 0; i < N; i++) p[i] = fma(p[i], r[i], 1.66666666666666666667e-01);
#pragma unroll
  for (int i = 0; i < N; i++) p[i] = fma(p[i], r[i], 0.5);
#pragma unroll
  for (int i = 0; i < N; i++) p[i] = fma(p[i], r[i], 1.0);
#pragma unroll
  for (int i = 0; i < N; i++) p[i] = fma(p[i], r[i], 1.0);
#pragma unroll
  for (int i = 0; i < N; i++) p[i] *= p[i];
#pragma unroll
  for (int i = 0; i < N; i++) p[i] *= p[i];
#pragma unroll
  for (int i = 0; i < N; i++) p[i] *= p[i];
#pragma unroll
  for (int i = 0; i < N; i++) x[i] = p[i] * __hiloint2double((__double2loint(t[i]) + 1023) << 20, 0);
}

// K(r2) for N squared distances at once (in place: r2 -> K); optionally dK/dr2 as well
template <int KIND, int N, bool WITH_DK>
__device__ __forceinline__ void kern_n(double (&v)[N], double (&dk)[N], double variance) {
  if (KIND == IWVI_KERN_RBF) {
#pragma unroll
    for (int i = 0; i < N; i++) v[i] *= -0.5;
    exp_bf_n<N>(v);
#pragma unroll
    for (int i = 0; i < N; i++) { v[i] *= variance; if (WITH_DK) dk[i] = -0.5 * v[i]; }
    return;
  }
  const double c = KIND == IWVI_KERN_MATERN52 ? 2.23606797749978969641 : KIND == IWVI_KERN_MATERN32 ? 1.73205080756887729353 : 1.0;
  double r[N], e[N];
  bool clamped[N];
#pragma unroll
  for (int i = 0; i < N; i++) { clamped[i] = v[i] < 1e-40; r[i] = sqrt(fmax(v[i], 1e-40)); e[i] = -c * r[i]; }
  exp_bf_n<N>(e);
#pragma unroll
  for (int i = 0; i < N; i++) {
    if (KIND == IWVI_KERN_MATERN52) {
      v[i] = variance * (1.0 + c * r[i] + (5.0 / 3.0) * r[i] * r[i]) * e[i];
      if (WITH_DK) dk[i] = -(5.0 / 6.0) * variance * (1.0 + c * r[i]) * e[i];
    } else if (KIND == IWVI_KERN_MATERN32) {
      v[i] = variance * (1.0 + c * r[i]) * e[i];
      if (WITH_DK) dk[i] = -1.5 * variance * e[i];
    } else {
      v[i] = variance * e[i];
      if (WITH_DK) dk[i] = -variance * e[i] / (2.0 * r[i]);
    }
    if (WITH_DK && clamped[i]) dk[i] = 0.0;
  }
}

__device__ __forceinline__ double kern_k(int kind, double r2, double variance) {
  if (kind == IWVI_KERN_RBF) return variance * exp_bf(-0.5 * r2);
  const double r = sqrt(fmax(r2, 1e-40));
  if (kind == IWVI_KERN_MATERN52) {
    const double s5 = 2.23606797749978969641;
    return variance * (1.0 + s5 * r + (5.0 / 3.0) * r * r) * exp_bf(-s5 * r);
  }
  if (kind == IWVI_KERN_MATERN32) {
    const double s3 = 1.73205080756887729353;
    return variance * (1.0 + s3 * r) * exp_bf(-s3 * r);
  }
  return variance * exp_bf(-r);
}
__device__ __forceinline__ void kern_k_dk(int kind, double r2, double variance, double& K, double& dK) {
  if (kind == IWVI_KERN_RBF) { K = variance * exp_bf(-0.5 * r2); dK = -0.5 * K; return; }
  const bool clamped = r2 < 1e-40;
  const double r = sqrt(fmax(r2, 1e-40));
  if (kind == IWVI_KERN_MATERN52) {
    const double s5 = 2.23606797749978969641;
    const double e = exp_bf(-s5 * r);
    K = variance * (1.0 + s5 * r + (5.0 / 3.0) * r * r) * e;
    dK = -(5.0 / 6.0) * variance * (1.0 + s5 * r) * e;
  } else if (kind == IWVI_KERN_MATERN32) {
    const double s3 = 1.73205080756887729353;
    const double e = exp_bf(-s3 * r);
    K = variance * (1.0 + s3 * r) * e;
    dK = -1.5 * variance * e;
  } else {
    const double e = exp_bf(-r);
    K = variance * e;
    dK = -variance * e / (2.0 * r);
  }
  if (clamped) dK = 0.0;
}

// ---------------------------------------------------------------------------------------------
// reductions
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// block-wide sum, result valid in thread 0; `red` is >= 32 doubles of shared memory
__device__ __forceinline__ double block_sum(double v, double* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double s = 0.0;
  if (threadIdx.x == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    for (int i = 0; i < nw; i++) s += red[i];
  }
  return s;
}

// Fire-and-forget add to global memory (no return value, so the issuing thread does not wait for L2).  Used only where
// ONE thread owns the address for the whole kernel, so the additions happen in program order: still deterministic.
__device__ __forceinline__ void red_add(double* p, double v) {
  asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}

// inter-CTA hand-off through global memory (producer: data stores, __syncthreads, thread 0: st_release; consumer:
// thread 0 spins on ld_acquire, __syncthreads, then everybody reads the data with L2-coherent loads)
__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

int iwvi_check_gp_desc(const iwvi_gp_desc* d);

// Optional phase timing (tools/phase_timing.py builds a second library with -DIWVI_PHASE_TIMING): thread 0 of every CTA
// accumulates clock64() deltas per phase; the totals land in a device array read back through iwvi_debug_phase_cycles.
#ifdef IWVI_PHASE_TIMING
extern __device__ unsigned long long iwvi_phase_cycles[3][16];
#define PHASE_DECL long long ph_t_ = clock64(); long long ph_acc_[16] = {0}
#define PHASE_MARK(k) do { if (threadIdx.x == 0) { const long long now_ = clock64(); ph_acc_[k] += now_ - ph_t_; ph_t_ = now_; } } while (0)
#define PHASE_FLUSH(kern) do { if (threadIdx.x == 0) for (int k_ = 0; k_ < 16; k_++) atomicAdd(&iwvi_phase_cycles[kern][k_], (unsigned long long)ph_acc_[k_]); } while (0)
#else
#define PHASE_DECL
#define PHASE_MARK(k)
#define PHASE_FLUSH(kern)
#endif

// the CUDA error behind the calling thread's last IWVI_ERR_LAUNCH (gp_prologue.cu; read through iwvi_last_cuda_error)
void iwvi_note_cuda_error(cudaError_t e);
#define IWVI_CHECK_LAUNCH() do { const cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) { iwvi_note_cuda_error(e_); return IWVI_ERR_LAUNCH; } } while (0)
