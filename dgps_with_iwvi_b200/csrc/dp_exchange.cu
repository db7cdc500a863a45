// dp_exchange.cu -- the exchange step of data-parallel training as ONE-SHOT all-reduce over NVLink peer memory.
//
// The reference has no multi-GPU path; data parallelism over minibatch rows is what SURVEY.md section 8(e) adds, and its
// only exchange is the sum of the parameter gradients (+ the ELBO slot) once per layer and step: a few hundred KB to
// 1.3 MB per segment, i.e. latency-bound.  Through NCCL a segment costs pack (gather of the entries that travel) +
// all-reduce + unpack = three launches and ~40 us at the tail of the step, where nothing hides it.  Here it is three
// small launches that do pack, transfer, reduction and unpack themselves over the peers' memory (mapped by
// torch.distributed._symmetric_memory, which is plumbing: allocation + handle exchange):
//
//   iwvi_dp_push    every rank gathers its segment through the index list and STORES it into slot [parity][rank] of
//                   every peer's receive buffer (8-byte stores over NVLink / NVSwitch); the last CTA to finish
//                   publishes the segment's epoch in every peer's flag word (release at system scope);
//   iwvi_dp_reduce  waits (a one-warp launch) until all ranks' flags carry this epoch (acquire at system scope), sums the `world` slots in
//                   RANK ORDER -- the same order on every rank, so all replicas hold bit-identical sums -- and scatters
//                   the result into the gradient bucket.
//
// Epochs live in device memory (one counter per segment, advanced by the reduce kernel), so both launches replay inside
// a CUDA graph.  Slots alternate with the epoch's parity: a rank can only push epoch e + 2 after it has completed its own
// reduce of epoch e + 1, which needed every peer's push of e + 1, which that peer issued after its reduce of epoch e
// (stream order) -- so a slot is never overwritten while a peer still reads it.
#include "common.cuh"

namespace {

struct DpPeers {
  double* recv[IWVI_DP_MAX_RANKS];
  unsigned long long* flags[IWVI_DP_MAX_RANKS];
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__global__ void __launch_bounds__(256) dp_push_kernel(const double* __restrict__ g, const int64_t* __restrict__ index,
                                                      int64_t n, int64_t seg_off, int64_t bucket_len, int seg, int rank,
                                                      int world, DpPeers peers, const unsigned long long* epoch,
                                                      unsigned int* counter) {
  const unsigned long long e = epoch[seg] + 1ull;
  const int64_t slot = ((int64_t)(e & 1ull) * world + rank) * bucket_len + seg_off;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
    const double v = g[index[j]];
    for (int p = 0; p < world; p++) peers.recv[p][slot + j] = v;
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int ticket = atomicAdd(counter, 1u);
    if (ticket == gridDim.x - 1) {            // every CTA's stores are out (each fenced before taking its ticket)
      *counter = 0u;
      __threadfence_system();
      for (int p = 0; p < world; p++) st_release_sys(peers.flags[p] + (size_t)seg * world + rank, e);
    }
  }
}

__global__ void __launch_bounds__(256) dp_reduce_kernel(double* __restrict__ g, const int64_t* __restrict__ index, int64_t n,
                                                        int64_t seg_off, int64_t bucket_len, int seg, int world,
                                                        const double* recv, const unsigned long long* flags,
                                                        unsigned long long* epoch, unsigned int* counter) {
  const unsigned long long e = epoch[seg] + 1ull;
  const double* base = recv + (int64_t)(e & 1ull) * world * bucket_len + seg_off;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int src = 0; src < world; src++) s += __ldcv(base + (int64_t)src * bucket_len + j);   // rank order, not through L1
    g[index[j]] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int ticket = atomicAdd(counter, 1u);
    if (ticket == gridDim.x - 1) {            // the last CTA to FINISH: every CTA has read the epoch by now
      *counter = 0u;
      epoch[seg] = e;
    }
  }
}

// The wait is a launch of its own, ONE warp: a spinning CTA pins its SM's registers, and the step's persistent DMMA
// kernels need whole SMs -- with the wait inside the (multi-CTA) reduce kernel, several segments waiting at once
// could cover every SM of two ranks with spinners that wait for each other's pushes, queued behind kernels that can no
// longer be scheduled.  One warp per waiting segment cannot.
__global__ void __launch_bounds__(32) dp_wait_kernel(int seg, int world, const unsigned long long* flags,
                                                     const unsigned long long* epoch) {
  const unsigned long long e = epoch[seg] + 1ull;
  if ((int)threadIdx.x < world) {
    const long long t0 = clock64();
    while (ld_acquire_sys(flags + (size_t)seg * world + threadIdx.x) < e) {
      // a peer that never arrives (crashed rank) must not leave this GPU spinning for ever: fail loudly after ~60 s
      if (clock64() - t0 > 120000000000ll) __trap();
    }
  }
}

inline int dp_grid(int64_t n) {
  int64_t b = (n + 255) / 256;
  return (int)(b < 1 ? 1 : (b > 64 ? 64 : b));     // a handful of SMs: the step's DMMA kernels keep the rest
}

}  // namespace

extern "C" int iwvi_dp_push(const double* g, const int64_t* index, int64_t n, int64_t seg_off, int64_t bucket_len,
                            int32_t seg, int32_t rank, int32_t world, const uint64_t* recv_ptrs, const uint64_t* flag_ptrs,
                            const uint64_t* epoch, uint32_t* counters, void* stream) {
  if (!g || !index || !recv_ptrs || !flag_ptrs || !epoch || !counters) return IWVI_ERR_NULL;
  if (world < 1 || world > IWVI_DP_MAX_RANKS || rank < 0 || rank >= world || n < 0 || seg < 0) return IWVI_ERR_BAD_DESC;
  if (n == 0) return IWVI_OK;
  DpPeers peers;
  for (int p = 0; p < world; p++) {
    peers.recv[p] = reinterpret_cast<double*>(recv_ptrs[p]);
    peers.flags[p] = reinterpret_cast<unsigned long long*>(flag_ptrs[p]);
  }
  dp_push_kernel<<<dp_grid(n), 256, 0, (cudaStream_t)stream>>>(g, index, n, seg_off, bucket_len, seg, rank, world, peers,
                                                                reinterpret_cast<const unsigned long long*>(epoch),
                                                                counters + 2 * seg);
  IWVI_CHECK_LAUNCH();
  return IWVI_OK;
}

extern "C" int iwvi_dp_reduce(double* g, const int64_t* index, int64_t n, int64_t seg_off, int64_t bucket_len, int32_t seg,
                              int32_t world, const double* recv_local, const uint64_t* flags_local, uint64_t* epoch,
                              uint32_t* counters, void* stream) {
  if (!g || !index || !recv_local || !flags_local || !epoch || !counters) return IWVI_ERR_NULL;
  if (world < 1 || world > IWVI_DP_MAX_RANKS || n < 0 || seg < 0) return IWVI_ERR_BAD_DESC;
  if (n == 0) return IWVI_OK;
  dp_wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(seg, world, reinterpret_cast<const unsigned long long*>(flags_local),
                                                     reinterpret_cast<const unsigned long long*>(epoch));
  IWVI_CHECK_LAUNCH();
  dp_reduce_kernel<<<dp_grid(n), 256, 0, (cudaStream_t)stream>>>(g, index, n, seg_off, bucket_len, seg, world, recv_local,
                                                                  reinterpret_cast<const unsigned long long*>(flags_local),
                                                                  reinterpret_cast<unsigned long long*>(epoch),
                                                                  counters + 2 * seg + 1);
  IWVI_CHECK_LAUNCH();
  return IWVI_OK;
}
