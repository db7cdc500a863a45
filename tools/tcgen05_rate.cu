// tcgen05.mma issue-rate probe: one CTA per SM, one thread issues REPS x (K / 8) TF32 MMAs (M = 128, N = 256, K = 8) on
// operands resident in shared memory, for two K-major layouts: no swizzle ("interleave", 8 x 16-byte core matrices) and
// 64-byte swizzle (16 TF32 per row).  Prints ns per MMA and the implied TFLOP/s per SM and per GPU.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tcgen05_rate tcgen05_rate.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)
constexpr int M = 128, N = 256, K = 16;
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)layout << 61);
}
__global__ void __launch_bounds__(128) rate_kernel(int reps, int swz, unsigned long long* out) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  float* A = reinterpret_cast<float*>(smem_raw);
  float* B = A + M * K;
  __shared__ uint64_t mbar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int e = tid; e < (M + N) * K; e += 128) A[e] = 1.0f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;
  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const long long t0 = clock64();
    for (int r = 0; r < reps; r++) {
      for (int k0 = 0; k0 < K; k0 += 8) {
        uint64_t ad, bd;
        if (swz) {   // 64-byte swizzle: rows of 64 bytes, 8-row groups of 512 bytes, k-step = 32 bytes inside the row
          ad = make_desc(smem_u32(A) + (k0 / 8) * 32, 16, 512, 4);
          bd = make_desc(smem_u32(B) + (k0 / 8) * 32, 16, 512, 4);
        } else {     // interleave: [row/8][k/4][8][4]: LBO = 128, SBO = (K/4) * 128
          ad = make_desc(smem_u32(A) + (k0 / 4) * 128, 128, (K / 4) * 128, 0);
          bd = make_desc(smem_u32(B) + (k0 / 4) * 128, 128, (K / 4) * 128, 0);
        }
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tmem_base), "l"(ad), "l"(bd), "r"(idesc), "r"(1u) : "memory");
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(smem_u32(&mbar)), "r"(0) : "memory");
    out[blockIdx.x] = (unsigned long long)(clock64() - t0);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(256));
}
int main() {
  cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
  const int sms = pr.multiProcessorCount, reps = 2000;
  unsigned long long* d; CK(cudaMalloc(&d, sms * 8));
  const int smem = (M + N) * K * 4 + 1024;
  CK(cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  for (int swz = 0; swz < 2; swz++) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    rate_kernel<<<sms, 128, smem>>>(reps, swz, d); CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0)); rate_kernel<<<sms, 128, smem>>>(reps, swz, d); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    const double n_mma = (double)reps * (K / 8);
    const double ns = ms * 1e6 / n_mma, tf = 2.0 * M * N * 8 / ns / 1e3;
    printf("{\"layout\": \"%s\", \"ns_per_mma\": %.1f, \"tflops_per_sm\": %.2f, \"tflops_gpu\": %.0f}\n",
           swz ? "swizzle64" : "interleave", ns, tf, tf * sms);
  }
  return 0;
}
