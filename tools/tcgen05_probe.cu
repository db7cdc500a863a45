// tcgen05 / TMEM feasibility probe (sm_100a): D[128 x N] = A[128 x K] * B[N x K]^T in TF32 with FP32 accumulation in
// tensor memory, operands written to shared memory by the threads themselves in the K-major no-swizzle ("interleave")
// canonical layout  [row / 8][k / 4][row % 8][k % 4]  (8 x 16-byte core matrices; LBO = 128 B between K chunks,
// SBO = (K / 4) * 128 B between 8-row groups), one elected thread issuing tcgen05.mma, completion through
// tcgen05.commit -> mbarrier, epilogue with tcgen05.ld.  Prints max |D - ref| for the plain TF32 product and for the
// 3xTF32 split (hi*hi + hi*lo + lo*hi) against a float64 reference.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tcgen05_probe tcgen05_probe.cu && ./tcgen05_probe
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

constexpr int M = 128, N = 64, K = 64;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version (Blackwell)
  return d;                 // layout type 0 (no swizzle), base offset 0
}

__device__ __forceinline__ float to_tf32(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

// split == 0: plain TF32 (operands rounded);  split == 1: 3xTF32
__global__ void __launch_bounds__(128) probe_kernel(const double* A, const double* B, float* D, int split) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  float* Ahi = reinterpret_cast<float*>(smem_raw);             // [M/8][K/4][8][4]
  float* Alo = Ahi + M * K;
  float* Bhi = Alo + M * K;                                    // [N/8][K/4][8][4]
  float* Blo = Bhi + N * K;
  __shared__ uint64_t mbar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  auto fill = [&](const double* src, float* hi, float* lo, int rows) {
    for (int e = tid; e < rows * K; e += 128) {
      const int r = e / K, k = e % K;
      const double x = src[e];
      const float h = to_tf32((float)x);
      const float l = to_tf32((float)(x - (double)h));
      const int off = ((r >> 3) * (K / 4) + (k >> 2)) * 32 + (r & 7) * 4 + (k & 3);
      hi[off] = h; lo[off] = l;
    }
  };
  fill(A, Ahi, Alo, M);
  fill(B, Bhi, Blo, N);
  // generic-proxy writes -> visible to the async proxy (the tensor core reads shared memory through it)
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(64));
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
    // instruction descriptor: D = F32, A = B = TF32, both K-major, N >> 3 at bit 17, M >> 4 at bit 24
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const uint32_t lbo = 128, sbo = (K / 4) * 128;
    int first = 1;
    const int nprod = split ? 3 : 1;
    for (int pr = 0; pr < nprod; pr++) {
      const float* a = (pr == 2) ? Alo : Ahi;
      const float* b = (pr == 1) ? Blo : Bhi;
      for (int k0 = 0; k0 < K; k0 += 8) {
        const uint64_t ad = make_desc(smem_u32(a) + (k0 / 4) * 128, lbo, sbo);
        const uint64_t bd = make_desc(smem_u32(b) + (k0 / 4) * 128, lbo, sbo);
        const uint32_t acc = first ? 0u : 1u;
        first = 0;
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem_base), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
  }
  // everybody waits for the MMAs
  {
    uint32_t done = 0;
    while (!done) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(smem_u32(&mbar)), "r"(0) : "memory");
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // epilogue: warp w reads TMEM lanes 32 w .. 32 w + 31 (rows of D), 8 columns at a time
  for (int c0 = 0; c0 < N; c0 += 8) {
    uint32_t r[8];
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 8; j++) D[(size_t)(warp * 32 + lane) * N + c0 + j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(64));
}

int main() {
  std::vector<double> A(M * K), B(N * K), ref(M * N);
  srand(1);
  for (auto& v : A) v = (rand() / (double)RAND_MAX - 0.5) * 2.0;
  for (auto& v : B) v = (rand() / (double)RAND_MAX - 0.5) * 2.0;
  for (int i = 0; i < M; i++)
    for (int j = 0; j < N; j++) {
      double s = 0;
      for (int k = 0; k < K; k++) s += A[i * K + k] * B[j * K + k];
      ref[i * N + j] = s;
    }
  double *dA, *dB; float* dD;
  CK(cudaMalloc(&dA, A.size() * 8)); CK(cudaMalloc(&dB, B.size() * 8)); CK(cudaMalloc(&dD, M * N * 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 8, cudaMemcpyHostToDevice));
  const int smem = (2 * M * K + 2 * N * K) * 4 + 1024;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  std::vector<float> D(M * N);
  for (int split = 0; split < 2; split++) {
    CK(cudaMemset(dD, 0, M * N * 4));
    probe_kernel<<<1, 128, smem>>>(dA, dB, dD, split);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(D.data(), dD, M * N * 4, cudaMemcpyDeviceToHost));
    double err = 0, mx = 0;
    for (int i = 0; i < M * N; i++) { err = fmax(err, fabs(D[i] - ref[i])); mx = fmax(mx, fabs(ref[i])); }
    printf("{\"probe\": \"tcgen05 tf32 %s\", \"M\": %d, \"N\": %d, \"K\": %d, \"max_abs_err\": %.3e, \"max_ref\": %.3f, \"D00\": %.6f, \"ref00\": %.6f}\n",
           split ? "3x split" : "plain", M, N, K, err, mx, D[0], ref[0]);
  }
  return 0;
}
