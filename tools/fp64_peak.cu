// FP64 peak probe for B200 (sm_100a): DMMA.8x8x4 issue rate, DFMA rate, LDS-fed DMMA.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
// Output: one JSON line per experiment on stdout. Roofline denominator for the gram/solve stages.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)

template<int ILP>
__global__ void dmma_rate(double* out, int iters) {
  double c[ILP][2];
  #pragma unroll
  for (int i = 0; i < ILP; i++) { c[i][0] = 0.0; c[i][1] = 0.0; }
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; it++) {
    #pragma unroll
    for (int i = 0; i < ILP; i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
  #pragma unroll
  for (int i = 0; i < ILP; i++) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template<int ILP>
__global__ void dfma_rate(double* out, int iters) {
  double c[ILP];
  #pragma unroll
  for (int i = 0; i < ILP; i++) c[i] = threadIdx.x * 1e-3 + i;
  double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9;
  for (int it = 0; it < iters; it++) {
    #pragma unroll
    for (int i = 0; i < ILP; i++) c[i] = fma(c[i], a, b);
  }
  double s = 0;
  #pragma unroll
  for (int i = 0; i < ILP; i++) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// smem-fed: each warp computes a 32x16 tile (4x2 m8n8 tiles) with fragments loaded from smem each k-step
__global__ void dmma_lds(double* out, int iters) {
  extern __shared__ double sm[];
  const int lda = 68;
  for (int i = threadIdx.x; i < 64 * lda * 2; i += blockDim.x) sm[i] = 1e-3 * (i % 7);
  __syncthreads();
  const double* As = sm; const double* Bs = sm + 64 * lda;
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int g = lane >> 2, t = lane & 3;
  int wm = (warp & 1) * 32, wn = ((warp >> 1) & 3) * 16;
  double c[4][2][2];
  #pragma unroll
  for (int i = 0; i < 4; i++) for (int j = 0; j < 2; j++) { c[i][j][0] = 0; c[i][j][1] = 0; }
  for (int it = 0; it < iters; it++) {
    #pragma unroll 4
    for (int k = 0; k < 64; k += 4) {
      double a[4], b[2];
      #pragma unroll
      for (int i = 0; i < 4; i++) a[i] = As[(wm + i * 8 + g) * lda + k + t];
      #pragma unroll
      for (int j = 0; j < 2; j++) b[j] = Bs[(wn + j * 8 + g) * lda + k + t];
      #pragma unroll
      for (int i = 0; i < 4; i++)
        #pragma unroll
        for (int j = 0; j < 2; j++)
          asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};"
                       : "+d"(c[i][j][0]), "+d"(c[i][j][1]) : "d"(a[i]), "d"(b[j]));
    }
  }
  double s = 0;
  #pragma unroll
  for (int i = 0; i < 4; i++) for (int j = 0; j < 2; j++) s += c[i][j][0] + c[i][j][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template<typename F> float time_ms(F f) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  f(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < 3; r++) {
    CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount;
  printf("{\"device\":\"%s\",\"sms\":%d,\"clock_khz\":%d}\n", p.name, sms, p.clockRate);
  double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 1024));
  const int iters = 20000;
  int warps_list[] = {1, 2, 4, 8, 16, 32};
  for (int wi = 0; wi < 6; wi++) {
    int warps = warps_list[wi];
    #define RUN_DMMA(ILP) { float ms = time_ms([&]{ dmma_rate<ILP><<<sms, warps*32>>>(out, iters); }); \
      double fl = 2.0*256*ILP*(double)iters*warps*sms; \
      printf("{\"exp\":\"dmma\",\"warps_per_sm\":%d,\"ilp\":%d,\"tflops\":%.2f,\"ms\":%.3f}\n", warps, ILP, fl/ms/1e9, ms); }
    RUN_DMMA(1) RUN_DMMA(2) RUN_DMMA(4) RUN_DMMA(8)
    #define RUN_DFMA(ILP) { float ms = time_ms([&]{ dfma_rate<ILP><<<sms, warps*32>>>(out, iters); }); \
      double fl = 2.0*32*ILP*(double)iters*warps*sms; \
      printf("{\"exp\":\"dfma\",\"warps_per_sm\":%d,\"ilp\":%d,\"tflops\":%.2f,\"ms\":%.3f}\n", warps, ILP, fl/ms/1e9, ms); }
    RUN_DFMA(1) RUN_DFMA(4) RUN_DFMA(8)
  }
  CK(cudaFuncSetAttribute(dmma_lds, cudaFuncAttributeMaxDynamicSharedMemorySize, 64*68*2*8));
  for (int warps = 8; warps <= 16; warps += 8) {
    float ms = time_ms([&]{ dmma_lds<<<sms, warps*32, 64*68*2*8>>>(out, 2000); });
    double fl = 2.0*256*8*16*2000.0*warps*sms;
    printf("{\"exp\":\"dmma_lds_32x16\",\"warps_per_sm\":%d,\"tflops\":%.2f,\"ms\":%.3f}\n", warps, fl/ms/1e9, ms);
  }
  CK(cudaFree(out));
  return 0;
}
