// Measurement utility, not part of the reference's interface: the FP64 tensor-pipe (DMMA.8x8x4) issue rate of the
// device, so that bench.py can state its roofline denominator from a measurement made in the same run
// (MEASURED_PEAKS.json carries bf16 and HBM figures only).  Register-only mma.sync.m8n8k4.f64 loop, 8 independent
// accumulator pairs per warp; flops = 2 * 8*8*4 * 8 * iters per warp.
#include "common.cuh"

namespace {
__global__ void probe_dmma_kernel(double* out, int iters) {
  double c[8][2];
#pragma unroll
  for (int i = 0; i < 8; i++) { c[i][0] = 0.0; c[i][1] = 0.0; }
  const double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void debug_stamp_kernel(unsigned long long* slots, int idx) {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  slots[idx] = t;
}
}  // namespace

// Measurement utility: writes the GPU's global nanosecond timer to slots[idx] when the stream reaches this point (a
// one-thread kernel, capturable into a CUDA graph): tools/graph_timeline.py brackets every launch of a replayed training
// step with these to see which chains are exposed.
extern "C" int iwvi_debug_stamp(unsigned long long* slots, int32_t idx, void* stream) {
  if (!slots) return IWVI_ERR_NULL;
  if (idx < 0) return IWVI_ERR_BAD_DESC;
  debug_stamp_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(slots, idx);
  IWVI_CHECK_LAUNCH();
  return IWVI_OK;
}

extern "C" int iwvi_probe_dmma(double* out, int32_t blocks, int32_t warps, int32_t iters, void* stream) {
  if (!out) return IWVI_ERR_NULL;
  if (blocks <= 0 || warps <= 0 || warps > 32 || iters <= 0) return IWVI_ERR_BAD_DESC;
  probe_dmma_kernel<<<blocks, warps * 32, 0, (cudaStream_t)stream>>>(out, iters);
  IWVI_CHECK_LAUNCH();
  return IWVI_OK;
}
