// lv_elbo.cu -- the HBM-bound stages of the IW-ELBO path, each a fused, coalesced elementwise/reduction kernel:
//
//   iwvi_lv_fwd / iwvi_lv_bwd          LatentVariableLayer.propagate + Encoder.__call__ (reference layers.py:72-105,
//                                      :137-152): the tanh MLP is evaluated ONCE per distinct row (the reference
//                                      evaluates it on K identical copies, models.py:113-116), then W = mu + eps*sigma,
//                                      [F, W] and log q(W) - log p(W) are produced per point.
//   iwvi_iwelbo_fwd / iwvi_iwelbo_bwd  Gaussian variational expectations + K-way logsumexp (models.py:133-150) or
//                                      the VI mean over samples (models.py:66-86); softmax weights kept for the adjoint.
//   iwvi_normal_fill                   counter-based N(0,1) noise (Philox4x32-10 + Box-Muller), keyed by the global
//                                      point index so that 1/2/4/8-GPU runs draw identical noise.
//   iwvi_adam_step                     Adam on the flat parameter buffer (tf.train.AdamOptimizer defaults,
//                                      build_models.py:293-295) fused with the gpflow `positive` transform chain rule.
// All cross-block reductions go through per-block partials summed in a fixed order (deterministic).
#include "common.cuh"

namespace {

#define LV_RB 8            // encoder rows per block iteration (one warp each)
#define LV_MAX_GRID 592    // 4 x 148
#define LV_W IWVI_MAX_ENC_WIDTH

__device__ __forceinline__ double softplus_d(double x) { return x > 0.0 ? x + log1p(exp(-x)) : log1p(exp(x)); }
__device__ __forceinline__ double sigmoid_d(double x) { return 1.0 / (1.0 + exp(-x)); }
// Encoder non-linearity (IWVI_ACT_*) and its derivative expressed through the OUTPUT t = act(a), which is what the backward
// pass has at hand (it recomputes the activations, not the pre-activations)
__device__ __forceinline__ double act_fwd(int act, double a) {
  switch (act) {
    case IWVI_ACT_TANH: return tanh(a);
    case IWVI_ACT_RELU: return a > 0.0 ? a : 0.0;
    case IWVI_ACT_SIGMOID: return sigmoid_d(a);
    case IWVI_ACT_SOFTPLUS: return softplus_d(a);
    case IWVI_ACT_ELU: return a > 0.0 ? a : expm1(a);
    default: return a;
  }
}
__device__ __forceinline__ double act_grad_from_output(int act, double t) {
  switch (act) {
    case IWVI_ACT_TANH: return 1.0 - t * t;
    case IWVI_ACT_RELU: return t > 0.0 ? 1.0 : 0.0;
    case IWVI_ACT_SIGMOID: return t * (1.0 - t);
    case IWVI_ACT_SOFTPLUS: return -expm1(-t);          // sigmoid(a) = 1 - exp(-softplus(a))
    case IWVI_ACT_ELU: return t > 0.0 ? 1.0 : t + 1.0;
    default: return 1.0;
  }
}

struct LvParams {
  iwvi_lv_desc d;
  const double *F, *enc_in, *params, *eps, *mu_in, *sigma_in, *d_samples, *d_kl, *d_mu, *d_sigma;
  double *samples, *kl, *mu, *sigma, *d_params, *dF, *ws;
  int n_params, n_groups;
  int rb;   // forward kernel: encoder rows per block iteration (<= LV_RB), chosen by the host so that the grid fills the SMs
};

// encoder forward for one row held by one warp.  acts: [n_layers+1][LV_W] of this warp (all layers kept).
__device__ __forceinline__ void encoder_row(const iwvi_lv_desc& d, const double* __restrict__ params,
                                            double* acts, int lane) {
  int off = 0;
  for (int l = 0; l < d.n_layers; l++) {
    const int din = d.dims[l], dout = d.dims[l + 1];
    const double* W = params + off;
    const double* b = W + din * dout;
    const double* h = acts + l * LV_W;
    double* o = acts + (l + 1) * LV_W;
    for (int u = lane; u < dout; u += 32) {
      double a = b[u];
      for (int i = 0; i < din; i++) a += h[i] * W[i * dout + u];
      if (l < d.n_layers - 1) a = act_fwd(d.act, a);
      if (din == dout) a += h[u];
      o[u] = a;
    }
    __syncwarp();
    off += din * dout + dout;
  }
}

__global__ void __launch_bounds__(256) lv_fwd_kernel(const LvParams p) {
  // Per block iteration: `rb` encoder rows (one warp each evaluates the MLP), then all 256 threads write the rows'
  // Kt points -- [F, W] and the local regulariser -- with consecutive threads on consecutive doubles of the row-major
  // outputs (fully coalesced 8-byte accesses; rows are Df + Lw doubles wide, odd at c3, so wider vectors would straddle
  // rows) and 32-bit index arithmetic.  rb is chosen by the host so that the grid covers the SMs: at c3 (512 rows) 8 rows
  // per block left 64 CTAs on 148 SMs.
  __shared__ double acts[LV_RB][(IWVI_MAX_ENC_LAYERS + 1) * LV_W];
  __shared__ double mu_s[LV_RB][IWVI_MAX_LW], sg_s[LV_RB][IWVI_MAX_LW];
  const iwvi_lv_desc& d = p.d;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int Lw = d.Lw, Df = d.Df, Kt = d.Kt, C = Df + Lw, rb = p.rb;
  for (int grp = blockIdx.x; grp < p.n_groups; grp += gridDim.x) {
    const int n0 = grp * rb;
    const int n = n0 + warp;
    __syncthreads();
    if (warp < rb && n < d.Be) {
      if (d.prior) {
        if (lane < Lw) { mu_s[warp][lane] = d.prior_mu; sg_s[warp][lane] = d.prior_sigma; }
      } else {
        for (int i = lane; i < d.Dxy; i += 32) acts[warp][i] = p.enc_in[(size_t)n * d.Dxy + i];
        __syncwarp();
        encoder_row(d, p.params, acts[warp], lane);
        const double* o = acts[warp] + d.n_layers * LV_W;
        if (lane < Lw) { mu_s[warp][lane] = o[lane]; sg_s[warp][lane] = softplus_d(o[Lw + lane] - 3.0); }
      }
      __syncwarp();
      if (lane < Lw) {
        if (p.mu) p.mu[(size_t)n * Lw + lane] = mu_s[warp][lane];
        if (p.sigma) p.sigma[(size_t)n * Lw + lane] = sg_s[warp][lane];
      }
    }
    __syncthreads();
    const int rows = min(rb, d.Be - n0);
    const size_t p0 = (size_t)n0 * Kt;
    const double* Fsrc = d.f_bcast ? p.F + (size_t)n0 * Df : p.F + p0 * Df;
    const double* eps = p.eps + p0 * Lw;
    double* smp = p.samples + p0 * C;
    double* klp = p.kl ? p.kl + p0 * Lw : nullptr;
    const int ne = rows * Kt * C;                       // <= 8 * Kt * C: 32-bit arithmetic throughout
    for (int e = tid; e < ne; e += 256) {
      const int pl = e / C, c = e - pl * C;
      const int nl = pl / Kt;
      double v;
      if (c < Df) {
        v = Fsrc[(d.f_bcast ? nl : pl) * Df + c];
      } else {
        const int j = c - Df;
        const double m = mu_s[nl][j], sg = sg_s[nl][j];
        const double z = eps[pl * Lw + j];
        v = m + z * sg;
        if (klp)
          klp[pl * Lw + j] = d.sampled ? -0.5 * z * z - log(sg) + 0.5 * v * v
                                       : 0.5 * m * m + 0.5 * (sg * sg - 1.0 - log(sg * sg));
      }
      smp[e] = v;
    }
  }
}

__global__ void __launch_bounds__(256) lv_bwd_kernel(const LvParams p) {
  __shared__ double acts[LV_RB][(IWVI_MAX_ENC_LAYERS + 1) * LV_W];
  __shared__ double dout_s[LV_RB][LV_W], dpre_s[LV_RB][LV_W];
  const iwvi_lv_desc& d = p.d;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int Lw = d.Lw, Df = d.Df, Kt = d.Kt, C = Df + Lw;
  double* part = p.ws + (size_t)blockIdx.x * p.n_params;
  if (!d.prior)
    for (int e = tid; e < p.n_params; e += 256) part[e] = 0.0;
  for (int grp = blockIdx.x; grp < p.n_groups; grp += gridDim.x) {
    const int n0 = grp * LV_RB;
    const int n = n0 + warp;
    const bool valid = n < d.Be;
    __syncthreads();
    // dF
    if (p.dF && p.d_samples) {
      const int rows = min(LV_RB, d.Be - n0);
      if (d.f_bcast) {
        for (int e = tid; e < rows * Df; e += 256) {
          const int nl = e / Df, c = e - nl * Df;
          double s = 0.0;
          for (int k = 0; k < Kt; k++) s += p.d_samples[((size_t)(n0 + nl) * Kt + k) * C + c];
          p.dF[(size_t)(n0 + nl) * Df + c] = s;
        }
      } else {
        const int64_t p0 = (int64_t)n0 * Kt;
        for (int64_t e = tid; e < (int64_t)rows * Kt * Df; e += 256) {
          const int64_t pl = e / Df;
          const int c = (int)(e - pl * Df);
          p.dF[(size_t)(p0 + pl) * Df + c] = p.d_samples[(size_t)(p0 + pl) * C + c];
        }
      }
    }
    if (d.prior) continue;
    // recompute the activations of this row, reduce the point cotangents over its Kt points
    for (int i = lane; i < LV_W; i += 32) { dout_s[warp][i] = 0.0; dpre_s[warp][i] = 0.0; }
    if (valid) {
      for (int i = lane; i < d.Dxy; i += 32) acts[warp][i] = p.enc_in[(size_t)n * d.Dxy + i];
      __syncwarp();
      encoder_row(d, p.params, acts[warp], lane);
      const double* o = acts[warp] + d.n_layers * LV_W;
      for (int j = 0; j < Lw; j++) {
        const double m = p.mu_in[(size_t)n * Lw + j], s = p.sigma_in[(size_t)n * Lw + j];
        double mub = 0.0, sgb = 0.0;
        for (int k = lane; k < Kt; k += 32) {
          const size_t pt = (size_t)n * Kt + k;
          const double z = p.eps[pt * Lw + j];
          double wb = p.d_samples ? p.d_samples[pt * C + Df + j] : 0.0;
          const double dk = p.d_kl ? p.d_kl[pt * Lw + j] : 0.0;
          if (d.sampled) {
            wb += dk * (m + z * s);
            mub += wb;
            sgb += wb * z - dk / s;
          } else {
            mub += wb + dk * m;
            sgb += wb * z + dk * (s - 1.0 / s);
          }
        }
        mub = warp_sum(mub); sgb = warp_sum(sgb);
        if (lane == 0) {
          if (p.d_mu) mub += p.d_mu[(size_t)n * Lw + j];
          if (p.d_sigma) sgb += p.d_sigma[(size_t)n * Lw + j];
          dout_s[warp][j] = mub;
          dout_s[warp][Lw + j] = sgb * sigmoid_d(o[Lw + j] - 3.0);
        }
      }
    } else {
      for (int i = lane; i < (IWVI_MAX_ENC_LAYERS + 1) * LV_W; i += 32) acts[warp][i] = 0.0;
    }
    // back-propagate through the layers; block-level partial sums of dW, db
    int off = p.n_params;
    for (int l = d.n_layers - 1; l >= 0; l--) {
      const int din = d.dims[l], dout = d.dims[l + 1];
      off -= din * dout + dout;
      const double* W = p.params + off;
      const double* h = acts[warp] + l * LV_W;
      const double* o = acts[warp] + (l + 1) * LV_W;
      const bool skip = din == dout;
      __syncwarp();
      for (int u = lane; u < dout; u += 32) {
        double g = dout_s[warp][u];
        if (l < d.n_layers - 1) { const double t = o[u] - (skip ? h[u] : 0.0); g *= act_grad_from_output(d.act, t); }
        dpre_s[warp][u] = g;
      }
      __syncthreads();
      for (int e = tid; e < din * dout + dout; e += 256) {
        double s = 0.0;
        if (e < din * dout) {
          const int i = e / dout, u = e - i * dout;
#pragma unroll
          for (int w = 0; w < LV_RB; w++) s += acts[w][l * LV_W + i] * dpre_s[w][u];
        } else {
          const int u = e - din * dout;
#pragma unroll
          for (int w = 0; w < LV_RB; w++) s += dpre_s[w][u];
        }
        part[off + e] += s;
      }
      // cotangent of this layer's input
      double nd[2] = {0.0, 0.0};
      for (int i = lane, q = 0; i < din; i += 32, q++) {
        double s = skip ? dout_s[warp][i] : 0.0;
        for (int u = 0; u < dout; u++) s += dpre_s[warp][u] * W[i * dout + u];
        nd[q] = s;
      }
      __syncthreads();
      for (int i = lane, q = 0; i < LV_W; i += 32, q++) dout_s[warp][i] = (i < din) ? nd[q] : 0.0;
    }
  }
}

__global__ void lv_bwd_final_kernel(const LvParams p, int nblocks) {
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < p.n_params; e += gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int b = 0; b < nblocks; b++) s += p.ws[(size_t)b * p.n_params + e];
    p.d_params[e] = s;
  }
}

int lv_check(const iwvi_lv_desc* d) {
  if (!d) return IWVI_ERR_NULL;
  if (d->Be < 0 || d->Kt < 1 || d->Df < 0 || d->Lw < 1 || d->Lw > IWVI_MAX_LW) return IWVI_ERR_BAD_DESC;
  if (!d->prior) {
    if (d->act < 0 || d->act > IWVI_ACT_IDENTITY) return IWVI_ERR_BAD_DESC;
    if (d->n_layers < 1 || d->n_layers > IWVI_MAX_ENC_LAYERS) return IWVI_ERR_UNSUPPORTED;
    if (d->dims[0] != d->Dxy || d->dims[d->n_layers] != 2 * d->Lw) return IWVI_ERR_BAD_DESC;
    for (int l = 0; l <= d->n_layers; l++)
      if (d->dims[l] < 1 || d->dims[l] > LV_W) return IWVI_ERR_UNSUPPORTED;
  }
  return IWVI_OK;
}
int lv_nparams(const iwvi_lv_desc* d) {
  int n = 0;
  if (!d->prior)
    for (int l = 0; l < d->n_layers; l++) n += d->dims[l] * d->dims[l + 1] + d->dims[l + 1];
  return n;
}

// ------------------------------------------------------------------------------------------------
// likelihood + logsumexp
// ------------------------------------------------------------------------------------------------
#define HALF_LOG_2PI 0.91893853320467274178

struct ElboParams {
  iwvi_elbo_desc d;
  const double *fmean, *fvar, *Y, *lik_var, *kl_local, *w_in, *d_elbo;
  double *elbo_data, *logp, *w, *dmean, *dvar, *dkl_local, *dlik, *ws;
  int nblocks;
};

__global__ void __launch_bounds__(256) elbo_fwd_kernel(const ElboParams p) {
  __shared__ double red[8];
  const iwvi_elbo_desc& d = p.d;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = blockIdx.x * 8 + warp;
  const int K = d.K, Dy = d.Dy, Lw = d.Lw, B = d.B;
  double lp = 0.0;
  if (n < B) {
    const double sv = p.lik_var[0];
    const double c0 = -HALF_LOG_2PI - 0.5 * log(sv);
    double mx = -INFINITY, se = 0.0, sl = 0.0;
    for (int k = lane; k < K; k += 32) {
      const size_t pt = d.data_major ? (size_t)n * K + k : (size_t)k * B + n;
      double L = 0.0;
      for (int j = 0; j < Dy; j++) {
        const double df = p.Y[(size_t)n * Dy + j] - p.fmean[pt * Dy + j];
        L += c0 - 0.5 * (df * df + p.fvar[pt * Dy + j]) / sv;
      }
      for (int j = 0; j < Lw; j++) L -= p.kl_local[pt * Lw + j];
      p.w[(size_t)n * K + k] = L;
      sl += L;
      if (L > mx) { se = se * exp(mx - L) + 1.0; mx = L; }
      else se += exp(L - mx);
    }
    if (d.iw) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double mx2 = __shfl_xor_sync(0xffffffffu, mx, o);
        const double se2 = __shfl_xor_sync(0xffffffffu, se, o);
        const double m = fmax(mx, mx2);
        const double a = (mx == -INFINITY) ? 0.0 : se * exp(mx - m);
        const double b = (mx2 == -INFINITY) ? 0.0 : se2 * exp(mx2 - m);
        se = a + b; mx = m;
      }
      const double lse = mx + log(se);
      lp = lse - log((double)K);
      for (int k = lane; k < K; k += 32) p.w[(size_t)n * K + k] = exp(p.w[(size_t)n * K + k] - lse);
    } else {
      sl = warp_sum(sl);
      lp = sl / (double)K;
      for (int k = lane; k < K; k += 32) p.w[(size_t)n * K + k] = 1.0 / (double)K;
    }
    if (lane == 0 && p.logp) p.logp[n] = lp;
  }
  if (lane == 0) red[warp] = (n < B) ? lp : 0.0;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < 8; i++) s += red[i];
    p.ws[blockIdx.x] = s;
  }
}

__global__ void elbo_fwd_final_kernel(const ElboParams p) {
  __shared__ double red[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < p.nblocks; i += blockDim.x) s += p.ws[i];
  s = block_sum(s, red);
  if (threadIdx.x == 0) p.elbo_data[0] = p.d.scale * s;
}

__global__ void __launch_bounds__(256) elbo_bwd_kernel(const ElboParams p) {
  __shared__ double red[32];
  const iwvi_elbo_desc& d = p.d;
  const int K = d.K, Dy = d.Dy, Lw = d.Lw, B = d.B;
  const int64_t T = (int64_t)B * K;
  const double sv = p.lik_var[0];
  const double ge = p.d_elbo[0] * d.scale;
  double dl = 0.0;
  for (int64_t pt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; pt < T; pt += (int64_t)gridDim.x * blockDim.x) {
    int n, k;
    if (d.data_major) { n = (int)(pt / K); k = (int)(pt - (int64_t)n * K); }
    else { k = (int)(pt / B); n = (int)(pt - (int64_t)k * B); }
    const double g = ge * p.w_in[(size_t)n * K + k];
    for (int j = 0; j < Dy; j++) {
      const double df = p.Y[(size_t)n * Dy + j] - p.fmean[pt * Dy + j];
      p.dmean[pt * Dy + j] = g * df / sv;
      p.dvar[pt * Dy + j] = -g / (2.0 * sv);
      dl += g * (-0.5 / sv + 0.5 * (df * df + p.fvar[pt * Dy + j]) / (sv * sv));
    }
    if (p.dkl_local)
      for (int j = 0; j < Lw; j++) p.dkl_local[pt * Lw + j] = -g;
  }
  const double tot = block_sum(dl, red);
  if (threadIdx.x == 0) p.ws[blockIdx.x] = tot;
}

__global__ void elbo_bwd_final_kernel(const ElboParams p) {
  __shared__ double red[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < p.nblocks; i += blockDim.x) s += p.ws[i];
  s = block_sum(s, red);
  if (threadIdx.x == 0) p.dlik[0] = s;
}

int elbo_check(const iwvi_elbo_desc* d) {
  if (!d) return IWVI_ERR_NULL;
  if (d->B < 1 || d->K < 1 || d->Dy < 1 || d->Lw < 0) return IWVI_ERR_BAD_DESC;
  return IWVI_OK;
}
#define ELBO_BWD_GRID 592

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 + Box-Muller
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// `state` (device, may be NULL): state[0] = optimiser steps completed so far.  With it the seed of the step being run is
// derived on the device, seed = seed_base * C1 + (state[0] + 1 + step_add) * C2 + (layer + 1) * C3 (mod 2^64), the same
// mix engine.layer_seed applies on the host -- so that a captured CUDA graph draws fresh noise on every replay.
__global__ void normal_fill_kernel(double* out, int64_t f0, int64_t n, uint64_t seed, const int64_t* state, int64_t step_add,
                                   int layer) {
  if (state)
    seed = seed * 0x9E3779B97F4A7C15ull + (uint64_t)(state[0] + 1 + step_add) * 0xBF58476D1CE4E5B9ull +
           (uint64_t)(layer + 1) * 0x94D049BB133111EBull;
  const int64_t q0 = f0 >> 1, q1 = (f0 + n - 1) >> 1;
  for (int64_t q = q0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q <= q1; q += (int64_t)gridDim.x * blockDim.x) {
    uint32_t r[4];
    philox4x32_10((uint32_t)q, (uint32_t)((uint64_t)q >> 32), 0u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
    const uint64_t a = ((uint64_t)r[1] << 32) | r[0], b = ((uint64_t)r[3] << 32) | r[2];
    const double u1 = ((double)(a >> 11) + 0.5) * (1.0 / 9007199254740992.0);
    const double u2 = ((double)(b >> 11) + 0.5) * (1.0 / 9007199254740992.0);
    const double rad = sqrt(-2.0 * log(u1));
    double sn, cs;
    sincospi(2.0 * u2, &sn, &cs);
    const int64_t e0 = 2 * q - f0, e1 = e0 + 1;
    if (e0 >= 0 && e0 < n) out[e0] = rad * cs;
    if (e1 >= 0 && e1 < n) out[e1] = rad * sn;
  }
}

// ------------------------------------------------------------------------------------------------
// Adam (+ positive transform)
// ------------------------------------------------------------------------------------------------
// `state` / `lr_dev` (device, may be NULL): step count and learning rate read on the device (CUDA-graph replay); then
// lr_t is the plain learning rate and the bias correction is formed here from t = state[0] + 1.
__global__ void adam_kernel(double* x, const double* g_elbo, double* m, double* v, const double* mask, double* theta_pos,
                            int64_t n, int64_t n_pos, double lr_t, double b1, double b2, double eps, const int64_t* state,
                            const double* lr_dev) {
  if (state) {
    const double t = (double)(state[0] + 1);
    lr_t = lr_dev[0] * sqrt(1.0 - pow(b2, t)) / (1.0 - pow(b1, t));
  }
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    // masked-out entries are not in the optimiser's var_list (set_trainable(False)) or belong to another segment of a
    // step that updates the parameters segment by segment: value, moments and constrained copy stay as they are
    if (mask && mask[i] == 0.0) continue;
    double xi = x[i];
    double g = -g_elbo[i];                         // the optimiser minimises -ELBO
    if (i < n_pos) g *= sigmoid_d(xi);             // d softplus(x) / dx
    const double mi = b1 * m[i] + (1.0 - b1) * g;
    const double vi = b2 * v[i] + (1.0 - b2) * g * g;
    m[i] = mi; v[i] = vi;
    xi -= lr_t * mi / (sqrt(vi) + eps);
    x[i] = xi;
    if (i < n_pos) theta_pos[i] = softplus_d(xi) + 1e-6;
  }
}
__global__ void step_counter_inc_kernel(int64_t* state) { state[0] += 1; }

// Minibatch assembly: rows idx[b] (idx == NULL: row b) of the resident data arrays into the plan's X, Y and [X, Y] buffers
// in one launch (Xb / Yb may be NULL when they are the sources themselves).
__global__ void batch_gather_kernel(const double* __restrict__ X, const double* __restrict__ Y,
                                    const int64_t* __restrict__ idx, int B, int Dx, int Dy, double* __restrict__ Xb,
                                    double* __restrict__ Yb, double* __restrict__ XYb) {
  const int W = Dx + Dy;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < B * W; e += gridDim.x * blockDim.x) {
    const int b = e / W, c = e - b * W;
    const int64_t row = idx ? idx[b] : b;
    const double v = c < Dx ? X[row * Dx + c] : Y[row * Dy + (c - Dx)];
    if (XYb) XYb[e] = v;
    if (c < Dx) { if (Xb) Xb[(size_t)b * Dx + c] = v; }
    else if (Yb) Yb[(size_t)b * Dy + (c - Dx)] = v;
  }
}
__global__ void positive_fwd_kernel(const double* x, double* theta, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    theta[i] = softplus_d(x[i]) + 1e-6;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" int64_t iwvi_lv_param_doubles(const iwvi_lv_desc* d) {
  if (lv_check(d) != IWVI_OK) return -1;
  return lv_nparams(d);
}
extern "C" int64_t iwvi_lv_bwd_ws_doubles(const iwvi_lv_desc* d) {
  if (lv_check(d) != IWVI_OK) return -1;
  return (int64_t)LV_MAX_GRID * lv_nparams(d) + 8;
}

extern "C" int iwvi_lv_fwd(const iwvi_lv_desc* d, const double* F, const double* enc_in, const double* params,
                           const double* eps, double* samples, double* kl, double* mu, double* sigma, void* stream) {
  int rc = lv_check(d);
  if (rc != IWVI_OK) return rc;
  if (!eps || !samples || (d->Df > 0 && !F)) return IWVI_ERR_NULL;
  if (!d->prior && (!enc_in || !params)) return IWVI_ERR_NULL;
  if (d->Be == 0) return IWVI_OK;
  LvParams p = {};
  p.d = *d; p.F = F; p.enc_in = enc_in; p.params = params; p.eps = eps;
  p.samples = samples; p.kl = kl; p.mu = mu; p.sigma = sigma;
  p.n_params = lv_nparams(d);
  // rows per block: as many as keep >= 2 x 148 blocks in flight (8 at most: one warp per row), and few enough that the
  // 32-bit element count of a block iteration (rows * Kt * (Df + Lw)) cannot overflow
  int rb = LV_RB;
  while (rb > 1 && (d->Be + rb - 1) / rb < 2 * 148) rb >>= 1;
  while (rb > 1 && (int64_t)rb * d->Kt * (d->Df + d->Lw) > (int64_t)1 << 30) rb >>= 1;
  if ((int64_t)rb * d->Kt * (d->Df + d->Lw) > (int64_t)1 << 30) return IWVI_ERR_UNSUPPORTED;
  p.rb = rb;
  p.n_groups = (d->Be + rb - 1) / rb;
  const int grid = p.n_groups < 4 * LV_MAX_GRID ? p.n_groups : 4 * LV_MAX_GRID;
  lv_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p);
  IWVI_CHECK_LAUNCH();
  return IWVI_OK;
}

extern "C" int iwvi_lv_bwd(const iwvi_lv_desc* d, const double* F, const double* enc_in, const double* params,
                           const double* eps, const double* mu, const double* sigma, const double* d_samples,
                           const double* d_kl, const double* d_mu, const double* d_sigma, double* d_params,
                           double* dF, double* ws, void* stream) {
  int rc = lv_check(d);
  if (rc != IWVI_OK) return rc;
  if (!eps) return IWVI_ERR_NULL;
  if (!d->prior && (!enc_in || !params || !mu || !sigma || !d_params || !ws)) return IWVI_ERR_NULL;
  if (d->Be == 0) return IWVI_ERR_BAD_DESC;
  LvParams p = {};
  p.d = *d; p.F = F; p.enc_in = enc_in; p.params = params; p.eps = eps; p.mu_in = mu; p.sigma_in = sigma;
  p.d_samples = d_samples; p.d_kl = d_kl; p.d_mu = d_mu; p.d_sigma = d_sigma;
  p.d_params = d_params; p.dF = dF; p.ws = ws;
  p.n_params = lv_nparams(d);
  p.n_groups = (d->Be + LV_RB - 1) / LV_RB;
  const int grid = p.n_groups < LV_MAX_GRID ? p.n_groups : LV_MAX_GRID;
  cudaStream_t st = (cudaStream_t)stream;
  lv_bwd_kernel<<<grid, 256, 0, st>>>(p);
  IWVI_CHECK_LAUNCH();
  if (!d->prior) {
    lv_bwd_final_kernel<<<(p.n_params + 127) / 128, 128, 0, st>>>(p, grid);
    IWVI_CHECK_LAUNCH();
  }
  return IWVI_OK;
}

extern "C" int64_t iwvi_elbo_ws_doubles(const iwvi_elbo_desc* d) {
  if (elbo_check(d) != IWVI_OK) return -1;
  const int64_t a = (d->B + 7) / 8;
  return a > ELBO_BWD_GRID ? a : ELBO_BWD_GRID;
}

extern "C" int iwvi_iwelbo_fwd(const iwvi_elbo_desc* d, const double* fmean, const double* fvar, const double* Y,
                               const double* lik_var, const double* kl_local, double* elbo_data, double* logp,
                               double* w, double* ws, void* stream) {
  int rc = elbo_check(d);
  if (rc != IWVI_OK) return rc;
  if (!fmean || !fvar || !Y || !lik_var || !elbo_data || !w || !ws) return IWVI_ERR_NULL;
  if (d->Lw > 0 && !kl_local) return IWVI_ERR_NULL;
  ElboParams p = {};
  p.d = *d; p.fmean = fmean; p.fvar = fvar; p.Y = Y; p.lik_var = lik_var; p.kl_local = kl_local;
  p.elbo_data = elbo_data; p.logp = logp; p.w = w; p.ws = ws;
  p.nblocks = (d->B + 7) / 8;
  cudaStream_t st = (cudaStream_t)stream;
  elbo_fwd_kernel<<<p.nblocks, 256, 0, st>>>(p);
  IWVI_CHECK_LAUNCH();
  elbo_fwd_final_kernel<<<1, 256, 0, st>>>(p);
  IWVI_CHECK_LAUNCH();
  return IWVI_OK;
}

extern "C" int iwvi_iwelbo_bwd(const iwvi_elbo_desc* d, const double* fmean, const double* fvar, const double* Y,
                               const double* lik_var, const double* w, const double* d_elbo, double* dmean,
                               double* dvar, double* dkl_local, double* dlik, double* ws, void* stream) {
  int rc = elbo_check(d);
  if (rc != IWVI_OK) return rc;
  if (!fmean || !fvar || !Y || !lik_var || !w || !d_elbo || !dmean || !dvar || !dlik || !ws) return IWVI_ERR_NULL;
  ElboParams p = {};
  p.d = *d; p.fmean = fmean; p.fvar = fvar; p.Y = Y; p.lik_var = lik_var; p.w_in = w; p.d_elbo = d_elbo;
  p.dmean = dmean; p.dvar = dvar; p.dkl_local = dkl_local; p.dlik = dlik; p.ws = ws;
  const int64_t T = (int64_t)d->B * d->K;
  int64_t nb = (T + 255) / 256;
  p.nblocks = (int)(nb < ELBO_BWD_GRID ? nb : ELBO_BWD_GRID);
  cudaStream_t st = (cudaStream_t)stream;
  elbo_bwd_kernel<<<p.nblocks, 256, 0, st>>>(p);
  IWVI_CHECK_LAUNCH();
  elbo_bwd_final_kernel<<<1, 256, 0, st>>>(p);
  IWVI_CHECK_LAUNCH();
  return IWVI_OK;
}

extern "C" int iwvi_normal_fill(double* out, int64_t n_points, int32_t C, int64_t first_point, uint64_t seed,
                                void* stream) {
  if (!out) return IWVI_ERR_NULL;
  if (n_points < 0 || C < 1 || first_point < 0) return IWVI_ERR_BAD_DESC;
  const int64_t n = n_points * C;
  if (n == 0) return IWVI_OK;
  const int64_t pairs = n / 2 + 2;
  int64_t nb = (pairs + 255) / 256;
  if (nb > 4 * 148) nb = 4 * 148;
  normal_fill_kernel<<<(int)nb, 256, 0, (cudaStream_t)stream>>>(out, first_point * C, n, seed, nullptr, 0, 0);
  IWVI_CHECK_LAUNCH();
  return IWVI_OK;
}

extern "C" int iwvi_normal_fill_counter(double* out, int64_t n_points, int32_t C, int64_t first_point, uint64_t seed_base,
                                        int32_t layer, int64_t step_add, const int64_t* state, void* stream) {
  if (!out || !state) return IWVI_ERR_NULL;
  if (n_points < 0 || C < 1 || first_point < 0 || layer < 0) return IWVI_ERR_BAD_DESC;
  const int64_t n = n_points * C;
  if (n == 0) return IWVI_OK;
  const int64_t pairs = n / 2 + 2;
  int64_t nb = (pairs + 255) / 256;
  if (nb > 4 * 148) nb = 4 * 148;
  normal_fill_kernel<<<(int)nb, 256, 0, (cudaStream_t)stream>>>(out, first_point * C, n, seed_base, state, step_add, layer);
  IWVI_CHECK_LAUNCH();
  return IWVI_OK;
}

extern "C" int iwvi_batch_gather(const double* X, const double* Y, const int64_t* idx, int32_t B, int32_t Dx, int32_t Dy,
                                 double* Xb, double* Yb, double* XYb, void* stream) {
  if (!X || !Y) return IWVI_ERR_NULL;
  if (B < 0 || Dx < 1 || Dy < 1) return IWVI_ERR_BAD_DESC;
  if (B == 0) return IWVI_OK;
  const int n = B * (Dx + Dy);
  int nb = (n + 255) / 256;
  if (nb > 4 * 148) nb = 4 * 148;
  batch_gather_kernel<<<nb, 256, 0, (cudaStream_t)stream>>>(X, Y, idx, B, Dx, Dy, Xb, Yb, XYb);
  IWVI_CHECK_LAUNCH();
  return IWVI_OK;
}

extern "C" int iwvi_positive_fwd(const double* x, double* theta, int64_t n, void* stream) {
  if (!x || !theta) return IWVI_ERR_NULL;
  if (n <= 0) return IWVI_OK;
  positive_fwd_kernel<<<(int)((n + 255) / 256 < 592 ? (n + 255) / 256 : 592), 256, 0, (cudaStream_t)stream>>>(x, theta, n);
  IWVI_CHECK_LAUNCH();
  return IWVI_OK;
}

extern "C" int iwvi_adam_step(double* x, const double* grad_elbo, double* m, double* v, const double* mask,
                              double* theta_pos, int64_t n, int64_t n_pos, double lr, double beta1, double beta2,
                              double eps, int64_t t, void* stream) {
  if (!x || !grad_elbo || !m || !v) return IWVI_ERR_NULL;
  if (n_pos > 0 && !theta_pos) return IWVI_ERR_NULL;
  if (n <= 0 || t < 1) return IWVI_ERR_BAD_DESC;
  const double lr_t = lr * sqrt(1.0 - pow(beta2, (double)t)) / (1.0 - pow(beta1, (double)t));
  int64_t nb = (n + 255) / 256;
  if (nb > 4 * 148) nb = 4 * 148;
  adam_kernel<<<(int)nb, 256, 0, (cudaStream_t)stream>>>(x, grad_elbo, m, v, mask, theta_pos, n, n_pos, lr_t, beta1,
                                                        beta2, eps, nullptr, nullptr);
  IWVI_CHECK_LAUNCH();
  return IWVI_OK;
}

extern "C" int iwvi_adam_step_counter_part(double* x, const double* grad_elbo, double* m, double* v, const double* mask,
                                           double* theta_pos, int64_t n, int64_t n_pos, const double* lr, double beta1,
                                           double beta2, double eps, int64_t* state, int32_t advance, void* stream) {
  if (!x || !grad_elbo || !m || !v || !lr || !state) return IWVI_ERR_NULL;
  if (n_pos > 0 && !theta_pos) return IWVI_ERR_NULL;
  if (n <= 0) return IWVI_ERR_BAD_DESC;
  int64_t nb = (n + 255) / 256;
  if (nb > 4 * 148) nb = 4 * 148;
  adam_kernel<<<(int)nb, 256, 0, (cudaStream_t)stream>>>(x, grad_elbo, m, v, mask, theta_pos, n, n_pos, 0.0, beta1, beta2,
                                                        eps, state, lr);
  IWVI_CHECK_LAUNCH();
  if (advance) {
    step_counter_inc_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(state);
    IWVI_CHECK_LAUNCH();
  }
  return IWVI_OK;
}

extern "C" int iwvi_adam_step_counter(double* x, const double* grad_elbo, double* m, double* v, const double* mask,
                                      double* theta_pos, int64_t n, int64_t n_pos, const double* lr, double beta1,
                                      double beta2, double eps, int64_t* state, void* stream) {
  return iwvi_adam_step_counter_part(x, grad_elbo, m, v, mask, theta_pos, n, n_pos, lr, beta1, beta2, eps, state, 1, stream);
}

#ifdef IWVI_PHASE_TIMING
__device__ unsigned long long iwvi_phase_cycles[3][16];
// debug only (not part of the ABI): copies the 3 x 16 phase totals to host memory and clears them
extern "C" __attribute__((visibility("default"))) int iwvi_debug_phase_cycles(unsigned long long* host_out) {
  unsigned long long zero[3][16] = {};
  if (cudaDeviceSynchronize() != cudaSuccess) return IWVI_ERR_LAUNCH;
  if (cudaMemcpyFromSymbol(host_out, iwvi_phase_cycles, sizeof(zero)) != cudaSuccess) return IWVI_ERR_LAUNCH;
  if (cudaMemcpyToSymbol(iwvi_phase_cycles, zero, sizeof(zero)) != cudaSuccess) return IWVI_ERR_LAUNCH;
  return IWVI_OK;
}
#endif
