// gp_bwd.cuh -- workspace layout and launch parameters shared by the launches of iwvi_gp_rows_bwd:
//   gp_rows_bwd.cu  (epilogue adjoint, tile kernel, split-K reduce, finalize)    gp_gram_bwd.cu  (gram adjoint)
#pragma once
#include <stdlib.h>
#include "common.cuh"

#define EPI_STRIDE 1344   // >= P*R + D*P + P + 1 at the maxima (32*8 + 32*32 + 32 + 1 = 1313)
#define EPI_PTS 32
#define TILE_PART_EXTRA 40  // dls[32], dvariance, pad
#ifndef GRAM_PTS
#define GRAM_PTS 16         // points per strip of the gram-adjoint kernel (32: register spills at 128 registers per thread)
#endif
#define GRAM_THREADS 256
#define GRAM_CTAS_PER_SM 2

struct BwdWs {   // workspace layout (doubles)
  int64_t off_bbar, off_gmb, off_gvb, off_gvt, off_epi, off_tile, off_red, off_qred, total;
  int Tp, n_epi, grid_tile, S, npairs, chunks_per_split, tile_stride;
  int gram_slots;   // per-CTA partial slots of the gram-adjoint kernel per point chain (its grid never exceeds this)
};
// host-side list-scheduling model behind the choice of S (see bwd_ws_layout); the last answer is cached per thread
// because the entry points recompute the layout on every call
inline int pick_reduce_split(int items, int npairs, int nchunks, int nsm) {
  if (const char* e = getenv("IWVI_REDUCE_S")) { const int v = atoi(e); if (v >= 1) return v < nchunks ? v : nchunks; }   // tuning aid
  static thread_local int key[4] = {-1, -1, -1, -1}, cached = 1;
  if (key[0] == items && key[1] == npairs && key[2] == nchunks && key[3] == nsm) return cached;
  int bestS = 1;
  double best = 1e300;
  const int ns = nsm < 256 ? (nsm > 0 ? nsm : 1) : 256;
  const int s_cap = nchunks >= 2048 ? 32 : 16;      // (c4-sized problems: 32 measured 3 % faster than 16; c3: 14)
  for (int S = 1; S <= s_cap && S <= nchunks; S++) {
    const int cps = (nchunks + S - 1) / S;
    double freeat[256];
    for (int k = 0; k < ns; k++) freeat[k] = 0.0;
    double makespan = 0.0;
    for (int it = 0; it < items * S; it++) {
      const int pair = it % npairs;
      int bi = 0;
      while ((bi + 1) * (bi + 2) / 2 <= pair) bi++;
      const bool diag = (pair - bi * (bi + 1) / 2) == bi;
      int k = 0;
      for (int k2 = 1; k2 < ns; k2++) if (freeat[k2] < freeat[k]) k = k2;
      freeat[k] += cps * (diag ? 0.6 : 1.0);
      if (freeat[k] > makespan) makespan = freeat[k];
    }
    const double cost = makespan + 0.5 * S;
    if (cost < best) { best = cost; bestS = S; }
  }
  key[0] = items; key[1] = npairs; key[2] = nchunks; key[3] = nsm; cached = bestS;
  return bestS;
}

inline bool fast_reduce_ok(const iwvi_gp_desc& d) {   // the tcgen05 variant tiles the output 128 x 256
  return (d.flags & IWVI_FLAG_FAST_REDUCE) && iwvi_round_up(d.M, IWVI_BLK) % 128 == 0;
}

inline BwdWs bwd_ws_layout(const iwvi_gp_desc& d, int nsm, bool fast = false) {
  BwdWs w;
  const AuxLayout al = iwvi_aux_layout(d.M, d.D, d.R);
  const SaveLayout sv = iwvi_save_layout(d.T, d.M, d.R);
  w.Tp = sv.Tp;
  w.n_epi = (w.Tp + EPI_PTS - 1) / EPI_PTS;
  w.grid_tile = nsm;
  w.npairs = al.NB * (al.NB + 1) / 2;
  const int nchunks = w.Tp / IWVI_BLK;
  const int items = (d.R + 1) * w.npairs;
  // Split the points into S ranges.  CTAs are dispatched in blockIdx order to whichever SM frees up first; a diagonal
  // pair costs ~0.6 of an off-diagonal one (reduce_diag).  Pick the S whose simulated makespan is smallest, with a
  // small charge per extra partial the finalize kernel has to sum.
  int bestS = pick_reduce_split(items, w.npairs, nchunks, nsm);
  if (fast) {
    // tcgen05 variant: (R + 1) x tiles CTAs per point range, one wave in all
    const int mts = (al.Mp + 255) / 256, tiles = mts * (mts + 1) / 2;   // 256 x 256 tiles of the lower block triangle
    bestS = nsm / (d.R * tiles);                                           // (dLm stays on the float64 kernel)
    const int min_s = (nchunks + 63) / 64;                                 // at most 4096 points per CTA (its scale table)
    if (bestS < min_s) bestS = min_s;
    if (bestS < 1) bestS = 1;
    if (bestS > nchunks) bestS = nchunks;
  }
  w.chunks_per_split = (nchunks + bestS - 1) / bestS;
  w.S = (nchunks + w.chunks_per_split - 1) / w.chunks_per_split;
  w.tile_stride = al.Mp * al.ldz + TILE_PART_EXTRA;
  int64_t o = 0;
  w.off_bbar = o; o += sv.u_stride;            // Bbar / 2, block-major like the saved A
  w.off_gmb = o;  o += (int64_t)w.Tp * IWVI_MAX_R;
  w.off_gvb = o;  o += (int64_t)w.Tp * IWVI_MAX_R;
  w.off_gvt = o;  o += (int64_t)w.Tp * IWVI_MAX_R;   // 2 gvar_bar transposed [r][Tp]: the reduce kernel's per-chunk scale vectors, one bulk copy each
  w.off_epi = o;  o += (int64_t)w.n_epi * EPI_STRIDE;
  w.gram_slots = GRAM_CTAS_PER_SM * nsm;
  w.off_tile = o; o += (int64_t)2 * w.gram_slots * w.tile_stride;  // per-CTA partials of the gram-adjoint kernel: two point chains (see iwvi_gp_rows_bwd_range)
  w.off_red = o;  o += (int64_t)(d.R + 1) * w.S * w.npairs * IWVI_BLK * IWVI_BLK;
  w.off_qred = o; o += (int64_t)w.S * al.NB * IWVI_BLK * IWVI_MAX_R;
  w.total = o;
  return w;
}

struct BwdParams {
  iwvi_gp_desc d;
  const double *Lm, *aux, *save, *X, *W, *mfA, *mfb, *eps, *d_sample, *d_mean, *d_var;
  double *dX, *dZ, *dls, *dvariance, *dq_mu, *dq_sqrt, *dLm, *dW, *dmfA, *dmfb, *ws;
  BwdWs wl;
  int ntiles, grid_tile;
  int tile0, tile1;   // tile kernel: tiles [tile0, tile1) of this launch
  int strip0, strip1; // gram kernel: strips of GRAM_PTS points [strip0, strip1) of this launch
  int slot0;          // gram kernel: first per-CTA partial slot of this launch (0: first chain, wl.gram_slots: second chain)
  int n_slots;        // gram kernel: slots per chain; a ranged launch zeroes the slots of its chain that its grid does not own
  int epi0;           // epilogue kernel: first 32-point CTA of this launch
  int q_lo;      // reduce kernel: first matrix index of this launch (0 .. R; R == dLm)
  int q_n;       // reduce kernel: number of matrices of this launch
  int qmu_only;  // reduce kernel: 1 = only the items (q_lo, block row bi, block column 0), the ones that also form dq_mu
  int fin_part;  // finalize kernel: 0 = everything, 1 = part A outputs, 2 = part B outputs
};

// gp_gram_bwd.cu: the gram adjoint (dX, per-CTA partials of dZ / dls / dvariance) of strips [p.tile0, p.tile1) of GRAM_PTS points
int iwvi_launch_gram_bwd(const BwdParams& p, int nsm, int max_smem, cudaStream_t st);
int iwvi_gram_bwd_grid(const BwdParams& p);   // CTAs (= partial slots written) of that launch
