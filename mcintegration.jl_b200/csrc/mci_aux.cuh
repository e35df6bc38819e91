/*
 * mci_aux.cuh — the small kernels either side of the fill kernels:
 *   build_maps_kernel   grids -> (grid, width) pair tables, Discrete tables -> staged image
 *   finalize_*_kernel   per-tile / per-CTA partials -> the packed per-iteration statistics
 *   refine_kernel       Dist.train! on the device (variable.jl:206-239, 369-382; common.jl:43-83)
 *   reweight_kernel     doReweight! (main.jl:322-346)
 *   fp64_peak_kernel    DFMA throughput probe (the roofline denominator MEASURED_PEAKS.json lacks)
 *
 * Packed statistics (raw sums, no floors), identical on every rank after the all-reduce:
 *   [hist(hist_len) | block_obs(n_block*obs_total) | block_cnt(n_block) | visited(nd) |
 *    propose(pa_len) | accept(pa_len) | neval(1)]
 */
#pragma once
#include "mci_device.cuh"

namespace mci {

struct PackedLayout {
  int off_hist, off_obs, off_cnt, off_visited, off_propose, off_accept, off_neval, len;
  int n_block, obs_total, hist_len, nd, pa_len;
};

/* staged image: Continuous leaf -> nbins x (g[i], g[i+1]-g[i]); Discrete -> dist[n] | accum[n+1] */
__global__ void build_maps_kernel(const __grid_constant__ DevProblem P, const double *__restrict__ grids,
                                  const double *__restrict__ dist, const double *__restrict__ accum,
                                  const int *__restrict__ leaf_grid_off, const int *__restrict__ leaf_dist_off,
                                  const int *__restrict__ leaf_acc_off, double *__restrict__ gmaps) {
  const int l = blockIdx.x;
  const DevLeaf &L = P.leaf[l];
  double *out = gmaps + L.smem_off;
  if (L.kind == MCI_LEAF_CONTINUOUS) {
    const double *g = grids + leaf_grid_off[l];
    for (int i = threadIdx.x; i < L.nbins; i += blockDim.x) {
      out[2 * i] = g[i];
      out[2 * i + 1] = __dsub_rn(g[i + 1], g[i]);
    }
  } else if (L.kind == MCI_LEAF_DISCRETE) { /* a FermiK component has no map */
    const double *d = dist + leaf_dist_off[l];
    const double *a = accum + leaf_acc_off[l];
    for (int i = threadIdx.x; i < L.nbins; i += blockDim.x) out[i] = d[i];
    for (int i = threadIdx.x; i <= L.nbins; i += blockDim.x) out[L.nbins + i] = a[i];
  }
}

/* VEGAS: one warp per output element; lanes stride over the partials, fixed shuffle tree */
__global__ void finalize_vegas_kernel(PackedLayout Y, const double *__restrict__ tile_obs,
                                      const double *__restrict__ tile_cnt, const double *__restrict__ cta_hist,
                                      int n_hist_copies, int n_local_blocks, int tiles_per_block, int rank, int world,
                                      uint64_t neval_per_block, double *__restrict__ packed) {
  const int lane = threadIdx.x & 31;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nw = (gridDim.x * blockDim.x) >> 5;
  const int n_obs_el = Y.n_block * Y.obs_total;
  const int total = Y.hist_len + n_obs_el + Y.n_block + 1;
  for (int e = gw; e < total; e += nw) {
    double t = 0.0;
    int dst;
    if (e < Y.hist_len) {
      for (int c = lane; c < n_hist_copies; c += 32) t += cta_hist[(size_t)c * Y.hist_len + e];
      dst = Y.off_hist + e;
    } else if (e < Y.hist_len + n_obs_el) {
      const int q = e - Y.hist_len;
      const int b = q / Y.obs_total, o = q - b * Y.obs_total;
      dst = Y.off_obs + q;
      if (b % world == rank) {
        const int lb = b / world;
        const double *src = tile_obs + (size_t)lb * tiles_per_block * Y.obs_total + o;
        for (int k = lane; k < tiles_per_block; k += 32) t += src[(size_t)k * Y.obs_total];
      }
    } else if (e < Y.hist_len + n_obs_el + Y.n_block) {
      const int b = e - Y.hist_len - n_obs_el;
      dst = Y.off_cnt + b;
      if (b % world == rank) {
        const int lb = b / world;
        for (int k = lane; k < tiles_per_block; k += 32) t += tile_cnt[(size_t)lb * tiles_per_block + k];
      }
    } else {
      dst = Y.off_neval;
      t = lane == 0 ? (double)((uint64_t)n_local_blocks * neval_per_block) : 0.0;
    }
    t = warp_sum(t);
    if (lane == 0) packed[dst] = t;
  }
  /* visited / propose / accept are not touched by the VEGAS solver */
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < Y.nd + 2 * Y.pa_len; i += gridDim.x * blockDim.x)
    packed[Y.off_visited + i] = 0.0;
}

/* chain solvers: per-CTA partial statistics [obs(OT) | norm | visited(nd) | propose | accept | neval],
 * ctas_per_block consecutive CTAs per local block */
__global__ void finalize_chain_kernel(PackedLayout Y, const double *__restrict__ cta_stats,
                                      const double *__restrict__ cta_hist, int n_hist_copies, int n_local_blocks,
                                      int ctas_per_block, int rank, int world, double *__restrict__ packed) {
  const int lane = threadIdx.x & 31;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nw = (gridDim.x * blockDim.x) >> 5;
  const int stat_len = Y.obs_total + 1 + Y.nd + 2 * Y.pa_len + 1;
  const int n_obs_el = Y.n_block * Y.obs_total;
  const int n_glob = Y.nd + 2 * Y.pa_len + 1;
  const int total = Y.hist_len + n_obs_el + Y.n_block + n_glob;
  const int n_cta = n_local_blocks * ctas_per_block;
  for (int e = gw; e < total; e += nw) {
    double t = 0.0;
    int dst;
    if (e < Y.hist_len) {
      for (int c = lane; c < n_hist_copies; c += 32) t += cta_hist[(size_t)c * Y.hist_len + e];
      dst = Y.off_hist + e;
    } else if (e < Y.hist_len + n_obs_el + Y.n_block) {
      int b, o;
      if (e < Y.hist_len + n_obs_el) {
        const int q = e - Y.hist_len;
        b = q / Y.obs_total; o = q - b * Y.obs_total;
        dst = Y.off_obs + q;
      } else {
        b = e - Y.hist_len - n_obs_el; o = Y.obs_total;
        dst = Y.off_cnt + b;
      }
      if (b % world == rank) {
        const int lb = b / world;
        for (int k = lane; k < ctas_per_block; k += 32) t += cta_stats[(size_t)(lb * ctas_per_block + k) * stat_len + o];
      }
    } else {
      const int q = e - (Y.hist_len + n_obs_el + Y.n_block); /* visited | propose | accept | neval: contiguous */
      dst = Y.off_visited + q;
      for (int k = lane; k < n_cta; k += 32) t += cta_stats[(size_t)k * stat_len + Y.obs_total + 1 + q];
    }
    t = warp_sum(t);
    if (lane == 0) packed[dst] = t;
  }
}

/* ------------------------------------------------------------------------------------------ */
/* train! on the device.  One CTA per leaf.                                                      */
/* Continuous leaves: the three running sums of variable.jl:206-239 (normalisation of rescale,   */
/* the mean increment weight, the regrid walk) are fixed-shape parallel reductions / a prefix    */
/* scan, and every new grid point is then computed independently from the prefix sums.  The      */
/* reference's own `sum` is pairwise and SIMD-reassociated, so no summation order is "the"       */
/* reference order; against a strictly sequential evaluation (the oracle) the new grid agrees   */
/* to ~1e-13 of an increment (asserted at rtol 1e-9, tests/test_gpu_parity.py).  A sequential   */
/* single-thread version of the same step took 0.25 ms per 999-bin leaf (division latency in the */
/* regrid loop): the whole cost of a small job.                                                   */
/* Discrete leaves (a handful of bins) keep the reference's order exactly.                       */
/* err[0] |= 1 when the reference would assert.                                                  */
/* ------------------------------------------------------------------------------------------ */

/* sum over the CTA in a fixed shape: xor-shuffle tree per warp, then the warp totals the same way; all threads get it */
__device__ __forceinline__ double cta_allsum(double v, double *sred32) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) sred32[warp] = v;
  __syncthreads();
  double t = lane < nwarp ? sred32[lane] : 0.0;
  t = warp_sum(t);
  return t;
}

__global__ void refine_kernel(const __grid_constant__ DevProblem P, PackedLayout Y, const double *__restrict__ packed,
                              const double *__restrict__ alpha, const int *__restrict__ leaf_grid_off,
                              const int *__restrict__ leaf_dist_off, const int *__restrict__ leaf_acc_off,
                              double *__restrict__ grids, double *__restrict__ dist, double *__restrict__ accum,
                              int *__restrict__ err) {
  extern __shared__ double sh[];
  const int l = blockIdx.x;
  const DevLeaf &L = P.leaf[l];
  if (!L.adapt) return;
  const int n = L.nbins;
  double *hist = sh;          /* n   (later: the prefix sums of the rescaled weights) */
  double *avg = sh + n;       /* n   */
  double *og = sh + 2 * n;    /* n+1 old grid */
  __shared__ double s_sum;
  __shared__ double s_red[32];
  __shared__ int s_bad;
  const double floor_ = __dmul_rn((double)(Y.n_block + 1), 1.0e-10); /* (B+1) x clearStatistics! floor, variable.jl:565 */
  if (threadIdx.x == 0) s_bad = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double h = __dadd_rn(floor_, packed[Y.off_hist + L.hist_off + i]);
    hist[i] = h;
    if (!isfinite(h) || !(h > 0.0)) s_bad = 1; /* variable.jl:212-213 */
  }
  __syncthreads();
  const double a = alpha[l];
  if (L.kind == MCI_LEAF_CONTINUOUS) {
    double *g = grids + leaf_grid_off[l];
    for (int i = threadIdx.x; i <= n; i += blockDim.x) og[i] = g[i];
    /* smooth(hist, 6.0): common.jl:43-54 */
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      double v;
      if (n <= 1) v = hist[i];
      else if (i == 0) v = __ddiv_rn(__dadd_rn(__dmul_rn(hist[0], 7.0), hist[1]), 8.0);
      else if (i == n - 1) v = __ddiv_rn(__dadd_rn(__dmul_rn(hist[n - 1], 7.0), hist[n - 2]), 8.0);
      else v = __ddiv_rn(__dadd_rn(__dadd_rn(hist[i - 1], __dmul_rn(hist[i], 6.0)), hist[i + 1]), 8.0);
      avg[i] = v;
    }
    __syncthreads();
    /* rescale(avg, alpha): common.jl:67-83 */
    if (n > 1) {
      double part = 0.0;
      for (int i = threadIdx.x; i < n; i += blockDim.x) part = __dadd_rn(part, avg[i]);
      const double s = cta_allsum(part, s_red);
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        if (!(avg[i] > 0.0)) s_bad = 1;
        double dd = __ddiv_rn(avg[i], s);
        if (dd > 0.0 && dd <= 0.99999999) dd = pow(__ddiv_rn(-(__dsub_rn(1.0, dd)), log(dd)), a);
        if (!isfinite(dd)) s_bad = 1;
        avg[i] = dd;
      }
      __syncthreads();
    }
    /* regrid: variable.jl:218-236.  The walk `while acc < f: j += 1; acc += avg[j]` ... `acc -= f` leaves, for new
     * grid point i, j = the first increment whose prefix sum S_j reaches (i - 1) f and acc = S_j - (i - 1) f: with the
     * prefix sums in shared memory every point is independent */
    {
      const int T = blockDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = (T + 31) >> 5;
      const int chunk = (n + T - 1) / T;
      const int lo = min(n, tid * chunk), hi = min(n, lo + chunk);
      double local = 0.0;
      for (int j = lo; j < hi; ++j) local = __dadd_rn(local, avg[j]);
      double incl = local; /* inclusive scan of the chunk totals: within the warp, then over the warps */
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const double up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl = __dadd_rn(incl, up);
      }
      __syncthreads();
      if (lane == 31) s_red[warp] = incl;
      __syncthreads();
      if (warp == 0) {
        double t = lane < nwarp ? s_red[lane] : 0.0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const double up = __shfl_up_sync(0xffffffffu, t, o);
          if (lane >= o) t = __dadd_rn(t, up);
        }
        if (lane < nwarp) s_red[lane] = t;
      }
      __syncthreads();
      double run = __dadd_rn(warp > 0 ? s_red[warp - 1] : 0.0, __dsub_rn(incl, local));
      double *S = hist; /* the raw histogram is no longer needed */
      for (int j = lo; j < hi; ++j) { run = __dadd_rn(run, avg[j]); S[j] = run; }
      __syncthreads();
      if (!s_bad) {
        const double f_ninc = __ddiv_rn(S[n - 1], (double)n);
        for (int i = 2 + tid; i <= n; i += T) {
          const double target = __dmul_rn((double)(i - 1), f_ninc);
          int jl = 0, jh = n - 1; /* first j with S[j] >= target (clamped to the last increment) */
          while (jl < jh) {
            const int jm = (jl + jh) >> 1;
            if (S[jm] >= target) jh = jm; else jl = jm + 1;
          }
          const double acc_f = __dsub_rn(S[jl], target);
          g[i - 1] = __dsub_rn(og[jl + 1], __dmul_rn(__ddiv_rn(acc_f, avg[jl]), __dsub_rn(og[jl + 1], og[jl])));
        }
      }
    }
  } else {
    /* train!(::Discrete): variable.jl:369-382 */
    double *d = dist + leaf_dist_off[l];
    double *ac = accum + leaf_acc_off[l];
    for (int i = threadIdx.x; i < n; i += blockDim.x) avg[i] = hist[i];
    __syncthreads();
    if (n > 1) {
      if (threadIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < n; ++i) s = __dadd_rn(s, avg[i]);
        s_sum = s;
      }
      __syncthreads();
      const double s = s_sum;
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double dd = __ddiv_rn(avg[i], s);
        if (dd > 0.0 && dd <= 0.99999999) dd = pow(__ddiv_rn(-(__dsub_rn(1.0, dd)), log(dd)), a);
        if (!isfinite(dd)) s_bad = 1;
        avg[i] = dd;
      }
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      double s = 0.0;
      for (int i = 0; i < n; ++i) s = __dadd_rn(s, avg[i]);
      s_sum = s;
    }
    __syncthreads();
    if (!s_bad) {
      const double s = s_sum;
      for (int i = threadIdx.x; i < n; i += blockDim.x) { avg[i] = __ddiv_rn(avg[i], s); d[i] = avg[i]; }
      __syncthreads();
      for (int i = threadIdx.x; i <= n; i += blockDim.x) { /* sum(distribution[1:i]) recomputed for every i (:376) */
        double t = 0.0;
        for (int k = 0; k < i; ++k) t = __dadd_rn(t, avg[k]);
        ac[i] = t;
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0 && s_bad) atomicOr(err, 1);
}

/* HIST_SMEM_FIXED: exponent of the next iteration's fixed-point scale from this iteration's (all-reduced) histogram:
 * 2^E M = 2^FIXED_CAP_LOG2 with M the largest bin; invalid (-> the fp64 kernel runs) while M is not settled within
 * 2^FIXED_STABLE_LOG2 of the iteration before.  state[0] = M of the previous call. */
__global__ void fixed_scale_kernel(PackedLayout Y, const double *__restrict__ packed, int *__restrict__ fixed_exp,
                                   double *__restrict__ state) {
  __shared__ double s_red[32];
  double m = 0.0;
  bool finite = true;
  for (int i = threadIdx.x; i < Y.hist_len; i += blockDim.x) {
    const double h = packed[Y.off_hist + i];
    finite = finite && isfinite(h);
    m = fmax(m, h);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  const int bad = __syncthreads_or(!finite);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)((blockDim.x + 31) >> 5); ++w) m = fmax(m, s_red[w]);
    const double prev = state[0];
    const double band = scalbn(1.0, FIXED_STABLE_LOG2);
    int e = FIXED_EXP_INVALID;
    if (!bad && m > 0.0 && prev > 0.0 && m <= prev * band && prev <= m * band) {
      int ex;
      frexp(m, &ex); /* m = f 2^ex, f in [0.5, 1) */
      e = FIXED_CAP_LOG2 - ex;
      if (e > 900 || e < -900) e = FIXED_EXP_INVALID;
    }
    *fixed_exp = e;
    state[0] = bad ? 0.0 : m;
  }
}

/* mci_set_maps on a fixed-point handle: move the uploaded maps from the staging area into place and, if they DIFFER from
 * the maps they replace, forget the fixed-point scale — it was learnt from the old map's histogram and says nothing about
 * the new one (the fp64 kernel then runs until the histogram has settled again).  A host that hands back the very maps
 * it read (the end-to-end loop of bench.py; a warm restart of the same handle) keeps the scale. */
__global__ void adopt_maps_kernel(const double *__restrict__ stage, double *__restrict__ maps, int n, int *__restrict__ fixed_exp,
                                  double *__restrict__ state) {
  int changed = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double v = stage[i];
    if (__double_as_longlong(v) != __double_as_longlong(maps[i])) changed = 1;
    maps[i] = v;
  }
  if (__syncthreads_or(changed) && threadIdx.x == 0) { *fixed_exp = FIXED_EXP_INVALID; state[0] = 0.0; }
}

/* mci_run_iterations: keep this iteration's packed statistics and flags in the per-call history (one launch) */
__global__ void snapshot_kernel(const double *__restrict__ packed, int len, const unsigned long long *__restrict__ flags,
                                double *__restrict__ rec) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < len; i += gridDim.x * blockDim.x) rec[i] = packed[i];
  if (blockIdx.x == 0 && threadIdx.x < 4) reinterpret_cast<unsigned long long *>(rec + len)[threadIdx.x] = flags[threadIdx.x];
}

/* doReweight!: main.jl:322-346; goal may be NULL */
__global__ void reweight_kernel(PackedLayout Y, const double *__restrict__ packed, double gamma,
                                const double *__restrict__ goal, double *__restrict__ reweight) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const int nd = Y.nd;
  const double floor_ = __dmul_rn((double)(Y.n_block + 1), 1.0e-8);
  double visited[MAX_ND];
  double avgstep = 0.0;
  for (int i = 0; i < nd; ++i) {
    visited[i] = __dadd_rn(floor_, packed[Y.off_visited + i]);
    avgstep = __dadd_rn(avgstep, visited[i]);
  }
  for (int i = 0; i < nd; ++i) {
    if (visited[i] <= 1) reweight[i] = __dmul_rn(reweight[i], pow(avgstep, gamma));
    else reweight[i] = __dmul_rn(reweight[i], pow(__ddiv_rn(avgstep, visited[i]), gamma));
  }
  if (goal) {
    double gs = 0.0;
    for (int i = 0; i < nd; ++i) gs = __dadd_rn(gs, goal[i]);
    for (int i = 0; i < nd; ++i) reweight[i] = __dmul_rn(reweight[i], __ddiv_rn(goal[i], gs));
  }
  double s = 0.0;
  for (int i = 0; i < nd; ++i) s = __dadd_rn(s, reweight[i]);
  for (int i = 0; i < nd; ++i) reweight[i] = __ddiv_rn(reweight[i], s);
}

/* DFMA throughput probe: 8 independent chains per thread */
__global__ void fp64_peak_kernel(double *out, int iters, double a, double b) {
  double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  double s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
  if (s == 12345.678) out[0] = s;
}

}  // namespace mci
