/*
 * mci_chain.cuh — the Markov-chain solvers: one walker (chain) per thread.
 *
 *   VegasMC.montecarlo  src/vegas_mc/montecarlo.jl:112-241 + changeVariable src/vegas_mc/updates.jl:45-106
 *   MCMC.montecarlo     src/mcmc/montecarlo.jl:72-205 + changeIntegrand / changeVariable / swapVariable
 *                       src/mcmc/updates.jl:1-147
 *   pool operations     create!/remove!/shift!/swap! and rollbacks, src/distribution/sampler.jl:13-97, 293-462
 *   probabilities       src/distribution/variable.jl:587-641
 *
 * The reference runs ONE chain of neval_per_block steps per block.  A GPU needs ~10^5 concurrent
 * chains, so a block is `walkers_per_block` independent chains of neval_per_block/walkers steps
 * (recorded in mci_problem_t; walkers = 1 is the reference's decomposition).  Walker w of block b
 * draws sequentially from the Philox stream (seed, iter, b, w).  Per-step arithmetic follows the
 * reference operation by operation (stateful shift! ratios, rollback caches, padding products).
 *
 * Statistics: observable / normalisation / visited / propose / accept are accumulated per thread
 * and reduced per CTA in a fixed order; all walkers of a CTA belong to one block.
 */
#pragma once
#include "mci_integrands.cuh"

namespace mci {

struct ChainArgs {
  const double *gmaps;
  const double *reweight; /* [nd] */
  double *cta_stats;      /* [n_cta][stat_len]                                                   */
  double *cta_hist;       /* [n_hist_copies][hist_len]                                           */
  unsigned long long *flags; /* [0] non-finite weights, [1] walkers without a start point (mcmc) */
  uint64_t iter;
  uint64_t steps;         /* steps per walker = neval_per_block / walkers_per_block              */
  int walkers_per_block, ctas_per_block;
  int n_local_blocks, rank, world;
  int hist_mode, n_hist_copies, maps_in_smem;
};

/* per-walker pool state.  Flat index of (leaf, slot) = leaf.base + (slot-1)*leaf.stride, the same
 * indexing the integrands use; tprob holds the CompositeVar products per (variable, slot). */
struct Walker {
  double data[MAX_DRAW];
  double prob[MAX_DRAW];
  double tprob[MAX_DRAW];
  int gidx[MAX_DRAW]; /* 0-based bin: grid increment (Continuous) / value - lower (Discrete) */
  /* rollback cache of the last shift! (the reference keeps it in the pool's last element) */
  double bk_data[MCI_MAX_LEAF], bk_prob[MCI_MAX_LEAF], bk_tprob;
  int bk_gidx[MCI_MAX_LEAF];
};

struct ChainCtx {
  const DevProblem &P;
  const double *maps;
  const PhiloxKey &K;
  Rng rng;
  __device__ __forceinline__ ChainCtx(const DevProblem &p, const double *m, const PhiloxKey &k) : P(p), maps(m), K(k) {}
};

__device__ __forceinline__ int fidx(const DevLeaf &L, int pos) { return L.base + (pos - 1) * L.stride; }

/* create! : sampler.jl:293-305 (Continuous), :13-22 (Discrete).  Returns 1/prob. */
__device__ __forceinline__ double leaf_create(ChainCtx &C, Walker &W, const DevLeaf &L, int pos) {
  const int f = fidx(L, pos);
  const double u = C.rng.rand(C.K);
  if (L.kind == MCI_LEAF_CONTINUOUS) {
    int iy;
    double dx;
    W.data[f] = map_continuous(reinterpret_cast<const double2 *>(C.maps + L.smem_off), L.nbins, L.nbins_d, u, iy, dx);
    W.gidx[f] = iy;
    W.prob[f] = __ddiv_rn(1.0, __dmul_rn(L.nbins_d, dx));
    return __ddiv_rn(1.0, W.prob[f]);
  } else {
    const double *dist = C.maps + L.smem_off;
    const int g = locate_clamped(dist + L.nbins, L.nbins, u);
    W.data[f] = (double)(L.lower_i + g);
    W.gidx[f] = g;
    W.prob[f] = dist[g];
    return __ddiv_rn(1.0, dist[g]);
  }
}

/* remove! : sampler.jl:317-322, :37-41 */
__device__ __forceinline__ double leaf_remove(ChainCtx &C, Walker &W, const DevLeaf &L, int pos) {
  const int f = fidx(L, pos);
  if (L.kind == MCI_LEAF_CONTINUOUS) {
    const double2 g = reinterpret_cast<const double2 *>(C.maps + L.smem_off)[W.gidx[f]];
    return __ddiv_rn(1.0, __dmul_rn(g.y, L.nbins_d));
  } else {
    return (C.maps + L.smem_off)[W.gidx[f]];
  }
}

/* shift! : sampler.jl:336-386, :57-71.  k = leaf number inside its variable (rollback slot). */
__device__ __forceinline__ double leaf_shift(ChainCtx &C, Walker &W, const DevLeaf &L, int pos, int k) {
  const int f = fidx(L, pos);
  W.bk_data[k] = W.data[f];
  W.bk_gidx[k] = W.gidx[f];
  W.bk_prob[k] = W.prob[f];
  const int curr = W.gidx[f];
  const double u = C.rng.rand(C.K);
  double ratio;
  if (L.kind == MCI_LEAF_CONTINUOUS) {
    const double2 *pair = reinterpret_cast<const double2 *>(C.maps + L.smem_off);
    int iy;
    double dx;
    W.data[f] = map_continuous(pair, L.nbins, L.nbins_d, u, iy, dx);
    W.gidx[f] = iy;
    ratio = __ddiv_rn(pair[curr].y, dx);
  } else {
    const double *dist = C.maps + L.smem_off;
    const int g = locate_clamped(dist + L.nbins, L.nbins, u);
    W.data[f] = (double)(L.lower_i + g);
    W.gidx[f] = g;
    ratio = __ddiv_rn(dist[g], dist[curr]);
  }
  W.prob[f] = __dmul_rn(W.prob[f], ratio);
  return __ddiv_rn(1.0, ratio);
}

/* probability of (variable v, slot pos): the leaf's own for plain variables, the cached product for CompositeVar */
__device__ __forceinline__ double top_prob(const DevProblem &P, const Walker &W, int v, int pos) {
  if (P.top_nleaf[v] == 1) return W.prob[fidx(P.leaf[P.top_first[v]], pos)];
  return W.tprob[P.top_tslot[v] + pos - 1];
}

/* top-level dispatch incl. CompositeVar: sampler.jl:410-462 */
__device__ __forceinline__ double var_create(ChainCtx &C, Walker &W, int v, int pos) {
  const DevProblem &P = C.P;
  const int nl = P.top_nleaf[v];
  if (nl == 1) return leaf_create(C, W, P.leaf[P.top_first[v]], pos);
  double prop = 1.0, tp = 1.0;
  for (int k = 0; k < nl; ++k) {
    const DevLeaf &L = P.leaf[P.top_first[v] + k];
    prop = __dmul_rn(prop, leaf_create(C, W, L, pos));
    tp = __dmul_rn(tp, W.prob[fidx(L, pos)]);
  }
  W.tprob[P.top_tslot[v] + pos - 1] = tp;
  return prop;
}
__device__ __forceinline__ double var_remove(ChainCtx &C, Walker &W, int v, int pos) {
  const DevProblem &P = C.P;
  double prop = 1.0;
  const int nl = P.top_nleaf[v];
  if (nl == 1) return leaf_remove(C, W, P.leaf[P.top_first[v]], pos);
  for (int k = 0; k < nl; ++k) prop = __dmul_rn(prop, leaf_remove(C, W, P.leaf[P.top_first[v] + k], pos));
  return prop;
}
__device__ __forceinline__ double var_shift(ChainCtx &C, Walker &W, int v, int pos) {
  const DevProblem &P = C.P;
  const int nl = P.top_nleaf[v];
  if (nl == 1) return leaf_shift(C, W, P.leaf[P.top_first[v]], pos, 0);
  double prop = 1.0, tp = 1.0;
  W.bk_tprob = W.tprob[P.top_tslot[v] + pos - 1];
  for (int k = 0; k < nl; ++k) {
    const DevLeaf &L = P.leaf[P.top_first[v] + k];
    prop = __dmul_rn(prop, leaf_shift(C, W, L, pos, k));
    tp = __dmul_rn(tp, W.prob[fidx(L, pos)]);
  }
  W.tprob[P.top_tslot[v] + pos - 1] = tp;
  return prop;
}
__device__ __forceinline__ void var_shift_rollback(ChainCtx &C, Walker &W, int v, int pos) {
  const DevProblem &P = C.P;
  const int nl = P.top_nleaf[v];
  for (int k = 0; k < nl; ++k) {
    const int f = fidx(P.leaf[P.top_first[v] + k], pos);
    W.data[f] = W.bk_data[k];
    W.gidx[f] = W.bk_gidx[k];
    W.prob[f] = W.bk_prob[k];
  }
  if (nl > 1) W.tprob[P.top_tslot[v] + pos - 1] = W.bk_tprob;
}
/* swap! (its own inverse): sampler.jl:85-97, 395-408, 443-462 */
__device__ __forceinline__ void var_swap(ChainCtx &C, Walker &W, int v, int p1, int p2) {
  const DevProblem &P = C.P;
  const int nl = P.top_nleaf[v];
  if (nl > 1) {
    double t = W.tprob[P.top_tslot[v] + p1 - 1];
    W.tprob[P.top_tslot[v] + p1 - 1] = W.tprob[P.top_tslot[v] + p2 - 1];
    W.tprob[P.top_tslot[v] + p2 - 1] = t;
  }
  for (int k = 0; k < nl; ++k) {
    const DevLeaf &L = P.leaf[P.top_first[v] + k];
    const int a = fidx(L, p1), b = fidx(L, p2);
    double t = W.data[a]; W.data[a] = W.data[b]; W.data[b] = t;
    int g = W.gidx[a]; W.gidx[a] = W.gidx[b]; W.gidx[b] = g;
    t = W.prob[a]; W.prob[a] = W.prob[b]; W.prob[b] = t;
  }
}

/* initialize!: variable.jl:576-580 / :497-507 — leaf-major, slots 1..maxdof */
__device__ __forceinline__ void var_initialize(ChainCtx &C, Walker &W, int v) {
  const DevProblem &P = C.P;
  const int nl = P.top_nleaf[v];
  for (int k = 0; k < nl; ++k)
    for (int pos = 1; pos <= P.maxdof[v]; ++pos) leaf_create(C, W, P.leaf[P.top_first[v] + k], pos);
  if (nl > 1)
    for (int pos = 1; pos <= P.maxdof[v]; ++pos) {
      double tp = 1.0;
      for (int k = 0; k < nl; ++k) tp = __dmul_rn(tp, W.prob[fidx(P.leaf[P.top_first[v] + k], pos)]);
      W.tprob[P.top_tslot[v] + pos - 1] = tp;
    }
}

/* variable.jl:606-619 / :628-641 */
__device__ __forceinline__ double probability(const DevProblem &P, const Walker &W, int i) {
  double pr = 1.0;
  for (int v = 0; v < P.nv; ++v)
    for (int pos = 1; pos <= P.dof[i * P.nv + v]; ++pos) pr = __dmul_rn(pr, top_prob(P, W, v, pos));
  return pr;
}
__device__ __forceinline__ double padding_probability(const DevProblem &P, const Walker &W, int i) {
  double pr = 1.0;
  for (int v = 0; v < P.nv; ++v)
    for (int pos = P.dof[i * P.nv + v] + 1; pos <= P.maxdof[v]; ++pos) pr = __dmul_rn(pr, top_prob(P, W, v, pos));
  return pr;
}

/* accumulate! on every adaptive leaf of (variable v, slot pos); warp-collective */
__device__ __forceinline__ void var_accumulate(const DevProblem &P, const Walker &W, int hmode, double *hist, int v,
                                               int pos, double wgt, bool active) {
  for (int k = 0; k < P.top_nleaf[v]; ++k) {
    const DevLeaf &L = P.leaf[P.top_first[v] + k];
    if (L.adapt) hist_add_rt(hmode, hist + L.hist_off, W.gidx[fidx(L, pos)], wgt, active);
  }
}

/* CTA reduction of one per-thread value in a fixed order; returns the total in thread 0 */
template <int THREADS>
__device__ __forceinline__ double cta_sum(double v, double *sred) {
  constexpr int NWARP = THREADS / 32;
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x == 0)
    for (int wv = 0; wv < NWARP; ++wv) t += sred[wv];
  return t;
}

struct ChainSmem {
  double *maps_s, *shist, *sred;
  const double *maps;
  double *myhist;
};

template <int THREADS>
__device__ __forceinline__ ChainSmem chain_setup(const DevProblem &P, const ChainArgs &A, double *smem) {
  constexpr int NWARP = THREADS / 32;
  ChainSmem S;
  S.maps_s = smem;
  const int map_smem = A.maps_in_smem ? P.map_doubles : 0;
  S.shist = smem + map_smem;
  const int nsh = A.hist_mode == HIST_WARP_PRIVATE ? NWARP * P.hist_len : (A.hist_mode == HIST_SMEM_ATOMIC ? P.hist_len : 0);
  S.sred = S.shist + nsh;
  if (A.maps_in_smem) stage_maps(P, S.maps_s, A.gmaps);
  for (int i = threadIdx.x; i < nsh; i += THREADS) S.shist[i] = 0.0;
  __syncthreads();
  S.maps = A.maps_in_smem ? S.maps_s : A.gmaps;
  const int warp = threadIdx.x >> 5;
  S.myhist = A.hist_mode == HIST_WARP_PRIVATE ? S.shist + warp * P.hist_len
             : A.hist_mode == HIST_SMEM_ATOMIC ? S.shist
                                               : A.cta_hist + (size_t)(blockIdx.x % A.n_hist_copies) * P.hist_len;
  return S;
}

template <int THREADS>
__device__ __forceinline__ void chain_flush_hist(const DevProblem &P, const ChainArgs &A, const ChainSmem &S) {
  constexpr int NWARP = THREADS / 32;
  if (A.hist_mode == HIST_WARP_PRIVATE || A.hist_mode == HIST_SMEM_ATOMIC) {
    __syncthreads();
    double *out = A.cta_hist + (size_t)blockIdx.x * P.hist_len;
    for (int i = threadIdx.x; i < P.hist_len; i += THREADS) {
      double t = 0.0;
      if (A.hist_mode == HIST_WARP_PRIVATE) {
        for (int wv = 0; wv < NWARP; ++wv) t += S.shist[wv * P.hist_len + i];
      } else {
        t = S.shist[i];
      }
      out[i] = t;
    }
  }
}

#define PA_IDX(P, u, i, j) (((u) * (P).nd + (i)) * (P).pa_dim + (j))
constexpr int MAX_PA = 3 * MAX_ND * (MCI_MAX_TOPVAR > MAX_ND ? MCI_MAX_TOPVAR : MAX_ND);

/* ------------------------------------------------------------------------------------------ */
/* VegasMC                                                                                      */
/* ------------------------------------------------------------------------------------------ */
template <int FID, int THREADS>
__global__ void __launch_bounds__(THREADS) vegasmc_kernel(const __grid_constant__ DevProblem P, const ChainArgs A) {
  using F = Integrand<FID>;
  constexpr int NI = F::NI;
  extern __shared__ __align__(16) double smem[];
  ChainSmem S = chain_setup<THREADS>(P, A, smem);
  const int tid = threadIdx.x;
  const int lb = blockIdx.x / A.ctas_per_block;
  const int cb = blockIdx.x - lb * A.ctas_per_block;
  const uint32_t block = (uint32_t)(A.rank + lb * A.world);
  const int walker = cb * THREADS + tid;
  const bool active = walker < A.walkers_per_block;
  const int N = P.ni;
  const int hmode = A.hist_mode;

  PhiloxKey K;
  philox_expand_key(P.seed, A.iter, K);
  ChainCtx C(P, S.maps, K);
  C.rng.init((uint64_t)walker, block);
  Walker W;
  FlatX X{W.data, &P};

  double rw[MAX_ND];
  for (int i = 0; i <= N; ++i) rw[i] = A.reweight[i];

  double obs[MCI_MAX_OBS];
  for (int o = 0; o < P.obs_total; ++o) obs[o] = 0.0;
  double visited[MAX_ND];
  for (int i = 0; i <= N; ++i) visited[i] = 0.0;
  double normalization = 0.0;
  unsigned int n_prop[MCI_MAX_TOPVAR], n_acc[MCI_MAX_TOPVAR]; /* propose/accept[2, 1, vi] (updates.jl:90,92) */
  for (int v = 0; v < P.nv; ++v) { n_prop[v] = 0; n_acc[v] = 0; }
  unsigned long long neval = 0, bad = 0;

  double weights[NI], nweights[NI], pad[NI + 1], npad[NI + 1];
  /* montecarlo.jl:151-166 */
  for (int v = 0; v < P.nv; ++v) var_initialize(C, W, v);
  F::template eval<0>(X, P, nweights);
  for (int i = 0; i <= N; ++i) pad[i] = padding_probability(P, W, i);
  double prob_cur = __dmul_rn(rw[N], pad[N]);
  for (int i = 0; i < N; ++i) {
    weights[i] = nweights[i];
    prob_cur = __dadd_rn(prob_cur, __dmul_rn(__dmul_rn(fabs(nweights[i]), rw[i]), pad[i]));
  }
  const double meas_from = (double)A.steps / 100; /* ne >= neval/100 (montecarlo.jl:213) */

  for (uint64_t ne = 1; ne <= A.steps; ++ne) {
    /* changeVariable: updates.jl:45-106 */
    if (active) {
      const int vi = C.rng.randint(K, P.nv) - 1;
      if (!P.top_skip[vi] && P.maxdof[vi] > 0) {
        const int pos = C.rng.randint(K, P.maxdof[vi]);
        const double prop = var_shift(C, W, vi, pos);
        if (prop > 4.9406564584124654e-324) {
          F::template eval<0>(X, P, nweights);
          neval++;
          bool finite = true;
          for (int i = 0; i < N; ++i) finite = finite && isfinite(nweights[i]);
          if (!finite) bad++;
          for (int i = 0; i <= N; ++i) npad[i] = padding_probability(P, W, i);
          double prob_new = __dmul_rn(rw[N], npad[N]);
          for (int i = 0; i < N; ++i)
            prob_new = __dadd_rn(prob_new, __dmul_rn(__dmul_rn(fabs(nweights[i]), rw[i]), npad[i]));
          const double R = __ddiv_rn(__dmul_rn(prop, prob_new), prob_cur);
          n_prop[vi]++;
          if (C.rng.rand(K) < R) {
            n_acc[vi]++;
            for (int i = 0; i < N; ++i) weights[i] = nweights[i];
            for (int i = 0; i <= N; ++i) pad[i] = npad[i];
            prob_cur = prob_new;
          } else {
            var_shift_rollback(C, W, vi, pos);
          }
        }
      }
    }
    __syncwarp();
    /* histogram of every step: montecarlo.jl:198-211 */
    if (hmode != HIST_NONE) {
      for (int i = 0; i < N; ++i) {
        const double aw = fabs(weights[i]);
        const double f2 = __ddiv_rn(__dmul_rn(aw, aw), probability(P, W, i));
        const double wf2 = __ddiv_rn(__dmul_rn(f2, pad[i]), prob_cur);
        for (int v = 0; v < P.nv; ++v)
          for (int pos = 1; pos <= P.dof[i * P.nv + v]; ++pos) var_accumulate(P, W, hmode, S.myhist, v, pos, wf2, active);
      }
    }
    /* measurement: montecarlo.jl:213-232 */
    if (active && (ne % (uint64_t)P.measurefreq) == 0 && (double)ne >= meas_from) {
      for (int i = 0; i < N; ++i) {
        visited[i] += __ddiv_rn(fabs(__dmul_rn(__dmul_rn(weights[i], pad[i]), rw[i])), prob_cur);
        const double relw = __ddiv_rn(__dmul_rn(weights[i], pad[i]), prob_cur);
        if (P.measure_kind == MCI_MEASURE_DEFAULT) obs[P.obs_off[i]] += relw;
        else if (i == 0) obs[P.obs_off[0] + W.gidx[P.leaf[P.nleaf - 1].base]] += relw;
      }
      normalization += __ddiv_rn(pad[N], prob_cur);
      visited[N] += __ddiv_rn(__dmul_rn(rw[N], pad[N]), prob_cur);
    }
  }
  if (bad) atomicAdd(A.flags, bad);

  /* per-CTA statistics, fixed order */
  const int stat_len = P.obs_total + 1 + P.nd + 2 * P.pa_len_() + 1;
  double *out = A.cta_stats + (size_t)blockIdx.x * stat_len;
  for (int i = tid; i < stat_len; i += THREADS) out[i] = 0.0;
  __syncthreads();
  for (int o = 0; o < P.obs_total; ++o) { double t = cta_sum<THREADS>(obs[o], S.sred); if (tid == 0) out[o] = t; }
  { double t = cta_sum<THREADS>(normalization, S.sred); if (tid == 0) out[P.obs_total] = t; }
  for (int i = 0; i <= N; ++i) { double t = cta_sum<THREADS>(visited[i], S.sred); if (tid == 0) out[P.obs_total + 1 + i] = t; }
  double *prop_o = out + P.obs_total + 1 + P.nd, *acc_o = prop_o + P.pa_len_();
  for (int v = 0; v < P.nv; ++v) {
    double t = cta_sum<THREADS>((double)n_prop[v], S.sred); if (tid == 0) prop_o[PA_IDX(P, 1, 0, v)] = t;
    t = cta_sum<THREADS>((double)n_acc[v], S.sred); if (tid == 0) acc_o[PA_IDX(P, 1, 0, v)] = t;
  }
  { double t = cta_sum<THREADS>((double)neval, S.sred); if (tid == 0) out[stat_len - 1] = t; }
  chain_flush_hist<THREADS>(P, A, S);
}

/* ------------------------------------------------------------------------------------------ */
/* MCMC                                                                                         */
/* ------------------------------------------------------------------------------------------ */
template <int FID>
__device__ __forceinline__ double eval_one(const DevProblem &P, const FlatX &X, int idx) {
  double w[Integrand<FID>::NI];
  Integrand<FID>::template eval<0>(X, P, w);
  double r = w[0];
#pragma unroll
  for (int i = 1; i < Integrand<FID>::NI; ++i) if (i == idx) r = w[i];
  return r;
}

template <int FID, int THREADS>
__global__ void __launch_bounds__(THREADS) mcmc_kernel(const __grid_constant__ DevProblem P, const ChainArgs A) {
  extern __shared__ __align__(16) double smem[];
  ChainSmem S = chain_setup<THREADS>(P, A, smem);
  /* propose/accept counters of the CTA: [2][pa_len] u32 in shared memory after the reduction scratch */
  unsigned int *s_prop = reinterpret_cast<unsigned int *>(S.sred + THREADS / 32);
  const int pa_len = P.pa_len_();
  unsigned int *s_acc = s_prop + pa_len;
  const int tid = threadIdx.x;
  for (int i = tid; i < 2 * pa_len; i += THREADS) s_prop[i] = 0u;
  __syncthreads();
  const int lb = blockIdx.x / A.ctas_per_block;
  const int cb = blockIdx.x - lb * A.ctas_per_block;
  const uint32_t block = (uint32_t)(A.rank + lb * A.world);
  const int walker = cb * THREADS + tid;
  const bool active = walker < A.walkers_per_block;
  const int N = P.ni;
  const int hmode = A.hist_mode;
  const double TINY = 4.9406564584124654e-324 * 1e50;
  const double EPS0 = 4.9406564584124654e-324;

  PhiloxKey K;
  philox_expand_key(P.seed, A.iter, K);
  ChainCtx C(P, S.maps, K);
  C.rng.init((uint64_t)walker, block);
  Walker W;
  FlatX X{W.data, &P};
  double rw[MAX_ND];
  for (int i = 0; i <= N; ++i) rw[i] = A.reweight[i];

  double obs[MCI_MAX_OBS];
  for (int o = 0; o < P.obs_total; ++o) obs[o] = 0.0;
  unsigned int visited[MAX_ND];
  for (int i = 0; i <= N; ++i) visited[i] = 0u;
  double normalization = 0.0;
  unsigned long long neval = 0, bad = 0;

  /* start point: montecarlo.jl:99-113, initialize! :190-205 */
  int curr = 0;
  double weight = 0.0, prob_cur = 1.0;
  if (active) {
    for (int t = 0; t < 10000; ++t) {
      for (int v = 0; v < P.nv; ++v) var_initialize(C, W, v);
      weight = eval_one<FID>(P, X, curr);
      prob_cur = __dmul_rn(fabs(weight), rw[curr]);
      if (prob_cur > TINY) break;
    }
    if (prob_cur == 0.0) atomicAdd(A.flags + 1, 1ull);
  }
  const int nupdates = 2 * P.nv + 2; /* [changeIntegrand, swapVariable, changeVariable x 2Nv] :117-120 */
  const uint64_t nburnin = (uint64_t)floor(__dmul_rn((double)A.steps, P.thermal_ratio));
  const uint64_t total = A.steps + nburnin;

  for (uint64_t it = 1; it <= total; ++it) {
    if (active) {
      visited[curr] += 1u;
      const int u = C.rng.randint(K, nupdates);
      if (u == 1) {
        /* changeIntegrand: updates.jl:1-70 */
        const int nnb = P.nb_ptr[curr + 1] - P.nb_ptr[curr];
        const int nw = P.nb_idx[P.nb_ptr[curr] + C.rng.randint(K, nnb) - 1];
        if (nw != curr) {
          double prop = __ddiv_rn((double)nnb, (double)(P.nb_ptr[nw + 1] - P.nb_ptr[nw]));
          for (int v = 0; v < P.nv; ++v) {
            const int cd = P.dof[curr * P.nv + v], nd_ = P.dof[nw * P.nv + v];
            if (cd < nd_) { for (int pos = cd + 1; pos <= nd_; ++pos) prop = __dmul_rn(prop, var_create(C, W, v, pos)); }
            else if (cd > nd_) { for (int pos = nd_ + 1; pos <= cd; ++pos) prop = __dmul_rn(prop, var_remove(C, W, v, pos)); }
          }
          if (prop > EPS0) {
            const double nwgt = (nw == N) ? 0.0 : eval_one<FID>(P, X, nw);
            neval++;
            if (!isfinite(nwgt)) bad++;
            const double nprob = (nw == N) ? rw[nw] : __dmul_rn(fabs(nwgt), rw[nw]);
            const double R = __ddiv_rn(__dmul_rn(prop, nprob), prob_cur);
            atomicAdd(&s_prop[PA_IDX(P, 0, curr, nw)], 1u);
            if (C.rng.rand(K) < R) {
              atomicAdd(&s_acc[PA_IDX(P, 0, curr, nw)], 1u);
              curr = nw; weight = nwgt; prob_cur = nprob;
            }
          }
        }
      } else if (curr != N) {
        if (u == 2) {
          /* swapVariable: updates.jl:111-147 */
          const int vi = C.rng.randint(K, P.nv) - 1;
          const int cd = P.dof[curr * P.nv + vi];
          if (cd > 0) {
            const int p1 = C.rng.randint(K, cd);
            const int p2 = C.rng.randint(K, cd);
            if (p1 != p2) {
              var_swap(C, W, vi, p1, p2);
              const double nwgt = eval_one<FID>(P, X, curr);
              neval++;
              if (!isfinite(nwgt)) bad++;
              const double nprob = __dmul_rn(fabs(nwgt), rw[curr]);
              const double R = __ddiv_rn(__dmul_rn(1.0, nprob), prob_cur);
              atomicAdd(&s_prop[PA_IDX(P, 2, curr, vi)], 1u);
              if (C.rng.rand(K) < R) { atomicAdd(&s_acc[PA_IDX(P, 2, curr, vi)], 1u); weight = nwgt; prob_cur = nprob; }
              else var_swap(C, W, vi, p1, p2);
            }
          }
        } else {
          /* changeVariable: updates.jl:72-109 */
          const int vi = C.rng.randint(K, P.nv) - 1;
          const int cd = P.dof[curr * P.nv + vi];
          if (!P.top_skip[vi] && cd > 0) {
            const int pos = C.rng.randint(K, cd);
            const double prop = var_shift(C, W, vi, pos);
            if (prop > EPS0) {
              const double nwgt = eval_one<FID>(P, X, curr);
              neval++;
              if (!isfinite(nwgt)) bad++;
              const double nprob = __dmul_rn(fabs(nwgt), rw[curr]);
              const double R = __ddiv_rn(__dmul_rn(prop, nprob), prob_cur);
              atomicAdd(&s_prop[PA_IDX(P, 1, curr, vi)], 1u);
              if (C.rng.rand(K) < R) { atomicAdd(&s_acc[PA_IDX(P, 1, curr, vi)], 1u); weight = nwgt; prob_cur = nprob; }
              else var_shift_rollback(C, W, vi, pos);
            }
          }
        }
      }
    }
    __syncwarp();
    /* measurement: montecarlo.jl:145-170 */
    if ((it % (uint64_t)P.measurefreq) == 0 && it >= nburnin) {
      if (hmode != HIST_NONE) {
        /* every lane walks the slots of ITS integrand; lanes on shorter dofs idle (warp-collective scatter) */
        for (int v = 0; v < P.nv; ++v)
          for (int pos = 1; pos <= P.maxdof[v]; ++pos) {
            const bool on = active && curr != N && pos <= P.dof[curr * P.nv + v];
            var_accumulate(P, W, hmode, S.myhist, v, pos, 1.0, on);
          }
      }
      if (active) {
        if (curr == N) normalization += __ddiv_rn(1.0, rw[N]);
        else {
          const double relw = __ddiv_rn(weight, prob_cur);
          if (P.measure_kind == MCI_MEASURE_DEFAULT) obs[P.obs_off[curr]] += relw;
          else obs[P.obs_off[0] + W.gidx[P.leaf[P.nleaf - 1].base]] += relw;
        }
      }
    }
  }
  if (bad) atomicAdd(A.flags, bad);

  const int stat_len = P.obs_total + 1 + P.nd + 2 * pa_len + 1;
  double *out = A.cta_stats + (size_t)blockIdx.x * stat_len;
  __syncthreads();
  for (int o = 0; o < P.obs_total; ++o) { double t = cta_sum<THREADS>(obs[o], S.sred); if (tid == 0) out[o] = t; }
  { double t = cta_sum<THREADS>(normalization, S.sred); if (tid == 0) out[P.obs_total] = t; }
  for (int i = 0; i <= N; ++i) { double t = cta_sum<THREADS>((double)visited[i], S.sred); if (tid == 0) out[P.obs_total + 1 + i] = t; }
  __syncthreads();
  double *prop_o = out + P.obs_total + 1 + P.nd, *acc_o = prop_o + pa_len;
  for (int i = tid; i < pa_len; i += THREADS) { prop_o[i] = (double)s_prop[i]; acc_o[i] = (double)s_acc[i]; }
  { double t = cta_sum<THREADS>((double)neval, S.sred); if (tid == 0) out[stat_len - 1] = t; }
  chain_flush_hist<THREADS>(P, A, S);
}

}  // namespace mci
