/*
 * mci_chain.cuh — the Markov-chain solvers: one walker (chain) per thread.
 *
 *   VegasMC.montecarlo  src/vegas_mc/montecarlo.jl:112-241 + changeVariable src/vegas_mc/updates.jl:45-106
 *   MCMC.montecarlo     src/mcmc/montecarlo.jl:72-205 + changeIntegrand / changeVariable / swapVariable
 *                       src/mcmc/updates.jl:1-147
 *   pool operations     create!/remove!/shift!/swap! and rollbacks, src/distribution/sampler.jl:13-97, 293-462
 *   probabilities       src/distribution/variable.jl:587-641
 *
 * The reference runs ONE chain of neval_per_block steps per block.  A GPU needs ~10^5 concurrent chains, so a
 * block is `walkers_per_block` independent chains of neval_per_block/walkers steps (recorded in mci_problem_t;
 * walkers = 1 is the reference's decomposition).  Walker w of block b draws sequentially from the Philox stream
 * (seed, iter, b, w).  Per-step arithmetic follows the reference operation by operation (stateful shift! ratios,
 * rollback caches, padding products).
 *
 * Walker state lives in SHARED memory as rows [row][thread]: data (x), gidx (bin) and — VegasMC only — the stateful
 * prob of every (leaf, slot), plus one backup row per leaf of a variable for the rollback of shift!.  A thread only
 * ever touches column `tid`, and the row stride is a multiple of 16 doubles, so every access is bank-conflict free
 * whatever slot a lane picks at random (round 1's first cut kept 64-entry arrays per thread in local memory:
 * ncu long_scoreboard 4.9-5.7 warps per issue, 12 % occupancy).  A CompositeVar's probability is the product of its
 * leaves' probabilities in leaf order, recomputed on demand (the reference caches the same product).
 *
 * MCMC evaluates ONE integrand per proposal whatever the update type, so the step is written as
 * propose (divergent, cheap) -> evaluate (convergent, the expensive part) -> accept/rollback (divergent, cheap):
 * the three update types share one integrand call site instead of serialising three copies of it.
 *
 * Statistics: observable / normalisation / visited / propose / accept are accumulated per thread and reduced per CTA
 * in a fixed order; all walkers of a CTA belong to one block.
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
  int rows;               /* walker rows in shared memory = n_draw (+ the leaves of the widest variable when some
                             variable has several: the rollback rows of shift!)                    */
  long long burnin;       /* extra unmeasured steps before the reference's loop (mci_problem_t.chain_burnin) */
  int coop_ci;            /* mcmc: changeIntegrand creates/removes enough slots for the warp-cooperative path */
  PhiloxKey key;          /* the ten round keys of (seed, iter), expanded on the host: operands from the constant bank */
};

/* a walker's columns of the shared state; T = row stride (threads per CTA) */
/* bin indices are kept as 16-bit values (mci_create refuses chain problems with more than 65535 bins in a leaf; the
 * reference's default is 999): 10 bytes per (slot, walker) instead of 12, which for BASELINE C5 (20 slots) is the
 * difference between 512 and 1024 resident walkers per SM */
typedef unsigned short gidx_t;

template <bool HAS_PROB>
struct Walker {
  double *data;
  double *prob; /* HAS_PROB only */
  gidx_t *gidx; /* 0-based bin: grid increment (Continuous) / value - lower (Discrete) */
  int T;
  int bk0;      /* first backup row (problems with a multi-leaf variable only) */
  /* rollback cache of shift! (sampler.jl:338-340): in registers when every variable is a single leaf — the common case,
   * and one shared-memory row per walker less — else one backup row per leaf of the variable */
  bool bk_regs;
  double bk_x, bk_p;
  int bk_g;
};

template <class RNG, bool FK = false>
struct ChainCtxT {
  static constexpr bool fermik = FK; /* the problem may hold FermiK variables (mcmc only): compile their proposals in */
  const Problem &P;
  const double *maps;
  const PhiloxKey &K;
  RNG rng;
  __device__ __forceinline__ ChainCtxT(const Problem &p, const double *m, const PhiloxKey &k) : P(p), maps(m), K(k) {}
};

/* integrand accessor on the shared walker rows: X(leaf, pos), pos 1-based like the reference's x[pos] */
struct ChainX {
  const double *data;
  const Problem *P;
  int T;
  __device__ __forceinline__ double operator()(int leaf, int pos) const {
    return data[(P->leaf[leaf].base + (pos - 1) * P->leaf[leaf].stride) * T];
  }
};

__device__ __forceinline__ int frow(const DevLeaf &L, int pos) { return L.base + (pos - 1) * L.stride; }

/* create! : sampler.jl:293-305 (Continuous), :13-22 (Discrete).  Returns 1/prob. */
template <bool HP, class CTX>
__device__ __forceinline__ double leaf_create(CTX &C, Walker<HP> &W, const DevLeaf &L, int pos) {
  const int f = frow(L, pos) * W.T;
  const double u = C.rng.rand(C.K);
  if (L.kind == MCI_LEAF_CONTINUOUS) {
    int iy;
    double dx;
    W.data[f] = map_continuous(reinterpret_cast<const double2 *>(C.maps + L.smem_off), L.nbins, L.nbins_d, u, iy, dx);
    W.gidx[f] = (gidx_t)iy;
    if (!HP) return __dmul_rn(L.nbins_d, dx); /* 1/prob = N dx (the reference's 1/(1/(N dx)) to 1 ulp) */
    const double pr = __ddiv_rn(1.0, __dmul_rn(L.nbins_d, dx));
    W.prob[f] = pr;
    return __ddiv_rn(1.0, pr);
  } else {
    const double *dist = C.maps + L.smem_off;
    const int g = locate_clamped(dist + L.nbins, L.nbins, u);
    W.data[f] = (double)(L.lower_i + g);
    W.gidx[f] = (gidx_t)g;
    if (HP) W.prob[f] = dist[g];
    return __ddiv_rn(1.0, dist[g]);
  }
}

/* remove! : sampler.jl:317-322, :37-41 */
template <bool HP, class CTX>
__device__ __forceinline__ double leaf_remove(CTX &C, Walker<HP> &W, const DevLeaf &L, int pos) {
  const int f = frow(L, pos) * W.T;
  if (L.kind == MCI_LEAF_CONTINUOUS) {
    const double2 g = reinterpret_cast<const double2 *>(C.maps + L.smem_off)[W.gidx[f]];
    return __ddiv_rn(1.0, __dmul_rn(g.y, L.nbins_d));
  } else {
    return (C.maps + L.smem_off)[W.gidx[f]];
  }
}

/* shift! : sampler.jl:336-386, :57-71.  k = leaf number inside its variable (backup row). */
template <bool HP, class CTX>
__device__ __forceinline__ double leaf_shift(CTX &C, Walker<HP> &W, const DevLeaf &L, int pos, int k) {
  const int f = frow(L, pos) * W.T;
  const int curr = W.gidx[f];
  if (W.bk_regs) {
    W.bk_x = W.data[f]; W.bk_g = curr;
    if (HP) W.bk_p = W.prob[f];
  } else {
    const int b = (W.bk0 + k) * W.T;
    W.data[b] = W.data[f];
    W.gidx[b] = (gidx_t)curr;
    if (HP) W.prob[b] = W.prob[f];
  }
  const double u = C.rng.rand(C.K);
  double ratio;
  if (L.kind == MCI_LEAF_CONTINUOUS) {
    const double2 *pair = reinterpret_cast<const double2 *>(C.maps + L.smem_off);
    int iy;
    double dx;
    W.data[f] = map_continuous(pair, L.nbins, L.nbins_d, u, iy, dx);
    W.gidx[f] = (gidx_t)iy;
    ratio = __ddiv_rn(pair[curr].y, dx);
  } else {
    const double *dist = C.maps + L.smem_off;
    const int g = locate_clamped(dist + L.nbins, L.nbins, u);
    W.data[f] = (double)(L.lower_i + g);
    W.gidx[f] = (gidx_t)g;
    ratio = __ddiv_rn(dist[g], dist[curr]);
  }
  if (HP) W.prob[f] = __dmul_rn(W.prob[f], ratio);
  return __ddiv_rn(1.0, ratio);
}

/* ---- FermiK{D} (variable.jl:1-21, sampler.jl:109-281): a momentum kept near the Fermi surface.  The D = 2 or 3
 * leaves of the variable are its Cartesian components, rows f0, f0 + T[, f0 + 2T] of the walker.  The reference keeps
 * a `prob` per slot that only :vegas/:vegasmc would read (and shift! corrupts it, sampler.jl:209), so :mcmc — the one
 * solver it works under — needs none.  Arithmetic in the reference's association order, never contracted. ---- */
#define MCI_2PI 6.283185307179586476925286766559
template <bool FK>
__device__ __forceinline__ bool is_fermik(const Problem &P, int v) {
  if constexpr (FK) return P.leaf[P.top_first[v]].kind == MCI_LEAF_FERMIK;
  else return false;
}
__device__ __forceinline__ double fermik_shell(double dk) { return __dmul_rn(__dmul_rn(2.0, dk), MCI_2PI); } /* 2 dk 2pi */

/* create!: sampler.jl:109-149.  |K| uniform in [kF - dk, kF + dk), direction uniform in (phi, theta); returns the
 * proposal volume 2 dk 2pi pi sin(theta) |K|^2 (D = 3) or 2 dk 2pi |K| (D = 2); 0 when |K| <= 0 */
template <class CTX>
__device__ __forceinline__ double fermik_create(CTX &C, Walker<false> &W, int v, int pos) {
  const Problem &P = C.P;
  const DevLeaf &L = P.leaf[P.top_first[v]];
  const int D = P.top_nleaf[v], T = W.T, f = frow(L, pos) * T;
  const double Kamp = __dadd_rn(L.p0, __dmul_rn(__dmul_rn(__dsub_rn(C.rng.rand(C.K), 0.5), 2.0), L.p1));
  if (Kamp <= 0.0) return 0.0;
  double sp, cp;
  sincos(__dmul_rn(MCI_2PI, C.rng.rand(C.K)), &sp, &cp);
  if (D == 3) {
    double st, ct;
    sincos(__dmul_rn(MCI_PI, C.rng.rand(C.K)), &st, &ct);
    W.data[f] = __dmul_rn(__dmul_rn(Kamp, cp), st);
    W.data[f + T] = __dmul_rn(__dmul_rn(Kamp, sp), st);
    W.data[f + 2 * T] = __dmul_rn(Kamp, ct);
    return __dmul_rn(__dmul_rn(fermik_shell(L.p1), MCI_PI), __dmul_rn(st, __dmul_rn(Kamp, Kamp)));
  }
  W.data[f] = __dmul_rn(Kamp, cp);
  W.data[f + T] = __dmul_rn(Kamp, sp);
  return __dmul_rn(fermik_shell(L.p1), Kamp);
}

/* remove!: sampler.jl:160-189.  0 unless kF - dk < |K| < kF + dk (and sin(theta) >= 1e-15), else 1 / the volume */
template <class CTX>
__device__ __forceinline__ double fermik_remove(CTX &C, Walker<false> &W, int v, int pos) {
  const Problem &P = C.P;
  const DevLeaf &L = P.leaf[P.top_first[v]];
  const int D = P.top_nleaf[v], T = W.T, f = frow(L, pos) * T;
  const double k1 = W.data[f], k2 = W.data[f + T];
  const double r2 = __dadd_rn(__dmul_rn(k1, k1), __dmul_rn(k2, k2));
  if (D == 3) {
    const double k3 = W.data[f + 2 * T];
    const double Kamp = sqrt(__dadd_rn(r2, __dmul_rn(k3, k3)));
    if (!(__dsub_rn(L.p0, L.p1) < Kamp && Kamp < __dadd_rn(L.p0, L.p1))) return 0.0;
    const double st = __ddiv_rn(sqrt(r2), Kamp);
    if (st < 1.0e-15) return 0.0;
    return __ddiv_rn(1.0, __dmul_rn(__dmul_rn(__dmul_rn(fermik_shell(L.p1), MCI_PI), st), __dmul_rn(Kamp, Kamp)));
  }
  const double Kamp = sqrt(r2);
  if (!(__dsub_rn(L.p0, L.p1) < Kamp && Kamp < __dadd_rn(L.p0, L.p1))) return 0.0;
  return __ddiv_rn(1.0, __dmul_rn(fermik_shell(L.p1), Kamp));
}

/* shift!: sampler.jl:199-251, one of three moves with probability 1/3 each — rescale |K| by a factor uniform in
 * [1/1.5, 1.5) (returns the factor for D = 3, 1 for D = 2), redraw the direction uniformly on the sphere / circle, or
 * displace every component by up to dk/2.  The old K goes to the backup rows (K[end] = K[idx]). */
template <class CTX>
__device__ __forceinline__ double fermik_shift(CTX &C, Walker<false> &W, int v, int pos) {
  const Problem &P = C.P;
  const DevLeaf &L = P.leaf[P.top_first[v]];
  const int D = P.top_nleaf[v], T = W.T, f = frow(L, pos) * T, b = W.bk0 * T;
  for (int c = 0; c < D; ++c) { W.data[b + c * T] = W.data[f + c * T]; W.gidx[b + c * T] = 0; }
  const double x = C.rng.rand(C.K);
  if (x < 1.0 / 3) {
    const double lam = 1.5, inv = 1.0 / 1.5;
    const double ratio = __dadd_rn(inv, __dmul_rn(C.rng.rand(C.K), __dsub_rn(lam, inv)));
    for (int c = 0; c < D; ++c) W.data[f + c * T] = __dmul_rn(W.data[f + c * T], ratio);
    return D == 2 ? 1.0 : ratio;
  } else if (x < 2.0 / 3) {
    double sp, cp;
    sincos(__dmul_rn(C.rng.rand(C.K), MCI_2PI), &sp, &cp);
    const double k1 = W.data[f], k2 = W.data[f + T];
    const double r2 = __dadd_rn(__dmul_rn(k1, k1), __dmul_rn(k2, k2));
    if (D == 3) {
      double st, ct;
      sincos(acos(__dsub_rn(1.0, __dmul_rn(2.0, C.rng.rand(C.K)))), &st, &ct);
      const double k3 = W.data[f + 2 * T];
      const double Kamp = sqrt(__dadd_rn(r2, __dmul_rn(k3, k3)));
      W.data[f] = __dmul_rn(__dmul_rn(Kamp, cp), st);
      W.data[f + T] = __dmul_rn(__dmul_rn(Kamp, sp), st);
      W.data[f + 2 * T] = __dmul_rn(Kamp, ct);
    } else {
      const double Kamp = sqrt(r2);
      W.data[f] = __dmul_rn(Kamp, cp);
      W.data[f + T] = __dmul_rn(Kamp, sp);
    }
    return 1.0;
  }
  for (int c = 0; c < D; ++c)
    W.data[f + c * T] = __dadd_rn(W.data[f + c * T], __dmul_rn(__dsub_rn(C.rng.rand(C.K), 0.5), L.p1));
  return 1.0;
}

/* the constructor's fill (variable.jl:13): every component kF / sqrt(D) */
__device__ __forceinline__ void fermik_reset(const Problem &P, Walker<false> &W) {
  for (int v = 0; v < P.nv; ++v)
    if (P.leaf[P.top_first[v]].kind == MCI_LEAF_FERMIK) {
      const int D = P.top_nleaf[v];
      const double k0 = __ddiv_rn(P.leaf[P.top_first[v]].p0, sqrt((double)D));
      for (int r = P.top_base[v]; r < P.top_base[v] + D * P.maxdof[v]; ++r) { W.data[r * W.T] = k0; W.gidx[r * W.T] = 0; }
    }
}

/* probability of (variable v, slot pos): the leaf's own for plain variables, the product over the leaves in leaf
 * order for a CompositeVar (sampler.jl:431-440 keeps exactly this product in its cache) */
__device__ __forceinline__ double top_prob(const Problem &P, const Walker<true> &W, int v, int pos) {
  const int nl = P.top_nleaf[v];
  if (nl == 1) return W.prob[frow(P.leaf[P.top_first[v]], pos) * W.T];
  double tp = 1.0;
  MCI_UNROLL_LAYOUT
  for (int k = 0; k < nl; ++k) tp = __dmul_rn(tp, W.prob[frow(P.leaf[P.top_first[v] + k], pos) * W.T]);
  return tp;
}

/* top-level dispatch incl. CompositeVar: sampler.jl:410-462 */
template <bool HP, class CTX>
__device__ __forceinline__ double var_create(CTX &C, Walker<HP> &W, int v, int pos) {
  const Problem &P = C.P;
  const int nl = P.top_nleaf[v];
  if constexpr (CTX::fermik && !HP) { if (is_fermik<true>(P, v)) return fermik_create(C, W, v, pos); }
  if (nl == 1) return leaf_create(C, W, P.leaf[P.top_first[v]], pos);
  double prop = 1.0;
  for (int k = 0; k < nl; ++k) prop = __dmul_rn(prop, leaf_create(C, W, P.leaf[P.top_first[v] + k], pos));
  return prop;
}
template <bool HP, class CTX>
__device__ __forceinline__ double var_remove(CTX &C, Walker<HP> &W, int v, int pos) {
  const Problem &P = C.P;
  const int nl = P.top_nleaf[v];
  if constexpr (CTX::fermik && !HP) { if (is_fermik<true>(P, v)) return fermik_remove(C, W, v, pos); }
  if (nl == 1) return leaf_remove(C, W, P.leaf[P.top_first[v]], pos);
  double prop = 1.0;
  for (int k = 0; k < nl; ++k) prop = __dmul_rn(prop, leaf_remove(C, W, P.leaf[P.top_first[v] + k], pos));
  return prop;
}
template <bool HP, class CTX>
__device__ __forceinline__ double var_shift(CTX &C, Walker<HP> &W, int v, int pos) {
  const Problem &P = C.P;
  const int nl = P.top_nleaf[v];
  if constexpr (CTX::fermik && !HP) { if (is_fermik<true>(P, v)) return fermik_shift(C, W, v, pos); }
  if (nl == 1) return leaf_shift(C, W, P.leaf[P.top_first[v]], pos, 0);
  double prop = 1.0;
  for (int k = 0; k < nl; ++k) prop = __dmul_rn(prop, leaf_shift(C, W, P.leaf[P.top_first[v] + k], pos, k));
  return prop;
}
template <bool HP>
__device__ __forceinline__ void var_shift_rollback(const Problem &P, Walker<HP> &W, int v, int pos) {
  const int nl = P.top_nleaf[v];
  if (W.bk_regs) { /* single-leaf variables only */
    const int f = frow(P.leaf[P.top_first[v]], pos) * W.T;
    W.data[f] = W.bk_x; W.gidx[f] = (gidx_t)W.bk_g;
    if (HP) W.prob[f] = W.bk_p;
    return;
  }
  for (int k = 0; k < nl; ++k) {
    const int f = frow(P.leaf[P.top_first[v] + k], pos) * W.T;
    const int b = (W.bk0 + k) * W.T;
    W.data[f] = W.data[b];
    W.gidx[f] = W.gidx[b];
    if (HP) W.prob[f] = W.prob[b];
  }
}
/* swap! (its own inverse): sampler.jl:85-97, 395-408, 443-462 */
template <bool HP>
__device__ __forceinline__ void var_swap(const Problem &P, Walker<HP> &W, int v, int p1, int p2) {
  const int nl = P.top_nleaf[v];
  for (int k = 0; k < nl; ++k) {
    const DevLeaf &L = P.leaf[P.top_first[v] + k];
    const int a = frow(L, p1) * W.T, b = frow(L, p2) * W.T;
    double t = W.data[a]; W.data[a] = W.data[b]; W.data[b] = t;
    gidx_t g = W.gidx[a]; W.gidx[a] = W.gidx[b]; W.gidx[b] = g;
    if (HP) { t = W.prob[a]; W.prob[a] = W.prob[b]; W.prob[b] = t; }
  }
}

/* initialize!: variable.jl:576-580 / :497-507 — leaf-major, slots 1..maxdof */
template <bool HP, class CTX>
__device__ __forceinline__ void var_initialize(CTX &C, Walker<HP> &W, int v) {
  const Problem &P = C.P;
  const int nl = P.top_nleaf[v];
  if constexpr (CTX::fermik && !HP) {
    if (is_fermik<true>(P, v)) {
      for (int pos = 1; pos <= P.maxdof[v]; ++pos) fermik_create(C, W, v, pos);
      return;
    }
  }
  for (int k = 0; k < nl; ++k)
    for (int pos = 1; pos <= P.maxdof[v]; ++pos) leaf_create(C, W, P.leaf[P.top_first[v] + k], pos);
}

/* variable.jl:606-619 / :628-641 */
__device__ __forceinline__ double probability(const Problem &P, const Walker<true> &W, int i) {
  double pr = 1.0;
  MCI_UNROLL_LAYOUT
  for (int v = 0; v < P.nv; ++v) {
    MCI_UNROLL_LAYOUT
    for (int pos = 1; pos <= P.dof[i * P.nv + v]; ++pos) pr = __dmul_rn(pr, top_prob(P, W, v, pos));
  }
  return pr;
}
__device__ __forceinline__ double padding_probability(const Problem &P, const Walker<true> &W, int i) {
  double pr = 1.0;
  MCI_UNROLL_LAYOUT
  for (int v = 0; v < P.nv; ++v) {
    MCI_UNROLL_LAYOUT
    for (int pos = P.dof[i * P.nv + v] + 1; pos <= P.maxdof[v]; ++pos) pr = __dmul_rn(pr, top_prob(P, W, v, pos));
  }
  return pr;
}

/* hist[bin] += v on the CTA's shared histogram (32-bit shared address) or on an L2-resident global partial */
__device__ __forceinline__ void chain_hist_add(int hmode, uint32_t hist_saddr, double *ghist, int bin, double v) {
  if (hmode == HIST_SMEM_ATOMIC) asm volatile("red.shared.add.f64 [%0], %1;" ::"r"(hist_saddr + 8u * (uint32_t)bin), "d"(v) : "memory");
  else atomicAdd(ghist + bin, v);
}

/* CTA reduction of one per-thread value in a fixed order; returns the total in thread 0 */
__device__ __forceinline__ double cta_sum(double v, double *sred) {
  const int nwarp = blockDim.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x == 0)
    for (int wv = 0; wv < nwarp; ++wv) t += sred[wv];
  return t;
}

/* shared-memory carve-up of the chain kernels:
 *   [maps (if they fit)] [histogram (HIST_SMEM_ATOMIC)] [reduction scratch: 32 doubles] [mcmc: 2*pa_len u32 counters]
 *   [walker data rows*T doubles] [walker prob rows*T doubles (VegasMC)] [walker gidx rows*T u16]                   */
struct ChainSmem {
  const double *maps;
  double *shist, *sred;
  unsigned int *scount;
  uint32_t hist_saddr;
  double *ghist;
};

template <bool HP>
__device__ __forceinline__ ChainSmem chain_setup(const Problem &P, const ChainArgs &A, double *smem, int n_count, Walker<HP> &W) {
  const int T = blockDim.x, tid = threadIdx.x;
  ChainSmem S;
  double *maps_s = smem;
  const int map_smem = A.maps_in_smem ? P.map_doubles : 0;
  S.shist = smem + map_smem;
  const int nsh = A.hist_mode == HIST_SMEM_ATOMIC ? P.hist_len : 0;
  S.sred = S.shist + nsh;
  S.scount = reinterpret_cast<unsigned int *>(S.sred + 32);
  double *wbase = S.sred + 32 + ((n_count + 1) >> 1);
  W.T = T;
  W.bk0 = P.n_draw;
  W.bk_regs = A.rows == P.n_draw; /* no backup rows were reserved: every variable is a single leaf */
  W.bk_x = W.bk_p = 0.0; W.bk_g = 0;
  W.data = wbase + tid;
  W.prob = HP ? wbase + (size_t)A.rows * T + tid : nullptr;
  W.gidx = reinterpret_cast<gidx_t *>(wbase + (size_t)(HP ? 2 : 1) * A.rows * T) + tid;
  if (A.maps_in_smem) stage_maps(P, maps_s, A.gmaps);
  for (int i = tid; i < nsh; i += T) S.shist[i] = 0.0;
  for (int i = tid; i < n_count; i += T) S.scount[i] = 0u;
  __syncthreads();
  S.maps = A.maps_in_smem ? maps_s : A.gmaps;
  S.hist_saddr = (uint32_t)__cvta_generic_to_shared(S.shist);
  S.ghist = A.cta_hist + (size_t)(blockIdx.x % A.n_hist_copies) * P.hist_len;
  return S;
}

/* COUNTS: the shared histogram holds u32 counters (MCMC: every contribution is exactly 1.0, variable.jl:196-200 with
 * weight = 1.0 from mcmc/montecarlo.jl:148-155, so a native ATOMS.ADD counter is the same number as the fp64 sum, bit
 * for bit, without the LDS/DADD/CAS loop and its ~21 shared-memory wavefronts per update) */
template <bool COUNTS>
__device__ __forceinline__ void chain_flush_hist(const Problem &P, const ChainArgs &A, const ChainSmem &S) {
  if (A.hist_mode == HIST_SMEM_ATOMIC) {
    __syncthreads();
    double *out = A.cta_hist + (size_t)blockIdx.x * P.hist_len;
    const unsigned int *cnt = reinterpret_cast<const unsigned int *>(S.shist);
    for (int i = threadIdx.x; i < P.hist_len; i += blockDim.x) out[i] = COUNTS ? (double)cnt[i] : S.shist[i];
  }
}

/* largest CTA the chain kernels are compiled for: 1024 threads cap the kernel at 64 registers.  (Measured on the
 * bubble under mcmc: a 512-thread / 128-register build removes the spills but halves the resident walkers and is
 * 1.3x slower, so the cap stays; a JIT program may still override it with a #define.) */
#ifndef MCI_CHAIN_MAXT
#define MCI_CHAIN_MAXT 1024
#endif
/* uniforms the MCMC kernel generates ahead per step (RngAhead): update type, variable, slot(s), shift, accept */
constexpr int MCMC_DRAWS_AHEAD = 6;
/* ... per kernel: a step draws the update kind, the variable, one or two slots / one uniform per leaf of the shifted
 * variable, and the acceptance — 5 uniforms when every variable is a single leaf, which RngAhead<5> serves with THREE
 * Philox calls per step for either parity of the stream position (6 take four: ncu, round 2: the look-ahead is 19 % of the
 * instructions of the C5 kernel).  Variables of several leaves (CompositeVar, FermiK) keep 6 */
template <int FID>
__device__ __host__ constexpr int mcmc_draws_ahead() {
#ifdef MCI_JIT
  return JitProblem::max_nleaf == 1 ? 5 : MCMC_DRAWS_AHEAD;
#else
  return (FID == MCI_F_GENZ_CORNER_COMP || FID == MCI_F_LATTICE_COMP || FID == MCI_F_DISCRETE2_CONST || FID == MCI_F_BUBBLE_FERMIK)
             ? MCMC_DRAWS_AHEAD : 5;
#endif
}
/* bin-by-Discrete observables of the chain kernels are kept in this many per-thread accumulators (mci_create checks) */
constexpr int CHAIN_MAX_BINNED_OBS = 8;
/* per-thread observable accumulators: one per integrand (default measure), CHAIN_MAX_BINNED_OBS for the single binned
 * integrand, the whole flattened observable for a user measure (JIT builds only: its length is a compile-time constant) */
template <int NI, int WD>
struct ChainNobs {
#ifdef MCI_JIT /* a JIT program knows the observable layout: exactly as many accumulators as the measure writes */
  static constexpr int value = JitProblem::measure_kind != MCI_MEASURE_DEFAULT ? JitProblem::obs_total : NI * WD;
#else
  static constexpr int value = (NI > 1 ? NI : CHAIN_MAX_BINNED_OBS) * WD;
#endif
};
#define PA_IDX(P, u, i, j) (((u) * (P).nd + (i)) * (P).pa_dim + (j))
/* FermiK proposals are compiled into the mcmc kernel of the integrand that takes one, and into a JIT program whose
 * problem holds one: every other mcmc kernel stays as it was */
template <int FID>
__device__ __host__ constexpr bool chain_has_fermik() {
#ifdef MCI_JIT
  return JitProblem::has_fermik;
#else
  return FID == MCI_F_BUBBLE_FERMIK;
#endif
}

/* ------------------------------------------------------------------------------------------ */
/* VegasMC                                                                                      */
/* ------------------------------------------------------------------------------------------ */
template <int FID>
__global__ void __launch_bounds__(MCI_CHAIN_MAXT, 1) vegasmc_kernel(const __grid_constant__ DevProblem Pp, const ChainArgs A) {
  MCI_VIEW(P, Pp);
  using F = Integrand<FID>;
  constexpr int NI = F::NI;
  constexpr int WD = IntegrandWD<FID>::value; /* doubles per integrand value: 2 = complex (re, im) */
  extern __shared__ __align__(16) double smem[];
  Walker<true> W;
  const int pa_len = P.pa_len_();
  /* propose/accept[2, 1, vi] counters of the CTA (updates.jl:90,92): u32 in shared memory (native ATOMS.ADD) */
  ChainSmem S = chain_setup<true>(P, A, smem, 2 * pa_len, W);
  unsigned int *s_prop = S.scount, *s_acc = S.scount + pa_len;
  const int tid = threadIdx.x;
  const int lb = blockIdx.x / A.ctas_per_block;
  const int cb = blockIdx.x - lb * A.ctas_per_block;
  const uint32_t block = (uint32_t)(A.rank + lb * A.world);
  const int walker = cb * blockDim.x + tid;
  const bool active = walker < A.walkers_per_block;
  constexpr int N = NI;
  const int hmode = A.hist_mode;

  MCI_ROUND_KEYS(K, P, A);
  ChainCtxT<Rng> C(P, S.maps, K);
  C.rng.init((uint64_t)walker, block);
  ChainX X{W.data, &P, W.T};

  double rw[NI + 1];
#pragma unroll
  for (int i = 0; i <= N; ++i) rw[i] = A.reweight[i];

  /* per-thread statistics.  Default measure: one accumulator per integrand; bin-by-Discrete: one per bin, selected
   * with a short compare chain so that the accumulators stay in registers */
  constexpr int NOBS = ChainNobs<NI, WD>::value;
  double obs[NOBS];
#pragma unroll
  for (int o = 0; o < NOBS; ++o) obs[o] = 0.0;
  double visited[NI + 1];
#pragma unroll
  for (int i = 0; i <= N; ++i) visited[i] = 0.0;
  double normalization = 0.0;
  unsigned long long neval = 0, bad = 0;

  double weights[NI * WD], nweights[NI * WD], pad[NI + 1], npad[NI + 1];
  /* montecarlo.jl:151-166 */
  for (int v = 0; v < P.nv; ++v) var_initialize(C, W, v);
  F::template eval<0>(X, P, nweights);
#pragma unroll
  for (int i = 0; i <= N; ++i) pad[i] = P.need_pad || i == N ? padding_probability(P, W, i) : 1.0;
  double prob_cur = __dmul_rn(rw[N], pad[N]);
#pragma unroll
  for (int i = 0; i < N * WD; ++i) weights[i] = nweights[i];
#pragma unroll
  for (int i = 0; i < N; ++i) prob_cur = __dadd_rn(prob_cur, __dmul_rn(__dmul_rn(wabs<WD>(nweights, i), rw[i]), pad[i]));
  const double meas_from = (double)A.steps / 100; /* ne >= neval/100 (montecarlo.jl:213) */

  for (long long ne = 1 - A.burnin; ne <= (long long)A.steps; ++ne) { /* ne <= 0: burn-in, no statistics */
    /* changeVariable: updates.jl:45-106 */
    if (active) {
      const int vi = C.rng.randint(K, P.nv) - 1;
      if (!P.top_skip[vi] && P.maxdof[vi] > 0) {
        const int pos = C.rng.randint(K, P.maxdof[vi]);
        const double prop = var_shift(C, W, vi, pos);
        if (prop > 4.9406564584124654e-324) {
          F::template eval<0>(X, P, nweights);
          const bool counted = ne >= 1; /* burn-in steps (ne <= 0) touch no statistic, counters included */
          if (counted) neval++;
          bool finite = true;
#pragma unroll
          for (int i = 0; i < N * WD; ++i) finite = finite && isfinite(nweights[i]);
          if (!finite && counted) bad++;
#pragma unroll
          for (int i = 0; i <= N; ++i) npad[i] = P.need_pad || i == N ? padding_probability(P, W, i) : 1.0;
          double prob_new = __dmul_rn(rw[N], npad[N]);
#pragma unroll
          for (int i = 0; i < N; ++i)
            prob_new = __dadd_rn(prob_new, __dmul_rn(__dmul_rn(wabs<WD>(nweights, i), rw[i]), npad[i]));
          const double R = __ddiv_rn(__dmul_rn(prop, prob_new), prob_cur);
          if (counted) atomicAdd(&s_prop[PA_IDX(P, 1, 0, vi)], 1u);
          if (C.rng.rand(K) < R) {
            if (counted) atomicAdd(&s_acc[PA_IDX(P, 1, 0, vi)], 1u);
#pragma unroll
            for (int i = 0; i < N * WD; ++i) weights[i] = nweights[i];
#pragma unroll
            for (int i = 0; i <= N; ++i) pad[i] = npad[i];
            prob_cur = prob_new;
          } else {
            var_shift_rollback(P, W, vi, pos);
          }
        }
      }
    }
    __syncwarp();
    /* histogram of every step: montecarlo.jl:198-211.  hist[bin of (v, pos)] += wf2_i for every integrand i with
     * pos <= dof[i][v]: the contributions of the integrands to one slot are summed first, one update per slot */
    if (hmode != HIST_NONE && active && ne >= 1) {
      double wf2[NI];
#pragma unroll
      for (int i = 0; i < N; ++i) {
        const double aw = wabs<WD>(weights, i);
        const double f2 = __ddiv_rn(__dmul_rn(aw, aw), probability(P, W, i));
        wf2[i] = __ddiv_rn(__dmul_rn(f2, pad[i]), prob_cur);
      }
      MCI_UNROLL_LAYOUT
      for (int v = 0; v < P.nv; ++v) {
        MCI_UNROLL_LAYOUT
        for (int pos = 1; pos <= P.maxdof[v]; ++pos) {
          double t = 0.0;
          bool any = false;
#pragma unroll
          for (int i = 0; i < N; ++i)
            if (pos <= P.dof[i * P.nv + v]) { t += wf2[i]; any = true; }
          if (any) {
            MCI_UNROLL_LAYOUT
            for (int k = 0; k < P.top_nleaf[v]; ++k) {
              const DevLeaf &L = P.leaf[P.top_first[v] + k];
              if (L.adapt) chain_hist_add(hmode, S.hist_saddr + 8u * (uint32_t)L.hist_off, S.ghist + L.hist_off, W.gidx[frow(L, pos) * W.T], t);
            }
          }
        }
      }
    }
    __syncwarp();
    /* measurement: montecarlo.jl:213-232 */
    if (active && ne >= 1 && (P.measurefreq == 1 || ((uint64_t)ne % (uint64_t)P.measurefreq) == 0) && (double)ne >= meas_from) {
      double relw_all[NI * WD];
#pragma unroll
      for (int i = 0; i < N; ++i) {
        /* abs(weights[i] * pad * reweight) / probability: a complex value is scaled component by component first */
        double wpr[WD];
#pragma unroll
        for (int c = 0; c < WD; ++c) wpr[c] = __dmul_rn(__dmul_rn(weights[WD * i + c], pad[i]), rw[i]);
        visited[i] += __ddiv_rn(wabs<WD>(wpr, 0), prob_cur);
#pragma unroll
        for (int c = 0; c < WD; ++c) {
          const double relw = __ddiv_rn(__dmul_rn(weights[WD * i + c], pad[i]), prob_cur);
          relw_all[WD * i + c] = relw;
          if (P.measure_kind == MCI_MEASURE_DEFAULT) obs[WD * i + c < NOBS ? WD * i + c : 0] += relw;
          else if (P.measure_kind == MCI_MEASURE_BIN_BY_DISCRETE && i == 0) {
            const int b = WD * W.gidx[P.leaf[P.nleaf - 1].base * W.T] + c;
#pragma unroll
            for (int o = 0; o < NOBS; ++o) obs[o] += (o == b) ? relw : 0.0;
          }
        }
      }
      if (P.measure_kind == MCI_MEASURE_USER) UserMeasureT<0>::call(X, obs, relw_all, P.userdata);
      normalization += __ddiv_rn(pad[N], prob_cur);
      visited[N] += __ddiv_rn(__dmul_rn(rw[N], pad[N]), prob_cur);
    }
  }
  if (bad) atomicAdd(A.flags, bad);

  /* per-CTA statistics, fixed order */
  const int stat_len = P.obs_total + 1 + P.nd + 2 * pa_len + 1;
  double *out = A.cta_stats + (size_t)blockIdx.x * stat_len;
  for (int i = tid; i < stat_len; i += blockDim.x) out[i] = 0.0;
  __syncthreads();
  if (P.measure_kind == MCI_MEASURE_DEFAULT) {
#pragma unroll
    for (int i = 0; i < N * WD; ++i) { double t = cta_sum(obs[i < NOBS ? i : 0], S.sred); if (tid == 0) out[P.obs_off[i / WD] + i % WD] = t; }
  } else if (P.measure_kind == MCI_MEASURE_BIN_BY_DISCRETE) {
#pragma unroll
    for (int o = 0; o < NOBS; ++o)
      if (o < P.obs_len[0]) { double t = cta_sum(obs[o], S.sred); if (tid == 0) out[P.obs_off[0] + o] = t; }
  } else {
#pragma unroll
    for (int o = 0; o < NOBS; ++o)
      if (o < P.obs_total) { double t = cta_sum(obs[o], S.sred); if (tid == 0) out[o] = t; }
  }
  { double t = cta_sum(normalization, S.sred); if (tid == 0) out[P.obs_total] = t; }
#pragma unroll
  for (int i = 0; i <= N; ++i) { double t = cta_sum(visited[i], S.sred); if (tid == 0) out[P.obs_total + 1 + i] = t; }
  __syncthreads();
  double *prop_o = out + P.obs_total + 1 + P.nd, *acc_o = prop_o + pa_len;
  for (int i = tid; i < pa_len; i += blockDim.x) { prop_o[i] = (double)s_prop[i]; acc_o[i] = (double)s_acc[i]; }
  { double t = cta_sum((double)neval, S.sred); if (tid == 0) out[stat_len - 1] = t; }
  chain_flush_hist<false>(P, A, S);
}

/* ------------------------------------------------------------------------------------------ */
/* MCMC                                                                                         */
/* ------------------------------------------------------------------------------------------ */
/* weight of integrand `idx` at the walker's point.  Integrands whose components are independent pay for one. */
template <int FID, class XA>
__device__ __forceinline__ void eval_one(const Problem &P, const XA &X, int idx, double *out) {
  constexpr int WD = IntegrandWD<FID>::value;
  if constexpr (FID == MCI_F_PRODUCT_4) {
    const int d = P.dof[0];
    const double a = P.userdata[idx];
    double p = 1.0;
    for (int i = 0; i < d; ++i) p *= (1 + a * (X(0, i + 1) - 0.5));
    out[0] = p;
  } else {
    double w[Integrand<FID>::NI * WD];
    Integrand<FID>::template eval<0>(X, P, w);
#pragma unroll
    for (int c = 0; c < WD; ++c) {
      double r = w[c];
#pragma unroll
      for (int i = 1; i < Integrand<FID>::NI; ++i) if (i == idx) r = w[WD * i + c];
      out[c] = r;
    }
  }
}

/* Warp-cooperative create!/remove! of changeIntegrand (updates.jl:14-33).  Moving between integrands of different
 * dof creates or removes up to maxdof slots; a lane doing that alone serialises ~90 instructions per slot while the
 * other 31 lanes wait (ncu: 1.0 active threads on those lines, 70 % of the kernel for the 20-D C5 workload).  Because
 * the stream is counter based, the slots of the proposing lane `src` are instead spread over the 32 lanes: lane j
 * draws uniform q0 + j of src's stream, maps it, writes src's column of the shared walker rows, and the 1/prob (create)
 * or prob (remove) factors are multiplied with a shuffle tree.  The product is associated differently from the
 * reference's left-to-right loop (last-bit differences in the acceptance ratio, inside the envelope of the integrand's
 * own rounding).  Returns the product on every lane; advances *q past the draws it consumed and leaves in
 * *b_out the buffered half when the new position is odd. */
__device__ __forceinline__ double coop_change_dof(const Problem &P, const double *maps, const PhiloxKey &K, double *wdata0,
                                                  gidx_t *wgidx0, int T, int src_tid, int curr, int nw, uint32_t c0, uint32_t c1,
                                                  uint32_t c3, uint32_t *q, double *b_out) {
  const int lane = threadIdx.x & 31;
  double prop = 1.0;
  double *data = wdata0 + src_tid;
  gidx_t *gidx = wgidx0 + src_tid;
  for (int v = 0; v < P.nv; ++v) {
    const int cd = P.dof[curr * P.nv + v], nd_ = P.dof[nw * P.nv + v];
    if (cd == nd_) continue;
    const int nl = P.top_nleaf[v];
    const int lo = cd < nd_ ? cd : nd_;
    const int n = (cd < nd_ ? nd_ - cd : cd - nd_) * nl; /* items in the reference's order: slot outer, leaf inner */
    for (int base = 0; base < n; base += 32) {
      const int item = base + lane;
      double term = 1.0, other = 0.0;
      if (item < n) {
        const int pos = lo + 1 + item / nl;
        const DevLeaf &L = P.leaf[P.top_first[v] + item % nl];
        const int f = frow(L, pos) * T;
        if (cd < nd_) { /* create! */
          const double u = stream_uniform(c0, c1, c3, *q + (uint32_t)item, K, &other);
          if (L.kind == MCI_LEAF_CONTINUOUS) {
            int iy;
            double dx;
            data[f] = map_continuous(reinterpret_cast<const double2 *>(maps + L.smem_off), L.nbins, L.nbins_d, u, iy, dx);
            gidx[f] = (gidx_t)iy;
            term = __dmul_rn(L.nbins_d, dx);
          } else {
            const double *dist = maps + L.smem_off;
            const int g = locate_clamped(dist + L.nbins, L.nbins, u);
            data[f] = (double)(L.lower_i + g);
            gidx[f] = (gidx_t)g;
            term = __ddiv_rn(1.0, dist[g]);
          }
        } else { /* remove! */
          if (L.kind == MCI_LEAF_CONTINUOUS) {
            const double2 g = reinterpret_cast<const double2 *>(maps + L.smem_off)[gidx[f]];
            term = __ddiv_rn(1.0, __dmul_rn(g.y, L.nbins_d));
          } else {
            term = (maps + L.smem_off)[gidx[f]];
          }
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) term *= __shfl_xor_sync(0xffffffffu, term, o);
      prop *= term;
      if (cd < nd_) {
        const int cnt = n - base < 32 ? n - base : 32;
        /* the lane that drew the last uniform of this batch holds the other half of that Philox call */
        const double ob = __shfl_sync(0xffffffffu, other, cnt - 1);
        *q += (uint32_t)cnt;
        if (*q & 1u) *b_out = ob;
      }
    }
  }
  __syncwarp();
  return prop;
}

template <int FID>
__global__ void __launch_bounds__(MCI_CHAIN_MAXT, 1) mcmc_kernel(const __grid_constant__ DevProblem Pp, const ChainArgs A) {
  MCI_VIEW(P, Pp);
  constexpr int NI = Integrand<FID>::NI;
  constexpr int WD = IntegrandWD<FID>::value; /* doubles per integrand value: 2 = complex (re, im) */
  extern __shared__ __align__(16) double smem[];
  const int pa_len = P.pa_len_();
  Walker<false> W;
  /* propose/accept counters of the CTA: [2][pa_len] u32 in shared memory (native ATOMS.ADD) */
  ChainSmem S = chain_setup<false>(P, A, smem, 2 * pa_len, W);
  unsigned int *s_prop = S.scount, *s_acc = S.scount + pa_len;
  const int tid = threadIdx.x;
  const int lb = blockIdx.x / A.ctas_per_block;
  const int cb = blockIdx.x - lb * A.ctas_per_block;
  const uint32_t block = (uint32_t)(A.rank + lb * A.world);
  const int walker = cb * blockDim.x + tid;
  const bool active = walker < A.walkers_per_block;
  constexpr int N = NI;
  const int hmode = A.hist_mode;
  const double TINY = 4.9406564584124654e-324 * 1e50;
  const double EPS0 = 4.9406564584124654e-324;

  MCI_ROUND_KEYS(K, P, A);
  constexpr bool FK = chain_has_fermik<FID>();
  ChainCtxT<RngAhead<mcmc_draws_ahead<FID>()>, FK> C(P, S.maps, K);
  C.rng.init((uint64_t)walker, block);
  if constexpr (FK) fermik_reset(P, W);
  ChainX X{W.data, &P, W.T};
  double rw[NI + 1];
#pragma unroll
  for (int i = 0; i <= N; ++i) rw[i] = A.reweight[i];

  constexpr int NOBS = ChainNobs<NI, WD>::value;
  double obs[NOBS];
#pragma unroll
  for (int o = 0; o < NOBS; ++o) obs[o] = 0.0;
  unsigned int visited[NI + 1];
#pragma unroll
  for (int i = 0; i <= N; ++i) visited[i] = 0u;
  double normalization = 0.0;
  unsigned long long neval = 0, bad = 0;

  /* start point: montecarlo.jl:99-113, initialize! :190-205; `idx` = the integrand to start with (:76) */
  int curr = P.start_idx;
  double weight[WD], nwgt[WD], prob_cur = 1.0;
#pragma unroll
  for (int c = 0; c < WD; ++c) weight[c] = nwgt[c] = 0.0;
  if (active) {
    for (int t = 0; t < 10000; ++t) {
      for (int v = 0; v < P.nv; ++v) var_initialize(C, W, v);
      if (curr == N) { prob_cur = rw[N]; break; } /* the normalisation integrand: weight 1 * reweight (:196-198) */
      eval_one<FID>(P, X, curr, weight);
      prob_cur = __dmul_rn(wabs<WD>(weight, 0), rw[curr]);
      if (prob_cur > TINY) break;
    }
    if (prob_cur == 0.0) atomicAdd(A.flags + 1, 1ull);
  }
  const int nupdates = 2 * P.nv + 2; /* [changeIntegrand, swapVariable, changeVariable x 2Nv] :117-120 */
  const uint64_t nburnin = (uint64_t)floor(__dmul_rn((double)A.steps, P.thermal_ratio));
  const uint64_t total = A.steps + nburnin;

  for (long long it = 1 - A.burnin; it <= (long long)total; ++it) { /* it <= 0: burn-in, no statistics */
    C.rng.fill(K); /* convergent: the divergent code below only pops registers */
    /* ---- propose (divergent, cheap) ---- */
    int kind = 0;   /* 0 none, 1 changeIntegrand, 2 changeVariable, 3 swapVariable (= first index of propose/accept + 1) */
    int tgt = curr; /* integrand to evaluate */
    int vi = 0, p1 = 0, p2 = 0;
    double prop = 1.0;
    bool ci = false, ci_dof = false; /* changeIntegrand proposed (to tgt); it changes some dof */
    if (active) {
#pragma unroll
      for (int i = 0; i <= N; ++i) visited[i] += (i == curr && it >= 1);
      const int u = C.rng.randint(K, nupdates);
      if (u == 1) {
        /* changeIntegrand: updates.jl:1-70.  The create!/remove! of the dof difference is done by the whole warp below */
        const int nnb = P.nb_ptr[curr + 1] - P.nb_ptr[curr];
        const int nw = P.nb_idx[P.nb_ptr[curr] + C.rng.randint(K, nnb) - 1];
        if (nw != curr) {
          prop = __ddiv_rn((double)nnb, (double)(P.nb_ptr[nw + 1] - P.nb_ptr[nw]));
          ci = true; tgt = nw;
          if (A.coop_ci) { /* many slots to create!/remove!: the whole warp does it below */
            MCI_UNROLL_LAYOUT
            for (int v = 0; v < P.nv; ++v) ci_dof = ci_dof || (P.dof[curr * P.nv + v] != P.dof[nw * P.nv + v]);
          } else {
            MCI_UNROLL_LAYOUT
            for (int v = 0; v < P.nv; ++v) {
              const int cd = P.dof[curr * P.nv + v], nd_ = P.dof[nw * P.nv + v];
              if (cd < nd_) { for (int pos = cd + 1; pos <= nd_; ++pos) prop = __dmul_rn(prop, var_create(C, W, v, pos)); }
              else if (cd > nd_) { for (int pos = nd_ + 1; pos <= cd; ++pos) prop = __dmul_rn(prop, var_remove(C, W, v, pos)); }
            }
          }
        }
      } else if (curr != N) {
        if (u == 2) {
          /* swapVariable: updates.jl:111-147 */
          vi = C.rng.randint(K, P.nv) - 1;
          const int cd = P.dof[curr * P.nv + vi];
          if (cd > 0) {
            p1 = C.rng.randint(K, cd);
            p2 = C.rng.randint(K, cd);
            if (p1 != p2) { var_swap(P, W, vi, p1, p2); kind = 3; }
          }
        } else {
          /* changeVariable: updates.jl:72-109 */
          vi = C.rng.randint(K, P.nv) - 1;
          const int cd = P.dof[curr * P.nv + vi];
          if (!P.top_skip[vi] && cd > 0) {
            p1 = C.rng.randint(K, cd);
            prop = var_shift(C, W, vi, p1);
            if (prop > EPS0) kind = 2;
          }
        }
      }
    }
    /* ---- changeIntegrand: create!/remove! the dof difference, one proposing lane at a time, 32 lanes per lane ---- */
    __syncwarp();
    {
      unsigned need = A.coop_ci ? __ballot_sync(0xffffffffu, ci_dof) : 0u;
      while (need) {
        const int src = __ffs(need) - 1;
        need &= need - 1;
        const int s_curr = __shfl_sync(0xffffffffu, curr, src), s_nw = __shfl_sync(0xffffffffu, tgt, src);
        const uint32_t s_c0 = __shfl_sync(0xffffffffu, C.rng.c0, src), s_c1 = __shfl_sync(0xffffffffu, C.rng.c1, src);
        uint32_t s_q = __shfl_sync(0xffffffffu, C.rng.q, src);
        double s_b = 0.0;
        const double f = coop_change_dof(P, S.maps, K, W.data - tid, W.gidx - tid, W.T, (tid & ~31) + src, s_curr, s_nw, s_c0, s_c1,
                                         C.rng.c3, &s_q, &s_b);
        if ((tid & 31) == src) {
          prop = __dmul_rn(prop, f);
          if (s_q != C.rng.q) C.rng.jump_to(s_q);
        }
        (void)s_b;
      }
      if (ci) { if (prop > EPS0) kind = 1; else tgt = curr; }
    }
    /* ---- evaluate (convergent): one integrand call site for all update types.  Independent thread scheduling does
     * not reconverge the lanes by itself after the divergent proposal code: without the barrier the warp drifts apart
     * for good (ncu: 4 active threads per instruction) ---- */
    __syncwarp();
#pragma unroll
    for (int c = 0; c < WD; ++c) nwgt[c] = 0.0;
    if (kind != 0 && tgt != N) eval_one<FID>(P, X, tgt, nwgt);
    __syncwarp();
    /* ---- accept / rollback (divergent, cheap) ---- */
    if (kind != 0) {
      const bool counted = it >= 1; /* burn-in steps (it <= 0) touch no statistic, counters included */
      if (counted) neval++;
      bool finite = true;
#pragma unroll
      for (int c = 0; c < WD; ++c) finite = finite && isfinite(nwgt[c]);
      if (!finite && counted) bad++;
      const double nprob = (tgt == N) ? rw[N] : __dmul_rn(wabs<WD>(nwgt, 0), rw[tgt]);
      const double R = __ddiv_rn(__dmul_rn(prop, nprob), prob_cur);
      const int j = kind == 1 ? tgt : vi;
      const int pa = PA_IDX(P, kind - 1, curr, j);
      if (counted) atomicAdd(&s_prop[pa], 1u);
      if (C.rng.rand(K) < R) {
        if (counted) atomicAdd(&s_acc[pa], 1u);
        curr = tgt; prob_cur = nprob;
#pragma unroll
        for (int c = 0; c < WD; ++c) weight[c] = nwgt[c];
      } else if (kind == 2) {
        var_shift_rollback(P, W, vi, p1);
      } else if (kind == 3) {
        var_swap(P, W, vi, p1, p2);
      }
    }
    __syncwarp();
    /* measurement: montecarlo.jl:145-170 */
    if (active && it >= 1 && (P.measurefreq == 1 || ((uint64_t)it % (uint64_t)P.measurefreq) == 0) && (uint64_t)it >= nburnin) {
      if (curr == N) {
        normalization += __ddiv_rn(1.0, rw[N]);
      } else {
        /* accumulate!(var, pos, 1.0) for every slot of the current integrand (:148-155).  The policy test stays outside
         * the slot loops and the shared counters are addressed through their 32-bit shared address: 4 instructions per
         * slot (index, LDS.U16, address, ATOMS) — the generic-pointer form with the test inside cost 12, 14 % of the C5
         * kernel (ncu source view, round 2) */
        if (hmode == HIST_SMEM_ATOMIC) {
          MCI_UNROLL_LAYOUT
          for (int v = 0; v < P.nv; ++v) {
            const int cd = P.dof[curr * P.nv + v];
            MCI_UNROLL_LAYOUT
            for (int k = 0; k < P.top_nleaf[v]; ++k) {
              const DevLeaf &L = P.leaf[P.top_first[v] + k];
              if (!L.adapt) continue;
              const uint32_t cells = S.hist_saddr + 4u * (uint32_t)L.hist_off;
              const gidx_t *g = W.gidx + frow(L, 1) * W.T;
              const int step = L.stride * W.T;
#pragma unroll 4
              for (int pos = 0; pos < cd; ++pos)
                asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(cells + 4u * (uint32_t)g[pos * step]), "r"(1u) : "memory");
            }
          }
        } else if (hmode != HIST_NONE) {
          for (int v = 0; v < P.nv; ++v) {
            const int cd = P.dof[curr * P.nv + v];
            for (int pos = 1; pos <= cd; ++pos)
              for (int k = 0; k < P.top_nleaf[v]; ++k) {
                const DevLeaf &L = P.leaf[P.top_first[v] + k];
                if (L.adapt) atomicAdd(S.ghist + L.hist_off + W.gidx[frow(L, pos) * W.T], 1.0);
              }
          }
        }
        double relw[WD]; /* relativeWeight = state.weight / state.probability (:163), component by component */
#pragma unroll
        for (int c = 0; c < WD; ++c) relw[c] = __ddiv_rn(weight[c], prob_cur);
        if (P.measure_kind == MCI_MEASURE_DEFAULT) {
#pragma unroll
          for (int i = 0; i < N * WD; ++i) obs[i < NOBS ? i : 0] += (i / WD == curr) ? relw[i % WD] : 0.0;
        } else if (P.measure_kind == MCI_MEASURE_BIN_BY_DISCRETE) {
          const int b = WD * W.gidx[P.leaf[P.nleaf - 1].base * W.T];
#pragma unroll
          for (int o = 0; o < NOBS; ++o) obs[o] += (o / WD * WD == b) ? relw[o % WD] : 0.0;
        } else { /* measure(idx, x, obs, relw, config): a measure is linear in relw, so pass zero for the other integrands */
          double relw_all[NI * WD];
#pragma unroll
          for (int i = 0; i < N * WD; ++i) relw_all[i] = (i / WD == curr) ? relw[i % WD] : 0.0;
          UserMeasureT<0>::call(X, obs, relw_all, P.userdata);
        }
      }
    }
  }
  if (bad) atomicAdd(A.flags, bad);

  const int stat_len = P.obs_total + 1 + P.nd + 2 * pa_len + 1;
  double *out = A.cta_stats + (size_t)blockIdx.x * stat_len;
  for (int i = tid; i < stat_len; i += blockDim.x) out[i] = 0.0;
  __syncthreads();
  if (P.measure_kind == MCI_MEASURE_DEFAULT) {
#pragma unroll
    for (int i = 0; i < N * WD; ++i) { double t = cta_sum(obs[i < NOBS ? i : 0], S.sred); if (tid == 0) out[P.obs_off[i / WD] + i % WD] = t; }
  } else if (P.measure_kind == MCI_MEASURE_BIN_BY_DISCRETE) {
#pragma unroll
    for (int o = 0; o < NOBS; ++o)
      if (o < P.obs_len[0]) { double t = cta_sum(obs[o], S.sred); if (tid == 0) out[P.obs_off[0] + o] = t; }
  } else {
#pragma unroll
    for (int o = 0; o < NOBS; ++o)
      if (o < P.obs_total) { double t = cta_sum(obs[o], S.sred); if (tid == 0) out[o] = t; }
  }
  { double t = cta_sum(normalization, S.sred); if (tid == 0) out[P.obs_total] = t; }
#pragma unroll
  for (int i = 0; i <= N; ++i) { double t = cta_sum((double)visited[i], S.sred); if (tid == 0) out[P.obs_total + 1 + i] = t; }
  __syncthreads();
  double *prop_o = out + P.obs_total + 1 + P.nd, *acc_o = prop_o + pa_len;
  for (int i = tid; i < pa_len; i += blockDim.x) { prop_o[i] = (double)s_prop[i]; acc_o[i] = (double)s_acc[i]; }
  { double t = cta_sum((double)neval, S.sred); if (tid == 0) out[stat_len - 1] = t; }
  chain_flush_hist<true>(P, A, S);
}

}  // namespace mci
