/*
 * mci_device.cuh — device-side core of libmcib200 (sm_100a): problem description as a
 * kernel parameter, Philox4x32-10 counter RNG, the adaptive-map draws (Continuous /
 * Discrete), the histogram scatter policies and the warp/CTA reductions.
 *
 * Reference arithmetic restated here (paths relative to numericalEFT/MCIntegration.jl v0.4.2):
 *   Continuous create!/shift! map   src/distribution/sampler.jl:293-305, 336-386
 *   Discrete   create!/shift!       src/distribution/sampler.jl:13-22, 57-71
 *   locate (bisection)              src/distribution/common.jl:8-36
 *   accumulate!                     src/distribution/variable.jl:196-200, 362-367, 474-478
 *
 * The map arithmetic uses explicit round-to-nearest intrinsics (never contracted into
 * FMA) so that `x` and the bin index are bit-identical to a plain IEEE evaluation of the
 * reference's expressions; see DESIGN.md "Numerics".
 */
#pragma once
#ifndef __CUDACC_RTC__
#include <cuda_runtime.h>
#include <stdint.h>
#endif

#include "../../include/mcib200.h"

namespace mci {

constexpr int MAX_DRAW = 64;     /* sum over top-level vars of maxdof * n_leaf                     */
constexpr int MAX_ND = MCI_MAX_INTEGRAND + 1;
constexpr int WARP = 32;

/* histogram scatter policies (MCI_HIST_*) */
constexpr int HIST_NONE = 0;         /* no adaptive leaf                                            */
constexpr int HIST_SMEM_ATOMIC = 1;  /* one CTA-shared fp64 histogram, shared-memory CAS atomics    */
constexpr int HIST_GLOBAL_RED = 2;   /* RED.ADD.F64 into replicated global copies (L2)              */
constexpr int HIST_WARP_PRIVATE = 3; /* one private histogram per warp, conflict-ranked plain RMW   */
constexpr int HIST_SMEM_STRIPED = 4; /* CTA-shared fp64 histogram in 16 bank-striped copies: lane l updates copy l%16, which
                                        lives in bank pair l%16, so neither the LDS.64 nor the CAS of a random-bin scatter
                                        ever has a bank conflict (2 wavefronts per instruction instead of ~10)           */
constexpr int HIST_SMEM_FIXED = 5;   /* the striped copies as 64-bit fixed-point cells (two u32 limbs, native ATOMS.ADD) */
#ifndef MCI_STRIPE_LOG2
#define MCI_STRIPE_LOG2 4
#endif
constexpr int STRIPE_LOG2 = MCI_STRIPE_LOG2;
constexpr int STRIPE = 1 << STRIPE_LOG2;
/* fixed-point histogram: a weight w2 is added as round(w2 * 2^E) into a 64-bit cell.  E is chosen on the device after
 * every iteration from M = the largest bin of that iteration's (all-reduced) histogram: 2^E M = 2^FIXED_CAP_LOG2 = 2^58.
 * A cell holds part of ONE bin, so it cannot overflow unless a bin grows beyond 2^6 M — and an overflow of the high limb
 * is detected exactly, at the add that causes it.  One unit is 2^-58 M: the n adds of a bin lose at most n/2 units
 * (typically 0.02 n, the bias of rounding a decreasing density), 2^-40 M for the 1e7 adds per bin of a BASELINE-size
 * iteration — 1e-12 of a bin of a trained map, whose bins are within a small factor of each other; fp64 accumulation
 * itself is no better than n 2^-53.  The scale is only USED when the histogram is settled: if M moved by more than
 * 2^FIXED_STABLE_LOG2 against the iteration before (the first two iterations of an untrained map, normally), E is
 * marked invalid and the fp64 kernel runs that iteration instead (both are launched; the device-side flag picks). */
constexpr int FIXED_CAP_LOG2 = 58;
constexpr int FIXED_STABLE_LOG2 = 8;
constexpr int FIXED_EXP_INVALID = -100000;
#ifndef MCI_POOL_MAP_REP
#define MCI_POOL_MAP_REP 4
#endif
/* copies of the pool's (edge, width) table in the striped variant, interleaved [increment][copy] (PoolMap, mci_vegas.cuh):
 * 4 copies cut the bank conflicts of the ten random LDS.128 lookups of a sample from ~10.9 to ~7.8 wavefronts each at no
 * extra instruction — measured (A/B in one job, B200) C2 38.9 -> 41.3, C3-A 25.2 -> 26.3, C5 integrand 19.3 -> 20.3 G samples/s */
constexpr int POOL_MAP_REP = MCI_POOL_MAP_REP;

struct DevLeaf {
  int kind, nbins, adapt, lower_i;
  int map_off;  /* Continuous: offset (in bins) into the pair table; Discrete: offset into dist     */
  int acc_off;  /* Discrete: offset into accumulation                                               */
  int hist_off; /* offset into the concatenated histogram                                           */
  int top;      /* owning top-level variable                                                        */
  int base;     /* flat index of (this leaf, slot 1) in the per-sample x[] / bin[] arrays           */
  int stride;   /* flat-index stride between consecutive slots (= n_leaf of the owning variable)    */
  int smem_off; /* offset (in doubles) of this leaf's map inside the CTA's shared-memory map area   */
  int pad_;
  double nbins_d;
  double p0, p1; /* FermiK component: kF, dk                                                         */
};

/* Passed by value as a __grid_constant__ kernel parameter (constant bank, uniform loads). */
struct DevProblem {
  int solver, fid, measure_kind, ni, nd, nv, nleaf, n_draw;
  int obs_total, hist_len, pa_dim, measurefreq;
  int need_pad;   /* some dof[i] != maxdof  (vegas/montecarlo.jl:82)                                */
  int any_adapt;
  int map_doubles; /* doubles of shared memory holding all maps                                     */
  int n_tslot;     /* sum of maxdof over top-level variables                                        */
  int wdim;        /* doubles per integrand value / observable entry: 1 real, 2 complex (re, im)    */
  int start_idx;   /* mcmc: integrand every chain starts in, 0-based (mcmc/montecarlo.jl:76)        */
  int top_nleaf[MCI_MAX_TOPVAR], top_first[MCI_MAX_TOPVAR], top_base[MCI_MAX_TOPVAR], maxdof[MCI_MAX_TOPVAR];
  int top_tslot[MCI_MAX_TOPVAR]; /* first top-level slot index of the variable                      */
  int top_skip[MCI_MAX_TOPVAR];  /* single Discrete leaf with one value: updates skip it            */
  int dof[MAX_ND * MCI_MAX_TOPVAR];
  int same_dof[MCI_MAX_INTEGRAND]; /* dof[i] == maxdof                                               */
  int obs_off[MCI_MAX_INTEGRAND], obs_len[MCI_MAX_INTEGRAND];
  int nb_ptr[MAX_ND + 1], nb_idx[MAX_ND * MAX_ND];
  DevLeaf leaf[MCI_MAX_LEAF];
  double userdata[MCI_MAX_USERDATA];
  double fconst[4]; /* integrand constants evaluated on the host with the C library              */
  double pool_npow[2]; /* single-pool layout: N^h and N^(d-h), N = increments of the pool's grid, d = maxdof, h = ceil(d/2):
                          the two halves of N^d that PoolJac multiplies into the Jacobian (constant-bank operands)  */
  double thermal_ratio;
  uint64_t seed;
  __host__ __device__ int pa_len_() const { return 3 * nd * pa_dim; }
};

/* `Problem` is what the device code reads the problem description through.  Ahead-of-time kernels read the
 * __grid_constant__ kernel parameter (constant bank).  A JIT build (mci_jit.cu: NVRTC, one program per problem) defines
 * MCI_JIT and supplies "mci_jit_problem.h", whose JitProblem serves every layout field from compile-time tables — all
 * loops over variables / leaves / slots / integrands then have constant bounds and unroll, per-sample arrays become
 * registers — and forwards only the run-time values (userdata, seed, ...) from the kernel parameter. */
#ifdef MCI_JIT
}  // namespace mci
#include "mci_jit_problem.h"
namespace mci {
using Problem = JitProblem;
#define MCI_VIEW(P, Pp) const ::mci::JitProblem P(Pp)
#define MCI_UNROLL_LAYOUT _Pragma("unroll") /* loops over the layout have constant bounds: unroll them completely */
#else
using Problem = DevProblem;
#define MCI_VIEW(P, Pp) const ::mci::DevProblem &P = Pp
#define MCI_UNROLL_LAYOUT
#endif

/* MCI_MEASURE_USER: the `measure` callback of integrate (main.jl:35-37) as CUDA C++ source.  The primary template does
 * nothing; a JIT program whose user_source defines
 *   template <class XA> __device__ void mci_user_measure(const XA &X, double *obs, const double *relw, const double *userdata);
 * specialises UserMeasureT<0> to forward to it (mci_jit.cu).  obs = this thread's accumulators, laid out like
 * config.observable flattened (integrand i at obs_off[i], obs_len[i] entries); relw[i] = relative weight of integrand i
 * (mcmc: zero except for the current integrand — a measure is linear in relw). */
template <int K>
struct UserMeasureT {
  template <class XA>
  __device__ __forceinline__ static void call(const XA &, double *, const double *, const double *) {}
};

/* |w_i| of integrand i: abs of a real value, hypot(re, im) of a complex one (Julia's abs(::Complex)) — what the
 * reference takes wherever it writes abs(weights[i]) (vegas/montecarlo.jl:172, vegas_mc/montecarlo.jl:159,200,
 * vegas_mc/updates.jl:86, mcmc/updates.jl:40,95,130) */
template <int WD>
__device__ __forceinline__ double wabs(const double *w, int i) {
  if constexpr (WD == 1) return fabs(w[i]);
  else return hypot(w[2 * i], w[2 * i + 1]);
}

/* ---------------------------------------------------------------------------------------- */
/* Philox4x32-10 (Salmon et al. 2011).  Stream layout (DESIGN.md "RNG"):                     */
/*   key = (seed_lo, seed_hi ^ iter*0x9E3779B1), counter = (idx_lo, idx_hi, pair, block)     */
/* ---------------------------------------------------------------------------------------- */
struct PhiloxKey {
  uint32_t k[10][2];
};

__host__ __device__ inline void philox_expand_key(uint64_t seed, uint64_t iter, PhiloxKey &K) {
  uint32_t k0 = (uint32_t)seed;
  uint32_t k1 = (uint32_t)(seed >> 32) ^ ((uint32_t)iter * 0x9E3779B1u);
  for (int r = 0; r < 10; ++r) {
    K.k[r][0] = k0;
    K.k[r][1] = k1;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}

/* The round keys depend on (seed, iter) only.  Expanded in the kernel they are either kept in 20 registers or — what
 * ptxas does at the 64-register cap of the 1024-thread kernels — recomputed with 18 additions in front of every Philox
 * call (4-6 % of the instructions of the chain and generic kernels, ncu source view).  The host expands them into the
 * launch arguments instead: every round then reads its two keys straight from the constant bank. */
#ifndef MCI_KEY_IN_ARGS
#define MCI_KEY_IN_ARGS 1
#endif
#if MCI_KEY_IN_ARGS
#define MCI_ROUND_KEYS(K, P, A) const PhiloxKey &K = (A).key
#else
#define MCI_ROUND_KEYS(K, P, A) PhiloxKey K; philox_expand_key((P).seed, (A).iter, K)
#endif

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const PhiloxKey &K,
                                               uint32_t &o0, uint32_t &o1, uint32_t &o2, uint32_t &o3) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ K.k[r][0];
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ K.k[r][1];
    c1 = (uint32_t)p1;
    c3 = (uint32_t)p0;
    c0 = n0;
    c2 = n2;
  }
  o0 = c0; o1 = c1; o2 = c2; o3 = c3;
}

/* k * 2^-52 with k the top 52 bits of (hi:lo): exponent trick, exact (no int->fp conversion) */
__device__ __forceinline__ double u52(uint32_t hi, uint32_t lo) {
  uint32_t h = (hi >> 12) | 0x3FF00000u;
  uint32_t l = __funnelshift_r(lo, hi, 12);
  return __dsub_rn(__hiloint2double((int)h, (int)l), 1.0);
}

/* sequential stream for the Markov-chain solvers: rand(config.rng).  Uniform number q of the stream is half (q & 1)
 * of Philox call (q >> 1) with counter (idx_lo, idx_hi, q >> 1, block): the position is explicit, so other lanes of
 * the warp can compute draws of this stream on the owner's behalf (cooperative create!, mci_chain.cuh). */
struct Rng {
  uint32_t c0, c1, c3;
  uint32_t q;  /* uniforms consumed so far */
  double b;    /* second half of call q >> 1, valid while q is odd */
  __device__ __forceinline__ void init(uint64_t idx, uint32_t block) {
    c0 = (uint32_t)idx; c1 = (uint32_t)(idx >> 32); c3 = block; q = 0; b = 0.0;
  }
  __device__ __forceinline__ double rand(const PhiloxKey &K) {
    if (q & 1u) { q++; return b; }
    uint32_t o0, o1, o2, o3;
    philox4x32_10(c0, c1, q >> 1, c3, K, o0, o1, o2, o3);
    b = u52(o2, o3);
    q++;
    return u52(o0, o1);
  }
  /* rand(rng, 1:n) -> 1 + floor(u*n) */
  __device__ __forceinline__ int randint(const PhiloxKey &K, int n) {
    int k = (int)floor(__dmul_rn(rand(K), (double)n)) + 1;
    return k > n ? n : k;
  }
};

/* The same stream with the next KB uniforms generated AHEAD, by all lanes of the warp together.  In the MCMC kernel
 * every draw sits inside divergent code (three update types, different variables): a Philox call inlined there runs
 * once per divergent group — ncu on the bubble: the Philox rounds executed with 1.5 active threads on average, 15 % of
 * all instructions.  fill() is called once per step at a convergent point; rand() then only pops a register.  A step
 * that needs more than KB draws (create! of several slots) falls back to computing the call in place.  The stream
 * positions are exactly those of Rng, so the oracle sees the same numbers. */
template <int KB>
struct RngAhead {
  uint32_t c0, c1, c3;
  uint32_t q;     /* uniforms consumed so far */
  int avail;      /* valid entries at the front of buf */
  double buf[KB];
  __device__ __forceinline__ void init(uint64_t idx, uint32_t block) {
    c0 = (uint32_t)idx; c1 = (uint32_t)(idx >> 32); c3 = block; q = 0; avail = 0;
  }
  /* generate the uniforms at positions q .. q+KB-1 (KB/2 + 1 Philox calls, selected by the parity of q) */
  __device__ __forceinline__ void fill(const PhiloxKey &K) {
    constexpr int NC = KB / 2 + 1;
    double h[2 * NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) {
      uint32_t o0, o1, o2, o3;
      philox4x32_10(c0, c1, (q >> 1) + k, c3, K, o0, o1, o2, o3);
      h[2 * k] = u52(o0, o1);
      h[2 * k + 1] = u52(o2, o3);
    }
    const bool odd = q & 1u;
#pragma unroll
    for (int i = 0; i < KB; ++i) buf[i] = odd ? h[i + 1] : h[i];
    avail = KB;
  }
  __device__ __forceinline__ double rand(const PhiloxKey &K) {
    double r;
    if (avail > 0) {
      r = buf[0];
#pragma unroll
      for (int i = 0; i + 1 < KB; ++i) buf[i] = buf[i + 1];
      avail--;
    } else {
      double other;
      uint32_t o0, o1, o2, o3;
      philox4x32_10(c0, c1, q >> 1, c3, K, o0, o1, o2, o3);
      other = u52(o2, o3);
      r = (q & 1u) ? other : u52(o0, o1);
    }
    q++;
    return r;
  }
  __device__ __forceinline__ int randint(const PhiloxKey &K, int n) {
    int k = (int)floor(__dmul_rn(rand(K), (double)n)) + 1;
    return k > n ? n : k;
  }
  /* the position was moved by someone else (cooperative create!): whatever was generated ahead is stale */
  __device__ __forceinline__ void jump_to(uint32_t newq) { q = newq; avail = 0; }
};

/* uniform number `q` of the stream (c0, c1, ., c3); *other = the other half of the same Philox call */
__device__ __forceinline__ double stream_uniform(uint32_t c0, uint32_t c1, uint32_t c3, uint32_t q, const PhiloxKey &K,
                                                  double *other) {
  uint32_t o0, o1, o2, o3;
  philox4x32_10(c0, c1, q >> 1, c3, K, o0, o1, o2, o3);
  const double u0 = u52(o0, o1), u1 = u52(o2, o3);
  *other = (q & 1u) ? u0 : u1;
  return (q & 1u) ? u1 : u0;
}

/* ---------------------------------------------------------------------------------------- */
/* maps                                                                                      */
/* ---------------------------------------------------------------------------------------- */

/* One draw on a Continuous map (sampler.jl:295-303): pair[i] = (grid[i], grid[i+1]-grid[i]).
 * Returns x; iy = 0-based increment; dx = width of the increment. */
__device__ __forceinline__ double map_continuous(const double2 *__restrict__ pair, int nbins, double nbins_d, double u,
                                                 int &iy, double &dx) {
  const double MAGIC = 4503599627370496.0; /* 2^52: (t + 2^52) rounded down holds floor(t) in its low bits */
  double t = __dmul_rn(u, nbins_d);
  double f = __dadd_rd(t, MAGIC);
  int i = __double2loint(f);
  double fl = __dsub_rn(f, MAGIC);
  double dy = __dsub_rn(t, fl);
  i = min(i, nbins - 1);
  double2 g = pair[i];
  iy = i;
  dx = g.y;
  return __dadd_rn(g.x, __dmul_rn(dy, g.y));
}

/* the same draw with the table lookup left to the caller: increment iy (0-based) and the offset dy inside it */
__device__ __forceinline__ void map_index(int nbins, double nbins_d, double u, int &iy, double &dy) {
  const double MAGIC = 4503599627370496.0;
  double t = __dmul_rn(u, nbins_d);
  double f = __dadd_rd(t, MAGIC);
  int i = __double2loint(f);
  double fl = __dsub_rn(f, MAGIC);
  dy = __dsub_rn(t, fl);
  iy = min(i, nbins - 1);
}

/* The pool kernel's form of the same draw, two instructions shorter and bit-identical (tests: pool trace == oracle):
 *  - the uniform is u = m - 1 with m in [1, 2) built from the top 52 bits of the Philox words (u52 without its
 *    subtraction).  m - 1 is exact, so t = RN(u * N) = RN(m * N - N) = fma(m, N, -N): one DFMA replaces DADD + DMUL;
 *  - no clamp of the increment: u <= 1 - 2^-52, so the exact product is <= N - N 2^-52 <= N - ulp(N) (ulp(N) = 2^(e-52)
 *    <= N 2^-52 for 2^e <= N), which is representable or rounds to a value below N: floor(RN(u N)) <= N - 1 for every N. */
__device__ __forceinline__ double m52(uint32_t hi, uint32_t lo) { /* 1 + k 2^-52, k = top 52 bits of (hi:lo) */
  return __hiloint2double((int)((hi >> 12) | 0x3FF00000u), (int)__funnelshift_r(lo, hi, 12));
}
__device__ __forceinline__ void map_index_m(double nbins_d, double neg_nbins_d, double m, int &iy, double &dy) {
  const double MAGIC = 4503599627370496.0;
  const double t = __fma_rn(m, nbins_d, neg_nbins_d);
  const double f = __dadd_rd(t, MAGIC);
  iy = __double2loint(f);
  dy = __dsub_rn(t, __dsub_rn(f, MAGIC));
}

/* locate (common.jl:8-36) clamped into the table instead of throwing; returns 0-based index g
 * with acc[g] <= u < acc[g+1]; acc has n+1 entries. */
__device__ __forceinline__ int locate_clamped(const double *__restrict__ acc, int n, double u) {
  if (u < acc[0]) return 0;
  int jl = 1, ju = n + 1;
  while (ju - jl > 1) {
    int jm = (jl + ju) >> 1;
    if (u < acc[jm - 1]) ju = jm; else jl = jm;
  }
  return jl - 1;
}

/* ---------------------------------------------------------------------------------------- */
/* warp helpers                                                                              */
/* ---------------------------------------------------------------------------------------- */
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ unsigned lane_id() {
  unsigned l;
  asm volatile("mov.u32 %0, %%laneid;" : "=r"(l));
  return l;
}

/* ---------------------------------------------------------------------------------------- */
/* histogram scatter: hist[bin] += v for the 32 lanes of a converged warp                    */
/* ---------------------------------------------------------------------------------------- */

/* HIST_WARP_PRIVATE: `wh` is this warp's private copy.  Lanes hitting the same bin are ranked
 * with MATCH.ANY and applied in lane order, one plain LDS/DADD/STS round per rank, so the
 * result is deterministic.  All 32 lanes must call; inactive lanes pass active=false. */
__device__ __forceinline__ void hist_add_private(double *wh, int bin, double v, bool active) {
  const unsigned lane = lane_id();
  int key = active ? bin : (int)(0x80000000u | lane);
  unsigned peers = __match_any_sync(0xffffffffu, key);
  int rank = __popc(peers & ((1u << lane) - 1u));
  if (__all_sync(0xffffffffu, rank == 0)) {
    if (active) wh[bin] += v;
  } else {
    int maxrank = __reduce_max_sync(0xffffffffu, rank);
    for (int r = 0; r <= maxrank; ++r) {
      if (active && rank == r) wh[bin] += v;
      __syncwarp();
    }
  }
}

template <int MODE>
__device__ __forceinline__ void hist_add(double *h, int bin, double v, bool active) {
  if (MODE == HIST_WARP_PRIVATE) {
    hist_add_private(h, bin, v, active);
  } else if (MODE == HIST_SMEM_ATOMIC || MODE == HIST_GLOBAL_RED) {
    if (active) atomicAdd(h + bin, v); /* shared: CAS loop; global: RED.E.ADD.F64 */
  }
}

__device__ __forceinline__ void hist_add_rt(int mode, double *h, int bin, double v, bool active) {
  if (mode == HIST_WARP_PRIVATE) hist_add_private(h, bin, v, active);
  else if (mode != HIST_NONE) { if (active) atomicAdd(h + bin, v); }
}

/* stage all maps into shared memory with 16-byte loads: Continuous leaves as (grid, width)
 * pairs, Discrete leaves as [dist(n) | accumulation(n+1)] */
__device__ __forceinline__ void stage_maps(const Problem &P, double *smaps, const double *__restrict__ gmaps) {
  const int n2 = P.map_doubles >> 1; /* map area is padded to an even number of doubles */
  const double2 *src = reinterpret_cast<const double2 *>(gmaps);
  double2 *dst = reinterpret_cast<double2 *>(smaps);
  for (int i = threadIdx.x; i < n2; i += blockDim.x) dst[i] = __ldg(src + i);
}

}  // namespace mci
