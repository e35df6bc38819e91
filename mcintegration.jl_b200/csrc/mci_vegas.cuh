/*
 * mci_vegas.cuh — the VEGAS fill kernels: Philox -> adaptive-map draw of every variable ->
 * integrand -> Jacobian-weighted observable + per-bin variance histogram, fused in one
 * persistent kernel.
 *
 * Replaces the hot loop of Vegas.montecarlo (src/vegas/montecarlo.jl:117-187):
 *   :122-131  redraw every pool slot <= maxdof of every variable, jac = prod 1/prob
 *   :133-137  padding probabilities when dof differs between integrands
 *   :140-144  weights = integrand(x, config)
 *   :148-165  observable[i] += w_i * pad_i * jac ; normalization += 1   (every measurefreq-th sample)
 *   :170-185  histogram += (abs(w_i) * jac)^2 for every (variable, integrand, slot <= dof)
 *
 * Work decomposition: one iteration = n_block logical blocks (src/main.jl:121, the unit of the
 * error estimate) x neval_per_block samples.  A block is cut into tiles of THREADS x spt samples;
 * persistent CTAs walk the (block, tile) list round-robin.  Sample s of block b draws from the
 * Philox stream (seed, iter, b, s), so the set of samples does not depend on the launch shape
 * or on the number of GPUs.  Per-tile observable sums go to tile_obs[] (summed per block, in
 * tile order, by the finalize kernel -> deterministic); histograms are kept in shared memory
 * for the CTA's whole life and flushed once to the CTA's global partial.
 */
#pragma once
#include "mci_integrands.cuh"

/* updates of one sample issued together (loads / adds back to back, then the dependent halves) */
#ifndef MCI_POOL_KEY_IN_ARGS
#define MCI_POOL_KEY_IN_ARGS 1
#endif
#ifndef MCI_HIST_BATCH
#define MCI_HIST_BATCH 5
#endif
/* pool draws: t = fma(m, N, -N) without clamp (1) or the generic map_index (0); Jacobian as N^d prod dx (1) or prod N dx (0) */
#ifndef MCI_POOL_FMA_INDEX
#define MCI_POOL_FMA_INDEX 1
#endif
#ifndef MCI_POOL_HALF_JAC
#define MCI_POOL_HALF_JAC 1
#endif
/* lost-race check of a batch of compare-and-swaps: OR of XORs (1) or a chain of 64-bit compares (0) */
#ifndef MCI_LOST_LOP3
#define MCI_LOST_LOP3 1
#endif
/* fixed-point cells: add the high limb unconditionally (1) or only when it is non-zero (0) */
#ifndef MCI_FIXED_HI_ALWAYS
#define MCI_FIXED_HI_ALWAYS 1
#endif

namespace mci {

struct FillArgs {
  const double *gmaps;   /* staged map image (see stage_maps)                                      */
  double *tile_obs;      /* [n_tiles_total][obs_total]                                             */
  double *tile_cnt;      /* [n_tiles_total] measured samples of the tile                            */
  double *cta_hist;      /* [n_hist_copies][hist_len] partial histograms (zeroed before launch)     */
  unsigned long long *flags; /* [0] = number of non-finite weights seen, [2] = fixed-point histogram range exceeded,
                                [3] = ... and the iteration was recomputed by the fp64 rescue launch                */
  const int *fixed_exp;  /* HIST_SMEM_FIXED: the exponent E of this iteration's scale 2^E (device memory)      */
  int fixed_bias;        /* added to E (0; the tests raise it through MCIB200_FIXED_EXP_BIAS to provoke an overflow) */
  int rescue;            /* generic kernel, second launch of an iteration: work only if the first one left the fixed-point range */
  uint64_t iter;
  uint64_t neval_per_block;
  int n_local_blocks, rank, world; /* this rank runs blocks b = rank + k*world                      */
  int tiles_per_block, spt;
  int hist_mode, n_hist_copies;
  int maps_in_smem;
  PhiloxKey key;         /* the ten round keys of (seed, iter), expanded on the host: operands from the constant bank */
};

/* ------------------------------------------------------------------------------------------ */
/* pool kernel: ONE top-level variable that is a single Continuous leaf (the reference's         */
/* idiomatic `var = Continuous(a, b), dof = [[d1], [d2], ...]`).  Everything about a sample      */
/* lives in registers.                                                                          */
/*   MAXD   compile-time bound of the pool slots; EXACT: every integrand uses exactly MAXD slots */
/*          (no run-time predication on the dof at all)                                         */
/*   HMODE  HIST_NONE | HIST_SMEM_ATOMIC (one CTA-shared fp64 histogram, red.shared.add.f64)     */
/*          | HIST_SMEM_STRIPED (the same in 16 bank-striped copies, one 1024-thread CTA per SM)  */
/*          | HIST_GLOBAL_RED (RED.E.ADD.F64 into replicated L2-resident copies)                 */
/* ------------------------------------------------------------------------------------------ */

/* hist[bin] += v on the CTA's shared histogram.  ptxas expands red.shared.add.f64 into
 * LDS.64 / DADD / ATOMS.CAST.SPIN.64 / retry-branch; handing it a ready 32-bit shared address keeps the
 * expansion at 7 instructions (the generic-pointer form recomputes the shared window base every time). */
__device__ __forceinline__ void red_shared_f64(uint32_t saddr, double v) {
  asm volatile("red.shared.add.f64 [%0], %1;" ::"r"(saddr), "d"(v) : "memory");
}

/* hist[bin[p]] += v for p0 <= p < min(p0 + B, n) on the striped shared histogram, B updates in flight together:
 * the B loads are issued back to back, then the B compare-and-swaps, so one sample pays ~2 shared-memory round trips
 * per batch instead of 2 per update (ncu: the serial LDS -> DADD -> CAS chains were 1/3 of a warp's time, stall
 * short_scoreboard).  A CAS that lost a race — another thread, or an earlier slot of this thread on the same bin —
 * falls back to the looping red.shared.add.f64. */
template <int B, int MAXD, bool EXACT>
__device__ __forceinline__ void hist_batch(uint32_t hist_saddr, const int *bin, int p0, int n, double v) {
  uint32_t a[B];
  unsigned long long old[B], got[B];
#pragma unroll
  for (int q = 0; q < B; ++q)
    if (p0 + q < MAXD && (EXACT || p0 + q < n)) {
      a[q] = hist_saddr + (8u * STRIPE) * (uint32_t)bin[p0 + q];
      asm volatile("ld.shared.b64 %0, [%1];" : "=l"(old[q]) : "r"(a[q]) : "memory");
    }
#pragma unroll
  for (int q = 0; q < B; ++q)
    if (p0 + q < MAXD && (EXACT || p0 + q < n)) {
      const unsigned long long nw = (unsigned long long)__double_as_longlong(__longlong_as_double((long long)old[q]) + v);
      asm volatile("atom.shared.cas.b64 %0, [%1], %2, %3;" : "=l"(got[q]) : "r"(a[q]), "l"(old[q]), "l"(nw) : "memory");
    }
#if MCI_LOST_LOP3
  /* did any of the B swaps lose?  OR of the XORs: two LOP3 per update (the 64-bit compares cost 3.6 ISETP each) */
  unsigned long long diff = 0;
#pragma unroll
  for (int q = 0; q < B; ++q)
    if (p0 + q < MAXD && (EXACT || p0 + q < n)) diff |= got[q] ^ old[q];
  const bool lost = diff != 0;
#else
  bool lost = false;
#pragma unroll
  for (int q = 0; q < B; ++q)
    if (p0 + q < MAXD && (EXACT || p0 + q < n)) lost = lost || (got[q] != old[q]);
#endif
  if (lost) {
#pragma unroll
    for (int q = 0; q < B; ++q)
      if (p0 + q < MAXD && (EXACT || p0 + q < n))
        if (got[q] != old[q]) red_shared_f64(a[q], v);
  }
}

/* HIST_SMEM_FIXED: hist[bin] += X on a 64-bit fixed-point cell of two u32 limbs (lo at saddr, hi at saddr + 4).  The
 * low limb is added with the returning form of the native shared-memory integer add; its carry — (old + lo) wrapped —
 * joins the high limb, which a second, non-returning add applies.  Integer adds commute: the cell ends up with the
 * exact sum whatever the interleaving of threads, so the histogram is bit-reproducible. */
/* one such update (the generic kernel: bins of different leaves, run-time layout) */
__device__ __forceinline__ void fixed_add1(uint32_t saddr, uint32_t xlo, uint32_t xhi, uint32_t &ovf) {
  uint32_t old;
  asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(saddr), "r"(xlo) : "memory");
  asm volatile("{\n\t.reg .u32 t, h, oh;\n\tadd.cc.u32 t, %1, %2;\n\taddc.u32 h, %3, 0;\n\t"
               "atom.shared.add.u32 oh, [%4+4], h;\n\tadd.cc.u32 t, oh, h;\n\taddc.u32 %0, %0, 0;\n\t}"
               : "+r"(ovf) : "r"(old), "r"(xlo), "r"(xhi), "r"(saddr) : "memory");
}

template <int B, int MAXD, bool EXACT>
__device__ __forceinline__ void fixed_batch(uint32_t hist_saddr, const int *bin, int p0, int n, uint32_t xlo, uint32_t xhi, uint32_t &ovf) {
  uint32_t a[B], old[B];
#pragma unroll
  for (int q = 0; q < B; ++q) /* the B low-limb adds go out back to back: their return values are independent */
    if (p0 + q < MAXD && (EXACT || p0 + q < n)) {
      a[q] = hist_saddr + (8u * STRIPE) * (uint32_t)bin[p0 + q];
      asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old[q]) : "r"(a[q]), "r"(xlo) : "memory");
    }
#pragma unroll
  for (int q = 0; q < B; ++q) /* carry + high limb */
    if (p0 + q < MAXD && (EXACT || p0 + q < n)) {
#if MCI_FIXED_HI_ALWAYS
      /* unconditional (ptxas turns a predicated shared-memory atomic into a branch around it, +4 instructions, which
       * costs more than the add it skips); the returning form, so that an overflow of the cell — the high limb wrapping —
       * is seen by the add that causes it: 7 instructions per update */
      asm volatile("{\n\t.reg .u32 t, h, oh;\n\tadd.cc.u32 t, %1, %2;\n\taddc.u32 h, %3, 0;\n\t"
                   "atom.shared.add.u32 oh, [%4+4], h;\n\tadd.cc.u32 t, oh, h;\n\taddc.u32 %0, %0, 0;\n\t}"
                   : "+r"(ovf) : "r"(old[q]), "r"(xlo), "r"(xhi), "r"(a[q]) : "memory");
#else
      asm volatile("{\n\t.reg .u32 t, h;\n\t.reg .pred q;\n\tadd.cc.u32 t, %0, %1;\n\taddc.u32 h, %2, 0;\n\t"
                   "setp.ne.u32 q, h, 0;\n\t@q red.shared.add.u32 [%3+4], h;\n\t}"
                   ::"r"(old[q]), "r"(xlo), "r"(xhi), "r"(a[q]) : "memory");
#endif
    }
}



/* The pool's map in shared memory.
 *   REP == 1: one table of (edge, width) pairs, one LDS.128 per draw.  32 lanes at random increments conflict on the 8
 *             16-byte bank groups: ~10 wavefronts per instruction — with the histogram the shared-memory pipe was 80 %
 *             busy (ncu, profiles/r1_fill_ncu.json), THE limiter of the kernel.
 *   REP == 4: the same pairs in 4 copies interleaved [increment][copy]; lane l reads copy l % 4.  The 8 lanes of one
 *             LDS.128 phase fall on the 4 copies two by two, and a copy owns the 16-byte bank groups c and c + 4, so a
 *             phase takes 2 wavefronts whenever one of its four lane pairs differs in increment parity: ~7.8 wavefronts
 *             per draw instead of ~10.9, no extra instruction, 48 KB more shared memory (affordable in the striped
 *             variant, which owns the SM).  An earlier split edge / width variant (two LDS.64 per draw, 6.4
 *             wavefronts) was measured neutral: C2 38.5 vs 38.8 G samples/s. */
template <int REP>
struct PoolMap {
  const double2 *pair; /* this lane's copy of pair 0; pairs of one copy are REP entries apart */
  __device__ __forceinline__ void lookup(int iy, double &g, double &dx) const {
    const double2 v = pair[iy * REP];
    g = v.x; dx = v.y;
  }
};

/* jac = prod_p N dx_p accumulated as (N^h dx_1 ... dx_h) (N^(d-h) dx_(h+1) ... dx_d), h = ceil(d / 2): one DMUL per
 * slot instead of two (N dx, then jac *= it).  The powers of N enter in two halves so that the partial products stay
 * as far from the over- / underflow thresholds as the per-slot products N dx themselves (N^12 ~ 1e36 at most for the
 * largest pool); the association differs from the reference's prod 1/prob by a few ulp (L0: jac rtol 1e-13). */
template <int MAXD, bool EXACT>
struct PoolJac {
  int h; /* the powers themselves are read from the problem descriptor (constant bank): no registers held across the loop */
  __device__ __forceinline__ void init(int d) { h = (d + 1) >> 1; }
};

/* the draws of one sample on the pool: slots 1..d from Philox calls (s, j, block), two slots per call
 * (vegas/montecarlo.jl:122-131).  Shared by the fill kernel and by the pool-path trace kernel (L0 parity probe). */
template <int MAXD, bool EXACT, bool CAN_PAD, int REP>
__device__ __forceinline__ void pool_draw(const Problem &P, const PhiloxKey &K, const PoolMap<REP> &M, const PoolJac<MAXD, EXACT> &J,
                                          uint64_t s, uint32_t block, PoolX<MAXD> &X, int *bin, double *prob, double &jac) {
  /* Which of the two trims pays depends on the register budget (64 at 1024 threads per CTA): the fma index keeps N in a
   * register pair for the whole sample, the halved Jacobian product changes ptxas' schedule.  Measured per pool size
   * (A/B of the four combinations in one job, B200, G samples/s, fp64 striped / fixed-point histogram; profiles/
   * r2_pool_trim_variants.jsonl): d = 10 both (43.9 / 53.0 against 41.1 / 51.1 with neither); d = 16 the fma index alone
   * (27.0 / 34.8; both: 22.8 / 34.6 — the fp64 variant spills); d = 20 the halved product alone (20.0 / 21.9; both: 16.3 / 21.7) */
  constexpr bool FMA_INDEX = MCI_POOL_FMA_INDEX && MAXD <= 16;
  constexpr bool HALF_JAC = MCI_POOL_HALF_JAC && (MAXD <= 12 || MAXD > 16);
  const double nbins_d = P.leaf[0].nbins_d;
  const double neg_nbins_d = -nbins_d;
  const int d = EXACT ? MAXD : P.maxdof[0];
  jac = HALF_JAC ? P.pool_npow[0] : 1.0;
#pragma unroll
  for (int j = 0; j < (MAXD + 1) / 2; ++j) {
    if (2 * j < d) {
      uint32_t o0, o1, o2, o3;
      philox4x32_10((uint32_t)s, (uint32_t)(s >> 32), (uint32_t)j, block, K, o0, o1, o2, o3);
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int p = 2 * j + e;
        if (p < MAXD && p < d) {
          double dx, dy, g;
          if constexpr (FMA_INDEX) map_index_m(nbins_d, neg_nbins_d, e ? m52(o2, o3) : m52(o0, o1), bin[p], dy);
          else map_index(P.leaf[0].nbins, nbins_d, e ? u52(o2, o3) : u52(o0, o1), bin[p], dy);
          M.lookup(bin[p], g, dx);
          X.v[p] = __dadd_rn(g, __dmul_rn(dy, dx));
          if constexpr (HALF_JAC) {
            if (p == J.h) jac *= P.pool_npow[1];
            jac *= dx;
            if (CAN_PAD && P.need_pad) prob[p] = 1.0 / __dmul_rn(nbins_d, dx);
          } else {
            const double ndx = __dmul_rn(nbins_d, dx);
            jac *= ndx;
            if (CAN_PAD && P.need_pad) prob[p] = 1.0 / ndx;
          }
        }
      }
    }
  }
  if (HALF_JAC && J.h >= d) jac *= P.pool_npow[1]; /* d == 1: the second half is empty (N^0 = 1) */
}

template <int FID, int MAXD, int HMODE, bool EXACT, bool GUARD, int REP>
__device__ __forceinline__ void pool_sample(const Problem &P, const PhiloxKey &K, const PoolMap<REP> &M, const PoolJac<MAXD, EXACT> &J,
                                            uint32_t hist_saddr, double *__restrict__ ghist, uint64_t s, uint32_t block,
                                            bool active, double *acc, int &cnt, unsigned long long &bad, double fx_scale,
                                            uint32_t &fx_ovf) {
  using F = Integrand<FID>;
  constexpr int NI = F::NI;
  constexpr bool CAN_PAD = NI > 1 && !EXACT; /* one integrand, or EXACT: dof == maxdof by construction */
  const int d = EXACT ? MAXD : P.maxdof[0];
  PoolX<MAXD> X;
  int bin[MAXD];
  double prob[CAN_PAD ? MAXD : 1]; /* only live when P.need_pad */
  double jac;
  pool_draw<MAXD, EXACT, CAN_PAD, REP>(P, K, M, J, s, block, X, bin, prob, jac);
  double w[NI];
  F::template eval<MAXD, PoolX<MAXD>, EXACT>(X, P, w);
  double pad[NI];
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    pad[i] = 1.0;
    if (CAN_PAD && P.need_pad && !P.same_dof[i]) {
      const int di = P.dof[i * P.nv];
#pragma unroll
      for (int p = 0; p < MAXD; ++p)
        if (p >= di && p < d) pad[i] *= prob[p];
    }
  }
  bool finite = true;
#pragma unroll
  for (int i = 0; i < NI; ++i) finite = finite && isfinite(w[i]);
  if ((!GUARD || active) && !finite) bad++;
  if (!GUARD || active) { /* measurefreq == 1 on this kernel (the planner sends other values to the generic kernel) */
#pragma unroll
    for (int i = 0; i < NI; ++i) acc[i] += w[i] * pad[i] * jac;
    cnt++;
  }
  if (HMODE != HIST_NONE && (!GUARD || active)) {
    /* hist[bin_p] += (|w_i| jac)^2 for every integrand i with p < dof_i.  When every integrand uses every slot (EXACT)
     * the contributions to a slot are summed first: one update per slot instead of one per (integrand, slot) */
    constexpr int NPASS = EXACT ? 1 : NI;
#pragma unroll
    for (int i = 0; i < NPASS; ++i) {
      double w2;
      if (EXACT) {
        w2 = 0.0;
#pragma unroll
        for (int k = 0; k < NI; ++k) { const double a = fabs(w[k]) * jac; w2 += a * a; }
      } else {
        const double a = fabs(w[i]) * jac;
        w2 = a * a;
      }
      const int di = EXACT ? MAXD : P.dof[i * P.nv];
      if constexpr (HMODE == HIST_SMEM_FIXED) {
        /* one conversion per (sample, pass): X = round-to-nearest(w2 2^E); every slot then adds the same integer */
        const double xs = w2 * fx_scale;
        const bool in_range = xs < 9.0e18; /* below 2^63 (false for NaN too): a larger weight is reported, not wrapped */
        const unsigned long long Xi = in_range ? __double2ull_rn(xs) : 0ull;
        if (!in_range) fx_ovf = 1u;
        const uint32_t xlo = (uint32_t)Xi, xhi = (uint32_t)(Xi >> 32);
#pragma unroll
        for (int p0 = 0; p0 < MAXD; p0 += MCI_HIST_BATCH) fixed_batch<MCI_HIST_BATCH, MAXD, EXACT>(hist_saddr, bin, p0, di, xlo, xhi, fx_ovf);
      } else if constexpr (HMODE == HIST_SMEM_STRIPED && MCI_HIST_BATCH > 1) {
#pragma unroll
        for (int p0 = 0; p0 < MAXD; p0 += MCI_HIST_BATCH) hist_batch<MCI_HIST_BATCH, MAXD, EXACT>(hist_saddr, bin, p0, di, w2);
      } else {
#pragma unroll
        for (int p = 0; p < MAXD; ++p)
          if (p < di) {
            if (HMODE == HIST_SMEM_ATOMIC) red_shared_f64(hist_saddr + 8u * (uint32_t)bin[p], w2);
            else if (HMODE == HIST_SMEM_STRIPED) red_shared_f64(hist_saddr + (8u * STRIPE) * (uint32_t)bin[p], w2); /* base holds the lane's stripe */
            else atomicAdd(ghist + bin[p], w2); /* RED.E.ADD.F64 */
          }
      }
    }
  }
}

template <int MAXD, int THREADS>
struct PoolMinBlocks {
  static constexpr int value = THREADS >= 1024 ? 1 : (MAXD <= 12 ? 1024 / THREADS : (MAXD <= 16 ? 768 / THREADS : 512 / THREADS));
};

template <int FID, int MAXD, int HMODE, int THREADS, bool EXACT>
__global__ void __launch_bounds__(THREADS, PoolMinBlocks<MAXD, THREADS>::value) vegas_pool_kernel(const __grid_constant__ DevProblem Pp, const FillArgs A) {
  MCI_VIEW(P, Pp);
  using F = Integrand<FID>;
  constexpr int NI = F::NI;
  constexpr int NWARP = THREADS / 32;
  /* HIST_SMEM_FIXED handles launch the fixed-point kernel AND, behind it, the fp64 striped kernel every iteration.  The scale
   * the device chose after the previous iteration says whose turn it is: valid -> fixed-point, invalid -> fp64, the other
   * returns at once.  The fp64 kernel is also the RESCUE: if the fixed-point kernel before it reported that a cell or a weight
   * left the 64-bit range (flags[2]), it recomputes the whole iteration — same samples, same block sums, fp64 histogram —
   * and says so in flags[3].  No host round trip, so the asynchronous and the fused drivers are covered as well. */
  bool rescue = false;
  if (A.fixed_exp != nullptr) {
    const bool valid = *A.fixed_exp != FIXED_EXP_INVALID;
    if (HMODE == HIST_SMEM_FIXED) { if (!valid) return; }
    else {
      rescue = valid && *reinterpret_cast<volatile unsigned long long *>(A.flags + 2) != 0ull;
      if (valid && !rescue) return;
      if (rescue && blockIdx.x == 0 && threadIdx.x == 0) A.flags[3] = 1ull;
    }
  }
  static_assert(HMODE == HIST_NONE || HMODE == HIST_SMEM_ATOMIC || HMODE == HIST_SMEM_STRIPED || HMODE == HIST_GLOBAL_RED ||
                    HMODE == HIST_SMEM_FIXED, "pool kernel histogram modes");
  static_assert(IntegrandWD<FID>::value == 1, "complex integrands run on the generic kernel");
  constexpr bool STRIPED = HMODE == HIST_SMEM_STRIPED || HMODE == HIST_SMEM_FIXED;
  constexpr int NCOPY = STRIPED ? STRIPE : (HMODE == HIST_SMEM_ATOMIC ? 1 : 0);
  constexpr int REP = STRIPED ? POOL_MAP_REP : 1; /* the striped variants own the SM: room for map copies */
  extern __shared__ __align__(16) double smem[];
  double *smaps = smem;                                 /* REP * P.map_doubles                    */
  double *shist = smem + REP * P.map_doubles;           /* the CTA's histogram: [bin][copy]       */
  double *sred = shist + NCOPY * P.hist_len;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool adapt = P.leaf[0].adapt != 0;

  PoolMap<REP> M;
  if (REP == 1) {
    stage_maps(P, smaps, A.gmaps);
    M.pair = reinterpret_cast<const double2 *>(smaps);
  } else { /* (edge, width) pairs in REP copies, interleaved [increment][copy]: lane l reads copy l % REP */
    const int nb = P.leaf[0].nbins;
    const double2 *src = reinterpret_cast<const double2 *>(A.gmaps);
    double2 *dst = reinterpret_cast<double2 *>(smaps);
    for (int i = tid; i < nb * REP; i += THREADS) dst[i] = __ldg(src + i / REP);
    M.pair = dst + (lane & (REP - 1));
  }
  for (int i = tid; i < NCOPY * P.hist_len; i += THREADS) shist[i] = 0.0;
  __syncthreads();
  const uint32_t hist_saddr = (uint32_t)__cvta_generic_to_shared(shist) + (STRIPED ? 8u * (lane & (STRIPE - 1)) : 0u);
  /* fixed-point cells: scale 2^E of this iteration, chosen on the device after the previous one (fixed_scale_kernel) */
  int fx_exp = 0;
  double fx_scale = 1.0;
  uint32_t fx_ovf = 0u;
  if (HMODE == HIST_SMEM_FIXED) {
    fx_exp = *A.fixed_exp + A.fixed_bias;
    fx_scale = scalbn(1.0, fx_exp);
  }
  double *ghist = A.cta_hist + (size_t)(blockIdx.x % A.n_hist_copies) * P.hist_len;

#if MCI_POOL_KEY_IN_ARGS
  MCI_ROUND_KEYS(K, P, A);
#else
  PhiloxKey K;
  philox_expand_key(P.seed, A.iter, K);
#endif
  unsigned long long bad = 0;
  PoolJac<MAXD, EXACT> J;
  J.init(EXACT ? MAXD : P.maxdof[0]);

  const int n_tiles = A.n_local_blocks * A.tiles_per_block;
  const uint64_t tile_samples = (uint64_t)THREADS * A.spt;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int lb = tile / A.tiles_per_block;
    const int tb = tile - lb * A.tiles_per_block;
    const uint32_t block = (uint32_t)(A.rank + lb * A.world);
    const uint64_t s0 = (uint64_t)tb * tile_samples;
    double acc[NI];
#pragma unroll
    for (int i = 0; i < NI; ++i) acc[i] = 0.0;
    int cnt = 0;
    if (s0 + tile_samples <= A.neval_per_block) { /* full tile: no per-lane guards anywhere */
      if (adapt) {
        for (int k = 0; k < A.spt; ++k)
          pool_sample<FID, MAXD, HMODE, EXACT, false, REP>(P, K, M, J, hist_saddr, ghist, s0 + (uint64_t)k * THREADS + tid, block,
                                                       true, acc, cnt, bad, fx_scale, fx_ovf);
      } else {
        for (int k = 0; k < A.spt; ++k)
          pool_sample<FID, MAXD, HIST_NONE, EXACT, false, REP>(P, K, M, J, hist_saddr, ghist, s0 + (uint64_t)k * THREADS + tid,
                                                           block, true, acc, cnt, bad, fx_scale, fx_ovf);
      }
    } else { /* ragged last tile of a block */
      for (int k = 0; k < A.spt; ++k) {
        const uint64_t s = s0 + (uint64_t)k * THREADS + tid;
        const bool active = s < A.neval_per_block;
        if (adapt) pool_sample<FID, MAXD, HMODE, EXACT, true, REP>(P, K, M, J, hist_saddr, ghist, s, block, active, acc, cnt, bad, fx_scale, fx_ovf);
        else pool_sample<FID, MAXD, HIST_NONE, EXACT, true, REP>(P, K, M, J, hist_saddr, ghist, s, block, active, acc, cnt, bad, fx_scale, fx_ovf);
      }
    }
    /* tile flush: warp shuffle -> shared -> one thread sums the warps in order */
#pragma unroll
    for (int i = 0; i < NI; ++i) acc[i] = warp_sum(acc[i]);
    int wc = __reduce_add_sync(0xffffffffu, cnt);
    __syncthreads();
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < NI; ++i) sred[warp * (NI + 1) + i] = acc[i];
      sred[warp * (NI + 1) + NI] = (double)wc;
    }
    __syncthreads();
    if (tid <= NI) {
      double t = 0.0;
      for (int wv = 0; wv < NWARP; ++wv) t += sred[wv * (NI + 1) + tid];
      const size_t gt = (size_t)lb * A.tiles_per_block + tb;
      if (tid < NI) {
        /* default measure: observable i is obs_off[i] (obs_len == 1 on this path) */
        A.tile_obs[gt * P.obs_total + P.obs_off[tid]] = t;
      } else {
        A.tile_cnt[gt] = t;
      }
    }
  }
  if (bad && !rescue) atomicAdd(A.flags, bad); /* (the fixed-point launch of this iteration has counted them already) */
  if (HMODE == HIST_SMEM_FIXED && fx_ovf) atomicAdd(A.flags + 2, 1ull); /* a weight beyond 2^63 units or a cell that wrapped */
  /* histogram flush: one global partial per CTA (stripes summed in a fixed order) */
  if (NCOPY > 0) {
    __syncthreads();
    double *out = A.cta_hist + (size_t)blockIdx.x * P.hist_len;
    if (HMODE == HIST_SMEM_FIXED) { /* integer sum of the stripes (exact), one conversion, one exact power-of-two scaling */
      const unsigned long long *cells = reinterpret_cast<const unsigned long long *>(shist); /* lo limb first: little endian */
      const double inv = scalbn(1.0, -fx_exp);
      for (int i = tid; i < P.hist_len; i += THREADS) {
        unsigned long long t = 0ull;
#pragma unroll
        for (int c = 0; c < NCOPY; ++c) t += cells[i * NCOPY + c];
        out[i] = (double)t * inv;
      }
    } else {
      for (int i = tid; i < P.hist_len; i += THREADS) {
        double t = 0.0;
#pragma unroll
        for (int c = 0; c < NCOPY; ++c) t += shist[i * NCOPY + c];
        out[i] = t;
      }
    }
  }
}

/* ------------------------------------------------------------------------------------------ */
/* one sample on an arbitrary layout (several variables, CompositeVar, Discrete leaves).         */
/* Shared by the generic fill kernel and by the trace kernel (mci_trace_samples).                */
/* Draw order = top-level variable, pool slot, leaf (vegas/montecarlo.jl:122-131,                */
/* sampler.jl:431-440).  tprob[] = probability of each (variable, slot), only when need_pad.     */
/* ------------------------------------------------------------------------------------------ */
template <int FID>
__device__ __forceinline__ void vegas_sample_generic(const Problem &P, const double *maps, const PhiloxKey &K,
                                                     uint64_t s, uint32_t block, double *x, int *bin, double *tprob,
                                                     double &jac, double *w, double *pad) {
  using F = Integrand<FID>;
  jac = 1.0;
  int k = 0;
  uint32_t o0 = 0, o1 = 0, o2 = 0, o3 = 0;
  MCI_UNROLL_LAYOUT
  for (int v = 0; v < P.nv; ++v) {
    MCI_UNROLL_LAYOUT
    for (int p = 0; p < P.maxdof[v]; ++p) {
      double tp = 1.0;
      MCI_UNROLL_LAYOUT
      for (int l = 0; l < P.top_nleaf[v]; ++l, ++k) {
        const DevLeaf &L = P.leaf[P.top_first[v] + l];
        double u;
        if ((k & 1) == 0) {
          philox4x32_10((uint32_t)s, (uint32_t)(s >> 32), (uint32_t)(k >> 1), block, K, o0, o1, o2, o3);
          u = u52(o0, o1);
        } else {
          u = u52(o2, o3);
        }
        if (L.kind == MCI_LEAF_CONTINUOUS) {
          double dx;
          x[k] = map_continuous(reinterpret_cast<const double2 *>(maps + L.smem_off), L.nbins, L.nbins_d, u, bin[k], dx);
          double ndx = __dmul_rn(L.nbins_d, dx);
          jac *= ndx;
          if (P.need_pad) tp *= 1.0 / ndx;
        } else {
          const double *dist = maps + L.smem_off;
          const double *acc = dist + L.nbins;
          int g = locate_clamped(acc, L.nbins, u);
          x[k] = (double)(L.lower_i + g);
          bin[k] = g;
          double pr = dist[g];
          jac /= pr;
          if (P.need_pad) tp *= pr;
        }
      }
      if (P.need_pad) tprob[P.top_tslot[v] + p] = tp;
    }
  }
  FlatX X{x, &P};
  F::template eval<0>(X, P, w);
  MCI_UNROLL_LAYOUT
  for (int i = 0; i < P.ni; ++i) {
    double pd = 1.0;
    if (P.need_pad && !P.same_dof[i]) {
      MCI_UNROLL_LAYOUT
      for (int v = 0; v < P.nv; ++v) {
        MCI_UNROLL_LAYOUT
        for (int p = P.dof[i * P.nv + v]; p < P.maxdof[v]; ++p) pd *= tprob[P.top_tslot[v] + p];
      }
    }
    pad[i] = pd;
  }
}

template <int FID, int THREADS>
__global__ void __launch_bounds__(THREADS) vegas_generic_kernel(const __grid_constant__ DevProblem Pp, const FillArgs A) {
  MCI_VIEW(P, Pp);
  using F = Integrand<FID>;
  constexpr int NI = F::NI;
  constexpr int WD = IntegrandWD<FID>::value; /* doubles per value: 2 = complex (re, im) */
  constexpr int NWARP = THREADS / 32;
  extern __shared__ __align__(16) double smem[];
  double *smaps = smem;
  const int map_smem = A.maps_in_smem ? P.map_doubles : 0;
  double *shist = smem + map_smem;
  /* HIST_SMEM_FIXED: one CTA-shared histogram of 64-bit fixed-point cells (native integer adds, mci_device.cuh) while the
   * device-side scale is valid; the same shared array as fp64 cells with compare-and-swap adds while it is not */
  int hmode = A.hist_mode, fx_exp = 0;
  double fx_scale = 1.0;
  uint32_t fx_ovf = 0u;
  const bool fixed_req = hmode == HIST_SMEM_FIXED;
  if (fixed_req) {
    fx_exp = *A.fixed_exp;
    if (A.rescue) { /* second launch of the iteration: recompute in fp64 if the first one left the fixed-point range */
      if (fx_exp == FIXED_EXP_INVALID || *reinterpret_cast<volatile unsigned long long *>(A.flags + 2) == 0ull) return;
      if (blockIdx.x == 0 && threadIdx.x == 0) A.flags[3] = 1ull;
      hmode = HIST_SMEM_ATOMIC;
    } else if (fx_exp == FIXED_EXP_INVALID) hmode = HIST_SMEM_ATOMIC;
    else { fx_exp += A.fixed_bias; fx_scale = scalbn(1.0, fx_exp); }
  }
  const bool one_smem_copy = hmode == HIST_SMEM_ATOMIC || hmode == HIST_SMEM_FIXED;
  double *sred = shist + (hmode == HIST_WARP_PRIVATE ? NWARP * P.hist_len : (one_smem_copy ? P.hist_len : 0));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (A.maps_in_smem) stage_maps(P, smaps, A.gmaps);
  const int nsh = hmode == HIST_WARP_PRIVATE ? NWARP * P.hist_len : (one_smem_copy ? P.hist_len : 0);
  for (int i = tid; i < nsh; i += THREADS) shist[i] = 0.0;
  __syncthreads();
  const double *maps = A.maps_in_smem ? smaps : A.gmaps;
  double *myhist = hmode == HIST_WARP_PRIVATE ? shist + warp * P.hist_len
                   : one_smem_copy ? shist
                                               : A.cta_hist + (size_t)(blockIdx.x % A.n_hist_copies) * P.hist_len;
  MCI_ROUND_KEYS(K, P, A);
  unsigned long long bad = 0;
  const int OT = P.obs_total;

  const int n_tiles = A.n_local_blocks * A.tiles_per_block;
  const uint64_t tile_samples = (uint64_t)THREADS * A.spt;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int lb = tile / A.tiles_per_block;
    const int tb = tile - lb * A.tiles_per_block;
    const uint32_t block = (uint32_t)(A.rank + lb * A.world);
    const uint64_t s0 = (uint64_t)tb * tile_samples;
    double acc[MCI_MAX_OBS];
    for (int o = 0; o < OT; ++o) acc[o] = 0.0;
    int cnt = 0;
    for (int k = 0; k < A.spt; ++k) {
      const uint64_t s = s0 + (uint64_t)k * THREADS + tid;
      const bool active = s < A.neval_per_block;
      if (!__any_sync(0xffffffffu, active)) break;
      double x[MAX_DRAW], tprob[MAX_DRAW], w[NI * WD], pad[NI], jac;
      int bin[MAX_DRAW];
      vegas_sample_generic<FID>(P, maps, K, s, block, x, bin, tprob, jac, w, pad);
      bool finite = true;
#pragma unroll
      for (int i = 0; i < NI * WD; ++i) finite = finite && isfinite(w[i]);
      if (active && !finite) bad++;
      const bool meas = active && (P.measurefreq == 1 || ((s + 1) % (uint64_t)P.measurefreq) == 0);
      if (meas) {
        if (P.measure_kind == MCI_MEASURE_DEFAULT) {
#pragma unroll
          for (int i = 0; i < NI; ++i)
#pragma unroll
            for (int c = 0; c < WD; ++c) acc[P.obs_off[i] + c] += w[WD * i + c] * pad[i] * jac;
        } else if (P.measure_kind == MCI_MEASURE_BIN_BY_DISCRETE) {
          /* obs[0][Ext - lower] += relw[0]; the binning leaf is the last leaf, slot 1 (example/bubble.jl:81-84) */
#pragma unroll
          for (int c = 0; c < WD; ++c) acc[P.obs_off[0] + WD * bin[P.leaf[P.nleaf - 1].base] + c] += w[c] * pad[0] * jac;
        } else { /* the caller's measure(x, obs, relw, config) (main.jl:35-37), compiled from source */
          double relw[NI * WD];
#pragma unroll
          for (int i = 0; i < NI * WD; ++i) relw[i] = w[i] * pad[i / WD] * jac;
          FlatX X{x, &P};
          UserMeasureT<0>::call(X, acc, relw, P.userdata);
        }
        cnt++;
      }
      if (hmode == HIST_SMEM_FIXED) {
        const uint32_t cells = (uint32_t)__cvta_generic_to_shared(shist);
        MCI_UNROLL_LAYOUT
        for (int i = 0; i < NI; ++i) {
          const double a = wabs<WD>(w, i) * jac;
          const double xs = a * a * fx_scale; /* one conversion per (sample, integrand): every slot adds the same integer */
          const bool in_range = xs < 9.0e18;
          const unsigned long long Xi = in_range ? __double2ull_rn(xs) : 0ull;
          if (active && !in_range) fx_ovf = 1u;
          const uint32_t xlo = (uint32_t)Xi, xhi = (uint32_t)(Xi >> 32);
          MCI_UNROLL_LAYOUT
          for (int v = 0; v < P.nv; ++v)
            MCI_UNROLL_LAYOUT
            for (int p = 0; p < P.dof[i * P.nv + v]; ++p)
              MCI_UNROLL_LAYOUT
              for (int l = 0; l < P.top_nleaf[v]; ++l) {
                const DevLeaf &L = P.leaf[P.top_first[v] + l];
                if (L.adapt && active) fixed_add1(cells + 8u * (uint32_t)(L.hist_off + bin[L.base + p * L.stride]), xlo, xhi, fx_ovf);
              }
        }
      } else if (hmode != HIST_NONE) {
        MCI_UNROLL_LAYOUT
        for (int v = 0; v < P.nv; ++v)
          MCI_UNROLL_LAYOUT
          for (int i = 0; i < NI; ++i) {
            const double a = wabs<WD>(w, i) * jac;
            const double w2 = a * a;
            MCI_UNROLL_LAYOUT
            for (int p = 0; p < P.dof[i * P.nv + v]; ++p)
              MCI_UNROLL_LAYOUT
              for (int l = 0; l < P.top_nleaf[v]; ++l) {
                const DevLeaf &L = P.leaf[P.top_first[v] + l];
                if (L.adapt) hist_add_rt(hmode, myhist + L.hist_off, bin[L.base + p * L.stride], w2, active);
              }
          }
      }
    }
    int wc = __reduce_add_sync(0xffffffffu, cnt);
    __syncthreads();
    for (int o = 0; o < OT; ++o) {
      double t = warp_sum(acc[o]);
      if (lane == 0) sred[warp * (OT + 1) + o] = t;
    }
    if (lane == 0) sred[warp * (OT + 1) + OT] = (double)wc;
    __syncthreads();
    const size_t gt = (size_t)lb * A.tiles_per_block + tb;
    for (int o = tid; o <= OT; o += THREADS) {
      double t = 0.0;
      for (int wv = 0; wv < NWARP; ++wv) t += sred[wv * (OT + 1) + o];
      if (o < OT) A.tile_obs[gt * OT + o] = t;
      else A.tile_cnt[gt] = t;
    }
  }
  if (bad && !(fixed_req && A.rescue)) atomicAdd(A.flags, bad);
  if (fx_ovf) atomicAdd(A.flags + 2, 1ull);
  if (hmode == HIST_SMEM_FIXED) { /* one conversion per bin and an exact power-of-two scaling */
    __syncthreads();
    double *out = A.cta_hist + (size_t)blockIdx.x * P.hist_len;
    const unsigned long long *cells = reinterpret_cast<const unsigned long long *>(shist);
    const double inv = scalbn(1.0, -fx_exp);
    for (int i = tid; i < P.hist_len; i += THREADS) out[i] = (double)cells[i] * inv;
  } else if (hmode == HIST_WARP_PRIVATE || hmode == HIST_SMEM_ATOMIC) {
    __syncthreads();
    double *out = A.cta_hist + (size_t)blockIdx.x * P.hist_len;
    for (int i = tid; i < P.hist_len; i += THREADS) {
      double t = 0.0;
      if (hmode == HIST_WARP_PRIVATE) {
        for (int wv = 0; wv < NWARP; ++wv) t += shist[wv * P.hist_len + i];
      } else {
        t = shist[i];
      }
      out[i] = t;
    }
  }
}

/* L0 parity probe (mci_trace_samples): regenerate `count` samples with the generic sampler */
template <int FID>
__global__ void vegas_trace_kernel(const __grid_constant__ DevProblem Pp, const double *gmaps, uint64_t iter, uint32_t block,
                                   uint64_t first, int count, double *xo, int *bo, double *jo, double *wo) {
  MCI_VIEW(P, Pp);
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count) return;
  PhiloxKey K;
  philox_expand_key(P.seed, iter, K);
  constexpr int WD = IntegrandWD<FID>::value;
  double x[MAX_DRAW], tprob[MAX_DRAW], w[Integrand<FID>::NI * WD], pad[Integrand<FID>::NI], jac;
  int bin[MAX_DRAW];
  vegas_sample_generic<FID>(P, gmaps, K, first + t, block, x, bin, tprob, jac, w, pad);
  for (int k = 0; k < P.n_draw; ++k) { xo[(size_t)t * P.n_draw + k] = x[k]; bo[(size_t)t * P.n_draw + k] = bin[k]; }
  jo[t] = jac;
  for (int i = 0; i < P.ni * WD; ++i) wo[(size_t)t * P.ni * WD + i] = w[i]; /* complex: (re, im) interleaved */
}

/* L0 parity probe of the POOL path: the very draw code of vegas_pool_kernel (pool_draw: fma map index, unclamped
 * increment, halved Jacobian products) on the staged map image in global memory, plus the integrand */
template <int FID, int MAXD>
__global__ void vegas_pool_trace_kernel(const __grid_constant__ DevProblem Pp, const double *gmaps, uint64_t iter, uint32_t block,
                                        uint64_t first, int count, double *xo, int *bo, double *jo, double *wo) {
  MCI_VIEW(P, Pp);
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count) return;
  constexpr int NI = Integrand<FID>::NI;
  PhiloxKey K;
  philox_expand_key(P.seed, iter, K);
  PoolMap<1> M;
  M.pair = reinterpret_cast<const double2 *>(gmaps);
  PoolJac<MAXD, false> J;
  J.init(P.maxdof[0]);
  PoolX<MAXD> X;
  int bin[MAXD];
  double prob[1], jac, w[NI];
#pragma unroll
  for (int p = 0; p < MAXD; ++p) { X.v[p] = 0.0; bin[p] = 0; }
  pool_draw<MAXD, false, false, 1>(P, K, M, J, first + t, block, X, bin, prob, jac);
  Integrand<FID>::template eval<MAXD, PoolX<MAXD>, false>(X, P, w);
  const int d = P.maxdof[0];
#pragma unroll
  for (int p = 0; p < MAXD; ++p)
    if (p < d) { xo[(size_t)t * d + p] = X.v[p]; bo[(size_t)t * d + p] = bin[p]; }
  jo[t] = jac;
  for (int i = 0; i < NI; ++i) wo[(size_t)t * NI + i] = w[i];
}

}  // namespace mci
