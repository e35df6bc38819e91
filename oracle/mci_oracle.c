/*
 * mci_oracle.c — CPU ORACLE (test infrastructure, NOT product; see mci_oracle.h).
 *
 * Plain-C restatement of the sampling-and-accumulation hot path of
 * numericalEFT/MCIntegration.jl v0.4.2.  Citations "file:line" are relative to
 * the reference repository.  Operation order follows the reference; floating
 * point contraction must be OFF when compiling (-ffp-contract=off) because
 * Julia does not fuse a*b+c.
 *
 * Deliberate, documented differences from the reference:
 *  D1  RNG: Philox4x32-10 counter streams replace Random.MersenneTwister
 *      (configuration.jl:40).  Stream = (seed, iteration, block, index) where
 *      index = sample number (vegas) or walker number (vegasmc/mcmc).
 *      52-bit uniforms k*2^-52 like Julia's rand(Float64).
 *      rand(rng, 1:n) -> 1 + floor(u*n).
 *  D2  A block of the chain solvers is `walkers_per_block` independent chains of
 *      neval_per_block / walkers_per_block steps each (walkers=1 is the
 *      reference's single chain).  Burn-in and the 1 % measurement delay apply
 *      per chain.
 *  D3  initialize! draws only slots 1..maxdof (the reference draws every pool
 *      slot, variable.jl:576-580; the extra slots are never read).
 *  D4  locate() clamps instead of throwing when u >= accumulation[end]
 *      (common.jl:10-12; accumulation[end] is only approximately 1).
 *  D5  vegas shift!: default shift_mode 0 recomputes prob = 1/(N*dx) (the
 *      create! formula, sampler.jl:303) for every draw; shift_mode 1 keeps the
 *      reference's running ratio product (sampler.jl:383-384).
 *  D6  Floors of clearStatistics! (configuration.jl:238-250) are applied once per
 *      iteration as (n_block+1)*floor on the summed raw statistics.
 *  D7  With a fixed seed the reference gives every MPI rank the same stream
 *      (main.jl:136); here streams are keyed by the global block index.
 *  D8  FermiK (sampler.jl:109-281) is accepted under :mcmc only, the one solver
 *      the reference runs it under (test/bubble_FermiK.jl:2,139-140); K.prob,
 *      which only :vegas/:vegasmc would read, is not kept.
 */
#include "mci_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define TINY (4.9406564584124654e-324 * 1e50) /* MCIntegration.jl:11  eps(0.0)*1e50 */
#define EPS0 4.9406564584124654e-324          /* eps(0.0) */
#define PI 3.14159265358979323846

/* ------------------------------------------------------------------------- */
/* RNG (D1)                                                                   */
/* ------------------------------------------------------------------------- */
void orc_philox4x32_10(const uint32_t ctr_in[4], const uint32_t key_in[2], uint32_t out[4]) {
  uint32_t c0 = ctr_in[0], c1 = ctr_in[1], c2 = ctr_in[2], c3 = ctr_in[3];
  uint32_t k0 = key_in[0], k1 = key_in[1];
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

typedef struct {
  uint32_t key[2];
  uint32_t ctr[4]; /* ctr[2] = next pair number */
  uint32_t buf[4];
  int have; /* 0 or 1 buffered uniforms */
} orng;

static void orng_init(orng *r, uint64_t seed, uint64_t iter, uint32_t block, uint64_t idx) {
  r->key[0] = (uint32_t)seed;
  r->key[1] = (uint32_t)(seed >> 32) ^ ((uint32_t)iter * 0x9E3779B1u);
  r->ctr[0] = (uint32_t)idx;
  r->ctr[1] = (uint32_t)(idx >> 32);
  r->ctr[2] = 0;
  r->ctr[3] = block;
  r->have = 0;
}

static inline double u52(uint32_t hi, uint32_t lo) {
  uint64_t bits = (((uint64_t)hi << 32) | lo) >> 12;
  return (double)bits * 0x1p-52;
}

/* rand(config.rng) */
static double orng_rand(orng *r) {
  if (r->have) {
    r->have = 0;
    return u52(r->buf[2], r->buf[3]);
  }
  orc_philox4x32_10(r->ctr, r->key, r->buf);
  r->ctr[2] += 1;
  r->have = 1;
  return u52(r->buf[0], r->buf[1]);
}

/* rand(config.rng, 1:n) */
static int orng_int(orng *r, int n) {
  int k = (int)floor(orng_rand(r) * n) + 1;
  return k > n ? n : k;
}

void orc_uniforms(uint64_t seed, uint64_t iter, uint32_t block, uint64_t idx, int32_t n, double *out) {
  orng r;
  orng_init(&r, seed, iter, block, idx);
  for (int i = 0; i < n; ++i) out[i] = orng_rand(&r);
}

/* ------------------------------------------------------------------------- */
/* distribution/common.jl                                                     */
/* ------------------------------------------------------------------------- */

/* locate: common.jl:8-36.  Returns 1-based jl with acc[jl] <= p < acc[jl+1]; -1 where the reference throws. */
int32_t orc_locate(const double *acc, int32_t n, double p) {
  if (acc[0] > p || acc[n - 1] <= p) return -1;
  int jl = 1, ju = n + 1;
  while (ju - jl > 1) {
    int jm = (jl + ju) / 2;
    if (p < acc[jm - 1]) ju = jm; else jl = jm;
  }
  return jl;
}

/* D4: same bisection, clamped into [1, n-1] */
static int locate_clamped(const double *acc, int n, double p) {
  int jl = 1, ju = n; /* acc has n entries; valid answers 1..n-1 */
  if (p < acc[0]) return 1;
  while (ju - jl > 1) {
    int jm = (jl + ju) / 2;
    if (p < acc[jm - 1]) ju = jm; else jl = jm;
  }
  return jl;
}

/* smooth: common.jl:43-54 */
void orc_smooth(const double *dist, int32_t n, double factor, double *out) {
  if (n <= 1) { for (int i = 0; i < n; ++i) out[i] = dist[i]; return; }
  out[0] = (dist[0] * (factor + 1) + dist[1]) / (factor + 2);
  out[n - 1] = (dist[n - 1] * (factor + 1) + dist[n - 2]) / (factor + 2);
  for (int i = 1; i < n - 1; ++i) out[i] = (dist[i - 1] + dist[i] * factor + dist[i + 1]) / (factor + 2);
}

/* rescale: common.jl:67-83.  Returns -1 where the reference asserts. */
int orc_rescale(double *dist, int32_t n, double alpha) {
  if (n == 1) return 0;
  for (int i = 0; i < n; ++i) if (!(dist[i] > 0)) return -1;
  double s = 0.0;
  for (int i = 0; i < n; ++i) s += dist[i];
  for (int i = 0; i < n; ++i) dist[i] /= s;
  for (int i = 0; i < n; ++i)
    if (dist[i] > 0 && dist[i] <= 0.99999999) dist[i] = pow(-(1 - dist[i]) / log(dist[i]), alpha);
  for (int i = 0; i < n; ++i) if (!isfinite(dist[i])) return -1;
  return 0;
}

/* ------------------------------------------------------------------------- */
/* distribution/variable.jl : train!                                          */
/* ------------------------------------------------------------------------- */

/* train!(::Continuous): variable.jl:206-239 */
int orc_train_continuous(const double *grid, int32_t npts, const double *hist, double alpha, double *newgrid) {
  int n = npts - 1;
  for (int i = 0; i < n; ++i) if (!isfinite(hist[i]) || !(hist[i] > 0)) return -1; /* :212-213 */
  double *avg_f = (double *)malloc(sizeof(double) * n);
  orc_smooth(hist, n, 6.0, avg_f);
  if (orc_rescale(avg_f, n, alpha) != 0) { free(avg_f); return -1; }
  newgrid[0] = grid[0];
  newgrid[npts - 1] = grid[npts - 1];
  int j = 0;          /* 1-based index into avg_f as in the reference */
  double acc_f = 0.0;
  double sum = 0.0;
  for (int i = 0; i < n; ++i) sum += avg_f[i];
  double f_ninc = sum / (npts - 1); /* :226 */
  for (int i = 2; i <= npts - 1; ++i) {
    while (acc_f < f_ninc) {
      j += 1;
      if (j > n) { j = n; break; } /* the reference has no guard (Appendix B.14); never hit for valid input */
      acc_f += avg_f[j - 1];
    }
    acc_f -= f_ninc;
    newgrid[i - 1] = grid[j] - (acc_f / avg_f[j - 1]) * (grid[j] - grid[j - 1]); /* :233 */
  }
  newgrid[npts - 1] = grid[npts - 1];
  free(avg_f);
  return 0;
}

/* train!(::Discrete): variable.jl:369-382 */
int orc_train_discrete(const double *hist, int32_t n, double alpha, double *dist, double *accum) {
  for (int i = 0; i < n; ++i) dist[i] = hist[i];
  if (orc_rescale(dist, n, alpha) != 0) return -1;
  double s = 0.0;
  for (int i = 0; i < n; ++i) s += dist[i];
  for (int i = 0; i < n; ++i) dist[i] /= s;
  accum[0] = 0.0;
  for (int i = 1; i <= n; ++i) { /* sum(distribution[1:i]) recomputed for every i (:376) */
    double a = 0.0;
    for (int k = 0; k < i; ++k) a += dist[k];
    accum[i] = a;
  }
  return 0;
}

/* create!(::Continuous) map: sampler.jl:293-305 */
void orc_map_continuous(const double *grid, int32_t npts, double y, int32_t *iy_out, double *x, double *prob) {
  int N = npts - 1;
  int iy = (int)floor(y * N) + 1;
  if (iy > N) iy = N; /* cannot happen for y<1 (checked), kept as a guard */
  double dy = y * N - (iy - 1);
  *x = grid[iy - 1] + dy * (grid[iy] - grid[iy - 1]);
  *iy_out = iy;
  *prob = 1.0 / (N * (grid[iy] - grid[iy - 1]));
}

/* _maxdof: configuration.jl:229-236 */
void orc_maxdof(const int32_t *dof, int32_t nd, int32_t nv, int32_t *maxdof) {
  for (int v = 0; v < nv; ++v) {
    int m = dof[v];
    for (int d = 1; d < nd; ++d) if (dof[d * nv + v] > m) m = dof[d * nv + v];
    maxdof[v] = m;
  }
}

/* _mean_std: main.jl:296-320 (real observables) */
void orc_mean_std(const double *sum, const double *sqsum, int32_t n, int32_t block, double *mean, double *std) {
  for (int i = 0; i < n; ++i) {
    mean[i] = sum[i] / block;
    if (block > 1) {
      double v = (sqsum[i] / block - mean[i] * mean[i]) / (block - 1);
      std[i] = v < 0.0 ? 0.0 : sqrt(v);
    } else {
      std[i] = 0.0;
    }
  }
}

/* average: statistics.jl:186-220 (real data).  init is 1-based; max = niter. */
void orc_average(const double *mean, const double *std, int32_t niter, int32_t init, double *avg, double *err,
                 double *chi2) {
  int max = niter;
  if (max <= init) { *avg = mean[0]; *err = std[0]; *chi2 = 0.0; return; } /* :189-191 */
  double wsum = 0.0;
  for (int i = init; i <= max; ++i) { double s = std[i - 1] + 1.0e-10; wsum += 1.0 / (s * s); }
  double m = 0.0;
  for (int i = init; i <= max; ++i) { double s = std[i - 1] + 1.0e-10; m += mean[i - 1] * (1.0 / (s * s)) / wsum; }
  double c = 0.0;
  for (int i = init; i <= max; ++i) {
    double s = std[i - 1] + 1.0e-10;
    double d = mean[i - 1] - m;
    c += (1.0 / (s * s)) * d * d;
  }
  *avg = m;
  *err = 1.0 / sqrt(wsum);
  *chi2 = c / ((max - init + 1) - 1);
}

/* doReweight!: main.jl:322-346 */
void orc_do_reweight(double *reweight, const double *visited, int32_t nd, double gamma, const double *goal) {
  double avgstep = 0.0;
  for (int i = 0; i < nd; ++i) avgstep += visited[i];
  for (int i = 0; i < nd; ++i) {
    if (visited[i] <= 1) reweight[i] *= pow(avgstep, gamma);
    else reweight[i] *= pow(avgstep / visited[i], gamma);
  }
  if (goal) {
    double gs = 0.0;
    for (int i = 0; i < nd; ++i) gs += goal[i];
    for (int i = 0; i < nd; ++i) reweight[i] *= goal[i] / gs;
  }
  double s = 0.0;
  for (int i = 0; i < nd; ++i) s += reweight[i];
  for (int i = 0; i < nd; ++i) reweight[i] /= s;
}

/* ------------------------------------------------------------------------- */
/* variable pools (variable.jl structs) and sampler.jl operations             */
/* ------------------------------------------------------------------------- */
typedef struct {
  int kind, nbins, adapt, lower_i;
  double lower, upper, alpha;
  const double *grid;  /* nbins+1 (Continuous) */
  const double *dist;  /* nbins   (Discrete)   */
  const double *accum; /* nbins+1 (Discrete)   */
  double *hist;        /* nbins, private to the running config */
  /* pool: 1-based index i -> [i-1]; last element is the rollback cache (variable.jl:139,301) */
  double *data;
  int *gidx;
  double *prob;
  int len;
  int off; /* offset of the owning top-level variable */
} oleaf;

typedef struct {
  int n_leaf, first_leaf, offset, len, composite, adapt;
  int fermik; /* FermiK{D}: the n_leaf = D leaves are its Cartesian components (variable.jl:1-21) */
  double *prob; /* composite: own array; otherwise aliases the leaf's */
  double prob_cache;
} otop;

struct orc_handle;

typedef struct oconfig {
  const struct orc_handle *h;
  int ni, nd, nv, nleaf;
  otop top[MCI_MAX_TOPVAR];
  oleaf leaf[MCI_MAX_LEAF];
  const int *dof;    /* [nd][nv] incl. normalisation row */
  const int *maxdof; /* [nv] */
  const double *reweight;
  const double *userdata;
  double *hist;      /* concat */
  double *obs;       /* concat, sum obs_len */
  double normalization;
  double *visited, *propose, *accept;
  uint64_t neval;
  orng rng;
  int shift_mode;
  int wd;       /* doubles per integrand value: 1 real, 2 complex (re, im) — Configuration(type=ComplexF64) */
  int counting; /* 0 during the extra burn-in steps of mci_problem_t.chain_burnin: they touch no statistic */
  void *mem; /* single allocation backing the arrays */
} oconfig;

/* abs(weights[i]): Julia's abs(::Complex) is hypot(re, im) */
static inline double wabs(const oconfig *c, const double *w, int i) {
  return c->wd == 1 ? fabs(w[i]) : hypot(w[2 * i], w[2 * i + 1]);
}

/* create!(::Continuous): sampler.jl:293-305 */
static double cont_create(oleaf *T, int idx, orng *rng) {
  int N = T->nbins;
  double y = orng_rand(rng);
  int iy = (int)floor(y * N) + 1;
  if (iy > N) iy = N;
  double dy = y * N - (iy - 1);
  double x = T->grid[iy - 1] + dy * (T->grid[iy] - T->grid[iy - 1]);
  T->data[idx - 1] = x;
  T->gidx[idx - 1] = iy;
  T->prob[idx - 1] = 1.0 / (N * (T->grid[iy] - T->grid[iy - 1]));
  return 1.0 / T->prob[idx - 1];
}
/* remove!(::Continuous): sampler.jl:317-322 */
static double cont_remove(oleaf *T, int idx) {
  int iy = T->gidx[idx - 1];
  return 1 / ((T->grid[iy] - T->grid[iy - 1]) * T->nbins);
}
/* shift!(::Continuous): sampler.jl:336-386 */
static double cont_shift(oleaf *T, int idx, orng *rng) {
  int e = T->len - 1;
  T->data[e] = T->data[idx - 1];
  T->gidx[e] = T->gidx[idx - 1];
  T->prob[e] = T->prob[idx - 1];
  int curr = T->gidx[idx - 1];
  int N = T->nbins;
  double y = orng_rand(rng);
  int iy = (int)floor(y * N) + 1;
  if (iy > N) iy = N;
  double dy = y * N - (iy - 1);
  double x = T->grid[iy - 1] + dy * (T->grid[iy] - T->grid[iy - 1]);
  T->data[idx - 1] = x;
  T->gidx[idx - 1] = iy;
  double prob_ratio = (T->grid[curr] - T->grid[curr - 1]) / (T->grid[iy] - T->grid[iy - 1]);
  T->prob[idx - 1] *= prob_ratio;
  return 1.0 / prob_ratio;
}
static void cont_shift_rollback(oleaf *T, int idx) { /* sampler.jl:388-393 */
  int e = T->len - 1;
  T->data[idx - 1] = T->data[e];
  T->gidx[idx - 1] = T->gidx[e];
  T->prob[idx - 1] = T->prob[e];
}
/* create!(::Discrete): sampler.jl:13-22 (+D4) */
static double disc_create(oleaf *d, int idx, orng *rng) {
  int g = locate_clamped(d->accum, d->nbins + 1, orng_rand(rng));
  d->data[idx - 1] = (double)(d->lower_i + g - 1);
  d->gidx[idx - 1] = g;
  d->prob[idx - 1] = d->dist[g - 1];
  return 1.0 / d->dist[g - 1];
}
static double disc_remove(oleaf *d, int idx) { /* sampler.jl:37-41 */
  int g = (int)d->data[idx - 1] - d->lower_i + 1;
  return d->dist[g - 1];
}
/* shift!(::Discrete): sampler.jl:57-71 */
static double disc_shift(oleaf *d, int idx, orng *rng) {
  int e = d->len - 1;
  d->data[e] = d->data[idx - 1];
  d->gidx[e] = d->gidx[idx - 1];
  d->prob[e] = d->prob[idx - 1];
  int curr = (int)d->data[idx - 1] - d->lower_i + 1;
  int g = locate_clamped(d->accum, d->nbins + 1, orng_rand(rng));
  d->data[idx - 1] = (double)(d->lower_i + g - 1);
  d->gidx[idx - 1] = g;
  double ratio = d->dist[g - 1] / d->dist[curr - 1];
  d->prob[idx - 1] *= ratio;
  return 1.0 / ratio;
}
static void disc_shift_rollback(oleaf *d, int idx) { /* sampler.jl:73-77 */
  int e = d->len - 1;
  d->data[idx - 1] = d->data[e];
  d->gidx[idx - 1] = d->gidx[e];
  d->prob[idx - 1] = d->prob[e];
}
static void leaf_swap(oleaf *T, int i1, int i2) { /* sampler.jl:85-97, 395-408 */
  double t = T->data[i1 - 1]; T->data[i1 - 1] = T->data[i2 - 1]; T->data[i2 - 1] = t;
  int g = T->gidx[i1 - 1]; T->gidx[i1 - 1] = T->gidx[i2 - 1]; T->gidx[i2 - 1] = g;
  double p = T->prob[i1 - 1]; T->prob[i1 - 1] = T->prob[i2 - 1]; T->prob[i2 - 1] = p;
}
static double leaf_create(oleaf *l, int idx, orng *rng) { return l->kind == MCI_LEAF_CONTINUOUS ? cont_create(l, idx, rng) : disc_create(l, idx, rng); }
static double leaf_remove(oleaf *l, int idx) { return l->kind == MCI_LEAF_CONTINUOUS ? cont_remove(l, idx) : disc_remove(l, idx); }
static double leaf_shift(oleaf *l, int idx, orng *rng) { return l->kind == MCI_LEAF_CONTINUOUS ? cont_shift(l, idx, rng) : disc_shift(l, idx, rng); }
static void leaf_shift_rollback(oleaf *l, int idx) { if (l->kind == MCI_LEAF_CONTINUOUS) cont_shift_rollback(l, idx); else disc_shift_rollback(l, idx); }

/* ---- FermiK{D}: sampler.jl:109-281.  Component c of K[idx] is leaf[first + c].data[idx - 1]; leaf.lower = kF,
 * leaf.upper = dk.  K.prob is read by :vegas/:vegasmc only, under which the reference itself declares FermiK broken
 * (test/bubble_FermiK.jl:2) and which mci_create / orc_create refuse, so it is not kept. ---- */
static double fermik_create(oleaf *K, int D, int idx, orng *rng) { /* :109-149 */
  double kF = K[0].lower, dk = K[0].upper;
  double Kamp = kF + (orng_rand(rng) - 0.5) * 2.0 * dk;
  if (Kamp <= 0.0) return 0.0;
  double phi = 2 * PI * orng_rand(rng);
  if (D == 3) {
    double theta = PI * orng_rand(rng);
    K[0].data[idx - 1] = Kamp * cos(phi) * sin(theta);
    K[1].data[idx - 1] = Kamp * sin(phi) * sin(theta);
    K[2].data[idx - 1] = Kamp * cos(theta);
    return 2 * dk * (2 * PI) * PI * (sin(theta) * (Kamp * Kamp));
  }
  K[0].data[idx - 1] = Kamp * cos(phi);
  K[1].data[idx - 1] = Kamp * sin(phi);
  return 2 * dk * (2 * PI) * Kamp;
}
static double fermik_remove(oleaf *K, int D, int idx) { /* :160-189 */
  double kF = K[0].lower, dk = K[0].upper;
  double k1 = K[0].data[idx - 1], k2 = K[1].data[idx - 1];
  if (D == 3) {
    double k3 = K[2].data[idx - 1];
    double Kamp = sqrt(k1 * k1 + k2 * k2 + k3 * k3);
    if (!(kF - dk < Kamp && Kamp < kF + dk)) return 0.0;
    double sint = sqrt(k1 * k1 + k2 * k2) / Kamp;
    if (sint < 1.0e-15) return 0.0;
    return 1.0 / (2 * dk * (2 * PI) * PI * sint * (Kamp * Kamp));
  }
  double Kamp = sqrt(k1 * k1 + k2 * k2);
  if (!(kF - dk < Kamp && Kamp < kF + dk)) return 0.0;
  return 1.0 / (2 * dk * (2 * PI) * Kamp);
}
static double fermik_shift(oleaf *K, int D, int idx, orng *rng) { /* :199-251 */
  int e = K[0].len - 1;
  for (int c = 0; c < D; ++c) K[c].data[e] = K[c].data[idx - 1]; /* K[end] = K[idx] */
  double x = orng_rand(rng);
  if (x < 1.0 / 3) {
    double lambda = 1.5;
    double ratio = 1.0 / lambda + orng_rand(rng) * (lambda - 1.0 / lambda);
    for (int c = 0; c < D; ++c) K[c].data[idx - 1] *= ratio;
    return D == 2 ? 1.0 : ratio;
  } else if (x < 2.0 / 3) {
    double phi = orng_rand(rng) * (2 * PI);
    double k1 = K[0].data[idx - 1], k2 = K[1].data[idx - 1];
    if (D == 3) {
      double theta = acos(1 - 2 * orng_rand(rng));
      double k3 = K[2].data[idx - 1];
      double Kamp = sqrt(k1 * k1 + k2 * k2 + k3 * k3);
      K[0].data[idx - 1] = Kamp * cos(phi) * sin(theta);
      K[1].data[idx - 1] = Kamp * sin(phi) * sin(theta);
      K[2].data[idx - 1] = Kamp * cos(theta);
    } else {
      double Kamp = sqrt(k1 * k1 + k2 * k2);
      K[0].data[idx - 1] = Kamp * cos(phi);
      K[1].data[idx - 1] = Kamp * sin(phi);
    }
    return 1.0;
  }
  for (int c = 0; c < D; ++c) K[c].data[idx - 1] = K[c].data[idx - 1] + (orng_rand(rng) - 0.5) * K[0].upper;
  return 1.0;
}
static void fermik_shift_rollback(oleaf *K, int D, int idx) { /* :253-257 */
  int e = K[0].len - 1;
  for (int c = 0; c < D; ++c) K[c].data[idx - 1] = K[c].data[e];
}

/* top-level dispatch incl. CompositeVar: sampler.jl:410-462 */
static double var_create(oconfig *c, int vi, int idx) {
  otop *T = &c->top[vi];
  if (T->fermik) return fermik_create(&c->leaf[T->first_leaf], T->n_leaf, idx, &c->rng);
  if (!T->composite) return leaf_create(&c->leaf[T->first_leaf], idx, &c->rng);
  double prop = 1.0;
  T->prob[idx - 1] = 1.0;
  for (int k = 0; k < T->n_leaf; ++k) {
    oleaf *v = &c->leaf[T->first_leaf + k];
    prop *= leaf_create(v, idx, &c->rng);
    T->prob[idx - 1] *= v->prob[idx - 1];
  }
  return prop;
}
static double var_remove(oconfig *c, int vi, int idx) {
  otop *T = &c->top[vi];
  if (T->fermik) return fermik_remove(&c->leaf[T->first_leaf], T->n_leaf, idx);
  if (!T->composite) return leaf_remove(&c->leaf[T->first_leaf], idx);
  double prop = 1.0;
  for (int k = 0; k < T->n_leaf; ++k) prop *= leaf_remove(&c->leaf[T->first_leaf + k], idx);
  return prop;
}
static double var_shift(oconfig *c, int vi, int idx) {
  otop *T = &c->top[vi];
  if (T->fermik) return fermik_shift(&c->leaf[T->first_leaf], T->n_leaf, idx, &c->rng);
  if (!T->composite) return leaf_shift(&c->leaf[T->first_leaf], idx, &c->rng);
  double prop = 1.0;
  T->prob_cache = T->prob[idx - 1];
  T->prob[idx - 1] = 1.0;
  for (int k = 0; k < T->n_leaf; ++k) {
    oleaf *v = &c->leaf[T->first_leaf + k];
    prop *= leaf_shift(v, idx, &c->rng);
    T->prob[idx - 1] *= v->prob[idx - 1];
  }
  return prop;
}
static void var_shift_rollback(oconfig *c, int vi, int idx) {
  otop *T = &c->top[vi];
  if (T->fermik) { fermik_shift_rollback(&c->leaf[T->first_leaf], T->n_leaf, idx); return; }
  for (int k = 0; k < T->n_leaf; ++k) leaf_shift_rollback(&c->leaf[T->first_leaf + k], idx);
  if (T->composite) T->prob[idx - 1] = T->prob_cache;
}
static double var_swap(oconfig *c, int vi, int i1, int i2) {
  otop *T = &c->top[vi];
  if (T->composite) { double p = T->prob[i1 - 1]; T->prob[i1 - 1] = T->prob[i2 - 1]; T->prob[i2 - 1] = p; }
  for (int k = 0; k < T->n_leaf; ++k) leaf_swap(&c->leaf[T->first_leaf + k], i1, i2);
  return 1.0;
}
/* a swap is its own inverse: swapRollback! sampler.jl:92-97,402-408,457-462 */
#define var_swap_rollback var_swap

/* accumulate!: variable.jl:196-200 (Continuous), :362-367 (Discrete), :474-478 (Composite) */
static void var_accumulate(oconfig *c, int vi, int idx, double w) {
  otop *T = &c->top[vi];
  for (int k = 0; k < T->n_leaf; ++k) {
    oleaf *v = &c->leaf[T->first_leaf + k];
    if (!v->adapt) continue;
    if (v->kind == MCI_LEAF_CONTINUOUS) v->hist[v->gidx[idx - 1] - 1] += w;
    else v->hist[(int)v->data[idx - 1] - v->lower_i] += w;
  }
}

/* initialize!: variable.jl:576-580 and :497-507 (Composite), with D3 */
static void var_initialize(oconfig *c, int vi) {
  otop *T = &c->top[vi];
  int last = T->offset + c->maxdof[vi];
  if (T->fermik) { /* the generic initialize!(T, config), variable.jl:576-580 */
    for (int i = 1 + T->offset; i <= last; ++i) fermik_create(&c->leaf[T->first_leaf], T->n_leaf, i, &c->rng);
    return;
  }
  for (int k = 0; k < T->n_leaf; ++k)
    for (int i = 1 + T->offset; i <= last; ++i) leaf_create(&c->leaf[T->first_leaf + k], i, &c->rng);
  if (T->composite)
    for (int i = 1 + T->offset; i <= last; ++i) {
      T->prob[i - 1] = 1.0;
      for (int k = 0; k < T->n_leaf; ++k) T->prob[i - 1] *= c->leaf[T->first_leaf + k].prob[i - 1];
    }
}

/* probability helpers: variable.jl:587-641 */
static double total_probability(const oconfig *c) {
  double prob = 1.0;
  for (int vi = 0; vi < c->nv; ++vi)
    for (int pos = 1; pos <= c->maxdof[vi]; ++pos) prob *= c->top[vi].prob[pos + c->top[vi].offset - 1];
  return prob;
}
static double probability(const oconfig *c, int i /*0-based integrand*/) {
  double prob = 1.0;
  for (int vi = 0; vi < c->nv; ++vi)
    for (int pos = 1; pos <= c->dof[i * c->nv + vi]; ++pos) prob *= c->top[vi].prob[pos + c->top[vi].offset - 1];
  return prob;
}
static double padding_probability(const oconfig *c, int i) {
  double prob = 1.0;
  for (int vi = 0; vi < c->nv; ++vi)
    for (int pos = c->dof[i * c->nv + vi] + 1; pos <= c->maxdof[vi]; ++pos)
      prob *= c->top[vi].prob[pos + c->top[vi].offset - 1];
  return prob;
}

/* ------------------------------------------------------------------------- */
/* integrands (restated from the reference's tests / examples)                */
/* ------------------------------------------------------------------------- */
/* X(c, leaf, pos): value of pool slot pos (1-based, before offset) of flat leaf index */
static inline double X(const oconfig *c, int leaf, int pos) {
  /* all leaves of one top variable share its offset (variable.jl:415-427) */
  return c->leaf[leaf].data[pos + c->leaf[leaf].off - 1];
}

/* example/bubble.jl:40-50 */
static double green(double tau, double w, double beta) {
  if (tau >= 0.0)
    return w > 0.0 ? exp(-w * tau) / (1 + exp(-w * beta)) : exp(w * (beta - tau)) / (1 + exp(w * beta));
  else
    return w > 0.0 ? -exp(-w * (tau + beta)) / (1 + exp(-w * beta)) : -exp(-w * tau) / (1 + exp(w * beta));
}

/* test/montecarlo.jl:201-204 */
static double volume_inverse(int d) {
  const double euler = 2.71828182845904523536028747135266249775724709369995957496696763;
  return pow(d / (2 * PI * euler), d / 2.0) * sqrt((double)d) * sqrt(PI);
}

/* evaluate all integrands (vegas / vegasmc convention f(x, c)); w has ni entries ((re, im) pairs when complex) */
static void eval_all(const oconfig *c, int id, double *w) {
  const double *ud = c->userdata;
  switch (id) {
  case MCI_F_GAUSS_ND: { /* BASELINE C2: (2pi)^(-d/2) exp(-sum x^2 / 2) */
    int d = c->dof[0];
    double s = 0.0;
    for (int i = 1; i <= d; ++i) { double x = X(c, 0, i); s += x * x; }
    w[0] = pow(2 * PI, -0.5 * d) * exp(-0.5 * s);
    break; }
  case MCI_F_LOG_OVER_SQRT: { /* test/montecarlo.jl:129-134 */
    double x = X(c, 0, 1);
    w[0] = log(x) / sqrt(x);
    break; }
  case MCI_F_GENZ_CORNER_POOL: { /* SURVEY §8d C3, layout A */
    int d = c->dof[0];
    double s = 1.0;
    for (int i = 1; i <= d; ++i) s += ud[i - 1] * X(c, 0, i);
    w[0] = pow(s, -(double)(d + 1));
    break; }
  case MCI_F_GENZ_CORNER_COMP: { /* layout B: d leaves, one slot each */
    int d = c->nleaf;
    double s = 1.0;
    for (int i = 0; i < d; ++i) s += ud[i] * X(c, i, 1);
    w[0] = pow(s, -(double)(d + 1));
    break; }
  case MCI_F_BUBBLE: { /* example/bubble.jl:50-75; userdata = [kF, beta, me, spin, dim, Qsize, extQx...] */
    double kF = ud[0], beta = ud[1], me = ud[2], spin = ud[3], dim = ud[4];
    double R = X(c, 0, 1), th = X(c, 1, 1), ph = X(c, 2, 1), T = X(c, 3, 1);
    int ext = (int)X(c, 4, 1);
    double r = R / (1 - R);
    double k0 = r * sin(th) * cos(ph), k1 = r * sin(th) * sin(ph), k2 = r * cos(th);
    double factor = 1.0 / pow(2 * PI, dim);
    factor *= r * r / ((1 - R) * (1 - R)) * sin(th);
    double q0 = ud[6 + ext - 1];
    double kq0 = k0 + q0, kq1 = k1 + 0.0, kq2 = k2 + 0.0;
    double tau = T - 0.0;
    double w1 = ((k0 * k0 + k1 * k1 + k2 * k2) - kF * kF) / (2 * me);
    double g1 = green(tau, w1, beta);
    double w2 = ((kq0 * kq0 + kq1 * kq1 + kq2 * kq2) - kF * kF) / (2 * me);
    double g2 = green(-tau, w2, beta);
    w[0] = g1 * g2 * spin * factor * cos(2 * PI * 0 * tau / beta);
    break; }
  case MCI_F_BUBBLE_FERMIK: { /* test/bubble_FermiK.jl:54-75, vars (T, K::FermiK{dim}, Ext); userdata as MCI_F_BUBBLE */
    int D = c->nleaf - 2;
    double kF = ud[0], beta = ud[1], me = ud[2], spin = ud[3];
    double tau = X(c, 0, 1), k0 = X(c, 1, 1), k1 = X(c, 2, 1), k2 = D == 3 ? X(c, 3, 1) : 0.0;
    int ext = (int)X(c, D + 1, 1);
    double kq0 = k0 + ud[6 + ext - 1];
    double w1 = ((k0 * k0 + k1 * k1 + k2 * k2) - kF * kF) / (2 * me);
    double g1 = green(tau, w1, beta);
    double w2 = ((kq0 * kq0 + k1 * k1 + k2 * k2) - kF * kF) / (2 * me);
    double g2 = green(-tau, w2, beta);
    double phase = 1.0 / pow(2 * PI, (double)D);
    w[0] = g1 * g2 * spin * phase * cos(2 * PI * 0 * tau / beta);
    break; }
  case MCI_F_PRODUCT_4: { /* SURVEY §8d C5 */
    int d = c->dof[0];
    for (int k = 0; k < 4; ++k) {
      double p = 1.0;
      for (int i = 1; i <= d; ++i) p *= (1 + ud[k] * (X(c, 0, i) - 0.5));
      w[k] = p;
    }
    break; }
  case MCI_F_SPHERE1: { /* test/montecarlo.jl:4-9 */
    double a = X(c, 0, 1), b = X(c, 0, 2);
    w[0] = (a * a + b * b < 1.0) ? 1.0 : 0.0;
    break; }
  case MCI_F_SPHERE2: { /* test/montecarlo.jl:19-24 */
    double a = X(c, 0, 1), b = X(c, 0, 2), z = X(c, 0, 3);
    w[0] = (a * a + b * b < 1.0) ? 1.0 : 0.0;
    w[1] = (a * a + b * b + z * z < 1.0) ? 1.0 : 0.0;
    break; }
  case MCI_F_DISCRETE_X: /* test/montecarlo.jl:104-111 */
    w[0] = X(c, 0, 1);
    break;
  case MCI_F_DISCRETE2_CONST: /* test/montecarlo.jl:113-120 */
    w[0] = 1.0;
    break;
  case MCI_F_LATTICE_POOL: /* test/montecarlo.jl:136-147 */
    w[0] = 1.0 / (1.0 - cos(X(c, 0, 1)) * cos(X(c, 0, 2)) * cos(X(c, 0, 3))) / (PI * PI * PI);
    break;
  case MCI_F_LATTICE_COMP: /* test/montecarlo.jl:149-164 */
    w[0] = 1.0 / (1.0 - cos(X(c, 0, 1)) * cos(X(c, 1, 1)) * cos(X(c, 2, 1))) / (PI * PI * PI);
    break;
  case MCI_F_HYPERSPHERE3: { /* test/montecarlo.jl:200-216 with N=3 */
    double x1 = X(c, 0, 1);
    double s = x1 * x1;
    for (int i = 1; i <= 3; ++i) {
      double xi = X(c, 0, i + 1);
      s += xi * xi;
      w[i - 1] = s < 1.0 ? volume_inverse(i + 1) : 0.0;
    }
    break; }
  case MCI_F_POLY3: { /* test/thread.jl: x^i */
    double x = X(c, 0, 1);
    w[0] = x; w[1] = x * x; w[2] = x * x * x;
    break; }
  case MCI_F_X2_PLUS_Y2: { /* vegas/montecarlo.jl:69 */
    double a = X(c, 0, 1), b = X(c, 0, 2);
    w[0] = a * a + b * b;
    break; }
  case MCI_F_COMPLEX1: { /* test/montecarlo.jl:166-170: x + x^2 im */
    double x = X(c, 0, 1);
    w[0] = x; w[1] = x * x;
    break; }
  case MCI_F_COMPLEX2: { /* test/montecarlo.jl:172-198: (x, x^2 im) */
    double x = X(c, 0, 1);
    w[0] = x; w[1] = 0.0; w[2] = 0.0; w[3] = x * x;
    break; }
  default:
    for (int i = 0; i < c->ni * c->wd; ++i) w[i] = NAN;
  }
}

/* mcmc convention f(idx, x, c): the reference's tests define it as the idx-th component; out has wd entries */
static void eval_one(const oconfig *c, int id, int idx /*0-based*/, double *out) {
  double w[2 * MCI_MAX_INTEGRAND];
  eval_all(c, id, w);
  for (int k = 0; k < c->wd; ++k) out[k] = w[c->wd * idx + k];
}

/* ------------------------------------------------------------------------- */
/* handle                                                                      */
/* ------------------------------------------------------------------------- */
struct orc_handle {
  mci_problem_t p; /* pointers re-targeted to owned copies */
  int nd, nv, ni, nleaf, pa_dim, pa_len, obs_total, hist_len, grid_len, disc_len;
  int topvar_nleaf[MCI_MAX_TOPVAR], topvar_offset[MCI_MAX_TOPVAR], topvar_first[MCI_MAX_TOPVAR];
  mci_leaf_t leaf[MCI_MAX_LEAF];
  int dof[(MCI_MAX_INTEGRAND + 1) * MCI_MAX_TOPVAR];
  int maxdof[MCI_MAX_TOPVAR];
  int obs_len[MCI_MAX_INTEGRAND], obs_off[MCI_MAX_INTEGRAND];
  int nb_ptr[MCI_MAX_INTEGRAND + 2], nb_idx[4 * (MCI_MAX_INTEGRAND + 1) * (MCI_MAX_INTEGRAND + 1)];
  double userdata[MCI_MAX_USERDATA];
  int leaf_grid_off[MCI_MAX_LEAF], leaf_disc_off[MCI_MAX_LEAF], leaf_accum_off[MCI_MAX_LEAF], leaf_hist_off[MCI_MAX_LEAF];
  double *grids, *disc_dist, *disc_accum, *reweight;
  int rank, world, shift_mode;
  int wd, start_idx; /* doubles per value (weight_kind); mcmc start integrand, 0-based */
  double *packed; int64_t packed_len;
  int off_hist, off_obs, off_cnt, off_visited, off_propose, off_accept, off_neval;
  char err[256];
};

static __thread char g_err[256];

const char *orc_last_error(const orc_handle *h) { return h ? h->err : g_err; }

int orc_create(const mci_problem_t *p, orc_handle **out) {
  if (!p || !out) { snprintf(g_err, sizeof g_err, "null argument"); return MCI_ERR_INVALID; }
  if (p->n_topvar < 1 || p->n_topvar > MCI_MAX_TOPVAR || p->n_leaf < 1 || p->n_leaf > MCI_MAX_LEAF ||
      p->n_integrand < 1 || p->n_integrand > MCI_MAX_INTEGRAND || p->n_block < 1 || p->measurefreq < 1) {
    snprintf(g_err, sizeof g_err, "problem dimensions out of range");
    return MCI_ERR_INVALID;
  }
  orc_handle *h = (orc_handle *)calloc(1, sizeof(orc_handle));
  h->p = *p;
  h->ni = p->n_integrand; h->nd = h->ni + 1; h->nv = p->n_topvar; h->nleaf = p->n_leaf;
  int first = 0;
  for (int v = 0; v < h->nv; ++v) {
    h->topvar_nleaf[v] = p->topvar_nleaf[v];
    h->topvar_offset[v] = p->topvar_offset ? p->topvar_offset[v] : 0;
    h->topvar_first[v] = first;
    first += p->topvar_nleaf[v];
  }
  if (first != h->nleaf) { snprintf(g_err, sizeof g_err, "sum(topvar_nleaf) != n_leaf"); free(h); return MCI_ERR_INVALID; }
  for (int l = 0; l < h->nleaf; ++l) {
    h->leaf[l] = p->leaf[l];
    if (p->leaf[l].kind == MCI_LEAF_FERMIK) { /* no map, no histogram; mcmc only (test/bubble_FermiK.jl:2,139-140) */
      h->leaf[l].nbins = 0; h->leaf[l].adapt = 0;
      if (p->solver != MCI_SOLVER_MCMC) {
        snprintf(g_err, sizeof g_err, "FermiK variables work under :mcmc only");
        free(h);
        return MCI_ERR_UNSUPPORTED;
      }
    }
  }
  for (int i = 0; i < h->ni; ++i)
    for (int v = 0; v < h->nv; ++v) h->dof[i * h->nv + v] = p->dof[i * h->nv + v];
  for (int v = 0; v < h->nv; ++v) h->dof[h->ni * h->nv + v] = 0; /* configuration.jl:153 */
  orc_maxdof(h->dof, h->nd, h->nv, h->maxdof);
  h->wd = p->weight_kind == MCI_WEIGHT_COMPLEX ? 2 : 1;
  h->start_idx = p->mcmc_start_idx > 0 ? p->mcmc_start_idx - 1 : 0; /* `idx` of MCMC.montecarlo, mcmc/montecarlo.jl:76 */
  h->obs_total = 0; /* in doubles: a complex entry is (re, im) */
  for (int i = 0; i < h->ni; ++i) { h->obs_len[i] = (p->obs_len ? p->obs_len[i] : 1) * h->wd; h->obs_off[i] = h->obs_total; h->obs_total += h->obs_len[i]; }
  /* neighbor: configuration.jl:201-209 default chain, or the given CSR */
  if (p->neighbor_ptr) {
    for (int d = 0; d <= h->nd; ++d) h->nb_ptr[d] = p->neighbor_ptr[d];
    for (int k = 0; k < h->nb_ptr[h->nd]; ++k) h->nb_idx[k] = p->neighbor_idx[k];
  } else {
    int Nd = h->nd, k = 0;
    for (int d = 1; d <= Nd; ++d) {
      h->nb_ptr[d - 1] = k;
      if (d == Nd) h->nb_idx[k++] = 0;                        /* norm -> first */
      else if (d == 1) { if (Nd == 2) h->nb_idx[k++] = 1; else { h->nb_idx[k++] = Nd - 1; h->nb_idx[k++] = 1; } }
      else if (d == Nd - 1) h->nb_idx[k++] = Nd - 3;          /* last diag -> second last */
      else { h->nb_idx[k++] = d - 2; h->nb_idx[k++] = d; }
    }
    h->nb_ptr[Nd] = k;
  }
  for (int i = 0; i < p->n_userdata && i < MCI_MAX_USERDATA; ++i) h->userdata[i] = p->userdata[i];
  int go = 0, dd = 0, da = 0, ho = 0;
  for (int l = 0; l < h->nleaf; ++l) {
    h->leaf_hist_off[l] = ho; ho += h->leaf[l].nbins;
    if (h->leaf[l].kind == MCI_LEAF_CONTINUOUS) { h->leaf_grid_off[l] = go; go += h->leaf[l].nbins + 1; }
    else if (h->leaf[l].kind == MCI_LEAF_DISCRETE) { h->leaf_disc_off[l] = dd; dd += h->leaf[l].nbins; h->leaf_accum_off[l] = da; da += h->leaf[l].nbins + 1; }
  }
  h->grid_len = go; h->disc_len = dd; h->hist_len = ho;
  h->grids = (double *)calloc(go + 1, sizeof(double));
  h->disc_dist = (double *)calloc(dd + 1, sizeof(double));
  h->disc_accum = (double *)calloc(da + 1, sizeof(double));
  h->reweight = (double *)calloc(h->nd, sizeof(double));
  for (int l = 0; l < h->nleaf; ++l) {
    mci_leaf_t *L = &h->leaf[l];
    if (L->kind == MCI_LEAF_CONTINUOUS) { /* collect(LinRange(lower, upper, ninc)) variable.jl:137 */
      int npts = L->nbins + 1;
      double *g = h->grids + h->leaf_grid_off[l];
      for (int i = 0; i < npts; ++i) {
        /* Julia's LinRange element: lerp (1-t)*a + t*b with t=i/(len-1) */
        double t = (double)i / (npts - 1);
        g[i] = (1 - t) * L->lower + t * L->upper;
      }
      g[0] = L->lower; g[npts - 1] = L->upper;
    } else if (L->kind == MCI_LEAF_DISCRETE) { /* variable.jl:305-315: uniform distribution and prefix sums */
      int n = L->nbins;
      double *d = h->disc_dist + h->leaf_disc_off[l], *a = h->disc_accum + h->leaf_accum_off[l];
      for (int i = 0; i < n; ++i) d[i] = 1.0 / n;
      a[0] = 0.0;
      for (int i = 1; i <= n; ++i) { double s = 0.0; for (int k = 0; k < i; ++k) s += d[k]; a[i] = s; }
    }
  }
  for (int d = 0; d < h->nd; ++d) h->reweight[d] = 1.0 / h->nd; /* configuration.jl:172-175 */
  h->pa_dim = h->nd > h->nv ? h->nd : h->nv;
  h->pa_len = 3 * h->nd * h->pa_dim;
  h->off_hist = 0;
  h->off_obs = h->off_hist + h->hist_len;
  h->off_cnt = h->off_obs + p->n_block * h->obs_total;
  h->off_visited = h->off_cnt + p->n_block;
  h->off_propose = h->off_visited + h->nd;
  h->off_accept = h->off_propose + h->pa_len;
  h->off_neval = h->off_accept + h->pa_len;
  h->packed_len = h->off_neval + 1;
  h->packed = (double *)calloc(h->packed_len, sizeof(double));
  h->rank = 0; h->world = 1; h->shift_mode = 0;
  /* retarget pointers */
  h->p.topvar_nleaf = h->topvar_nleaf; h->p.topvar_offset = h->topvar_offset; h->p.leaf = h->leaf;
  h->p.dof = h->dof; h->p.obs_len = h->obs_len; h->p.userdata = h->userdata;
  h->p.neighbor_ptr = h->nb_ptr; h->p.neighbor_idx = h->nb_idx;
  *out = h;
  return MCI_OK;
}

void orc_destroy(orc_handle *h) {
  if (!h) return;
  free(h->grids); free(h->disc_dist); free(h->disc_accum); free(h->reweight); free(h->packed);
  free(h);
}

int orc_sizes(const orc_handle *h, int32_t *grid_len, int32_t *disc_len, int32_t *hist_len, int32_t *obs_len,
              int32_t *nd, int32_t *pa_len) {
  if (grid_len) *grid_len = h->grid_len;
  if (disc_len) *disc_len = h->disc_len;
  if (hist_len) *hist_len = h->hist_len;
  if (obs_len) *obs_len = h->obs_total;
  if (nd) *nd = h->nd;
  if (pa_len) *pa_len = h->pa_len;
  return MCI_OK;
}

static int n_disc_accum(const orc_handle *h) {
  int n = 0;
  for (int l = 0; l < h->nleaf; ++l) if (h->leaf[l].kind == MCI_LEAF_DISCRETE) n += h->leaf[l].nbins + 1;
  return n;
}

int orc_set_maps(orc_handle *h, const double *grids, const double *dist, const double *accum, const double *rw) {
  if (grids) memcpy(h->grids, grids, sizeof(double) * h->grid_len);
  if (dist) memcpy(h->disc_dist, dist, sizeof(double) * h->disc_len);
  if (accum) memcpy(h->disc_accum, accum, sizeof(double) * n_disc_accum(h));
  if (rw) memcpy(h->reweight, rw, sizeof(double) * h->nd);
  return MCI_OK;
}
int orc_get_maps(orc_handle *h, double *grids, double *dist, double *accum, double *rw) {
  if (grids) memcpy(grids, h->grids, sizeof(double) * h->grid_len);
  if (dist) memcpy(dist, h->disc_dist, sizeof(double) * h->disc_len);
  if (accum) memcpy(accum, h->disc_accum, sizeof(double) * n_disc_accum(h));
  if (rw) memcpy(rw, h->reweight, sizeof(double) * h->nd);
  return MCI_OK;
}
int orc_set_partition(orc_handle *h, int32_t rank, int32_t world) {
  if (world < 1 || rank < 0 || rank >= world || h->p.n_block % world != 0) { /* main.jl:122 */
    snprintf(h->err, sizeof h->err, "invalid partition rank=%d world=%d n_block=%d", rank, world, h->p.n_block);
    return MCI_ERR_INVALID;
  }
  h->rank = rank; h->world = world;
  return MCI_OK;
}
int orc_set_shift_mode(orc_handle *h, int32_t mode) { h->shift_mode = mode; return MCI_OK; }
double *orc_packed_ptr(orc_handle *h) { return h->packed; }
int64_t orc_packed_len(const orc_handle *h) { return h->packed_len; }

/* build a private configuration (the deepcopy of main.jl:130-133 + clearStatistics! of main.jl:251, floors per D6) */
static oconfig *config_new(const orc_handle *h) {
  oconfig *c = (oconfig *)calloc(1, sizeof(oconfig));
  c->h = h; c->ni = h->ni; c->nd = h->nd; c->nv = h->nv; c->nleaf = h->nleaf;
  c->dof = h->dof; c->maxdof = h->maxdof; c->reweight = h->reweight; c->userdata = h->userdata;
  c->shift_mode = h->shift_mode;
  c->wd = h->wd; c->counting = 1;
  size_t nd = 0;
  int len[MCI_MAX_TOPVAR];
  for (int v = 0; v < h->nv; ++v) { len[v] = h->maxdof[v] + h->topvar_offset[v] + 2; nd += (size_t)len[v] * (1 + 3 * h->topvar_nleaf[v]); }
  nd += h->hist_len + h->obs_total + h->nd + 2 * h->pa_len;
  double *mem = (double *)calloc(nd + 8, sizeof(double));
  c->mem = mem;
  double *q = mem;
  for (int v = 0; v < h->nv; ++v) {
    otop *T = &c->top[v];
    T->n_leaf = h->topvar_nleaf[v]; T->first_leaf = h->topvar_first[v]; T->offset = h->topvar_offset[v];
    T->fermik = h->leaf[h->topvar_first[v]].kind == MCI_LEAF_FERMIK;
    T->len = len[v]; T->composite = T->n_leaf > 1 && !T->fermik; T->prob_cache = 1.0;
    for (int k = 0; k < T->n_leaf; ++k) {
      int l = T->first_leaf + k;
      oleaf *L = &c->leaf[l];
      const mci_leaf_t *P = &h->leaf[l];
      L->kind = P->kind; L->nbins = P->nbins; L->adapt = P->adapt; L->lower_i = P->lower_i;
      L->lower = P->lower; L->upper = P->upper; L->alpha = P->alpha;
      if (P->kind == MCI_LEAF_CONTINUOUS) L->grid = h->grids + h->leaf_grid_off[l];
      else if (P->kind == MCI_LEAF_DISCRETE) { L->dist = h->disc_dist + h->leaf_disc_off[l]; L->accum = h->disc_accum + h->leaf_accum_off[l]; }
      L->len = len[v];
      L->off = T->offset;
      L->data = q; q += len[v];
      L->prob = q; q += len[v];
      L->gidx = (int *)q; q += len[v]; /* one double slot per int: plenty */
      for (int i = 0; i < len[v]; ++i) { L->prob[i] = 1.0; L->gidx[i] = 1; L->data[i] = P->kind == MCI_LEAF_CONTINUOUS ? P->lower : (double)P->lower_i; }
      if (P->kind == MCI_LEAF_FERMIK) /* zeros(dim, size) .+ kF / sqrt(dim), variable.jl:13 */
        for (int i = 0; i < len[v]; ++i) L->data[i] = P->lower / sqrt((double)T->n_leaf);
    }
    if (T->composite) { T->prob = q; q += len[v]; for (int i = 0; i < len[v]; ++i) T->prob[i] = 1.0; }
    else T->prob = c->leaf[T->first_leaf].prob;
  }
  c->hist = q; q += h->hist_len;
  for (int l = 0; l < h->nleaf; ++l) c->leaf[l].hist = c->hist + h->leaf_hist_off[l];
  c->obs = q; q += h->obs_total;
  c->visited = q; q += h->nd;
  c->propose = q; q += h->pa_len;
  c->accept = q; q += h->pa_len;
  c->normalization = 0.0;
  c->neval = 0;
  return c;
}
static void config_free(oconfig *c) { free(c->mem); free(c); }

/* measure: default obs[i] += relw (vegas/montecarlo.jl:151-153) or the bubble rule (example/bubble.jl:81-84) */
static void do_measure_all(oconfig *c, const double *relw) {
  const orc_handle *h = c->h;
  const int wd = c->wd;
  if (h->p.measure_kind == MCI_MEASURE_DEFAULT) {
    for (int i = 0; i < c->ni; ++i)
      for (int k = 0; k < wd; ++k) c->obs[h->obs_off[i] + k] += relw[wd * i + k];
  } else {
    const oleaf *E = &c->leaf[c->nleaf - 1];
    int off = c->top[c->nv - 1].offset;
    int e = (int)E->data[1 + off - 1] - E->lower_i;
    for (int k = 0; k < wd; ++k) c->obs[h->obs_off[0] + wd * e + k] += relw[k];
  }
}
static void do_measure_one(oconfig *c, int curr, const double *relw) {
  const orc_handle *h = c->h;
  const int wd = c->wd;
  if (h->p.measure_kind == MCI_MEASURE_DEFAULT) { for (int k = 0; k < wd; ++k) c->obs[h->obs_off[curr] + k] += relw[k]; }
  else {
    const oleaf *E = &c->leaf[c->nleaf - 1];
    int off = c->top[c->nv - 1].offset;
    int e = (int)E->data[1 + off - 1] - E->lower_i;
    for (int k = 0; k < wd; ++k) c->obs[h->obs_off[0] + wd * e + k] += relw[k];
  }
}

#define PA(c, u, i, j) ((u) * (c)->nd * (c)->h->pa_dim + (i) * (c)->h->pa_dim + (j))

/* ------------------------------------------------------------------------- */
/* Vegas.montecarlo: vegas/montecarlo.jl:72-191                               */
/* ------------------------------------------------------------------------- */
typedef struct { double *x; int32_t *bins; double *jac; double *w; int n_draw; } otrace;

static void vegas_sample(oconfig *c, uint64_t ne /*1-based*/, double *weights, double *pad, otrace *tr, int64_t trow) {
  const orc_handle *h = c->h;
  int id = h->p.integrand_id;
  double jac = 1.0;
  int kdraw = 0;
  for (int vi = 0; vi < c->nv; ++vi) {            /* :122-131 */
    otop *T = &c->top[vi];
    for (int idx = 1; idx <= c->maxdof[vi]; ++idx) {
      if (c->shift_mode == 1) var_shift(c, vi, idx + T->offset);
      else var_create(c, vi, idx + T->offset);    /* D5 */
      jac /= T->prob[idx + T->offset - 1];
      if (tr)
        for (int k = 0; k < T->n_leaf; ++k, ++kdraw) {
          const oleaf *L = &c->leaf[T->first_leaf + k];
          tr->x[trow * tr->n_draw + kdraw] = L->data[idx + T->offset - 1];
          tr->bins[trow * tr->n_draw + kdraw] = L->kind == MCI_LEAF_CONTINUOUS ? L->gidx[idx + T->offset - 1] - 1 : (int)L->data[idx + T->offset - 1] - L->lower_i;
        }
    }
  }
  for (int i = 0; i < c->ni; ++i) {               /* :133-137 */
    int same = 1;
    for (int vi = 0; vi < c->nv; ++vi) if (c->dof[i * c->nv + vi] != c->maxdof[vi]) same = 0;
    pad[i] = same ? 1.0 : padding_probability(c, i);
  }
  eval_all(c, id, weights);                       /* :140-144 */
  c->neval += 1;
  const int wd = c->wd;
  if (tr) { tr->jac[trow] = jac; for (int i = 0; i < c->ni * wd; ++i) tr->w[trow * c->ni * wd + i] = weights[i]; }
  if (ne % (uint64_t)h->p.measurefreq == 0) {     /* :148-165 */
    double relw[2 * MCI_MAX_INTEGRAND];
    for (int i = 0; i < c->ni * wd; ++i) relw[i] = weights[i] * pad[i / wd] * jac;
    do_measure_all(c, relw);
    c->normalization += 1.0;
  }
  for (int vi = 0; vi < c->nv; ++vi) {            /* :170-185 */
    int offset = c->top[vi].offset;
    for (int i = 0; i < c->ni; ++i) {
      double w2 = wabs(c, weights, i);
      double j2 = jac;
      for (int pos = 1; pos <= c->dof[i * c->nv + vi]; ++pos) var_accumulate(c, vi, pos + offset, (w2 * j2) * (w2 * j2));
    }
  }
}

static void vegas_block(oconfig *c, uint64_t iter, int block) {
  const orc_handle *h = c->h;
  double weights[2 * MCI_MAX_INTEGRAND], pad[MCI_MAX_INTEGRAND];
  if (c->shift_mode == 1) { /* Dist.initialize! :108-110, on a stream no sample uses */
    orng_init(&c->rng, h->p.seed, iter, (uint32_t)block, ~(uint64_t)0);
    for (int vi = 0; vi < c->nv; ++vi) var_initialize(c, vi);
  }
  for (uint64_t ne = 1; ne <= h->p.neval_per_block; ++ne) { /* :117 */
    orng_init(&c->rng, h->p.seed, iter, (uint32_t)block, ne - 1);
    vegas_sample(c, ne, weights, pad, NULL, 0);
  }
}

/* ------------------------------------------------------------------------- */
/* VegasMC.montecarlo: vegas_mc/montecarlo.jl:112-241, updates.jl:45-106       */
/* ------------------------------------------------------------------------- */
static int leaf_is_single_discrete(const oconfig *c, int vi) {
  const otop *T = &c->top[vi];
  return !T->composite && c->leaf[T->first_leaf].kind == MCI_LEAF_DISCRETE && c->leaf[T->first_leaf].nbins == 1;
}

static double vegasmc_change_variable(oconfig *c, double currProbability, double *weights, double *_weights,
                                      double *pad, double *_pad) {
  const orc_handle *h = c->h;
  int N = c->ni;
  int vi = orng_int(&c->rng, c->nv) - 1;                   /* updates.jl:50 */
  if (leaf_is_single_discrete(c, vi)) return currProbability; /* :52-54 */
  if (c->maxdof[vi] <= 0) return currProbability;          /* :55-57 */
  int idx = c->top[vi].offset + orng_int(&c->rng, c->maxdof[vi]); /* :58 */
  double prop = var_shift(c, vi, idx);                     /* :60 */
  if (prop <= EPS0) return currProbability;                /* :62-64 */
  eval_all(c, h->p.integrand_id, _weights);                /* :66-74 */
  if (c->counting) c->neval += 1;
  for (int i = 0; i < N + 1; ++i) _pad[i] = padding_probability(c, i); /* :79-81 */
  double newProbability = c->reweight[N] * _pad[N];        /* :82 */
  for (int i = 0; i < N; ++i) newProbability += wabs(c, _weights, i) * c->reweight[i] * _pad[i];
  double R = prop * newProbability / currProbability;      /* :88 */
  if (c->counting) c->propose[PA(c, 1, 0, vi)] += 1.0;     /* :90 propose[2,1,vi] */
  if (orng_rand(&c->rng) < R) {
    if (c->counting) c->accept[PA(c, 1, 0, vi)] += 1.0;
    for (int i = 0; i < N * c->wd; ++i) weights[i] = _weights[i];
    for (int i = 0; i < N + 1; ++i) pad[i] = _pad[i];
    return newProbability;
  } else {
    var_shift_rollback(c, vi, idx);
    return currProbability;
  }
}

static void vegasmc_walker(oconfig *c, uint64_t iter, int block, uint64_t walker, uint64_t neval) {
  const orc_handle *h = c->h;
  int N = c->ni;
  const int wd = c->wd;
  double weights[2 * MCI_MAX_INTEGRAND], _weights[2 * MCI_MAX_INTEGRAND], relw[2 * MCI_MAX_INTEGRAND];
  double pad[MCI_MAX_INTEGRAND + 1] = {0.0}, _pad[MCI_MAX_INTEGRAND + 1] = {0.0};
  orng_init(&c->rng, h->p.seed, iter, (uint32_t)block, walker);
  for (int vi = 0; vi < c->nv; ++vi) var_initialize(c, vi);        /* montecarlo.jl:151-153 */
  eval_all(c, h->p.integrand_id, _weights);                         /* :155-159 */
  for (int i = 0; i < N + 1; ++i) pad[i] = padding_probability(c, i); /* :161 */
  double probability_ = c->reweight[N] * pad[N];                    /* :162 */
  for (int i = 0; i < N * wd; ++i) weights[i] = _weights[i];
  for (int i = 0; i < N; ++i) probability_ += wabs(c, _weights, i) * c->reweight[i] * pad[i];
  for (int64_t ne = 1 - (int64_t)h->p.chain_burnin; ne <= (int64_t)neval; ++ne) { /* :184; ne <= 0: extra burn-in (mcib200.h) */
    c->counting = ne >= 1;
    probability_ = vegasmc_change_variable(c, probability_, weights, _weights, pad, _pad); /* :186-188; single update */
    if (ne < 1) continue;
    for (int i = 0; i < N; ++i) {                                   /* :198-211 */
      double f2 = wabs(c, weights, i) * wabs(c, weights, i) / probability(c, i);
      double wf2 = f2 * pad[i] / probability_;
      for (int vi = 0; vi < c->nv; ++vi) {
        int offset = c->top[vi].offset;
        for (int pos = 1; pos <= c->dof[i * c->nv + vi]; ++pos) var_accumulate(c, vi, pos + offset, wf2);
      }
    }
    if ((uint64_t)ne % (uint64_t)h->p.measurefreq == 0 && (double)ne >= (double)neval / 100) { /* :213 */
      for (int i = 0; i < N; ++i) {
        double wpr[2];
        for (int k = 0; k < wd; ++k) wpr[k] = weights[wd * i + k] * pad[i] * c->reweight[i];
        c->visited[i] += wabs(c, wpr, 0) / probability_;
        for (int k = 0; k < wd; ++k) relw[wd * i + k] = weights[wd * i + k] * pad[i] / probability_;
      }
      do_measure_all(c, relw);
      c->normalization += 1.0 * pad[N] / probability_;                /* :229 */
      c->visited[N] += c->reweight[N] * pad[N] / probability_;        /* :230 */
    }
  }
}

/* ------------------------------------------------------------------------- */
/* MCMC.montecarlo: mcmc/montecarlo.jl:72-205, updates.jl:1-147               */
/* ------------------------------------------------------------------------- */
typedef struct { int curr; double weight[2]; double probability; } ostate; /* weight: (re, im) when complex */

static void mcmc_initialize(oconfig *c, ostate *s) { /* montecarlo.jl:190-205 */
  for (int vi = 0; vi < c->nv; ++vi) var_initialize(c, vi);
  if (s->curr != c->ni) {
    eval_one(c, c->h->p.integrand_id, s->curr, s->weight);
    s->probability = wabs(c, s->weight, 0) * c->reweight[s->curr];
  } else { s->weight[0] = s->weight[1] = 0.0; s->probability = c->reweight[s->curr]; }
}

static void mcmc_change_integrand(oconfig *c, ostate *s) { /* updates.jl:1-70 */
  const orc_handle *h = c->h;
  int curr = s->curr;
  int nnb = h->nb_ptr[curr + 1] - h->nb_ptr[curr];
  int new_ = h->nb_idx[h->nb_ptr[curr] + orng_int(&c->rng, nnb) - 1]; /* :6 */
  if (new_ == curr) return;
  const int *currdof = c->dof + curr * c->nv, *newdof = c->dof + new_ * c->nv;
  double prop = (double)nnb / (double)(h->nb_ptr[new_ + 1] - h->nb_ptr[new_]); /* :12 */
  for (int vi = 0; vi < c->nv; ++vi) {                       /* :15-26 */
    int offset = c->top[vi].offset;
    if (currdof[vi] < newdof[vi]) { for (int pos = currdof[vi] + 1; pos <= newdof[vi]; ++pos) prop *= var_create(c, vi, pos + offset); }
    else if (currdof[vi] > newdof[vi]) { for (int pos = newdof[vi] + 1; pos <= currdof[vi]; ++pos) prop *= var_remove(c, vi, pos + offset); }
  }
  if (prop <= EPS0) return;                                  /* :29-31 */
  double newWeight[2] = {0.0, 0.0};
  if (new_ != c->ni) eval_one(c, h->p.integrand_id, new_, newWeight); /* :35-38 */
  if (c->counting) c->neval += 1;
  double newProbability = (new_ == c->ni) ? c->reweight[new_] : wabs(c, newWeight, 0) * c->reweight[new_];
  double R = prop * newProbability / s->probability;         /* :46 */
  if (c->counting) c->propose[PA(c, 0, curr, new_)] += 1.0;
  if (orng_rand(&c->rng) < R) {
    if (c->counting) c->accept[PA(c, 0, curr, new_)] += 1.0;
    s->curr = new_; s->weight[0] = newWeight[0]; s->weight[1] = newWeight[1]; s->probability = newProbability;
  } /* createRollback!/removeRollback! are no-ops: sampler.jl:24,43,306,324 */
}

static void mcmc_change_variable(oconfig *c, ostate *s) { /* updates.jl:72-109 */
  if (s->curr == c->ni) return;
  int curr = s->curr;
  const int *currdof = c->dof + curr * c->nv;
  int vi = orng_int(&c->rng, c->nv) - 1;                     /* :78 */
  if (leaf_is_single_discrete(c, vi)) return;
  if (currdof[vi] <= 0) return;
  int idx = c->top[vi].offset + orng_int(&c->rng, currdof[vi]); /* :84 */
  double prop = var_shift(c, vi, idx);
  if (prop <= EPS0) return;
  double weight[2] = {0.0, 0.0};
  eval_one(c, c->h->p.integrand_id, curr, weight);
  if (c->counting) c->neval += 1;
  double newProbability = wabs(c, weight, 0) * c->reweight[curr];
  double R = prop * newProbability / s->probability;
  if (c->counting) c->propose[PA(c, 1, curr, vi)] += 1.0;
  if (orng_rand(&c->rng) < R) {
    if (c->counting) c->accept[PA(c, 1, curr, vi)] += 1.0;
    s->weight[0] = weight[0]; s->weight[1] = weight[1]; s->probability = newProbability;
  } else var_shift_rollback(c, vi, idx);
}

static void mcmc_swap_variable(oconfig *c, ostate *s) { /* updates.jl:111-147 */
  if (s->curr == c->ni) return;
  int curr = s->curr;
  const int *currdof = c->dof + curr * c->nv;
  int vi = orng_int(&c->rng, c->nv) - 1;                     /* :117 */
  if (currdof[vi] <= 0) return;
  int idx1 = c->top[vi].offset + orng_int(&c->rng, currdof[vi]);
  int idx2 = c->top[vi].offset + orng_int(&c->rng, currdof[vi]);
  if (idx1 == idx2) return;
  double prop = var_swap(c, vi, idx1, idx2);
  if (prop <= EPS0) return;
  double weight[2] = {0.0, 0.0};
  eval_one(c, c->h->p.integrand_id, curr, weight);
  if (c->counting) c->neval += 1;
  double newProbability = wabs(c, weight, 0) * c->reweight[curr];
  double R = prop * newProbability / s->probability;
  if (c->counting) c->propose[PA(c, 2, curr, vi)] += 1.0;
  if (orng_rand(&c->rng) < R) {
    if (c->counting) c->accept[PA(c, 2, curr, vi)] += 1.0;
    s->weight[0] = weight[0]; s->weight[1] = weight[1]; s->probability = newProbability;
  } else var_swap_rollback(c, vi, idx1, idx2);
}

/* returns 0 ok, MCI_ERR_MCMC_START when no start point exists (montecarlo.jl:109-110) */
static int mcmc_walker(oconfig *c, uint64_t iter, int block, uint64_t walker, uint64_t neval) {
  const orc_handle *h = c->h;
  orng_init(&c->rng, h->p.seed, iter, (uint32_t)block, walker);
  ostate s = {h->start_idx, {0.0, 0.0}, 1.0}; /* `idx`: the integral to start with (:76), default 1 */
  for (int i = 0; i < 10000; ++i) {                          /* :103-108 */
    mcmc_initialize(c, &s);
    if (s.curr == c->ni || s.probability > TINY) break;
  }
  if (s.curr != c->ni && s.probability == 0.0) return MCI_ERR_MCMC_START;
  int nupdates = 2 * c->nv + 2;                              /* :117-120 */
  uint64_t nburnin = (uint64_t)floor((double)neval * h->p.thermal_ratio); /* :133 */
  for (int64_t i = 1 - (int64_t)h->p.chain_burnin; i <= (int64_t)(neval + nburnin); ++i) { /* :134; i <= 0: extra burn-in */
    c->counting = i >= 1;
    if (i >= 1) c->visited[s.curr] += 1;
    int u = orng_int(&c->rng, nupdates);                     /* :138 */
    if (u == 1) mcmc_change_integrand(c, &s);
    else if (u == 2) mcmc_swap_variable(c, &s);
    else mcmc_change_variable(c, &s);
    if (i >= 1 && (uint64_t)i % (uint64_t)h->p.measurefreq == 0 && (uint64_t)i >= nburnin) { /* :145 */
      if (s.curr != c->ni)
        for (int vi = 0; vi < c->nv; ++vi) {
          int offset = c->top[vi].offset;
          for (int pos = 1; pos <= c->dof[s.curr * c->nv + vi]; ++pos) var_accumulate(c, vi, pos + offset, 1.0);
        }
      if (s.curr == c->ni) c->normalization += 1.0 / c->reweight[c->ni]; /* :158-159 */
      else { /* :161-169 */
        double relw[2] = {s.weight[0] / s.probability, s.weight[1] / s.probability};
        do_measure_one(c, s.curr, relw);
      }
    }
  }
  return 0;
}

/* ------------------------------------------------------------------------- */
/* iteration driver: the block loop of main.jl:144-180                         */
/* ------------------------------------------------------------------------- */
int orc_run_iteration(orc_handle *h, uint64_t iter, int32_t nthreads) {
  const mci_problem_t *p = &h->p;
  int nb = p->n_block;
  int nlocal = nb / h->world;
  memset(h->packed, 0, sizeof(double) * h->packed_len);
  oconfig **cfg = (oconfig **)calloc(nlocal, sizeof(oconfig *));
  int status = 0;
  if (nthreads < 1) nthreads = 1;
#ifdef _OPENMP
#pragma omp parallel for num_threads(nthreads) schedule(dynamic, 1)
#endif
  for (int k = 0; k < nlocal; ++k) {
    int b = h->rank + k * h->world; /* blocks b mod world == rank */
    oconfig *c = config_new(h);
    cfg[k] = c;
    int rc = 0;
    if (p->solver == MCI_SOLVER_VEGAS) vegas_block(c, iter, b);
    else {
      uint64_t W = p->walkers_per_block < 1 ? 1 : (uint64_t)p->walkers_per_block;
      uint64_t L = p->neval_per_block / W; /* D2 */
      for (uint64_t w = 0; w < W && rc == 0; ++w) {
        if (p->solver == MCI_SOLVER_VEGASMC) vegasmc_walker(c, iter, b, w, L);
        else rc = mcmc_walker(c, iter, b, w, L);
      }
    }
    if (rc != 0) {
#ifdef _OPENMP
#pragma omp critical
#endif
      status = rc;
    }
  }
  /* addConfig! in block order (configuration.jl:252-262): deterministic */
  for (int k = 0; k < nlocal; ++k) {
    int b = h->rank + k * h->world;
    oconfig *c = cfg[k];
    for (int i = 0; i < h->hist_len; ++i) h->packed[h->off_hist + i] += c->hist[i];
    for (int i = 0; i < h->obs_total; ++i) h->packed[h->off_obs + b * h->obs_total + i] = c->obs[i];
    h->packed[h->off_cnt + b] = c->normalization;
    for (int i = 0; i < h->nd; ++i) h->packed[h->off_visited + i] += c->visited[i];
    for (int i = 0; i < h->pa_len; ++i) { h->packed[h->off_propose + i] += c->propose[i]; h->packed[h->off_accept + i] += c->accept[i]; }
    h->packed[h->off_neval] += (double)c->neval;
    config_free(c);
  }
  free(cfg);
  if (status) { snprintf(h->err, sizeof h->err, "Cannot find the variables that makes the #%d integrand nonzero!", h->start_idx + 1); return status; }
  return MCI_OK;
}

int orc_fetch_iteration(orc_handle *h, mci_iteration_out_t *out) {
  int nb = h->p.n_block;
  double B1 = (double)(nb + 1);
  if (out->block_obs) memcpy(out->block_obs, h->packed + h->off_obs, sizeof(double) * nb * h->obs_total);
  if (out->block_norm) for (int b = 0; b < nb; ++b) out->block_norm[b] = 1.0e-10 + h->packed[h->off_cnt + b]; /* configuration.jl:243 */
  if (out->hist) for (int i = 0; i < h->hist_len; ++i) out->hist[i] = B1 * 1.0e-10 + h->packed[h->off_hist + i]; /* variable.jl:565 */
  if (out->visited) for (int i = 0; i < h->nd; ++i) out->visited[i] = B1 * 1.0e-8 + h->packed[h->off_visited + i];
  if (out->propose) for (int i = 0; i < h->pa_len; ++i) out->propose[i] = B1 * 1.0e-8 + h->packed[h->off_propose + i];
  if (out->accept) for (int i = 0; i < h->pa_len; ++i) out->accept[i] = B1 * 1.0e-10 + h->packed[h->off_accept + i];
  if (out->neval) *out->neval = (uint64_t)h->packed[h->off_neval];
  for (int b = 0; b < nb; ++b)
    if (!(1.0e-10 + h->packed[h->off_cnt + b] > 0.0)) { /* main.jl:269-271 */
      snprintf(h->err, sizeof h->err, "Block normalization = %g is not positively defined!", h->packed[h->off_cnt + b]);
      return MCI_ERR_NUMERIC;
    }
  return MCI_OK;
}

/* the `if adapt: train!(v)` step of main.jl:193-198 on the summed histogram */
int orc_refine(orc_handle *h) {
  double B1 = (double)(h->p.n_block + 1);
  for (int l = 0; l < h->nleaf; ++l) {
    mci_leaf_t *L = &h->leaf[l];
    if (!L->adapt) continue;
    int n = L->nbins;
    double *hist = (double *)malloc(sizeof(double) * n);
    for (int i = 0; i < n; ++i) hist[i] = B1 * 1.0e-10 + h->packed[h->off_hist + h->leaf_hist_off[l] + i];
    int rc;
    if (L->kind == MCI_LEAF_CONTINUOUS) {
      double *g = h->grids + h->leaf_grid_off[l];
      double *ng = (double *)malloc(sizeof(double) * (n + 1));
      rc = orc_train_continuous(g, n + 1, hist, L->alpha, ng);
      if (rc == 0) memcpy(g, ng, sizeof(double) * (n + 1));
      free(ng);
    } else {
      rc = orc_train_discrete(hist, n, L->alpha, h->disc_dist + h->leaf_disc_off[l], h->disc_accum + h->leaf_accum_off[l]);
    }
    free(hist);
    if (rc != 0) { snprintf(h->err, sizeof h->err, "train!: histogram of leaf %d not finite/positive", l); return MCI_ERR_NUMERIC; }
  }
  return MCI_OK;
}

int orc_reweight(orc_handle *h, double gamma, const double *goal) {
  double B1 = (double)(h->p.n_block + 1);
  double visited[MCI_MAX_INTEGRAND + 1];
  for (int i = 0; i < h->nd; ++i) visited[i] = B1 * 1.0e-8 + h->packed[h->off_visited + i];
  orc_do_reweight(h->reweight, visited, h->nd, gamma, goal);
  return MCI_OK;
}

int orc_trace_samples(orc_handle *h, uint64_t iter, int32_t block, uint64_t first, int32_t count, double *x,
                      int32_t *bins, double *jac, double *weights) {
  if (h->p.solver != MCI_SOLVER_VEGAS) return MCI_ERR_UNSUPPORTED;
  oconfig *c = config_new(h);
  c->shift_mode = 0;
  int n_draw = 0;
  for (int vi = 0; vi < h->nv; ++vi) n_draw += h->maxdof[vi] * h->topvar_nleaf[vi];
  otrace tr = {x, bins, jac, weights, n_draw};
  double w[2 * MCI_MAX_INTEGRAND], pad[MCI_MAX_INTEGRAND];
  for (int32_t s = 0; s < count; ++s) {
    orng_init(&c->rng, h->p.seed, iter, (uint32_t)block, first + s);
    vegas_sample(c, first + s + 1, w, pad, &tr, s);
  }
  config_free(c);
  return MCI_OK;
}

/* test/utility.jl:30-56 */
void orc_probability_identity(const double *gridx, const double *gridy, const double *dist6, uint64_t seed,
                              int32_t do_shift, double *total, double *prob2, double *pad2) {
  mci_problem_t p;
  memset(&p, 0, sizeof p);
  int nleaf[3] = {1, 1, 1}, off[3] = {0, 0, 0}, dof[6] = {1, 1, 1, 2, 3, 3}, ol[2] = {1, 1};
  mci_leaf_t leaf[3] = {{MCI_LEAF_CONTINUOUS, 3, 1, 0, 0.0, 1.0, 2.0}, {MCI_LEAF_CONTINUOUS, 3, 1, 0, 0.0, 1.0, 2.0},
                        {MCI_LEAF_DISCRETE, 6, 1, 1, 0.0, 0.0, 2.0}};
  p.solver = MCI_SOLVER_VEGAS; p.integrand_id = MCI_F_SPHERE1; p.n_integrand = 2; p.n_topvar = 3; p.n_leaf = 3;
  p.topvar_nleaf = nleaf; p.topvar_offset = off; p.leaf = leaf; p.dof = dof; p.obs_len = ol;
  p.n_block = 1; p.neval_per_block = 1; p.seed = seed; p.measurefreq = 1; p.walkers_per_block = 1;
  orc_handle *h;
  orc_create(&p, &h);
  double grids[8], dist[6], acc[7], s = 0.0;
  memcpy(grids, gridx, 4 * sizeof(double)); memcpy(grids + 4, gridy, 4 * sizeof(double));
  for (int i = 0; i < 6; ++i) s += dist6[i];
  for (int i = 0; i < 6; ++i) dist[i] = dist6[i] / s;
  acc[0] = 0.0;
  for (int i = 1; i <= 6; ++i) { double a = 0.0; for (int k = 0; k < i; ++k) a += dist[k]; acc[i] = a; }
  orc_set_maps(h, grids, dist, acc, NULL);
  oconfig *c = config_new(h);
  orng_init(&c->rng, seed, 0, 0, 0);
  for (int vi = 0; vi < c->nv; ++vi) var_initialize(c, vi);
  if (do_shift)
    for (int vi = 0; vi < c->nv; ++vi)
      for (int i = 1; i <= c->maxdof[vi]; ++i) var_shift(c, vi, i + c->top[vi].offset);
  *total = total_probability(c);
  for (int i = 0; i < 2; ++i) { prob2[i] = probability(c, i); pad2[i] = padding_probability(c, i); }
  config_free(c);
  orc_destroy(h);
}
