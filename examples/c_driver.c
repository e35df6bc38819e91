/*
 * c_driver.c — the reference's `integrate` iteration loop (src/main.jl:142-207) written against nothing but the C ABI
 * of libmcib200.so (include/mcib200.h): what a host in any language binds.  BASELINE C2:
 *
 *     integrate(x -> (2pi)^-5 exp(-|x|^2/2); var = Continuous(-5, 5), dof = [[10]], solver = :vegas, neval, niter)
 *
 * Build and run (B200):
 *     gcc -O2 -std=c11 -Iinclude examples/c_driver.c -o /tmp/c_driver -Lmcintegration.jl_b200 -lmcib200 \
 *         -Wl,-rpath,$PWD/mcintegration.jl_b200 -lm
 *     /tmp/c_driver [neval_per_iteration] [niter]
 * Prints one line per iteration and the weighted average over iterations (statistics.jl:186-220, ignore = 1).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mcib200.h"

#define CHECK(h, call)                                                                  \
  do {                                                                                  \
    int rc_ = (call);                                                                   \
    if (rc_ != MCI_OK) {                                                                \
      fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, mci_last_error(h));          \
      return 1;                                                                         \
    }                                                                                   \
  } while (0)

int main(int argc, char **argv) {
  const uint64_t neval = argc > 1 ? strtoull(argv[1], NULL, 10) : (1ull << 24);
  const int niter = argc > 2 ? atoi(argv[2]) : 10;
  enum { BLOCK = 16, DIM = 10 };

  /* Configuration(var = Continuous(-5, 5), dof = [[10]]): configuration.jl:105-194 */
  mci_leaf_t leaf = {MCI_LEAF_CONTINUOUS, 999 /* ninc = 1000 grid points */, 1 /* adapt */, 0, -5.0, 5.0, 2.0 /* alpha */};
  int32_t topvar_nleaf[1] = {1}, topvar_offset[1] = {0}, dof[1] = {DIM}, obs_len[1] = {1};
  mci_problem_t p;
  memset(&p, 0, sizeof p);
  p.solver = MCI_SOLVER_VEGAS;
  p.integrand_id = MCI_F_GAUSS_ND;
  p.measure_kind = MCI_MEASURE_DEFAULT;
  p.n_integrand = 1;
  p.n_topvar = 1;
  p.n_leaf = 1;
  p.topvar_nleaf = topvar_nleaf;
  p.topvar_offset = topvar_offset;
  p.leaf = &leaf;
  p.dof = dof;
  p.obs_len = obs_len;
  p.n_block = BLOCK;                    /* main.jl:121-122 */
  p.neval_per_block = neval / BLOCK;    /* main.jl:232 */
  p.seed = 1234;
  p.measurefreq = 1;
  p.walkers_per_block = 1;
  p.thermal_ratio = 0.1;

  mci_handle_t *h = NULL;
  CHECK(NULL, mci_create(&p, &h));

  double block_obs[BLOCK], block_norm[BLOCK];
  uint64_t nev = 0;
  mci_iteration_out_t out;
  memset(&out, 0, sizeof out);
  out.block_obs = block_obs;
  out.block_norm = block_norm;
  out.neval = &nev;

  double wsum = 0.0, wmean = 0.0; /* average(): weights 1/(sigma + 1e-10)^2, statistics.jl:217 */
  for (int iter = 0; iter < niter; ++iter) {
    CHECK(h, mci_run_iteration(h, (uint64_t)iter, &out)); /* all _block! calls of this iteration: main.jl:144-180 */
    CHECK(h, mci_refine(h));                              /* train! on every variable: main.jl:193-198 */
    double s = 0.0, s2 = 0.0;                             /* per-block means and _mean_std: main.jl:275-287, 296-320 */
    for (int b = 0; b < BLOCK; ++b) {
      const double m = block_obs[b] / block_norm[b];
      s += m;
      s2 += m * m;
    }
    const double mean = s / BLOCK;
    double var = (s2 / BLOCK - mean * mean) / (BLOCK - 1);
    const double sigma = sqrt(var > 0.0 ? var : 0.0);
    printf("iter %2d  %.10f +- %.3e   (%llu evaluations)\n", iter + 1, mean, sigma, (unsigned long long)nev);
    if (iter >= 1) { /* ignore = 1: the first iteration only trains the grid (main.jl:80) */
      const double w = 1.0 / ((sigma + 1e-10) * (sigma + 1e-10));
      wsum += w;
      wmean += w * mean;
    }
  }
  if (wsum > 0.0) {
    const double exact = pow(erf(5.0 / sqrt(2.0)), DIM);
    printf("result  %.10f +- %.3e   exact %.10f   kernel %s\n", wmean / wsum, 1.0 / sqrt(wsum), exact, mci_kernel_name(h));
  }
  mci_destroy(h);
  return 0;
}
