/*
 * mci_oracle.h — CPU ORACLE for the MCIntegration.jl hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may build, load or call it, and only as the checker
 * or the reported CPU baseline.  The product (libmcib200.so) never links or
 * falls back to this code.
 *
 * What it is: a plain-C restatement of the reference's algorithm (paths are
 * relative to the reference repo numericalEFT/MCIntegration.jl @ v0.4.2; each
 * function in mci_oracle.c cites the lines it follows), with ONE deliberate
 * substitution: uniforms come from Philox4x32-10 keyed by (seed, iteration)
 * with counter (sample-or-walker index, draw pair, block) instead of Julia's
 * stdlib MersenneTwister, which lives outside the reference repo.
 *
 * PARITY PIN STATUS: the reference is Julia and no Julia runtime exists in
 * this image, and the reference's tests hold no fixed-seed golden vectors, so
 * *stream-level parity is unpinned*.  What IS pinned (tests/test_oracle_*.py):
 *   - every deterministic fixture of the reference's own tests: locate
 *     (test/utility.jl:1-10), _maxdof (:12-15), the probability identity
 *     (:30-56), _mean_std (test/statistics.jl), the reweight fixed point
 *     (test/mpi_test.jl:148-168), reduce/add algebra of histograms;
 *   - the reference's statistical known answers with its own 7-sigma rule
 *     (test/montecarlo.jl:262-387, test/bubble.jl) for all three solvers;
 *   - golden vectors for the map, train! and rescale AND for the three whole
 *     solver loops (vegas, vegasmc, mcmc on a two-integrand pool problem:
 *     observables, normalisation, histograms, visited / propose / accept
 *     counters; the same loops on a CompositeVar of two Discrete leaves with
 *     non-uniform distributions; :vegas / :vegasmc over two top-level
 *     variables with padding across them; the :mcmc loop of test/bubble_FermiK.jl on
 *     (T, K::FermiK{3 or 2}, Ext)) produced by an independent pure-Python
 *     restatement of the Julia sources (tests/golden/make_golden.py), which
 *     tests/test_oracle_fuzz.py also runs against this file on seeded random
 *     problems;
 *   - Philox4x32-10 against the Random123 known-answer vectors.
 */
#ifndef MCI_ORACLE_H
#define MCI_ORACLE_H

#include <stdint.h>
#include "../include/mcib200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_handle orc_handle;

/* ---- problem-level API: mirrors the product ABI so tests drive both alike ---- */
int orc_create(const mci_problem_t *problem, orc_handle **out);
void orc_destroy(orc_handle *h);
const char *orc_last_error(const orc_handle *h);
int orc_sizes(const orc_handle *h, int32_t *grid_len, int32_t *disc_len, int32_t *hist_len, int32_t *obs_len,
              int32_t *nd, int32_t *pa_len);
int orc_set_maps(orc_handle *h, const double *grids, const double *disc_dist, const double *disc_accum,
                 const double *reweight);
int orc_get_maps(orc_handle *h, double *grids, double *disc_dist, double *disc_accum, double *reweight);
int orc_set_partition(orc_handle *h, int32_t rank, int32_t world);
/* shift_mode 0: stateless create!-form probabilities (what the device computes);
 * shift_mode 1: the reference's stateful ratio form of shift! inside the VEGAS loop (sampler.jl:383-384) */
int orc_set_shift_mode(orc_handle *h, int32_t mode);
/* run this rank's blocks of iteration `iter` with `nthreads` OpenMP threads (blocks are the parallel unit,
 * as with parallel=:thread, main.jl:153-158); raw sums go to the packed buffer */
int orc_run_iteration(orc_handle *h, uint64_t iter, int32_t nthreads);
/* packed statistics buffer (raw sums, no floors): [hist | block_obs | block_count | visited | propose | accept | neval] */
double *orc_packed_ptr(orc_handle *h);
int64_t orc_packed_len(const orc_handle *h);
/* statistics of the last iteration with the reference's floors applied */
int orc_fetch_iteration(orc_handle *h, mci_iteration_out_t *out);
int orc_refine(orc_handle *h);
int orc_reweight(orc_handle *h, double gamma, const double *goal);
int orc_trace_samples(orc_handle *h, uint64_t iter, int32_t block, uint64_t first, int32_t count, double *x,
                      int32_t *bins, double *jac, double *weights);

/* ---- function-level API, for the reference's unit fixtures and golden vectors ---- */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* n uniforms of stream (seed, iter, block, idx) starting at draw 0 */
void orc_uniforms(uint64_t seed, uint64_t iter, uint32_t block, uint64_t idx, int32_t n, double *out);
int32_t orc_locate(const double *accumulation, int32_t n, double p); /* 1-based like the reference; -1 on error */
void orc_smooth(const double *dist, int32_t n, double factor, double *out);
int orc_rescale(double *dist, int32_t n, double alpha);
int orc_train_continuous(const double *grid, int32_t npts, const double *hist, double alpha, double *newgrid);
int orc_train_discrete(const double *hist, int32_t n, double alpha, double *dist, double *accum);
/* Continuous create! map: returns 1-based iy, x and prob */
void orc_map_continuous(const double *grid, int32_t npts, double y, int32_t *iy, double *x, double *prob);
void orc_maxdof(const int32_t *dof, int32_t nd, int32_t nv, int32_t *maxdof);
void orc_mean_std(const double *sum, const double *sqsum, int32_t n, int32_t block, double *mean, double *std);
void orc_average(const double *mean, const double *std, int32_t niter, int32_t init /*1-based*/, double *avg,
                 double *err, double *chi2);
void orc_do_reweight(double *reweight, const double *visited, int32_t nd, double gamma, const double *goal);
/* probability identity probe (test/utility.jl:30-56): builds X,Y continuous with the given 4-point grids and a
 * 6-bin Discrete, dof=[[1,1,1],[2,3,3]], initialises, optionally shifts every slot, and returns
 * total_probability, probability(i), padding_probability(i) for i=1,2 */
void orc_probability_identity(const double *gridx, const double *gridy, const double *dist6, uint64_t seed,
                              int32_t do_shift, double *total, double *prob2, double *pad2);

#ifdef __cplusplus
}
#endif
#endif
