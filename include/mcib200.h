/*
 * mcib200.h — C ABI of libmcib200.so: the B200 (sm_100a) sampling-and-accumulation
 * hot path of MCIntegration.jl (VEGAS / VEGAS-MC / MCMC `montecarlo` drivers plus
 * the adaptive maps of src/distribution).
 *
 * This is the drop-in boundary.  Every entry point states which reference
 * interface it replaces (paths relative to the reference repo, v0.4.2).
 * A Julia host binds these with `ccall((:mci_xxx, "libmcib200"), ...)`; see
 * INTEGRATION.md and julia/MCIntegrationB200.jl.
 *
 * Conventions
 *   - plain C, `extern "C"`, no C++/torch types in any signature;
 *   - every pointer argument is HOST memory owned by the caller; the callee
 *     copies what it needs before returning and never frees caller memory;
 *   - return value 0 = MCI_OK, negative = error; text via mci_last_error();
 *   - no exception ever crosses the ABI, no callback into the host: the
 *     integrand and the `measure` rule run on the device and are selected by
 *     id (mci_integrand_id / mci_measure_kind);
 *   - a handle is bound to one GPU and is not thread-safe; distinct handles
 *     are independent (one per GPU / rank);
 *   - there is NO CPU fallback: without a CUDA device mci_create() fails with
 *     MCI_ERR_NO_DEVICE.
 *   - all arrays are little-endian f64 / i32 / u64, C row-major.
 */
#ifndef MCIB200_H
#define MCIB200_H

#ifdef __CUDACC_RTC__ /* compiled by NVRTC for a JIT-specialised kernel: no host headers there */
typedef int int32_t;
typedef unsigned int uint32_t;
typedef long long int64_t;
typedef unsigned long long uint64_t;
#else
#include <stdint.h>
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define MCI_ABI_VERSION 3

/* ---- limits of the POD problem description ------------------------------ */
#define MCI_MAX_TOPVAR 16    /* top-level variables (`config.var` tuple, configuration.jl:40) */
#define MCI_MAX_LEAF 32      /* leaf maps over all top-level variables                  */
#define MCI_MAX_INTEGRAND 16 /* Ni, without the normalisation integrand: the reference's own bound
                                MaxOrder = 16 (src/distribution/distribution.jl:59)        */
#define MCI_MAX_SLOT 24      /* max dof of one pool variable on the device              */
#define MCI_MAX_USERDATA 64  /* doubles                                                 */
#define MCI_MAX_OBS 64       /* total observable length (sum of obs_len)                */

/* ---- error codes ------------------------------------------------------- */
enum {
  MCI_OK = 0,
  MCI_ERR_INVALID = -1,     /* bad argument / problem description (Julia @assert sites)  */
  MCI_ERR_NO_DEVICE = -2,   /* no CUDA device or kernel image not loadable: no fallback  */
  MCI_ERR_CUDA = -3,        /* CUDA runtime error                                        */
  MCI_ERR_NUMERIC = -4,     /* non-finite weight / non-positive block normalisation
                               (main.jl:269-271, variable.jl:212-213, common.jl:71,79)   */
  MCI_ERR_UNSUPPORTED = -5, /* layout/integrand combination not built into this library  */
  MCI_ERR_COMM = -6,        /* NCCL failure                                              */
  MCI_ERR_MCMC_START = -7   /* no non-zero start weight found (mcmc/montecarlo.jl:109)   */
};

/* solver: `solver::Symbol` kwarg of integrate (main.jl:72, dispatch main.jl:253-264) */
enum { MCI_SOLVER_VEGAS = 0, MCI_SOLVER_VEGASMC = 1, MCI_SOLVER_MCMC = 2 };

/* leaf kinds: Continuous (variable.jl:87-153) / Discrete (variable.jl:272-325) / one Cartesian component of a
 * FermiK{D} momentum (variable.jl:1-21, sampler.jl:109-281).  A FermiK{D} variable is a top-level variable of D = 2
 * or 3 FERMIK leaves (k_x, k_y[, k_z]: X(leaf0 + c, pos) is component c of K[pos]); it has no map and no histogram
 * (nbins = 0, adapt = 0) and, like in the reference (test/bubble_FermiK.jl:2,139-140), works under :mcmc only. */
enum { MCI_LEAF_CONTINUOUS = 0, MCI_LEAF_DISCRETE = 1, MCI_LEAF_FERMIK = 2 };

/*
 * Built-in integrands.  A Julia closure cannot run on the device, so the
 * integrand `f(x, c)` / `f(idx, x, c)` (main.jl:25-33) becomes a device
 * function selected by id.  Each id fixes the variable layout it reads
 * (validated in mci_create).  `userdata` replaces `config.userdata`.
 */
enum mci_integrand_id {
  MCI_F_GAUSS_ND = 1,       /* (2pi)^(-d/2) exp(-|x|^2/2), one Continuous pool, dof=[[d]]     (BASELINE C2) */
  MCI_F_LOG_OVER_SQRT = 2,  /* log(x)/sqrt(x), pool, dof=[[1]]        test/montecarlo.jl:129-134 (C1)      */
  MCI_F_GENZ_CORNER_POOL = 3, /* (1+sum c_i x_i)^-(d+1), pool dof=[[d]], c_i=userdata[i]         (C3 A)   */
  MCI_F_GENZ_CORNER_COMP = 4, /* same, d-leaf CompositeVar, dof=[[1]]                             (C3 B)   */
  MCI_F_BUBBLE = 5,         /* example/bubble.jl:50-75, vars (R,theta,phi,T,Ext)                  (C4)     */
  MCI_F_PRODUCT_4 = 6,      /* f_k = prod_i (1+a_k (x_i-1/2)), 4 integrands, pool dof=[[d]]x4     (C5)     */
  MCI_F_SPHERE1 = 7,        /* x1^2+x2^2<1, pool dof=[[2]]            test/montecarlo.jl:4-9               */
  MCI_F_SPHERE2 = 8,        /* circle + sphere, dof=[[2],[3]]         test/montecarlo.jl:19-49             */
  MCI_F_DISCRETE_X = 9,     /* f = x[1], Discrete leaf               test/montecarlo.jl:104-111           */
  MCI_F_DISCRETE2_CONST = 10, /* f = 1, composite of two Discrete    test/montecarlo.jl:113-120           */
  MCI_F_LATTICE_POOL = 11,  /* 1/(1-cos x cos y cos z)/pi^3, pool dof=[[3]]   test/montecarlo.jl:136-147  */
  MCI_F_LATTICE_COMP = 12,  /* same with CompositeVar(X,Y,Z), dof=1           test/montecarlo.jl:149-164  */
  MCI_F_HYPERSPHERE3 = 13,  /* 3 nested unit balls, dof=[[2],[3],[4]] test/montecarlo.jl:200-216         */
  MCI_F_POLY3 = 14,         /* (x, x^2, x^3), dof=[[1],[1],[1]]       test/thread.jl                       */
  MCI_F_X2_PLUS_Y2 = 15,    /* x1^2+x2^2, pool dof=[[2]]              main.jl:64, vegas/montecarlo.jl:69   */
  MCI_F_BUBBLE_FERMIK = 16, /* the same polarization bubble on vars (T, K::FermiK{3}, Ext), mcmc  test/bubble_FermiK.jl:54-75 */
  MCI_F_COMPLEX1 = 17,      /* x + i x^2, one COMPLEX integrand, pool dof=[[1]]   test/montecarlo.jl:166-170 (TestComplex1) */
  MCI_F_COMPLEX2 = 18,      /* (x, i x^2), two COMPLEX integrands, dof=[[1],[1]]  test/montecarlo.jl:172-198 (TestComplex2) */
  MCI_F_USER = 100          /* the caller's own integrand: CUDA C++ source in mci_problem_t.user_source, compiled with
                               NVRTC at mci_create into a kernel specialised to the problem — the equivalent of the
                               reference taking any closure f(x, c) (main.jl:25-33).  The source defines
                                 template <class XA> __device__ void mci_user_integrand(const XA &X,
                                                                   const double *userdata, double *w);
                               X(leaf, pos): value of pool slot pos (1-based, the reference's x[pos]) of leaf variable
                               `leaf` (0-based: top-level order, then leaf order inside a CompositeVar);
                               w[0 .. n_integrand-1] receives the integrand values (weight_kind = MCI_WEIGHT_COMPLEX:
                               w[2i], w[2i+1] = real and imaginary part of integrand i).                            */
};

/* `measure` kwarg (main.jl:35-37, 83).  Default = obs[i] += relw[i]. */
enum mci_measure_kind {
  MCI_MEASURE_DEFAULT = 0,
  MCI_MEASURE_BIN_BY_DISCRETE = 1, /* obs[0][Ext-lower] += relw[0]; example/bubble.jl:81-84.
                                     The binning leaf is the LAST leaf of the problem.      */
  MCI_MEASURE_USER = 2 /* the caller's measure(x, obs, relw, config) / measure(idx, x, obs, relw, config) as CUDA C++
                          source in user_source (compiled by NVRTC, like MCI_F_USER):
                            template <class XA> __device__ void mci_user_measure(const XA &X, double *obs,
                                                               const double *relw, const double *userdata);
                          obs = the flattened config.observable (integrand i at the running offset of obs_len);
                          relw[i] = relative weight of integrand i (mcmc: zero except the current integrand)      */
};

/* one leaf map (Continuous ctor kwargs variable.jl:137, Discrete ctor :299) */
typedef struct mci_leaf_t {
  int32_t kind;    /* MCI_LEAF_*                                                    */
  int32_t nbins;   /* Continuous: number of increments = length(grid)-1 (`ninc`-1,
                      variable.jl:147); Discrete: upper-lower+1                     */
  int32_t adapt;   /* 0/1                                                           */
  int32_t lower_i; /* Discrete: `lower`                                             */
  double lower;    /* Continuous bounds; FermiK: kF                                 */
  double upper;    /*                    FermiK: dk (half width of the radial shell) */
  double alpha;    /* learning rate of train! (variable.jl:206-239, :369-382); FermiK: maxK (stored, never read: sampler.jl) */
} mci_leaf_t;

/*
 * POD mirror of the Configuration fields the solver loops consume
 * (configuration.jl:37-60) plus the integrate() kwargs that reach the hot
 * loop (main.jl:71-90).
 */
typedef struct mci_problem_t {
  int32_t solver;          /* MCI_SOLVER_*                                               */
  int32_t integrand_id;    /* enum mci_integrand_id                                      */
  int32_t measure_kind;    /* enum mci_measure_kind                                      */
  int32_t n_integrand;     /* config.N (without the normalisation integrand)             */
  int32_t n_topvar;        /* length(config.var)                                         */
  int32_t n_leaf;          /* sum of leaves                                              */
  const int32_t *topvar_nleaf;  /* [n_topvar] 1 for Continuous/Discrete, k for CompositeVar */
  const int32_t *topvar_offset; /* [n_topvar] `var.offset` (informational: device slots are compact) */
  const mci_leaf_t *leaf;  /* [n_leaf] in top-level order then leaf order                */
  const int32_t *dof;      /* [n_integrand][n_topvar] row-major; the all-zero normalisation
                              row (configuration.jl:153) is appended by the library      */
  const int32_t *obs_len;  /* [n_integrand] 1 for scalar observables                     */
  const int32_t *neighbor_ptr; /* CSR over Nd=n_integrand+1 nodes (0-based), or NULL for the
                                  default chain (configuration.jl:201-209); mcmc only   */
  const int32_t *neighbor_idx;
  const double *userdata;  /* [n_userdata]                                               */
  int32_t n_userdata;
  int32_t n_block;         /* total blocks of one iteration over ALL ranks (main.jl:121) */
  uint64_t neval_per_block;/* nevalperblock = neval / block (main.jl:232).  For the chain
                              solvers = walkers_per_block * steps per walker            */
  uint64_t seed;           /* Philox key                                                 */
  int32_t measurefreq;     /* main.jl:84                                                 */
  int32_t walkers_per_block; /* chain solvers: concurrent chains per block; 0 = as many whole CTAs
                                as fill the GPU with chains of >= 512 steps (depends on the
                                number of local blocks: read it back with mci_chain_info)   */
  double thermal_ratio;    /* mcmc burn-in (mcmc/montecarlo.jl:77,133)                   */
  int32_t device;          /* CUDA device ordinal                                        */
  int32_t hist_policy;     /* 0 = auto; see MCI_HIST_* (performance knob, not semantics) */
  int32_t jit;             /* MCI_JIT_*: compile the kernel for this problem with NVRTC, its layout baked in as
                              compile-time tables.  AUTO: when one iteration has >= 2^24 samples (and NVRTC loads),
                              else the ahead-of-time kernel; MCI_F_USER always jits                                */
  int32_t reserved_;
  const char *user_source; /* MCI_F_USER: NUL-terminated CUDA C++ source, copied by mci_create; else NULL         */
  uint64_t chain_burnin;   /* chain solvers: extra update steps every walker makes BEFORE the reference's loop starts,
                              touching no statistic (0 = the reference).  A GPU cuts a block into many short chains
                              and each starts from a draw of the VEGAS map, not of the target density: chains of
                              ~10^3 steps carry a visible start bias, which this burns away          */
  /* ---- ABI v3 ---- */
  int32_t mcmc_start_idx;  /* mcmc: the integrand every chain starts in, 1-based — the `idx` kwarg of MCMC.montecarlo
                              (mcmc/montecarlo.jl:76); 0 = the reference's default 1                                 */
  int32_t weight_kind;     /* MCI_WEIGHT_*: the `type` kwarg of Configuration (configuration.jl:108): Float64 or
                              ComplexF64 integrand values and observables (main.jl:276-286,302-305)                  */
} mci_problem_t;

/* weight_kind.  COMPLEX: every integrand value is (re, im); every observable entry is (re, im) — obs_len[i] counts
 * COMPLEX entries and block_obs holds 2 * sum(obs_len) doubles per block, interleaved re, im; the sampling weight of
 * a value is abs(z) = hypot(re, im) wherever the reference takes abs(weight) (vegas/montecarlo.jl:172,
 * vegas_mc/montecarlo.jl:159,200, mcmc/updates.jl:40,95,130).  Built-in real integrands stay real. */
enum { MCI_WEIGHT_REAL = 0, MCI_WEIGHT_COMPLEX = 1 };

enum { MCI_JIT_AUTO = 0, MCI_JIT_ON = 1, MCI_JIT_OFF = 2 };
/* where the per-bin variance histogram (accumulate!, variable.jl:196-200) is kept during a launch; every policy
 * produces the same sums up to floating-point order (DESIGN.md 5.1 has the measurements behind AUTO):
 *   SMEM_ATOMIC  one fp64 copy per CTA in shared memory, red.shared.add.f64
 *   GLOBAL_RED   RED.E.ADD.F64 into a few L2-resident copies (layouts whose histograms do not fit in shared memory)
 *   WARP_PRIVATE one copy per warp, MATCH.ANY conflict ranking, no atomics (generic VEGAS kernel only; kept for study)
 *   SMEM_STRIPED 16 bank-striped copies per CTA, one 1024-thread CTA per SM (single-pool VEGAS kernel; AUTO picks it)
 *   SMEM_FIXED   (single-pool VEGAS kernel, opt-in) the striped copies as 64-bit FIXED-POINT cells of two u32 limbs updated
 *                with the native shared-memory integer add (sm_100a has no native shared fp64 add: every fp64 update is a
 *                load + add + compare-and-swap).  A sample's weight is scaled by a power of two chosen on the device from
 *                the largest bin of the previous iteration's histogram (2^58 units), rounded to nearest once and added
 *                exactly: the sums are order-independent, i.e. bit-reproducible.  Iterations whose histogram has not settled
 *                within 2^8 of the one before (the first two of an untrained map) run the fp64 kernel; a bin that outgrows
 *                the 2^6 of headroom is detected at the add that wraps and the iteration is recomputed in fp64 at once
 *                (mci_hist_rescues counts them).  Also available on any layout of the generic VEGAS kernel (one shared copy). */
enum { MCI_HIST_AUTO = 0, MCI_HIST_SMEM_ATOMIC = 1, MCI_HIST_GLOBAL_RED = 2, MCI_HIST_WARP_PRIVATE = 3, MCI_HIST_SMEM_STRIPED = 4,
       MCI_HIST_SMEM_FIXED = 5 };

/*
 * Per-iteration statistics handed back to the host: what `_block!` +
 * `addConfig!` + MPIreduceConfig! leave in obsSum/summedConfig
 * (main.jl:236-292, configuration.jl:252-299).  Any pointer may be NULL.
 */
typedef struct mci_iteration_out_t {
  double *block_obs;   /* [n_block][sum obs_len]  config_n.observable of each block (un-normalised);
                          MCI_WEIGHT_COMPLEX: [n_block][sum obs_len][2] = (re, im)                       */
  double *block_norm;  /* [n_block]               config_n.normalization (incl. the 1e-10 floor)    */
  double *hist;        /* [sum_leaf nbins]        summed leaf histograms incl. (n_block+1)*1e-10     */
  double *visited;     /* [Nd]                    incl. floors                                       */
  double *propose;     /* [3][Nd][max(Nd,Nv)]     row-major (Julia's array is column-major)          */
  double *accept;      /* [3][Nd][max(Nd,Nv)]                                                        */
  uint64_t *neval;     /* [1] integrand evaluations (config.neval)                                   */
} mci_iteration_out_t;

typedef struct mci_handle_t mci_handle_t;

#ifndef __CUDACC_RTC__ /* host entry points: not visible to the device-only JIT compilation */

/* library / ABI version (MCI_ABI_VERSION) */
int mci_abi_version(void);

/*
 * Build device state for one problem on one GPU.
 * Replaces: Configuration(...) ctor marshalling (configuration.jl:105-194) +
 * the per-thread deepcopy of configs (main.jl:130-137).
 * Maps start as the reference constructors leave them: uniform grids
 * (variable.jl:137), Discrete distribution uniform (variable.jl:305-315),
 * reweight = 1/Nd (configuration.jl:172-175).
 */
int mci_create(const mci_problem_t *problem, mci_handle_t **out);

/* release everything owned by the handle (NULL is a no-op) */
void mci_destroy(mci_handle_t *h);

/* message of the last error on this handle; h == NULL -> last mci_create error of this thread */
const char *mci_last_error(const mci_handle_t *h);

/* sizes the caller needs to allocate the arrays below */
int mci_sizes(const mci_handle_t *h, int32_t *grid_len /* sum (nbins+1) over Continuous */,
              int32_t *disc_len /* sum nbins over Discrete */, int32_t *hist_len /* sum nbins all leaves */,
              int32_t *obs_len /* sum obs_len, in DOUBLES (x2 for MCI_WEIGHT_COMPLEX) */, int32_t *nd /* Ni+1 */,
              int32_t *pa_len /* 3*Nd*max(Nd,Nv) */);

/*
 * Upload maps (warm restart `config=result.config`, main.jl:95-97; and the
 * host-side train! path).  grids: concatenated Continuous grids (nbins+1
 * each); disc_dist: concatenated Discrete `distribution`; disc_accum:
 * concatenated `accumulation` (nbins+1 each, variable.jl:314-315);
 * reweight[Nd].  NULL = leave unchanged.
 */
int mci_set_maps(mci_handle_t *h, const double *grids, const double *disc_dist, const double *disc_accum,
                 const double *reweight);
int mci_get_maps(mci_handle_t *h, double *grids, double *disc_dist, double *disc_accum, double *reweight);

/*
 * One iteration's blocks: the `for _ in 1:block/nprocs  _block!(...)` loop
 * plus MPIreduce of the statistics (main.jl:144-180) collapsed into one call.
 * Runs this rank's blocks {b : b mod world == rank} on the device (fill
 * kernel), reduces per-CTA partials in a fixed order, all-reduces over ranks
 * when a communicator is attached, and copies the statistics to `out`.
 * `iter` is the 0-based iteration number: it enters the Philox key so every
 * iteration draws fresh samples.
 * Replaces: Vegas.montecarlo (vegas/montecarlo.jl:72-191), VegasMC.montecarlo
 * (vegas_mc/montecarlo.jl:112-241), MCMC.montecarlo (mcmc/montecarlo.jl:72-184).
 */
int mci_run_iteration(mci_handle_t *h, uint64_t iter, mci_iteration_out_t *out);

/*
 * Asynchronous halves of mci_run_iteration for drivers that keep the loop on
 * the device: launch only (no host sync) / fetch the statistics of the last
 * launched iteration.
 */
int mci_launch_iteration(mci_handle_t *h, uint64_t iter);
int mci_fetch_iteration(mci_handle_t *h, mci_iteration_out_t *out);

/*
 * Device-side grid refinement on the (all-reduced) histogram of the last
 * iteration: Dist.train! for every leaf with adapt=true, then histogram
 * reset.  Replaces variable.jl:206-239 (Continuous: smooth -> rescale ->
 * regrid), variable.jl:369-382 (Discrete) and common.jl:43-83.
 */
int mci_refine(mci_handle_t *h);
/* on != 0: mci_refine / mci_reweight only enqueue (no host synchronisation); a failed train! assert
 * then surfaces at the next mci_fetch_iteration.  For drivers that keep the whole loop on the device. */
int mci_set_async(mci_handle_t *h, int32_t on);

/*
 * doReweight! (main.jl:322-346) on the visited counts of the last iteration.
 * goal == NULL -> no reweight_goal.  Updates the device copy of reweight.
 */
int mci_reweight(mci_handle_t *h, double gamma, const double *goal);

/*
 * Multi-GPU: one process per GPU.  Rank 0 calls mci_comm_unique_id and the
 * host distributes the 128 bytes (MPI.bcast in Julia, torch.distributed in
 * the Python harness); every rank then calls mci_comm_init.  Afterwards
 * mci_run_iteration ends with ONE ncclAllReduce(sum, f64) of the packed
 * statistics buffer, replacing MPIreduce + MPIreduceConfig! + MPIbcastConfig!
 * + doReweightMPI!'s Bcast (main.jl:177-188, configuration.jl:264-321,
 * main.jl:348-356).  Blocks are partitioned b mod world == rank
 * (main.jl:122,152-166).
 */
int mci_comm_unique_id(void *id128);
/* Communicators are kept per process, keyed by the unique id: a later handle that passes an id this process has
 * already used (same rank, world, device) JOINS that communicator instead of creating one, so the successive
 * integrate() calls of a job pay for NCCL's start-up once.  Handles that share a communicator must not run
 * iterations concurrently.  mci_destroy leaves the communicator alone; mci_comm_finalize destroys them all. */
int mci_comm_init(mci_handle_t *h, const void *id128, int32_t rank, int32_t world);
/* collective: every rank calls it after destroying its last handle (optional — process exit also ends them) */
int mci_comm_finalize(void);
/* partition blocks without a communicator (the caller reduces, e.g. torch.distributed) */
int mci_set_partition(mci_handle_t *h, int32_t rank, int32_t world);
/* device pointer + length (doubles) of the packed statistics buffer, for external collectives */
int mci_stats_device_ptr(mci_handle_t *h, void **dptr, int64_t *n_doubles);
/* after an external collective on that buffer: refresh the host copy mci_fetch_iteration reads */
int mci_stats_sync_host(mci_handle_t *h);

/*
 * Parity probe (L0): regenerate samples [first, first+count) of block `block`
 * of iteration `iter` with the current maps, VEGAS sampler only.  Layout:
 * x[count][n_draw], bins[count][n_draw] (0-based grid bin / discrete index),
 * jac[count], weights[count][Ni] ([count][Ni][2] = (re, im) for MCI_WEIGHT_COMPLEX).  n_draw = draws per sample in the
 * reference's draw order (vegas/montecarlo.jl:122-131, sampler.jl:431-440).
 */
int mci_trace_samples(mci_handle_t *h, uint64_t iter, int32_t block, uint64_t first, int32_t count, double *x,
                      int32_t *bins, double *jac, double *weights);

/*
 * Compile-only check of the JIT path (no GPU needed): builds the device description of `problem`, generates the
 * specialised program for its solver and runs NVRTC on it.  Returns MCI_OK, or MCI_ERR_UNSUPPORTED with the compiler
 * log in `log` (NUL-terminated, at most log_cap bytes).  Lets a host validate user_source before it has a device.
 */
int mci_jit_check(const mci_problem_t *problem, char *log, int32_t log_cap);

/*
 * Number of non-finite integrand values seen by the fill kernel of the last fetched iteration, summed over this
 * rank's blocks.  Replaces the `debug = true` warnings of the solver loops ("abs of the integrand ... is not
 * finite", vegas/montecarlo.jl:176-178, vegas_mc/montecarlo.jl:189-191, mcmc/montecarlo.jl:142-144): the host prints
 * the warning when the count is non-zero.
 */
int mci_nonfinite_count(const mci_handle_t *h, uint64_t *n);

/*
 * Small jobs: run `niter` whole iterations — fill -> statistics [-> all-reduce] -> doReweight! (chain solvers) ->
 * train! -> staged maps — as ONE CUDA graph launch per iteration with no host synchronisation in between, and copy
 * every iteration's statistics out at the end.  out[k] receives iteration first_iter + k (any pointer may be NULL).
 * Equivalent to niter x (mci_run_iteration, mci_reweight(gamma, goal), mci_refine) — the body of the loop of
 * main.jl:142-207 with adapt = true — minus 2 host round trips per iteration.  `do_reweight` / `do_refine` select the
 * steps (main.jl:183, :193-198).
 */
int mci_run_iterations(mci_handle_t *h, uint64_t first_iter, int32_t niter, double gamma, const double *goal,
                       int32_t do_reweight, int32_t do_refine, mci_iteration_out_t *out);

/*
 * MCI_HIST_SMEM_FIXED: number of fetched iterations whose fixed-point histogram left its 64-bit range (a bin outgrew the
 * largest bin of the iteration before by more than 2^6) and were recomputed on the spot by the fp64 launch that follows
 * every fixed-point launch.  Informational: the statistics handed out are the fp64 ones.
 */
int mci_hist_rescues(const mci_handle_t *h, uint64_t *n);

/* measurement helpers ---------------------------------------------------- */
/* number of kernel launches issued by this handle so far */
int mci_launch_count(const mci_handle_t *h, uint64_t *n);
/* CUDA-event time (ms) of the fill kernel / of the whole device part of the last iteration */
int mci_last_timing(mci_handle_t *h, float *fill_ms, float *iteration_ms);
/* CUDA-event durations (ms) of the fill kernel of the most recent iterations, oldest first (at most 64 are kept and
 * at most `cap` returned; *n = how many).  Synchronises the handle's stream.  Lets a driver that keeps the loop on
 * the device (mci_launch_iteration) read per-launch kernel times of its timed region afterwards. */
int mci_fill_history(mci_handle_t *h, float *ms, int32_t cap, int32_t *n);
/* CUDA-event stopwatch on the handle's stream (all kernels of a handle run on its own stream, which
 * torch.cuda.Event cannot see): start records an event, stop records another, waits for it and returns ms */
int mci_timer_start(mci_handle_t *h);
int mci_timer_stop(mci_handle_t *h, float *ms);
/* FP64 DFMA peak micro-benchmark on `device` (TFLOP/s, FMA = 2 flop) */
int mci_fp64_peak(int32_t device, double *tflops);
/* launch shape of the fill kernel: CTAs, threads per CTA, dynamic shared memory, MCI_HIST_* actually used,
 * tiles per block and samples per thread (VEGAS; 0 for the chain solvers) */
int mci_launch_info(const mci_handle_t *h, int32_t *grid, int32_t *threads, int64_t *smem_bytes, int32_t *hist_mode,
                    int32_t *tiles_per_block, int32_t *samples_per_thread);
/* chain solvers: the decomposition in use — chains per block (as given, or as chosen for walkers_per_block = 0),
 * steps per chain, CTAs per block */
int mci_chain_info(const mci_handle_t *h, int32_t *walkers_per_block, uint64_t *steps_per_walker, int32_t *ctas_per_block);
/* name of the fill kernel variant chosen for this handle (static string) */
const char *mci_kernel_name(const mci_handle_t *h);

#endif /* !__CUDACC_RTC__ */

#ifdef __cplusplus
}
#endif
#endif /* MCIB200_H */
