/*
 * mci_abi.cu — the C ABI of libmcib200.so (include/mcib200.h): handle management, launch
 * planning, the per-iteration launch sequence and the NCCL all-reduce of the packed statistics.
 *
 * Host roles replaced (reference paths): the per-iteration block loop + statistics reduction of
 * integrate (src/main.jl:144-199), Configuration marshalling (src/configuration.jl:105-194),
 * MPIreduceConfig!/MPIbcastConfig! (src/configuration.jl:264-321).
 *
 * There is no CPU fallback anywhere in this file: without a CUDA device mci_create fails.
 */
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "mci_aux.cuh"
#include "mci_chain.cuh"
#include "mci_jit.h"
#include "mci_kernels.h"
#include "mci_vegas.cuh"

using namespace mci;

/* ------------------------------------------------------------------------------------------ */
/* NCCL, bound at run time (dlopen) so the library loads on hosts without libnccl              */
/* ------------------------------------------------------------------------------------------ */
namespace {
typedef struct { char internal[128]; } nccl_uid_t;
typedef void *nccl_comm_t;
struct NcclApi {
  void *lib = nullptr;
  int (*GetUniqueId)(nccl_uid_t *) = nullptr;
  int (*CommInitRank)(nccl_comm_t *, int, nccl_uid_t, int) = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
  int (*CommDestroy)(nccl_comm_t) = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
  bool ok = false;
};
NcclApi g_nccl;
std::once_flag g_nccl_once;

void nccl_load_once() {
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char *n : names) {
    g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.lib) break;
  }
  if (!g_nccl.lib) return;
  g_nccl.GetUniqueId = (int (*)(nccl_uid_t *))dlsym(g_nccl.lib, "ncclGetUniqueId");
  g_nccl.CommInitRank = (int (*)(nccl_comm_t *, int, nccl_uid_t, int))dlsym(g_nccl.lib, "ncclCommInitRank");
  g_nccl.AllReduce = (int (*)(const void *, void *, size_t, int, int, nccl_comm_t, cudaStream_t))dlsym(g_nccl.lib, "ncclAllReduce");
  g_nccl.CommDestroy = (int (*)(nccl_comm_t))dlsym(g_nccl.lib, "ncclCommDestroy");
  g_nccl.GetErrorString = (const char *(*)(int))dlsym(g_nccl.lib, "ncclGetErrorString");
  g_nccl.ok = g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.AllReduce && g_nccl.CommDestroy;
}
/* two handles on two host threads may attach communicators at the same time: bind libnccl exactly once */
bool nccl_load() {
  std::call_once(g_nccl_once, nccl_load_once);
  return g_nccl.ok;
}
constexpr int NCCL_DOUBLE = 8, NCCL_SUM = 0;

thread_local char g_err[256];
/* auto policy of the pool kernel = 16 bank-striped shared copies: measured 39.8 vs 28.0 G samples/s (one shared copy) on C2, profiles/r1_hist_policy_sweeps.jsonl */
constexpr bool MCI_AUTO_STRIPED = true;
/* auto policy of the generic kernel wherever the cells of all leaves fit in shared memory: fixed-point cells.  Measured
 * (B200, 16 blocks x 2^22 samples, G samples/s, fp64 automatic choice -> fixed point): 16-leaf composite Genz (BASELINE C3-B,
 * fp64 = L2 reductions) 9.07 -> 16.41; 6-leaf composite 23.50 -> 32.06; example/bubble.jl under :vegas 19.90 -> 22.12
 * (profiles/r2_hist_fixed_point.jsonl) */
#ifndef MCI_AUTO_FIXED_GENERIC_DEFAULT
#define MCI_AUTO_FIXED_GENERIC_DEFAULT 1
#endif
constexpr bool MCI_AUTO_FIXED_GENERIC = MCI_AUTO_FIXED_GENERIC_DEFAULT != 0;
/* auto policy of the pool kernel: the striped copies as fixed-point cells.  Measured (same job, G samples/s, fp64 striped
 * copies -> fixed point): C2 42.3 -> 53.0, C3-A 22.7 -> 34.5, C5 integrand under :vegas 16.3 -> 21.7
 * (profiles/r2_hist_fixed_point.jsonl) */
#ifndef MCI_AUTO_FIXED_POOL_DEFAULT
#define MCI_AUTO_FIXED_POOL_DEFAULT 1
#endif
constexpr bool MCI_AUTO_FIXED_POOL = MCI_AUTO_FIXED_POOL_DEFAULT != 0;
}  // namespace

struct mci_handle_t {
  mci_problem_t p;
  DevProblem dp;
  PackedLayout Y;
  int nd, nv, ni, nleaf, pa_len, obs_total, hist_len, grid_len, disc_len, accum_len;
  std::vector<int> leaf_grid_off, leaf_dist_off, leaf_acc_off;
  std::vector<double> alpha;
  int rank = 0, world = 1;
  int device = 0, n_sm = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev_iter0 = nullptr, ev_iter1 = nullptr, ev_t0 = nullptr, ev_t1 = nullptr;
  /* ring of event pairs around the fill kernel of the most recent iterations (mci_fill_history) */
  static constexpr int FILL_RING = 64;
  cudaEvent_t ev_fill0[FILL_RING] = {}, ev_fill1[FILL_RING] = {};
  uint64_t n_iterations = 0;
  /* device state */
  double *d_grids = nullptr, *d_dist = nullptr, *d_accum = nullptr, *d_reweight = nullptr, *d_goal = nullptr;
  double *d_alpha = nullptr, *d_gmaps = nullptr;
  int *d_leaf_grid_off = nullptr, *d_leaf_dist_off = nullptr, *d_leaf_acc_off = nullptr;
  double *d_tile_obs = nullptr, *d_tile_cnt = nullptr, *d_cta_hist = nullptr, *d_cta_stats = nullptr, *d_packed = nullptr;
  unsigned long long *d_flags = nullptr; /* [4]: non-finite weights, mcmc walkers without a start, fixed-point range, spare */
  int *d_err = nullptr, *d_fixed_exp = nullptr;
  double *d_fixed_state = nullptr; /* largest histogram bin of the previous iteration (fixed_scale_kernel) */
  double *d_stage = nullptr;       /* staging area of mci_set_maps on fixed-point handles: grids | dist */
  /* the buffers above are carved out of three allocations (cudaMalloc / cudaFree / cudaMallocHost cost 0.1-1 ms each
   * and synchronise the device: a job of the reference's default size, neval = 1e4, is over in 4 ms):
   * the fixed state, the per-launch partials (re-planned by mci_set_partition) and the pinned host mirrors */
  void *d_state_arena = nullptr, *d_launch_arena = nullptr, *h_arena = nullptr;
  size_t state_cap = 0, launch_cap = 0, host_cap = 0; /* capacities, for the block cache */
  /* mci_run_iterations: the statistics (and flags) of every iteration of one call, kept on the device until the end */
  void *d_history = nullptr, *h_history = nullptr;
  size_t history_cap = 0, h_history_cap = 0;
  /* pinned host mirrors */
  double *h_packed = nullptr;
  unsigned long long *h_flags = nullptr;
  int *h_err = nullptr;
  /* launch plan */
  const void *kernel = nullptr;
  const void *kernel_first = nullptr; /* HIST_SMEM_FIXED: the fp64 striped kernel that runs while no scale exists yet */
  const char *kernel_name = "";
  bool is_pool = false;
  int threads = 0, grid = 0, hist_mode = 0, n_hist_copies = 1, maps_in_smem = 1;
  size_t smem = 0;
  int tiles_per_block = 0, spt = 0, ctas_per_block = 0, max_grid = 0, chain_rows = 0, coop_ci = 0, walkers = 1;
  bool has_fermik = false; /* some variable is a FermiK (mcmc only; no cooperative create!) */
  uint64_t steps = 0;
  bool maps_dirty = true, have_iteration = false, async_mode = false;
  nccl_comm_t comm = nullptr;
  /* JIT specialisation (mci_jit.cu) */
  bool use_jit = false;
  std::string user_source, kernel_name_jit;
  bool kernel_name_is_jit() const { return kernel_name == kernel_name_jit.c_str() && !kernel_name_jit.empty(); }
  const void *jit_lib = nullptr, *trace_kernel = nullptr;
  bool jit_failed = false;
  int jit_mode = 0, chain_maxt = 1024;
  uint64_t launches = 0;
  uint64_t nonfinite = 0; /* non-finite integrand values of the last fetched iteration (flags[0]) */
  uint64_t hist_rescues = 0; /* fetched iterations whose fixed-point histogram was recomputed in fp64 (flags[3]) */
  float fill_ms = 0.f, iter_ms = 0.f;
  char err[256];
};

namespace {

int fail(mci_handle_t *h, int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(h ? h->err : g_err, 256, fmt, ap);
  va_end(ap);
  return code;
}

#define CU(h, call)                                                                                              \
  do {                                                                                                           \
    cudaError_t e_ = (call);                                                                                     \
    if (e_ != cudaSuccess) return fail(h, MCI_ERR_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

/* integrands per id: number of values and the (leaf, slots) they read; -1 slots = dof[0] of the pool */
struct FidInfo { int ni; int nleaf_min; int slots0; };
bool fid_info(int fid, FidInfo &I) {
  switch (fid) {
    case MCI_F_GAUSS_ND: I = {1, 1, -1}; return true;
    case MCI_F_LOG_OVER_SQRT: I = {1, 1, 1}; return true;
    case MCI_F_GENZ_CORNER_POOL: I = {1, 1, -1}; return true;
    case MCI_F_GENZ_CORNER_COMP: I = {1, 1, 1}; return true;
    case MCI_F_BUBBLE: I = {1, 5, 1}; return true;
    case MCI_F_BUBBLE_FERMIK: I = {1, 4, 1}; return true;
    case MCI_F_PRODUCT_4: I = {4, 1, -1}; return true;
    case MCI_F_SPHERE1: I = {1, 1, 2}; return true;
    case MCI_F_SPHERE2: I = {2, 1, 3}; return true;
    case MCI_F_DISCRETE_X: I = {1, 1, 1}; return true;
    case MCI_F_DISCRETE2_CONST: I = {1, 1, 1}; return true;
    case MCI_F_LATTICE_POOL: I = {1, 1, 3}; return true;
    case MCI_F_LATTICE_COMP: I = {1, 3, 1}; return true;
    case MCI_F_HYPERSPHERE3: I = {3, 1, 4}; return true;
    case MCI_F_POLY3: I = {3, 1, 1}; return true;
    case MCI_F_X2_PLUS_Y2: I = {1, 1, 2}; return true;
    case MCI_F_COMPLEX1: I = {1, 1, 1}; return true;
    case MCI_F_COMPLEX2: I = {2, 1, 1}; return true;
    case MCI_F_USER: I = {0, 1, 0}; return true; /* any number of values, any layout: the source decides */
  }
  return false;
}

/* doubles of userdata a built-in integrand reads (mci_integrands.cuh): a short blob would be zero-filled and give a
 * wrong integral without any error, so mci_create insists on at least this many */
int userdata_needed(const DevProblem &P) {
  switch (P.fid) {
    case MCI_F_GENZ_CORNER_POOL: return P.dof[0];
    case MCI_F_GENZ_CORNER_COMP: return P.nleaf;
    case MCI_F_BUBBLE:
    case MCI_F_BUBBLE_FERMIK: return 6 + P.leaf[P.nleaf - 1].nbins; /* [kF, beta, me, spin, dim, Qsize, extQ_x[Ext]...] */
    case MCI_F_PRODUCT_4: return 4;
  }
  return 0;
}

double volume_inverse(int d) { /* test/montecarlo.jl:201-204 */
  const double euler = 2.71828182845904523536028747135266249775724709369995957496696763;
  const double pi = 3.14159265358979323846;
  return pow(d / (2 * pi * euler), d / 2.0) * sqrt((double)d) * sqrt(pi);
}

/* A process-wide cache of released arenas.  cudaMalloc / cudaFree / cudaMallocHost / cudaFreeHost cost 0.1-1 ms each
 * and synchronise the device; a job of the reference's default size (neval = 1e4, main.jl:74) is over in ~1 ms of
 * device time, so a host that calls integrate() in a loop pays more for allocation than for integration.  Blocks
 * are rounded up to a power of two (>= 4 KB), at most 8 per kind are kept, none above 32 MB. */
struct BlockCache {
  std::mutex mu;
  struct Blk { void *p; size_t cap; int dev; };
  std::vector<Blk> dev, host;
  static size_t round_up(size_t n) { size_t c = 4096; while (c < n) c <<= 1; return c; }
  cudaError_t get(bool pinned, int device, size_t bytes, void **out, size_t *cap) {
    const size_t want = round_up(bytes);
    {
      std::lock_guard<std::mutex> lk(mu);
      std::vector<Blk> &v = pinned ? host : dev;
      for (size_t i = 0; i < v.size(); ++i)
        if (v[i].cap == want && (pinned || v[i].dev == device)) {
          *out = v[i].p; *cap = want;
          v.erase(v.begin() + i);
          return cudaSuccess;
        }
    }
    *cap = want;
    return pinned ? cudaMallocHost(out, want) : cudaMalloc(out, want);
  }
  void put(bool pinned, int device, void *p, size_t cap) {
    if (!p) return;
    if (cap <= ((size_t)32 << 20)) {
      std::lock_guard<std::mutex> lk(mu);
      std::vector<Blk> &v = pinned ? host : dev;
      if (v.size() < 8) { v.push_back({p, cap, device}); return; }
    }
    if (pinned) cudaFreeHost(p); else cudaFree(p);
  }
};
BlockCache g_blocks;

/* bump allocation out of one cudaMalloc: sizes are summed first (256-byte aligned), then the pointers are handed out */
struct Arena {
  size_t total = 0;
  char *base = nullptr;
  template <class T>
  size_t reserve(size_t n) { const size_t off = total; total += (std::max<size_t>(n, 1) * sizeof(T) + 255) & ~(size_t)255; return off; }
  template <class T>
  T *at(size_t off) const { return reinterpret_cast<T *>(base + off); }
};
int arena_device(mci_handle_t *h, Arena &A, void **keep, size_t *cap, bool zero = true) {
  CU(h, g_blocks.get(false, h->device, std::max<size_t>(A.total, 256), keep, cap));
  if (zero) CU(h, cudaMemsetAsync(*keep, 0, std::max<size_t>(A.total, 256), h->stream));
  A.base = static_cast<char *>(*keep);
  return MCI_OK;
}

/* kernel of one family for this handle: the ahead-of-time instantiation, or — user integrand / jit requested — the
 * NVRTC build of the same template specialised to the problem (mci_jit.cu).  kind: 0 pool, 1 generic, 2 vegasmc, 3 mcmc */
const void *resolve_kernel(mci_handle_t *h, int kind, int mode, int threads, bool exact) {
  const DevProblem &P = h->dp;
  if (!h->use_jit) {
    switch (kind) {
      case 0: return lookup_vegas_pool(P.fid, P.maxdof[0], exact, mode, &h->kernel_name);
      case 1: return lookup_vegas_generic(P.fid, &h->kernel_name, threads);
      case 2: return lookup_vegasmc(P.fid, &h->kernel_name);
      default: return lookup_mcmc(P.fid, &h->kernel_name);
    }
  }
  if (kind == 0 && h->jit_mode == MCI_JIT_AUTO && P.fid != MCI_F_USER) {
    /* an exact ahead-of-time pool instantiation is already what the JIT would build */
    const void *k = lookup_vegas_pool(P.fid, P.maxdof[0], exact, mode, &h->kernel_name);
    if (k && strstr(h->kernel_name, ",exact>")) return k;
  }
  char name[160];
  switch (kind) {
    case 0:
      if ((mode == HIST_SMEM_STRIPED || mode == HIST_SMEM_FIXED) && P.maxdof[0] > 20) return nullptr; /* 1024 threads x 64 registers */
      snprintf(name, sizeof name, "mci::vegas_pool_kernel<%d, %d, %d, %d, %s>", P.fid, P.maxdof[0], mode, threads, exact ? "true" : "false");
      break;
    case 1: snprintf(name, sizeof name, "mci::vegas_generic_kernel<%d, %d>", P.fid, threads); break;
    case 2: snprintf(name, sizeof name, "mci::vegasmc_kernel<%d>", P.fid); break;
    default: snprintf(name, sizeof name, "mci::mcmc_kernel<%d>", P.fid); break;
  }
  std::vector<std::string> exprs{name};
  if (kind == 0) { char t[80]; snprintf(t, sizeof t, "mci::vegas_pool_trace_kernel<%d, %d>", P.fid, P.maxdof[0]); exprs.push_back(t); }
  const bool with_twin = kind == 0 && mode == HIST_SMEM_FIXED; /* the fp64 striped twin of a fixed-point pool kernel: same program */
  if (with_twin) {
    char t[160];
    snprintf(t, sizeof t, "mci::vegas_pool_kernel<%d, %d, %d, %d, %s>", P.fid, P.maxdof[0], HIST_SMEM_STRIPED, threads, exact ? "true" : "false");
    exprs.push_back(t);
  }
  if (kind == 1) { char t[64]; snprintf(t, sizeof t, "mci::vegas_trace_kernel<%d>", P.fid); exprs.push_back(t); }
  std::vector<const void *> ks;
  std::string err;
  int rc = jit_compile(P, h->user_source.c_str(), exprs, ks, err);
  h->chain_maxt = 1024;
  if (rc != 0) {
    if (h->jit_mode == MCI_JIT_AUTO && P.fid != MCI_F_USER && P.measure_kind != MCI_MEASURE_USER) { /* auto: the ahead-of-time kernel serves as well */
      h->use_jit = false;
      return resolve_kernel(h, kind, mode, threads, exact);
    }
    snprintf(h->err, sizeof h->err, "%.*s", (int)sizeof h->err - 1, err.c_str());
    h->jit_failed = true;
    return nullptr;
  }
  if (ks.size() != exprs.size() + 1) { /* one kernel per name expression + the library handle */
    snprintf(h->err, sizeof h->err, "NVRTC build of %s returned %zu entries, expected %zu", name, ks.size(), exprs.size() + 1);
    h->jit_failed = true;
    return nullptr;
  }
  if (h->jit_lib) jit_unload(h->jit_lib);
  h->jit_lib = ks.back();
  h->trace_kernel = kind <= 1 ? ks[1] : nullptr;
  if (with_twin) h->kernel_first = ks[2];
  h->kernel_name_jit = std::string(name) + " [jit]";
  h->kernel_name = h->kernel_name_jit.c_str();
  return ks[0];
}

/* choose kernel, histogram policy, shared memory and grid for the current partition */
int plan(mci_handle_t *h) {
  const mci_problem_t &p = h->p;
  DevProblem &P = h->dp;
  const int n_local = p.n_block / h->world;
  const size_t map_bytes = (size_t)P.map_doubles * 8;
  const size_t hist_bytes = (size_t)P.hist_len * 8;
  const size_t SMEM_MAX = 227 * 1024;
  int policy = p.hist_policy;
  if (!P.any_adapt) policy = -1;
  int best_mis = 1;

  auto pick_hist = [&](int nwarp, size_t extra, bool allow_private) {
    /* returns mode and sets maps_in_smem */
    h->maps_in_smem = map_bytes + extra <= SMEM_MAX ? 1 : 0;
    size_t base = (h->maps_in_smem ? map_bytes : 0) + extra;
    if (policy == -1) return (int)HIST_NONE;
    if (policy == MCI_HIST_WARP_PRIVATE && allow_private && base + nwarp * hist_bytes <= SMEM_MAX) return (int)HIST_WARP_PRIVATE;
    if (policy == MCI_HIST_SMEM_ATOMIC && base + hist_bytes <= SMEM_MAX) return (int)HIST_SMEM_ATOMIC;
    if (policy == MCI_HIST_GLOBAL_RED) return (int)HIST_GLOBAL_RED;
    /* auto (measured on B200, C2, profiles/r1_hist_policy_sweeps.jsonl): one CTA-shared fp64 histogram with shared-memory CAS
     * beats per-warp private copies with MATCH.ANY conflict ranking (ADU-pipe bound) by 2.3x and L2 REDs by 1.5x */
    /* a shared histogram that leaves room for only one 256-thread CTA per SM starves the SM of warps: L2 REDs then win
     * (16-leaf composite: 9.1 vs 6.1 G samples/s) */
    if (base + hist_bytes <= (nwarp >= 16 ? SMEM_MAX : SMEM_MAX / 2)) return (int)HIST_SMEM_ATOMIC;
    return (int)HIST_GLOBAL_RED;
  };
  auto hist_smem = [&](int mode, int nwarp) -> size_t {
    return mode == HIST_WARP_PRIVATE ? nwarp * hist_bytes : (mode == HIST_SMEM_ATOMIC ? hist_bytes : 0);
  };

  if (p.solver == MCI_SOLVER_VEGAS) {
    bool pool_layout = P.nv == 1 && P.nleaf == 1 && P.leaf[0].kind == MCI_LEAF_CONTINUOUS &&
                       p.measure_kind == MCI_MEASURE_DEFAULT && P.obs_total == P.ni && P.maxdof[0] <= MCI_MAX_SLOT &&
                       p.measurefreq == 1 && P.wdim == 1;
    h->is_pool = false;
    if (pool_layout) {
      int nwarp = POOL_THREADS / 32;
      size_t red = (size_t)nwarp * (P.ni + 1) * 8;
      int mode = pick_hist(nwarp, red, false); /* the pool kernel has no per-warp private mode */
      int threads = POOL_THREADS;
      size_t hs = hist_smem(mode, nwarp);
      /* 16 bank-striped copies of the histogram, one 1024-thread CTA per SM, when they fit */
      const size_t red_s = (size_t)(POOL_THREADS_STRIPED / 32) * (P.ni + 1) * 8;
      const bool striped_fits = (size_t)POOL_MAP_REP * map_bytes + (size_t)STRIPE * hist_bytes + red_s <= SMEM_MAX;
      /* fixed-point cells (two kernels per handle: the fixed-point one and its fp64 twin): asked for, or the automatic choice */
      const bool want_fixed = (p.hist_policy == MCI_HIST_SMEM_FIXED || (p.hist_policy == MCI_HIST_AUTO && MCI_AUTO_FIXED_POOL)) &&
                              P.maxdof[0] <= 20;
      if (P.any_adapt && striped_fits && (p.hist_policy == MCI_HIST_SMEM_STRIPED || want_fixed || (p.hist_policy == MCI_HIST_AUTO && MCI_AUTO_STRIPED))) {
        mode = want_fixed ? HIST_SMEM_FIXED : HIST_SMEM_STRIPED; threads = POOL_THREADS_STRIPED; nwarp = threads / 32; red = red_s;
        hs = (size_t)STRIPE * hist_bytes + (size_t)(POOL_MAP_REP - 1) * map_bytes; h->maps_in_smem = 1;
      }
      if (h->maps_in_smem) {
        const void *k = resolve_kernel(h, 0, mode, threads, !P.need_pad);
        if (h->jit_failed) return MCI_ERR_UNSUPPORTED;
        if (k && mode == HIST_SMEM_FIXED && !h->kernel_name_is_jit()) {
          /* ahead-of-time build: the fp64 striped instantiation of the same shape is the twin */
          h->kernel_first = lookup_vegas_pool(P.fid, P.maxdof[0], !P.need_pad, HIST_SMEM_STRIPED, nullptr);
        }
        if (mode == HIST_SMEM_FIXED && (!k || !h->kernel_first)) { /* no fixed-point pair for this problem: plain striped copies */
          mode = HIST_SMEM_STRIPED; h->kernel_first = nullptr;
          k = resolve_kernel(h, 0, mode, threads, !P.need_pad);
          if (h->jit_failed) return MCI_ERR_UNSUPPORTED;
        }
        if (!k && (mode == HIST_SMEM_STRIPED || mode == HIST_SMEM_FIXED)) { /* no striped instantiation for this many slots: one shared copy */
          threads = POOL_THREADS; nwarp = threads / 32; red = (size_t)nwarp * (P.ni + 1) * 8;
          mode = pick_hist(nwarp, red, false); hs = hist_smem(mode, nwarp);
          if (h->maps_in_smem) k = resolve_kernel(h, 0, mode, threads, !P.need_pad);
          if (h->jit_failed) return MCI_ERR_UNSUPPORTED;
        }
        if (k) {
          h->kernel = k; h->is_pool = true; h->threads = threads; h->hist_mode = mode;
          h->smem = map_bytes + hs + red;
        }
      }
    }
    if (!h->is_pool) {
      int gthreads = GENERIC_THREADS;
      int nwarp = gthreads / 32;
      size_t red = (size_t)nwarp * (P.obs_total + 1) * 8;
      int mode = pick_hist(nwarp, red, true);
      /* fixed-point cells: ONE shared copy of every leaf's histogram per CTA, updated with native integer adds.  Asked for
       * explicitly, or chosen when the fp64 alternative is the L2 reductions (histograms too large to share an SM with a
       * second CTA).  Large histograms take one 1024-thread CTA per SM; the maps then stay in global memory (L1 / L2 hits)
       * unless they fit beside the cells */
      const bool fixed_ok = P.any_adapt && hist_bytes + (size_t)(GENERIC_THREADS_BIG / 32) * (P.obs_total + 1) * 8 + 1024 <= SMEM_MAX;
      if (fixed_ok && (p.hist_policy == MCI_HIST_SMEM_FIXED || (p.hist_policy == MCI_HIST_AUTO && mode != HIST_NONE && MCI_AUTO_FIXED_GENERIC))) {
        mode = HIST_SMEM_FIXED;
        const size_t red_big = (size_t)(GENERIC_THREADS_BIG / 32) * (P.obs_total + 1) * 8;
        h->maps_in_smem = map_bytes + hist_bytes + red_big <= SMEM_MAX ? 1 : 0;
        /* cells (and maps) limit the CTAs per SM: size the CTA so that an SM still holds ~1024 threads — one CTA per SM ->
         * 1024 threads, two -> 512, else 256 */
        {
          const size_t per_cta = (h->maps_in_smem ? map_bytes : 0) + hist_bytes + red;
          if (per_cta > SMEM_MAX / 2) gthreads = GENERIC_THREADS_BIG;
          else if (per_cta > SMEM_MAX / 3) gthreads = GENERIC_THREADS_MID;
        }
        nwarp = gthreads / 32;
        red = (size_t)nwarp * (P.obs_total + 1) * 8;
      }
      h->kernel = resolve_kernel(h, 1, mode, gthreads, false);
      if (h->jit_failed) return MCI_ERR_UNSUPPORTED;
      if (!h->kernel) return fail(h, MCI_ERR_UNSUPPORTED, "integrand id %d has no vegas kernel", P.fid);
      h->threads = gthreads; h->hist_mode = mode;
      h->smem = (h->maps_in_smem ? map_bytes : 0) + (mode == HIST_SMEM_FIXED ? hist_bytes : hist_smem(mode, nwarp)) + red;
    }
  } else {
    /* chain solvers: walker state in shared memory rows [row][thread]; pick the CTA size that keeps most walkers
     * resident per SM (the maps and the histogram are per CTA, so larger CTAs amortise them) */
    h->kernel = resolve_kernel(h, p.solver == MCI_SOLVER_VEGASMC ? 2 : 3, 0, 0, false);
    if (h->jit_failed) return MCI_ERR_UNSUPPORTED;
    if (!h->kernel) return fail(h, MCI_ERR_UNSUPPORTED, "integrand id %d has no chain kernel", P.fid);
    int max_nl = 1;
    for (int v = 0; v < P.nv; ++v) max_nl = std::max(max_nl, P.top_nleaf[v]);
    h->chain_rows = P.n_draw + (max_nl > 1 ? max_nl : 0); /* single-leaf variables keep shift!'s rollback cache in registers */
    { /* largest number of (slot, leaf) items one changeIntegrand creates or removes, over the neighbour graph */
      int worst = 0;
      for (int a = 0; a < P.nd; ++a)
        for (int k = P.nb_ptr[a]; k < P.nb_ptr[a + 1]; ++k) {
          int items = 0;
          for (int v = 0; v < P.nv; ++v) items += std::abs(P.dof[a * P.nv + v] - P.dof[P.nb_idx[k] * P.nv + v]) * P.top_nleaf[v];
          worst = std::max(worst, items);
        }
      int coop_min = 5; /* measured: 20 items 1.8x and 5 items 1.2x faster cooperatively, 2-3 items 2x slower */
      if (const char *e = getenv("MCIB200_COOP_MIN")) coop_min = atoi(e);
      h->coop_ci = worst >= coop_min && !h->has_fermik;
    }
    const size_t row_bytes = (size_t)h->chain_rows * (8 + 2 + (p.solver == MCI_SOLVER_VEGASMC ? 8 : 0)); /* x, u16 bin[, prob] */
    const size_t fixed = 32 * 8 + (((size_t)2 * h->pa_len + 1) / 2) * 8;
    CU(h, cudaFuncSetAttribute(h->kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_MAX));
    int best_score = -1, best_occ = 1;
    /* CTA size that keeps most walkers resident for W walkers per block (W <= 0: as many as wanted) */
    auto choose_T = [&](int W) {
      best_score = -1;
      for (int T = h->use_jit ? h->chain_maxt : 1024; T >= 128; T >>= 1) {
        if (W > 0 && T > 128 && T / 2 >= W) continue; /* more than half of the CTA would idle */
        const int nwarp = T / 32;
        const size_t extra = fixed + row_bytes * T;
        if (extra > SMEM_MAX) continue;
        int policy_mode = pick_hist(nwarp, extra, false);
        if (policy_mode == HIST_WARP_PRIVATE) policy_mode = HIST_SMEM_ATOMIC;
        /* mcmc keeps its shared histogram as u32 counters: a bin can receive at most T * steps * slots increments */
        if (p.solver == MCI_SOLVER_MCMC && policy_mode == HIST_SMEM_ATOMIC &&
            (double)T * ((double)(p.neval_per_block / (W > 0 ? W : T)) * (1.0 + P.thermal_ratio) + 1.0) * (double)P.n_draw >= 4.0e9)
          policy_mode = HIST_GLOBAL_RED;
        const int mis = h->maps_in_smem;
        const size_t smem = (mis ? map_bytes : 0) + hist_smem(policy_mode, nwarp) + extra;
        int occ = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, h->kernel, T, smem) != cudaSuccess) { cudaGetLastError(); continue; }
        const int score = occ * T * 4 + (mis ? 2 : 0) + (policy_mode == HIST_SMEM_ATOMIC ? 1 : 0);
        if (score > best_score) { best_score = score; best_occ = occ; h->threads = T; h->hist_mode = policy_mode; h->smem = smem; best_mis = mis; }
      }
    };
    if (p.walkers_per_block > 0) {
      h->walkers = p.walkers_per_block;
      choose_T(h->walkers);
    } else {
      /* walkers_per_block = 0: fill the GPU with whole CTAs — every block gets floor(resident CTAs / local blocks) of
       * them — unless that would cut the chains below 512 steps (start bias) */
      choose_T(0);
      if (best_score >= 0) {
        const int ctas = std::max(1, h->n_sm * best_occ / std::max(1, n_local));
        long long W = (long long)ctas * h->threads;
        const long long wmax = std::max<long long>(1, (long long)(p.neval_per_block / 512));
        if (W > wmax) { W = wmax >= 32 ? wmax / 32 * 32 : wmax; choose_T((int)W); }
        h->walkers = (int)W;
      }
    }
    if (best_score < 0) return fail(h, MCI_ERR_UNSUPPORTED, "walker state of %d rows does not fit in shared memory", h->chain_rows);
    h->maps_in_smem = best_mis;
  }
  CU(h, cudaFuncSetAttribute(h->kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem));
  if (h->hist_mode == HIST_SMEM_FIXED && h->is_pool) CU(h, cudaFuncSetAttribute(h->kernel_first, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem));
  int occ = 0;
  CU(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, h->kernel, h->threads, h->smem));
  if (occ < 1) return fail(h, MCI_ERR_CUDA, "kernel %s does not fit on an SM (smem %zu)", h->kernel_name, h->smem);

  if (p.solver == MCI_SOLVER_VEGAS) {
    const int resident = h->n_sm * occ;
    const uint64_t npb = p.neval_per_block;
    const uint64_t want_tiles_pb = std::max<uint64_t>(1, ((uint64_t)8 * resident + n_local - 1) / n_local);
    uint64_t spt = (npb + (uint64_t)h->threads * want_tiles_pb - 1) / ((uint64_t)h->threads * want_tiles_pb);
    spt = std::min<uint64_t>(std::max<uint64_t>(spt, 1), 64);
    h->spt = (int)spt;
    const uint64_t tile = (uint64_t)h->threads * spt;
    const uint64_t tpb = (npb + tile - 1) / tile;
    if (tpb * n_local > (uint64_t)1 << 30) return fail(h, MCI_ERR_INVALID, "neval_per_block too large for one launch");
    h->tiles_per_block = (int)tpb;
    h->grid = (int)std::min<uint64_t>((uint64_t)resident, tpb * n_local);
    h->ctas_per_block = 0;
  } else {
    const int W = std::max(1, h->walkers);
    h->steps = p.neval_per_block / W;
    /* propose / accept (and mcmc's visited) are counted in u32 per CTA and turned into the f64 statistics at the end */
    if ((double)std::min(W, h->threads) * ((double)h->steps * (1.0 + P.thermal_ratio) + (double)p.chain_burnin + 1.0) >= 4.0e9)
      return fail(h, MCI_ERR_UNSUPPORTED, "%llu steps per chain overflow the per-CTA counters: cut the block into more chains (walkers_per_block)",
                  (unsigned long long)h->steps);
    /* a walker's position in its Philox stream is a 32-bit count of uniforms: a step draws at most 4 (update kind,
     * variable, slot(s), acceptance) plus one per leaf it redraws — every slot of the largest integrand when an mcmc
     * changeIntegrand creates them all, three per FermiK proposal.  Refuse chains that would wrap it and replay */
    {
      int max_nl = 1;
      for (int v = 0; v < P.nv; ++v) max_nl = std::max(max_nl, P.top_nleaf[v]);
      const double per_step = 4.0 + (p.solver == MCI_SOLVER_MCMC ? (double)P.n_draw + 3.0 : (double)max_nl);
      const double total_steps = (double)h->steps * (1.0 + (p.solver == MCI_SOLVER_MCMC ? P.thermal_ratio : 0.0)) + (double)p.chain_burnin + 1.0;
      if (per_step * total_steps + 10000.0 * (double)P.n_draw >= 4294967296.0)
        return fail(h, MCI_ERR_UNSUPPORTED, "%llu steps per chain exhaust the 2^32 uniforms of a walker's stream: cut the block into more chains (walkers_per_block)",
                    (unsigned long long)h->steps);
    }
    h->ctas_per_block = (W + h->threads - 1) / h->threads;
    h->grid = h->ctas_per_block * n_local;
    h->tiles_per_block = 0;
  }
  h->n_hist_copies = (h->hist_mode == HIST_GLOBAL_RED) ? std::min(h->grid, 2 * h->n_sm) : h->grid;
  if (h->n_hist_copies < 1) h->n_hist_copies = 1;

  /* (re)allocate per-launch partial buffers */
  g_blocks.put(false, h->device, h->d_launch_arena, h->launch_cap);
  h->d_launch_arena = nullptr;
  h->d_tile_obs = h->d_tile_cnt = h->d_cta_hist = h->d_cta_stats = nullptr;
  {
    Arena L;
    size_t o_obs = 0, o_cnt = 0, o_stats = 0;
    const size_t nt = (size_t)h->tiles_per_block * n_local;
    const size_t stat_len = P.obs_total + 1 + P.nd + 2 * (size_t)h->pa_len + 1;
    if (p.solver == MCI_SOLVER_VEGAS) { o_obs = L.reserve<double>(nt * P.obs_total); o_cnt = L.reserve<double>(nt); }
    else o_stats = L.reserve<double>((size_t)h->grid * stat_len);
    const size_t o_hist = L.reserve<double>((size_t)h->n_hist_copies * std::max(1, P.hist_len));
    int rc;
    /* no memset: the fill kernels write every partial they own before the finalize kernel reads it, and the L2-RED
     * histogram copies are zeroed per launch */
    if ((rc = arena_device(h, L, &h->d_launch_arena, &h->launch_cap, false))) return rc;
    if (p.solver == MCI_SOLVER_VEGAS) { h->d_tile_obs = L.at<double>(o_obs); h->d_tile_cnt = L.at<double>(o_cnt); }
    else h->d_cta_stats = L.at<double>(o_stats);
    h->d_cta_hist = L.at<double>(o_hist);
  }
  return MCI_OK;
}

int enqueue_build_maps(mci_handle_t *h) {
  build_maps_kernel<<<h->nleaf, 128, 0, h->stream>>>(h->dp, h->d_grids, h->d_dist, h->d_accum, h->d_leaf_grid_off,
                                                     h->d_leaf_dist_off, h->d_leaf_acc_off, h->d_gmaps);
  CU(h, cudaGetLastError());
  h->launches++;
  h->maps_dirty = false;
  return MCI_OK;
}

}  // namespace

extern "C" {

int mci_abi_version(void) { return MCI_ABI_VERSION; }

const char *mci_last_error(const mci_handle_t *h) { return h ? h->err : g_err; }

void mci_destroy(mci_handle_t *h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  h->comm = nullptr; /* communicators are shared and outlive their handles (mci_comm_finalize) */
  if (h->jit_lib) jit_unload(h->jit_lib);
  g_blocks.put(false, h->device, h->d_state_arena, h->state_cap);
  g_blocks.put(false, h->device, h->d_launch_arena, h->launch_cap);
  g_blocks.put(false, h->device, h->d_history, h->history_cap);
  g_blocks.put(true, 0, h->h_arena, h->host_cap);
  g_blocks.put(true, 0, h->h_history, h->h_history_cap);
  if (h->ev_iter0) cudaEventDestroy(h->ev_iter0);
  for (int i = 0; i < mci_handle_t::FILL_RING; ++i) {
    if (h->ev_fill0[i]) cudaEventDestroy(h->ev_fill0[i]);
    if (h->ev_fill1[i]) cudaEventDestroy(h->ev_fill1[i]);
  }
  if (h->ev_iter1) cudaEventDestroy(h->ev_iter1);
  if (h->ev_t0) cudaEventDestroy(h->ev_t0);
  if (h->ev_t1) cudaEventDestroy(h->ev_t1);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

}  // extern "C"

/* validation of the POD problem + the device-side description (DevProblem, packed layout); no device is touched */
static int describe_problem(const mci_problem_t *p, mci_handle_t *h) {
  if (p->n_topvar < 1 || p->n_topvar > MCI_MAX_TOPVAR || p->n_leaf < 1 || p->n_leaf > MCI_MAX_LEAF ||
      p->n_integrand < 1 || p->n_integrand > MCI_MAX_INTEGRAND || p->n_block < 1 || p->measurefreq < 1 ||
      p->neval_per_block < 1)
    return fail(nullptr, MCI_ERR_INVALID, "problem dimensions out of range");
  if (p->solver < MCI_SOLVER_VEGAS || p->solver > MCI_SOLVER_MCMC) return fail(nullptr, MCI_ERR_INVALID, "Solver %d is not supported!", p->solver);
  FidInfo FI;
  if (!fid_info(p->integrand_id, FI)) return fail(nullptr, MCI_ERR_UNSUPPORTED, "unknown integrand id %d", p->integrand_id);
  if (p->integrand_id == MCI_F_USER && (!p->user_source || !*p->user_source))
    return fail(nullptr, MCI_ERR_INVALID, "integrand_id = MCI_F_USER needs user_source (CUDA C++ defining mci_user_integrand)");
  if (p->integrand_id != MCI_F_USER && FI.ni != p->n_integrand)
    return fail(nullptr, MCI_ERR_INVALID, "integrand %d returns %d values but n_integrand = %d", p->integrand_id, FI.ni, p->n_integrand);
  if (p->n_leaf < FI.nleaf_min) return fail(nullptr, MCI_ERR_INVALID, "integrand %d needs %d leaf variables", p->integrand_id, FI.nleaf_min);
  if (p->weight_kind != MCI_WEIGHT_REAL && p->weight_kind != MCI_WEIGHT_COMPLEX) return fail(nullptr, MCI_ERR_INVALID, "unknown weight_kind %d", p->weight_kind);
  {
    const bool cplx_id = p->integrand_id == MCI_F_COMPLEX1 || p->integrand_id == MCI_F_COMPLEX2;
    if (p->integrand_id != MCI_F_USER && cplx_id != (p->weight_kind == MCI_WEIGHT_COMPLEX))
      return fail(nullptr, MCI_ERR_INVALID, "integrand %d is %s: weight_kind must be %s", p->integrand_id, cplx_id ? "complex" : "real",
                  cplx_id ? "MCI_WEIGHT_COMPLEX" : "MCI_WEIGHT_REAL");
  }
  if (p->mcmc_start_idx < 0 || p->mcmc_start_idx > p->n_integrand + 1)
    return fail(nullptr, MCI_ERR_INVALID, "mcmc_start_idx = %d is not an integrand (1..%d)", p->mcmc_start_idx, p->n_integrand + 1);
  if (p->n_userdata < 0 || p->n_userdata > MCI_MAX_USERDATA || (p->n_userdata > 0 && !p->userdata))
    return fail(nullptr, MCI_ERR_INVALID, "userdata: %d doubles at %p (at most %d, and a non-NULL pointer)", p->n_userdata, (const void *)p->userdata, MCI_MAX_USERDATA);
  h->err[0] = 0;
  h->p = *p;
  h->device = p->device;
  DevProblem &P = h->dp;
  memset(&P, 0, sizeof P);
  const int ni = p->n_integrand, nd = ni + 1, nv = p->n_topvar, nleaf = p->n_leaf;
  h->ni = ni; h->nd = nd; h->nv = nv; h->nleaf = nleaf;
  P.solver = p->solver; P.fid = p->integrand_id; P.measure_kind = p->measure_kind;
  P.ni = ni; P.nd = nd; P.nv = nv; P.nleaf = nleaf; P.measurefreq = p->measurefreq;
  P.thermal_ratio = p->thermal_ratio; P.seed = p->seed;
  P.wdim = p->weight_kind == MCI_WEIGHT_COMPLEX ? 2 : 1;
  P.start_idx = p->mcmc_start_idx > 0 ? p->mcmc_start_idx - 1 : 0;
  int first = 0;
  for (int v = 0; v < nv; ++v) { P.top_nleaf[v] = p->topvar_nleaf[v]; P.top_first[v] = first; first += p->topvar_nleaf[v]; }
#define BAIL(code, ...) return fail(nullptr, code, __VA_ARGS__)
  if (first != nleaf) BAIL(MCI_ERR_INVALID, "sum(topvar_nleaf) != n_leaf");
  for (int i = 0; i < ni; ++i)
    for (int v = 0; v < nv; ++v) {
      P.dof[i * nv + v] = p->dof[i * nv + v];
      if (p->dof[i * nv + v] < 0) BAIL(MCI_ERR_INVALID, "negative dof");
    }
  for (int v = 0; v < nv; ++v) P.dof[ni * nv + v] = 0; /* configuration.jl:153 */
  for (int v = 0; v < nv; ++v) { int m = 0; for (int i = 0; i < nd; ++i) m = std::max(m, P.dof[i * nv + v]); P.maxdof[v] = m; }
  int base = 0, tslot = 0;
  for (int v = 0; v < nv; ++v) {
    P.top_base[v] = base; P.top_tslot[v] = tslot;
    base += P.maxdof[v] * P.top_nleaf[v]; tslot += P.maxdof[v];
  }
  P.n_draw = base; P.n_tslot = tslot;
  if (P.n_draw > MAX_DRAW) BAIL(MCI_ERR_UNSUPPORTED, "sum(maxdof * leaves) = %d exceeds %d", P.n_draw, MAX_DRAW);
  P.need_pad = 0;
  for (int i = 0; i < ni; ++i) {
    int same = 1;
    for (int v = 0; v < nv; ++v) if (P.dof[i * nv + v] != P.maxdof[v]) same = 0;
    P.same_dof[i] = same;
    if (!same) P.need_pad = 1;
  }
  int ot = 0;
  /* device-side observable layout in doubles: a complex entry is (re, im) */
  for (int i = 0; i < ni; ++i) {
    if (p->obs_len && p->obs_len[i] < 1) BAIL(MCI_ERR_INVALID, "obs_len[%d] = %d", i, p->obs_len[i]);
    P.obs_len[i] = (p->obs_len ? p->obs_len[i] : 1) * P.wdim; P.obs_off[i] = ot; ot += P.obs_len[i];
  }
  if (ot > MCI_MAX_OBS) BAIL(MCI_ERR_UNSUPPORTED, "total observable length %d doubles exceeds %d", ot, MCI_MAX_OBS);
  P.obs_total = ot; h->obs_total = ot;
  /* leaves */
  h->leaf_grid_off.assign(nleaf, 0); h->leaf_dist_off.assign(nleaf, 0); h->leaf_acc_off.assign(nleaf, 0);
  h->alpha.assign(nleaf, 2.0);
  int go = 0, dd = 0, da = 0, ho = 0, mo = 0;
  bool has_fermik = false, force_jit = false;
  for (int v = 0; v < nv; ++v)
    for (int k = 0; k < P.top_nleaf[v]; ++k) {
      const int l = P.top_first[v] + k;
      const mci_leaf_t &S = p->leaf[l];
      DevLeaf &L = P.leaf[l];
      if (S.kind == MCI_LEAF_FERMIK) {
        /* FermiK{D}: D component leaves, no map, no histogram (variable.jl:1-21; accumulate!/train! fall through to
         * the no-op methods, variable.jl:551,558) */
        const int D = P.top_nleaf[v];
        if (D != 2 && D != 3) BAIL(MCI_ERR_INVALID, "a FermiK variable has 2 or 3 component leaves, not %d", D);
        for (int kk = 0; kk < D; ++kk)
          if (p->leaf[P.top_first[v] + kk].kind != MCI_LEAF_FERMIK) BAIL(MCI_ERR_INVALID, "FermiK components cannot be bundled with other leaves");
        if (p->solver != MCI_SOLVER_MCMC)
          BAIL(MCI_ERR_UNSUPPORTED, "FermiK variables work under :mcmc only (the reference's own restriction, test/bubble_FermiK.jl:2,139-140)");
        if (p->integrand_id != MCI_F_USER && p->integrand_id != MCI_F_BUBBLE_FERMIK)
          BAIL(MCI_ERR_INVALID, "integrand %d does not take a FermiK variable", p->integrand_id);
        if (!(S.lower > 0.0) || !(S.upper > 0.0)) BAIL(MCI_ERR_INVALID, "FermiK needs kF > 0 and dk > 0");
        L.kind = S.kind; L.nbins = 0; L.adapt = 0; L.lower_i = 0; L.nbins_d = 0.0;
        L.top = v; L.base = P.top_base[v] + k; L.stride = D; L.hist_off = ho; L.smem_off = mo;
        L.p0 = S.lower; L.p1 = S.upper;
        h->alpha[l] = 0.0;
        has_fermik = true;
        continue;
      }
      if (S.nbins < 1) BAIL(MCI_ERR_INVALID, "leaf %d has no bins", l);
      if (p->solver != MCI_SOLVER_VEGAS && S.nbins > 65535)
        BAIL(MCI_ERR_UNSUPPORTED, "leaf %d has %d bins: the chain solvers keep bin indices in 16 bits (at most 65535)", l, S.nbins);
      L.kind = S.kind; L.nbins = S.nbins; L.adapt = S.adapt ? 1 : 0; L.lower_i = S.lower_i; L.nbins_d = (double)S.nbins;
      L.top = v; L.base = P.top_base[v] + k; L.stride = P.top_nleaf[v];
      L.hist_off = ho; ho += S.nbins;
      L.smem_off = mo;
      h->alpha[l] = S.alpha;
      if (S.kind == MCI_LEAF_CONTINUOUS) {
        if (!(S.upper > S.lower)) BAIL(MCI_ERR_INVALID, "leaf %d: upper <= lower", l);
        h->leaf_grid_off[l] = go; go += S.nbins + 1; mo += 2 * S.nbins;
      } else if (S.kind == MCI_LEAF_DISCRETE) {
        h->leaf_dist_off[l] = dd; dd += S.nbins; h->leaf_acc_off[l] = da; da += S.nbins + 1;
        mo += 2 * S.nbins + 1; mo += mo & 1;
      } else BAIL(MCI_ERR_INVALID, "leaf %d: unknown kind %d", l, S.kind);
      if (L.adapt) P.any_adapt = 1;
    }
  if (p->integrand_id == MCI_F_BUBBLE_FERMIK &&
      !(nv == 3 && P.top_nleaf[0] == 1 && P.top_nleaf[2] == 1 && P.leaf[0].kind == MCI_LEAF_CONTINUOUS &&
        P.leaf[1].kind == MCI_LEAF_FERMIK && P.leaf[nleaf - 1].kind == MCI_LEAF_DISCRETE))
    BAIL(MCI_ERR_INVALID, "integrand %d reads vars (T::Continuous, K::FermiK, Ext::Discrete)", p->integrand_id);
  h->has_fermik = has_fermik;
  P.map_doubles = mo; P.hist_len = ho;
  h->grid_len = go; h->disc_len = dd; h->accum_len = da; h->hist_len = ho;
  for (int v = 0; v < nv; ++v)
    P.top_skip[v] = (P.top_nleaf[v] == 1 && P.leaf[P.top_first[v]].kind == MCI_LEAF_DISCRETE && P.leaf[P.top_first[v]].nbins == 1);
  /* the slots the integrand reads must exist */
  {
    int need0 = P.fid == MCI_F_USER ? 0 : (FI.slots0 < 0 ? P.dof[0] : FI.slots0);
    if (P.maxdof[0] < need0 && FI.nleaf_min == 1) BAIL(MCI_ERR_INVALID, "integrand %d reads %d slots of variable 1 but maxdof = %d", P.fid, need0, P.maxdof[0]);
    if (FI.nleaf_min > 1)
      for (int l = 0; l < FI.nleaf_min; ++l) if (P.maxdof[P.leaf[l].top] < 1) BAIL(MCI_ERR_INVALID, "integrand %d needs one slot of every variable", P.fid);
    if (p->measure_kind == MCI_MEASURE_BIN_BY_DISCRETE) {
      const DevLeaf &E = P.leaf[nleaf - 1];
      if (E.kind != MCI_LEAF_DISCRETE || P.obs_len[0] < E.nbins * P.wdim || P.maxdof[E.top] < 1)
        BAIL(MCI_ERR_INVALID, "bin-by-discrete measure needs a Discrete last leaf and obs_len[0] >= its size");
      /* the measure bins integrand 1 only (example/bubble.jl:81-84); the chain kernels keep its bins in per-thread
       * accumulators, which for several integrands would silently drop bins: refuse instead */
      if (p->solver != MCI_SOLVER_VEGAS && ni != 1)
        BAIL(MCI_ERR_UNSUPPORTED, "bin-by-discrete measure under the chain solvers takes one integrand, not %d", ni);
      /* more bins than the ahead-of-time chain kernels keep accumulators for: the NVRTC build sizes them to the problem */
      if (p->solver != MCI_SOLVER_VEGAS && E.nbins > CHAIN_MAX_BINNED_OBS) force_jit = true;
    } else if (p->measure_kind == MCI_MEASURE_USER) {
      if (!p->user_source || !*p->user_source) BAIL(MCI_ERR_INVALID, "measure_kind = MCI_MEASURE_USER needs user_source defining mci_user_measure");
    } else if (p->measure_kind != MCI_MEASURE_DEFAULT) BAIL(MCI_ERR_INVALID, "unknown measure kind %d", p->measure_kind);
  }
  /* neighbor graph: configuration.jl:201-209 default, or the caller's CSR */
  if (p->neighbor_ptr) {
    if (p->neighbor_ptr[nd] > MAX_ND * MAX_ND) BAIL(MCI_ERR_INVALID, "neighbor graph too large");
    for (int d = 0; d <= nd; ++d) P.nb_ptr[d] = p->neighbor_ptr[d];
    for (int k = 0; k < P.nb_ptr[nd]; ++k) {
      P.nb_idx[k] = p->neighbor_idx[k];
      if (P.nb_idx[k] < 0 || P.nb_idx[k] >= nd) BAIL(MCI_ERR_INVALID, "neighbor index out of range");
    }
    for (int d = 0; d < nd; ++d) if (P.nb_ptr[d + 1] <= P.nb_ptr[d]) BAIL(MCI_ERR_INVALID, "integrand %d has no neighbor", d + 1);
  } else {
    int k = 0;
    for (int d = 1; d <= nd; ++d) {
      P.nb_ptr[d - 1] = k;
      if (d == nd) P.nb_idx[k++] = 0;
      else if (d == 1) { if (nd == 2) P.nb_idx[k++] = 1; else { P.nb_idx[k++] = nd - 1; P.nb_idx[k++] = 1; } }
      else if (d == nd - 1) P.nb_idx[k++] = nd - 3;
      else { P.nb_idx[k++] = d - 2; P.nb_idx[k++] = d; }
    }
    P.nb_ptr[nd] = k;
  }
  if (p->n_userdata < userdata_needed(P))
    BAIL(MCI_ERR_INVALID, "integrand %d reads %d doubles of userdata, got %d", P.fid, userdata_needed(P), p->n_userdata);
  for (int i = 0; i < p->n_userdata; ++i) P.userdata[i] = p->userdata[i];
  {
    const double pi = 3.14159265358979323846;
    if (P.fid == MCI_F_GAUSS_ND) P.fconst[0] = pow(2 * pi, -0.5 * P.dof[0]);
    if (P.fid == MCI_F_BUBBLE) P.fconst[0] = 1.0 / pow(2 * pi, P.userdata[4]);
    if (P.fid == MCI_F_BUBBLE_FERMIK) P.fconst[0] = 1.0 / pow(2 * pi, (double)P.top_nleaf[1]); /* phase = 1/(2pi)^dim, test/bubble_FermiK.jl:71 */
    if (P.fid == MCI_F_HYPERSPHERE3) for (int i = 1; i <= 3; ++i) P.fconst[i - 1] = volume_inverse(i + 1);
  }
  { /* halves of N^d for the pool kernel's Jacobian (PoolJac, mci_vegas.cuh) */
    const int d = P.maxdof[0], hh = (d + 1) >> 1;
    P.pool_npow[0] = P.pool_npow[1] = 1.0;
    for (int k = 0; k < hh; ++k) P.pool_npow[0] *= P.leaf[0].nbins_d;
    for (int k = 0; k < d - hh; ++k) P.pool_npow[1] *= P.leaf[0].nbins_d;
  }
  P.pa_dim = std::max(nd, nv);
  h->pa_len = 3 * nd * P.pa_dim;
  PackedLayout &Y = h->Y;
  Y.n_block = p->n_block; Y.obs_total = ot; Y.hist_len = ho; Y.nd = nd; Y.pa_len = h->pa_len;
  Y.off_hist = 0; Y.off_obs = ho; Y.off_cnt = Y.off_obs + p->n_block * ot; Y.off_visited = Y.off_cnt + p->n_block;
  Y.off_propose = Y.off_visited + nd; Y.off_accept = Y.off_propose + h->pa_len; Y.off_neval = Y.off_accept + h->pa_len;
  Y.len = Y.off_neval + 1;
  /* own the pointers of the problem copy */
  h->p.topvar_nleaf = nullptr; h->p.topvar_offset = nullptr; h->p.leaf = nullptr; h->p.dof = nullptr; h->p.obs_len = nullptr;
  h->p.neighbor_ptr = nullptr; h->p.neighbor_idx = nullptr; h->p.userdata = nullptr;
  if (p->user_source) h->user_source = p->user_source;
  h->p.user_source = nullptr;
  /* JIT policy: user source always; MCI_JIT_ON / MCI_JIT_OFF as told; auto = specialise when one iteration is big
   * enough (>= 2^24 samples) to repay ~1 s of NVRTC, except on the exact ahead-of-time pool kernels, which already are
   * what the JIT would build */
  h->jit_mode = p->jit;
  if (const char *e = getenv("MCIB200_JIT")) h->jit_mode = atoi(e) == 0 ? MCI_JIT_OFF : MCI_JIT_ON;
  if (force_jit && h->jit_mode == MCI_JIT_OFF)
    BAIL(MCI_ERR_UNSUPPORTED, "bin-by-discrete measure with more than %d bins under a chain solver needs the NVRTC build (jit != MCI_JIT_OFF)", CHAIN_MAX_BINNED_OBS);
  h->use_jit = force_jit || p->integrand_id == MCI_F_USER || p->measure_kind == MCI_MEASURE_USER || h->jit_mode == MCI_JIT_ON ||
               (h->jit_mode == MCI_JIT_AUTO && (double)p->n_block * (double)p->neval_per_block >= 16777216.0);
#undef BAIL

  return MCI_OK;
}

extern "C" {

int mci_create(const mci_problem_t *p, mci_handle_t **out) {
  if (!p || !out) return fail(nullptr, MCI_ERR_INVALID, "null argument");
  *out = nullptr;
  mci_handle_t *h = new mci_handle_t();
  {
    const int rc = describe_problem(p, h);
    if (rc) { delete h; return rc; }
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1 || p->device < 0 || p->device >= ndev)
    { delete h; return fail(nullptr, MCI_ERR_NO_DEVICE, "no CUDA device %d (found %d): libmcib200 has no CPU fallback", p->device, ndev); }
  int cc_major = 0, n_sm = 0; /* two attributes, not cudaGetDeviceProperties (milliseconds) */
  if (cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, p->device) != cudaSuccess ||
      cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, p->device) != cudaSuccess || cc_major < 10)
    { delete h; return fail(nullptr, MCI_ERR_NO_DEVICE, "device %d is not a Blackwell (sm_100a) GPU", p->device); }

  h->n_sm = n_sm;
  const int nd = h->nd, nleaf = h->nleaf, go = h->grid_len, dd = h->disc_len, da = h->accum_len, mo = h->dp.map_doubles;
  const PackedLayout &Y = h->Y;
  /* ---- device resources ---- */
  *out = h; /* so that failures below can be cleaned up by the caller through mci_destroy */
#define TRY(x) do { int rc_ = (x); if (rc_) { snprintf(g_err, 256, "%s", h->err); mci_destroy(h); *out = nullptr; return rc_; } } while (0)
#define CUT(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { fail(nullptr, MCI_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); mci_destroy(h); *out = nullptr; return MCI_ERR_CUDA; } } while (0)
  CUT(cudaSetDevice(h->device));
  CUT(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  CUT(cudaEventCreate(&h->ev_iter0)); CUT(cudaEventCreate(&h->ev_iter1));
  /* (the ring of fill-kernel events is created on first use, launch_fill) */
  {
    /* the whole fixed device state is laid out in one arena, built in a host image and uploaded with ONE copy */
    Arena D;
    const size_t o_grids = D.reserve<double>(go + 1), o_dist = D.reserve<double>(dd + 1), o_accum = D.reserve<double>(da + 1);
    const size_t o_rw = D.reserve<double>(nd), o_goal = D.reserve<double>(nd), o_alpha = D.reserve<double>(nleaf);
    const size_t o_gmaps = D.reserve<double>(mo + 2);
    const size_t o_lg = D.reserve<int>(nleaf), o_ld = D.reserve<int>(nleaf), o_la = D.reserve<int>(nleaf);
    const size_t o_packed = D.reserve<double>(Y.len), o_flags = D.reserve<unsigned long long>(4), o_err = D.reserve<int>(1);
    const size_t o_fx = D.reserve<int>(1), o_fxs = D.reserve<double>(1), o_stage = D.reserve<double>(go + dd + 2);
    TRY(arena_device(h, D, &h->d_state_arena, &h->state_cap, false));
    h->d_grids = D.at<double>(o_grids); h->d_dist = D.at<double>(o_dist); h->d_accum = D.at<double>(o_accum);
    h->d_reweight = D.at<double>(o_rw); h->d_goal = D.at<double>(o_goal); h->d_alpha = D.at<double>(o_alpha);
    h->d_gmaps = D.at<double>(o_gmaps);
    h->d_leaf_grid_off = D.at<int>(o_lg); h->d_leaf_dist_off = D.at<int>(o_ld); h->d_leaf_acc_off = D.at<int>(o_la);
    h->d_packed = D.at<double>(o_packed); h->d_flags = D.at<unsigned long long>(o_flags); h->d_err = D.at<int>(o_err);
    h->d_fixed_exp = D.at<int>(o_fx); h->d_fixed_state = D.at<double>(o_fxs); h->d_stage = D.at<double>(o_stage);
    Arena H;
    const size_t ho_packed = H.reserve<double>(Y.len), ho_flags = H.reserve<unsigned long long>(4), ho_err = H.reserve<int>(1);
    const size_t ho_init = H.reserve<char>(D.total); /* staging image of the device state */
    CUT(g_blocks.get(true, 0, H.total, &h->h_arena, &h->host_cap));
    H.base = static_cast<char *>(h->h_arena);
    h->h_packed = H.at<double>(ho_packed); h->h_flags = H.at<unsigned long long>(ho_flags); h->h_err = H.at<int>(ho_err);
    memset(h->h_packed, 0, Y.len * sizeof(double));
    h->h_flags[0] = h->h_flags[1] = h->h_flags[2] = h->h_flags[3] = 0; *h->h_err = 0;
    char *img = H.at<char>(ho_init);
    memset(img, 0, D.total);
    auto at = [&](size_t off) { return img + off; };
    memcpy(at(o_alpha), h->alpha.data(), nleaf * sizeof(double));
    memcpy(at(o_lg), h->leaf_grid_off.data(), nleaf * sizeof(int));
    memcpy(at(o_ld), h->leaf_dist_off.data(), nleaf * sizeof(int));
    memcpy(at(o_la), h->leaf_acc_off.data(), nleaf * sizeof(int));
    /* initial maps as the reference constructors leave them (variable.jl:137, 305-315; configuration.jl:172-175) */
    double *g = reinterpret_cast<double *>(at(o_grids)), *di = reinterpret_cast<double *>(at(o_dist));
    double *ac = reinterpret_cast<double *>(at(o_accum)), *rw = reinterpret_cast<double *>(at(o_rw));
    for (int l = 0; l < nleaf; ++l) {
      const mci_leaf_t &S = p->leaf[l];
      if (S.kind == MCI_LEAF_CONTINUOUS) {
        const int npts = S.nbins + 1;
        double *q = g + h->leaf_grid_off[l];
        for (int i = 0; i < npts; ++i) { double t = (double)i / (npts - 1); q[i] = (1 - t) * S.lower + t * S.upper; }
        q[0] = S.lower; q[npts - 1] = S.upper;
      } else if (S.kind == MCI_LEAF_DISCRETE) {
        const int n = S.nbins;
        double *d = di + h->leaf_dist_off[l], *a = ac + h->leaf_acc_off[l];
        for (int i = 0; i < n; ++i) d[i] = 1.0 / n;
        a[0] = 0.0;
        for (int i = 1; i <= n; ++i) { double s = 0.0; for (int k = 0; k < i; ++k) s += d[k]; a[i] = s; }
      }
    }
    for (int d = 0; d < nd; ++d) rw[d] = 1.0 / nd;
    *reinterpret_cast<int *>(at(o_fx)) = FIXED_EXP_INVALID;
    CUT(cudaMemcpyAsync(h->d_state_arena, img, D.total, cudaMemcpyHostToDevice, h->stream));
    /* the staging image lives in the pinned block until the handle dies, so the copy needs no synchronisation here */
  }
  TRY(plan(h));
#undef TRY
#undef CUT
  return MCI_OK;
}

int mci_sizes(const mci_handle_t *h, int32_t *grid_len, int32_t *disc_len, int32_t *hist_len, int32_t *obs_len,
              int32_t *nd, int32_t *pa_len) {
  if (!h) return MCI_ERR_INVALID;
  if (grid_len) *grid_len = h->grid_len;
  if (disc_len) *disc_len = h->disc_len;
  if (hist_len) *hist_len = h->hist_len;
  if (obs_len) *obs_len = h->obs_total;
  if (nd) *nd = h->nd;
  if (pa_len) *pa_len = h->pa_len;
  return MCI_OK;
}

int mci_set_maps(mci_handle_t *h, const double *grids, const double *dist, const double *accum, const double *rw) {
  if (!h) return MCI_ERR_INVALID;
  CU(h, cudaSetDevice(h->device));
  const bool fixed = h->hist_mode == HIST_SMEM_FIXED;
  if (grids && h->grid_len) {
    if (fixed) { /* through the staging area: the scale survives only if the maps are the ones already in place (mci_aux.cuh) */
      CU(h, cudaMemcpyAsync(h->d_stage, grids, h->grid_len * sizeof(double), cudaMemcpyHostToDevice, h->stream));
      adopt_maps_kernel<<<1, 256, 0, h->stream>>>(h->d_stage, h->d_grids, h->grid_len, h->d_fixed_exp, h->d_fixed_state);
      h->launches++;
    } else CU(h, cudaMemcpyAsync(h->d_grids, grids, h->grid_len * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  }
  if (dist && h->disc_len) {
    if (fixed) {
      CU(h, cudaMemcpyAsync(h->d_stage + h->grid_len, dist, h->disc_len * sizeof(double), cudaMemcpyHostToDevice, h->stream));
      adopt_maps_kernel<<<1, 256, 0, h->stream>>>(h->d_stage + h->grid_len, h->d_dist, h->disc_len, h->d_fixed_exp, h->d_fixed_state);
      h->launches++;
    } else CU(h, cudaMemcpyAsync(h->d_dist, dist, h->disc_len * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  }
  if (accum && h->accum_len) CU(h, cudaMemcpyAsync(h->d_accum, accum, h->accum_len * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  if (rw) CU(h, cudaMemcpyAsync(h->d_reweight, rw, h->nd * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CU(h, cudaGetLastError());
  CU(h, cudaStreamSynchronize(h->stream)); /* the caller may reuse its buffers */
  h->maps_dirty = true;
  return MCI_OK;
}

int mci_get_maps(mci_handle_t *h, double *grids, double *dist, double *accum, double *rw) {
  if (!h) return MCI_ERR_INVALID;
  CU(h, cudaSetDevice(h->device));
  if (grids && h->grid_len) CU(h, cudaMemcpyAsync(grids, h->d_grids, h->grid_len * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  if (dist && h->disc_len) CU(h, cudaMemcpyAsync(dist, h->d_dist, h->disc_len * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  if (accum && h->accum_len) CU(h, cudaMemcpyAsync(accum, h->d_accum, h->accum_len * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  if (rw) CU(h, cudaMemcpyAsync(rw, h->d_reweight, h->nd * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  return MCI_OK;
}

int mci_set_partition(mci_handle_t *h, int32_t rank, int32_t world) {
  if (!h) return MCI_ERR_INVALID;
  if (world < 1 || rank < 0 || rank >= world || h->p.n_block % world != 0) /* main.jl:122 */
    return fail(h, MCI_ERR_INVALID, "invalid partition rank=%d world=%d n_block=%d", rank, world, h->p.n_block);
  CU(h, cudaSetDevice(h->device));
  CU(h, cudaStreamSynchronize(h->stream));
  h->rank = rank; h->world = world;
  return plan(h);
}

/* enqueue one iteration on the handle's stream: [staged maps] -> fill -> finalize -> [all-reduce]; with_d2h also
 * queues the copy of the statistics into the pinned mirrors that mci_fetch_iteration reads */
static int enqueue_iteration(mci_handle_t *h, uint64_t iter, bool with_d2h) {
  const mci_problem_t &p = h->p;
  const int n_local = p.n_block / h->world;
  CU(h, cudaEventRecord(h->ev_iter0, h->stream));
  if (h->maps_dirty) { int rc = enqueue_build_maps(h); if (rc) return rc; }
  CU(h, cudaMemsetAsync(h->d_flags, 0, 4 * sizeof(unsigned long long), h->stream));
  if (h->hist_mode == HIST_GLOBAL_RED)
    CU(h, cudaMemsetAsync(h->d_cta_hist, 0, (size_t)h->n_hist_copies * h->hist_len * sizeof(double), h->stream));
  void *args[2];
  FillArgs FA;
  ChainArgs CA;
  if (p.solver == MCI_SOLVER_VEGAS) {
    FA.gmaps = h->d_gmaps; FA.tile_obs = h->d_tile_obs; FA.tile_cnt = h->d_tile_cnt; FA.cta_hist = h->d_cta_hist;
    FA.flags = h->d_flags; FA.fixed_exp = h->hist_mode == HIST_SMEM_FIXED ? h->d_fixed_exp : nullptr; FA.iter = iter;
    FA.fixed_bias = 0; FA.rescue = 0;
    if (const char *e = getenv("MCIB200_FIXED_EXP_BIAS")) FA.fixed_bias = atoi(e); /* test hook: shrink the headroom */
    FA.neval_per_block = p.neval_per_block; FA.n_local_blocks = n_local;
    philox_expand_key(h->dp.seed, iter, FA.key);
    FA.rank = h->rank; FA.world = h->world; FA.tiles_per_block = h->tiles_per_block; FA.spt = h->spt;
    FA.hist_mode = h->hist_mode; FA.n_hist_copies = h->n_hist_copies; FA.maps_in_smem = h->maps_in_smem;
    args[0] = &h->dp; args[1] = &FA;
  } else {
    CA.gmaps = h->d_gmaps; CA.reweight = h->d_reweight; CA.cta_stats = h->d_cta_stats; CA.cta_hist = h->d_cta_hist;
    philox_expand_key(h->dp.seed, iter, CA.key);
    CA.flags = h->d_flags; CA.iter = iter; CA.steps = h->steps; CA.walkers_per_block = std::max(1, h->walkers);
    CA.ctas_per_block = h->ctas_per_block; CA.n_local_blocks = n_local; CA.rank = h->rank; CA.world = h->world;
    CA.hist_mode = h->hist_mode; CA.n_hist_copies = h->n_hist_copies; CA.maps_in_smem = h->maps_in_smem;
    CA.rows = h->chain_rows; CA.coop_ci = h->coop_ci; CA.burnin = (long long)p.chain_burnin;
    args[0] = &h->dp; args[1] = &CA;
  }
  const int slot = (int)(h->n_iterations % mci_handle_t::FILL_RING);
  if (!h->ev_fill0[slot]) { CU(h, cudaEventCreate(&h->ev_fill0[slot])); CU(h, cudaEventCreate(&h->ev_fill1[slot])); }
  CU(h, cudaEventRecord(h->ev_fill0[slot], h->stream));
  CU(h, cudaLaunchKernel(h->kernel, dim3(h->grid), dim3(h->threads), args, h->smem, h->stream));
  if (h->hist_mode == HIST_SMEM_FIXED) {
    /* ... and behind it the fp64 launch of the same iteration: it works when the device-side scale is not settled, or when
     * the fixed-point launch reported that it left its range (rescue), and returns at once otherwise (mci_vegas.cuh).  The
     * pool layout has a twin kernel for it; the generic kernel is launched a second time in rescue mode */
    FA.rescue = 1;
    CU(h, cudaLaunchKernel(h->is_pool ? h->kernel_first : h->kernel, dim3(h->grid), dim3(h->threads), args, h->smem, h->stream));
    h->launches++;
  }
  CU(h, cudaEventRecord(h->ev_fill1[slot], h->stream));
  h->n_iterations++;
  h->launches++;
  const int fin_threads = 256;
  const int total_el = h->Y.len;
  const int fin_grid = std::max(1, std::min(4 * h->n_sm, (total_el * 32 + fin_threads - 1) / fin_threads));
  if (p.solver == MCI_SOLVER_VEGAS) {
    finalize_vegas_kernel<<<fin_grid, fin_threads, 0, h->stream>>>(h->Y, h->d_tile_obs, h->d_tile_cnt, h->d_cta_hist,
                                                                  h->hist_mode == HIST_NONE ? 0 : h->n_hist_copies, n_local,
                                                                  h->tiles_per_block, h->rank, h->world, p.neval_per_block,
                                                                  h->d_packed);
  } else {
    finalize_chain_kernel<<<fin_grid, fin_threads, 0, h->stream>>>(h->Y, h->d_cta_stats, h->d_cta_hist,
                                                                  h->hist_mode == HIST_NONE ? 0 : h->n_hist_copies, n_local,
                                                                  h->ctas_per_block, h->rank, h->world, h->d_packed);
  }
  CU(h, cudaGetLastError());
  h->launches++;
  if (h->comm) {
    int rc = g_nccl.AllReduce(h->d_packed, h->d_packed, (size_t)h->Y.len, NCCL_DOUBLE, NCCL_SUM, h->comm, h->stream);
    if (rc != 0) return fail(h, MCI_ERR_COMM, "ncclAllReduce: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "error");
  }
  if (h->hist_mode == HIST_SMEM_FIXED) { /* the scale of the next iteration, from the (all-reduced) statistics of this one */
    fixed_scale_kernel<<<1, 256, 0, h->stream>>>(h->Y, h->d_packed, h->d_fixed_exp, h->d_fixed_state);
    CU(h, cudaGetLastError());
    h->launches++;
  }
  if (with_d2h) {
    CU(h, cudaMemcpyAsync(h->h_packed, h->d_packed, h->Y.len * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaMemcpyAsync(h->h_flags, h->d_flags, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
  }
  CU(h, cudaEventRecord(h->ev_iter1, h->stream));
  h->have_iteration = true;
  return MCI_OK;
}

/* raw packed statistics + flags of one iteration -> the caller's arrays, with the floors of clearStatistics! /
 * addConfig! (configuration.jl:238-262, variable.jl:565: one per block config + the summed config) and the error
 * checks of the reference's loop (mcmc start point, block normalisation) */
static int unpack_iteration(mci_handle_t *h, const double *pk, const unsigned long long *flags, mci_iteration_out_t *out) {
  const PackedLayout &Y = h->Y;
  const int nb = Y.n_block;
  const double B1 = (double)(nb + 1);
  /* floors of clearStatistics! (configuration.jl:238-250, variable.jl:565): one per block config + the summed config */
  if (out->block_obs) memcpy(out->block_obs, pk + Y.off_obs, sizeof(double) * nb * Y.obs_total);
  if (out->block_norm) for (int b = 0; b < nb; ++b) out->block_norm[b] = 1.0e-10 + pk[Y.off_cnt + b];
  if (out->hist) for (int i = 0; i < Y.hist_len; ++i) out->hist[i] = B1 * 1.0e-10 + pk[Y.off_hist + i];
  if (out->visited) for (int i = 0; i < Y.nd; ++i) out->visited[i] = B1 * 1.0e-8 + pk[Y.off_visited + i];
  if (out->propose) for (int i = 0; i < Y.pa_len; ++i) out->propose[i] = B1 * 1.0e-8 + pk[Y.off_propose + i];
  if (out->accept) for (int i = 0; i < Y.pa_len; ++i) out->accept[i] = B1 * 1.0e-10 + pk[Y.off_accept + i];
  if (out->neval) *out->neval = (uint64_t)pk[Y.off_neval];
  h->nonfinite = flags[0];
  if (flags[3]) h->hist_rescues++; /* the fixed-point histogram left its range; the fp64 launch recomputed the iteration */
  if (flags[2] && !flags[3])
    return fail(h, MCI_ERR_NUMERIC, "fixed-point histogram (MCI_HIST_SMEM_FIXED): a bin left the 64-bit range and no fp64 rescue ran");
  if (flags[1]) /* mcmc/montecarlo.jl:109-113 */
    return fail(h, MCI_ERR_MCMC_START, "Cannot find the variables that makes the #%d integrand nonzero!", h->dp.start_idx + 1);
  for (int b = 0; b < nb; ++b)
    if (!(1.0e-10 + pk[Y.off_cnt + b] > 0.0)) /* main.jl:269-271 */
      return fail(h, MCI_ERR_NUMERIC, "Block normalization = %g is not positively defined!", pk[Y.off_cnt + b]);
  return MCI_OK;
}

/* a failed assert of train! on the device (variable.jl:212-213, common.jl:71,79) surfaces at the next host sync */
static int check_train_error(mci_handle_t *h) {
  if (!*h->h_err) return MCI_OK;
  *h->h_err = 0;
  cudaMemsetAsync(h->d_err, 0, sizeof(int), h->stream);
  return fail(h, MCI_ERR_NUMERIC, "train!: histogram should be all finite and positive");
}

int mci_launch_iteration(mci_handle_t *h, uint64_t iter) {
  if (!h) return MCI_ERR_INVALID;
  CU(h, cudaSetDevice(h->device));
  return enqueue_iteration(h, iter, true);
}

int mci_fetch_iteration(mci_handle_t *h, mci_iteration_out_t *out) {
  if (!h || !out) return MCI_ERR_INVALID;
  if (!h->have_iteration) return fail(h, MCI_ERR_INVALID, "no iteration has been launched");
  CU(h, cudaSetDevice(h->device));
  CU(h, cudaMemcpyAsync(h->h_err, h->d_err, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  {
    const int slot = (int)((h->n_iterations - 1) % mci_handle_t::FILL_RING);
    cudaEventElapsedTime(&h->fill_ms, h->ev_fill0[slot], h->ev_fill1[slot]);
  }
  cudaEventElapsedTime(&h->iter_ms, h->ev_iter0, h->ev_iter1);
  int rc = unpack_iteration(h, h->h_packed, h->h_flags, out);
  if (rc) return rc;
  return check_train_error(h);
}

int mci_run_iteration(mci_handle_t *h, uint64_t iter, mci_iteration_out_t *out) {
  int rc = mci_launch_iteration(h, iter);
  if (rc) return rc;
  mci_iteration_out_t none;
  memset(&none, 0, sizeof none);
  return mci_fetch_iteration(h, out ? out : &none);
}

int mci_refine(mci_handle_t *h) {
  if (!h) return MCI_ERR_INVALID;
  if (!h->have_iteration) return fail(h, MCI_ERR_INVALID, "mci_refine before any iteration");
  CU(h, cudaSetDevice(h->device));
  if (!h->dp.any_adapt) return MCI_OK;
  int nmax = 1;
  for (int l = 0; l < h->nleaf; ++l) nmax = std::max(nmax, h->dp.leaf[l].nbins);
  const size_t sm = (size_t)(3 * nmax + 2) * sizeof(double);
  if (sm > 48 * 1024) CU(h, cudaFuncSetAttribute((const void *)refine_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  refine_kernel<<<h->nleaf, 256, sm, h->stream>>>(h->dp, h->Y, h->d_packed, h->d_alpha, h->d_leaf_grid_off, h->d_leaf_dist_off,
                                                  h->d_leaf_acc_off, h->d_grids, h->d_dist, h->d_accum, h->d_err);
  CU(h, cudaGetLastError());
  h->launches++;
  int rc = enqueue_build_maps(h);
  if (rc) return rc;
  if (!h->async_mode) {
    CU(h, cudaMemcpyAsync(h->h_err, h->d_err, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    if (*h->h_err) {
      *h->h_err = 0;
      cudaMemsetAsync(h->d_err, 0, sizeof(int), h->stream);
      return fail(h, MCI_ERR_NUMERIC, "train!: histogram should be all finite and positive");
    }
  }
  return MCI_OK;
}

int mci_reweight(mci_handle_t *h, double gamma, const double *goal) {
  if (!h) return MCI_ERR_INVALID;
  if (!h->have_iteration) return fail(h, MCI_ERR_INVALID, "mci_reweight before any iteration");
  CU(h, cudaSetDevice(h->device));
  if (goal) CU(h, cudaMemcpyAsync(h->d_goal, goal, h->nd * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  reweight_kernel<<<1, 32, 0, h->stream>>>(h->Y, h->d_packed, gamma, goal ? h->d_goal : nullptr, h->d_reweight);
  CU(h, cudaGetLastError());
  h->launches++;
  if (goal) CU(h, cudaStreamSynchronize(h->stream));
  return MCI_OK;
}

int mci_run_iterations(mci_handle_t *h, uint64_t first_iter, int32_t niter, double gamma, const double *goal,
                       int32_t do_reweight, int32_t do_refine, mci_iteration_out_t *out) {
  if (!h || niter < 1) return MCI_ERR_INVALID;
  CU(h, cudaSetDevice(h->device));
  const PackedLayout &Y = h->Y;
  const size_t rec = (size_t)Y.len + 4; /* packed statistics + the four flag words, in 8-byte slots */
  const size_t bytes = rec * (size_t)niter * sizeof(double);
  if (h->history_cap < bytes) {
    g_blocks.put(false, h->device, h->d_history, h->history_cap);
    h->d_history = nullptr; h->history_cap = 0;
    CU(h, g_blocks.get(false, h->device, bytes, &h->d_history, &h->history_cap));
  }
  if (h->h_history_cap < bytes) {
    g_blocks.put(true, 0, h->h_history, h->h_history_cap);
    h->h_history = nullptr; h->h_history_cap = 0;
    CU(h, g_blocks.get(true, 0, bytes, &h->h_history, &h->h_history_cap));
  }
  double *dh = static_cast<double *>(h->d_history), *hh = static_cast<double *>(h->h_history);
  const bool chain = h->p.solver != MCI_SOLVER_VEGAS;
  if (goal && do_reweight && chain) CU(h, cudaMemcpyAsync(h->d_goal, goal, h->nd * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  int nmax = 1;
  for (int l = 0; l < h->nleaf; ++l) nmax = std::max(nmax, h->dp.leaf[l].nbins);
  const size_t sm = (size_t)(3 * nmax + 2) * sizeof(double);
  if (do_refine && h->dp.any_adapt && sm > 48 * 1024)
    CU(h, cudaFuncSetAttribute((const void *)refine_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  for (int k = 0; k < niter; ++k) {
    int rc = enqueue_iteration(h, first_iter + (uint64_t)k, false);
    if (rc) return rc;
    snapshot_kernel<<<std::max(1, std::min(8, (Y.len + 255) / 256)), 256, 0, h->stream>>>(h->d_packed, Y.len, h->d_flags, dh + rec * k);
    h->launches++;
    if (do_reweight && chain) { /* main.jl:183 */
      reweight_kernel<<<1, 32, 0, h->stream>>>(Y, h->d_packed, gamma, goal ? h->d_goal : nullptr, h->d_reweight);
      h->launches++;
    }
    if (do_refine && h->dp.any_adapt) { /* main.jl:193-198 */
      refine_kernel<<<h->nleaf, 256, sm, h->stream>>>(h->dp, Y, h->d_packed, h->d_alpha, h->d_leaf_grid_off, h->d_leaf_dist_off,
                                                      h->d_leaf_acc_off, h->d_grids, h->d_dist, h->d_accum, h->d_err);
      h->launches++;
      rc = enqueue_build_maps(h);
      if (rc) return rc;
    }
    CU(h, cudaGetLastError());
  }
  CU(h, cudaMemcpyAsync(hh, dh, bytes, cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaMemcpyAsync(h->h_err, h->d_err, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  /* leave the handle as after a fetch of the last iteration */
  memcpy(h->h_packed, hh + rec * (niter - 1), Y.len * sizeof(double));
  memcpy(h->h_flags, hh + rec * (niter - 1) + Y.len, 4 * sizeof(unsigned long long));
  {
    const int slot = (int)((h->n_iterations - 1) % mci_handle_t::FILL_RING);
    cudaEventElapsedTime(&h->fill_ms, h->ev_fill0[slot], h->ev_fill1[slot]);
  }
  uint64_t bad = 0;
  mci_iteration_out_t none;
  memset(&none, 0, sizeof none);
  for (int k = 0; k < niter; ++k) {
    const double *pk = hh + rec * k;
    const int rc = unpack_iteration(h, pk, reinterpret_cast<const unsigned long long *>(pk + Y.len), out ? &out[k] : &none);
    if (rc) return rc;
    bad += h->nonfinite;
  }
  h->nonfinite = bad; /* over the whole call */
  return check_train_error(h);
}

int mci_set_async(mci_handle_t *h, int32_t on) {
  if (!h) return MCI_ERR_INVALID;
  h->async_mode = on != 0;
  return MCI_OK;
}

int mci_comm_unique_id(void *id128) {
  if (!id128) return MCI_ERR_INVALID;
  if (!nccl_load()) return fail(nullptr, MCI_ERR_COMM, "libnccl.so.2 not loadable: %s", dlerror());
  nccl_uid_t id;
  int rc = g_nccl.GetUniqueId(&id);
  if (rc != 0) return fail(nullptr, MCI_ERR_COMM, "ncclGetUniqueId failed (%d)", rc);
  memcpy(id128, &id, 128);
  return MCI_OK;
}

/* Communicators live for the process, keyed by their unique id: creating one costs about a second (topology discovery,
 * then the connections at its first collective), which would otherwise be paid by every handle — i.e. by every
 * integrate() call of a multi-GPU job (measured on 4 B200: 1.1 s + 0.5 s against 15 ms of iterations for BASELINE C4).
 * A handle that names an id this process has already used (same rank, world and device) joins that communicator. */
namespace {
struct SharedComm { nccl_comm_t comm; int rank, world, device; };
std::mutex g_comm_mu;
std::map<std::string, SharedComm> g_comms;
}  // namespace

int mci_comm_init(mci_handle_t *h, const void *id128, int32_t rank, int32_t world) {
  if (!h || !id128) return MCI_ERR_INVALID;
  if (!nccl_load()) return fail(h, MCI_ERR_COMM, "libnccl.so.2 not loadable");
  int rc = mci_set_partition(h, rank, world);
  if (rc) return rc;
  nccl_uid_t id;
  memcpy(&id, id128, 128);
  CU(h, cudaSetDevice(h->device));
  std::lock_guard<std::mutex> lock(g_comm_mu);
  const std::string key((const char *)id128, 128);
  auto it = g_comms.find(key);
  if (it != g_comms.end()) {
    if (it->second.rank != rank || it->second.world != world || it->second.device != h->device)
      return fail(h, MCI_ERR_COMM, "this unique id already names a communicator of rank %d/%d on device %d", it->second.rank,
                  it->second.world, it->second.device);
    h->comm = it->second.comm;
    return MCI_OK;
  }
  nccl_comm_t c = nullptr;
  rc = g_nccl.CommInitRank(&c, world, id, rank);
  if (rc != 0) { h->comm = nullptr; return fail(h, MCI_ERR_COMM, "ncclCommInitRank: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "error"); }
  g_comms[key] = SharedComm{c, rank, world, h->device};
  h->comm = c;
  return MCI_OK;
}

/* destroy every communicator of this process (collective: every rank calls it, after its last handle is destroyed) */
int mci_comm_finalize(void) {
  std::lock_guard<std::mutex> lock(g_comm_mu);
  if (g_nccl.ok)
    for (auto &kv : g_comms) {
      cudaSetDevice(kv.second.device);
      g_nccl.CommDestroy(kv.second.comm);
    }
  g_comms.clear();
  return MCI_OK;
}

int mci_stats_device_ptr(mci_handle_t *h, void **dptr, int64_t *n_doubles) {
  if (!h) return MCI_ERR_INVALID;
  if (dptr) *dptr = h->d_packed;
  if (n_doubles) *n_doubles = h->Y.len;
  return MCI_OK;
}

/* re-read the packed statistics from the device (after an external collective on mci_stats_device_ptr) */
int mci_stats_sync_host(mci_handle_t *h) {
  if (!h) return MCI_ERR_INVALID;
  CU(h, cudaSetDevice(h->device));
  CU(h, cudaMemcpyAsync(h->h_packed, h->d_packed, h->Y.len * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  return MCI_OK;
}

int mci_trace_samples(mci_handle_t *h, uint64_t iter, int32_t block, uint64_t first, int32_t count, double *x,
                      int32_t *bins, double *jac, double *weights) {
  if (!h || count < 0) return MCI_ERR_INVALID;
  if (h->p.solver != MCI_SOLVER_VEGAS) return fail(h, MCI_ERR_UNSUPPORTED, "mci_trace_samples: VEGAS sampler only");
  if (count == 0) return MCI_OK;
  CU(h, cudaSetDevice(h->device));
  /* a handle whose fill kernel is the pool kernel is probed through the pool kernel's own draw code */
  const void *k = h->trace_kernel; /* set when the fill kernel is an NVRTC build */
  if (!k && h->is_pool) k = lookup_vegas_pool_trace(h->dp.fid, h->dp.maxdof[0]);
  if (!k) k = lookup_vegas_trace(h->dp.fid);
  if (!k) return fail(h, MCI_ERR_UNSUPPORTED, "no trace kernel for integrand %d", h->dp.fid);
  if (h->maps_dirty) { int rc = enqueue_build_maps(h); if (rc) return rc; }
  const int nd = h->dp.n_draw, ni = h->ni * h->dp.wdim;
  /* one allocation for the four outputs, released on every path */
  const size_t n_x = (size_t)count * nd, n_w = (size_t)count * ni; /* ni values per sample: (re, im) pairs when complex */
  const size_t bytes = sizeof(double) * (n_x + (size_t)count + n_w) + sizeof(int) * n_x;
  char *buf = nullptr;
  CU(h, cudaMalloc((void **)&buf, bytes));
  double *dx = reinterpret_cast<double *>(buf), *dj = dx + n_x, *dw = dj + count;
  int *db = reinterpret_cast<int *>(dw + n_w);
  const double *gm = h->d_gmaps;
  uint32_t blk = (uint32_t)block;
  void *args[] = {&h->dp, &gm, &iter, &blk, &first, &count, &dx, &db, &dj, &dw};
  cudaError_t e = cudaLaunchKernel(k, dim3((count + 127) / 128), dim3(128), args, 0, h->stream);
  h->launches++;
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  if (e == cudaSuccess && x) e = cudaMemcpy(x, dx, sizeof(double) * (size_t)count * nd, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && bins) e = cudaMemcpy(bins, db, sizeof(int) * (size_t)count * nd, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && jac) e = cudaMemcpy(jac, dj, sizeof(double) * (size_t)count, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && weights) e = cudaMemcpy(weights, dw, sizeof(double) * (size_t)count * ni, cudaMemcpyDeviceToHost);
  cudaFree(buf);
  if (e != cudaSuccess) return fail(h, MCI_ERR_CUDA, "mci_trace_samples: %s", cudaGetErrorString(e));
  return MCI_OK;
}

int mci_jit_check(const mci_problem_t *p, char *log, int32_t log_cap) {
  if (!p) return MCI_ERR_INVALID;
  if (log && log_cap > 0) log[0] = 0;
  mci_handle_t *h = new mci_handle_t();
  int rc = describe_problem(p, h);
  if (rc) {
    if (log && log_cap > 0) snprintf(log, log_cap, "%s", g_err);
    delete h;
    return rc;
  }
  const DevProblem &P = h->dp;
  /* the kernels mci_create would pick with the automatic policies */
  char name[160];
  std::vector<std::string> exprs;
  if (p->solver == MCI_SOLVER_VEGAS) {
    const bool pool = P.nv == 1 && P.nleaf == 1 && P.leaf[0].kind == MCI_LEAF_CONTINUOUS && p->measure_kind == MCI_MEASURE_DEFAULT &&
                      P.obs_total == P.ni && P.maxdof[0] <= MCI_MAX_SLOT && p->measurefreq == 1 && P.wdim == 1;
    if (pool) {
      const bool striped = P.any_adapt && P.maxdof[0] <= 20;
      snprintf(name, sizeof name, "mci::vegas_pool_kernel<%d, %d, %d, %d, %s>", P.fid, P.maxdof[0],
               !P.any_adapt ? HIST_NONE : (striped ? HIST_SMEM_STRIPED : HIST_SMEM_ATOMIC), striped ? POOL_THREADS_STRIPED : POOL_THREADS,
               P.need_pad ? "false" : "true");
    } else {
      snprintf(name, sizeof name, "mci::vegas_generic_kernel<%d, %d>", P.fid, GENERIC_THREADS);
    }
    exprs.push_back(name);
    if (pool) snprintf(name, sizeof name, "mci::vegas_pool_trace_kernel<%d, %d>", P.fid, P.maxdof[0]);
    else snprintf(name, sizeof name, "mci::vegas_trace_kernel<%d>", P.fid);
    exprs.push_back(name);
  } else {
    snprintf(name, sizeof name, p->solver == MCI_SOLVER_VEGASMC ? "mci::vegasmc_kernel<%d>" : "mci::mcmc_kernel<%d>", P.fid);
    exprs.push_back(name);
  }
  std::vector<const void *> ks;
  std::string err;
  rc = jit_compile(P, p->user_source, exprs, ks, err, std::string(), /*compile_only=*/true);
  delete h;
  if (rc != 0) {
    if (log && log_cap > 0) snprintf(log, log_cap, "%s", err.c_str());
    return MCI_ERR_UNSUPPORTED;
  }
  return MCI_OK;
}

int mci_nonfinite_count(const mci_handle_t *h, uint64_t *n) {
  if (!h || !n) return MCI_ERR_INVALID;
  *n = h->nonfinite;
  return MCI_OK;
}

int mci_hist_rescues(const mci_handle_t *h, uint64_t *n) {
  if (!h || !n) return MCI_ERR_INVALID;
  *n = h->hist_rescues;
  return MCI_OK;
}

int mci_launch_count(const mci_handle_t *h, uint64_t *n) {
  if (!h || !n) return MCI_ERR_INVALID;
  *n = h->launches;
  return MCI_OK;
}

int mci_last_timing(mci_handle_t *h, float *fill_ms, float *iteration_ms) {
  if (!h) return MCI_ERR_INVALID;
  if (fill_ms) *fill_ms = h->fill_ms;
  if (iteration_ms) *iteration_ms = h->iter_ms;
  return MCI_OK;
}

int mci_fill_history(mci_handle_t *h, float *ms, int32_t cap, int32_t *n) {
  if (!h || !ms || !n || cap < 0) return MCI_ERR_INVALID;
  CU(h, cudaSetDevice(h->device));
  CU(h, cudaStreamSynchronize(h->stream));
  const uint64_t have = std::min<uint64_t>(h->n_iterations, (uint64_t)mci_handle_t::FILL_RING);
  const int m = (int)std::min<uint64_t>(have, (uint64_t)cap);
  for (int k = 0; k < m; ++k) { /* ms[0] = oldest of the m most recent iterations */
    const int slot = (int)((h->n_iterations - m + k) % mci_handle_t::FILL_RING);
    CU(h, cudaEventElapsedTime(&ms[k], h->ev_fill0[slot], h->ev_fill1[slot]));
  }
  *n = m;
  return MCI_OK;
}

int mci_chain_info(const mci_handle_t *h, int32_t *walkers_per_block, uint64_t *steps_per_walker, int32_t *ctas_per_block) {
  if (!h) return MCI_ERR_INVALID;
  if (walkers_per_block) *walkers_per_block = h->p.solver == MCI_SOLVER_VEGAS ? 0 : h->walkers;
  if (steps_per_walker) *steps_per_walker = h->p.solver == MCI_SOLVER_VEGAS ? 0 : h->steps;
  if (ctas_per_block) *ctas_per_block = h->ctas_per_block;
  return MCI_OK;
}

int mci_launch_info(const mci_handle_t *h, int32_t *grid, int32_t *threads, int64_t *smem_bytes, int32_t *hist_mode,
                    int32_t *tiles_per_block, int32_t *samples_per_thread) {
  if (!h) return MCI_ERR_INVALID;
  if (grid) *grid = h->grid;
  if (threads) *threads = h->threads;
  if (smem_bytes) *smem_bytes = (int64_t)h->smem;
  if (hist_mode) *hist_mode = h->hist_mode;
  if (tiles_per_block) *tiles_per_block = h->tiles_per_block;
  if (samples_per_thread) *samples_per_thread = h->spt;
  return MCI_OK;
}

const char *mci_kernel_name(const mci_handle_t *h) { return h ? h->kernel_name : ""; }

/* CUDA-event stopwatch on the handle's stream (the stream every kernel of the handle is launched on) */
int mci_timer_start(mci_handle_t *h) {
  if (!h) return MCI_ERR_INVALID;
  CU(h, cudaSetDevice(h->device));
  if (!h->ev_t0) { CU(h, cudaEventCreate(&h->ev_t0)); CU(h, cudaEventCreate(&h->ev_t1)); }
  CU(h, cudaEventRecord(h->ev_t0, h->stream));
  return MCI_OK;
}
int mci_timer_stop(mci_handle_t *h, float *ms) {
  if (!h || !ms || !h->ev_t0) return MCI_ERR_INVALID;
  CU(h, cudaSetDevice(h->device));
  CU(h, cudaEventRecord(h->ev_t1, h->stream));
  CU(h, cudaEventSynchronize(h->ev_t1));
  CU(h, cudaEventElapsedTime(ms, h->ev_t0, h->ev_t1));
  return MCI_OK;
}

int mci_fp64_peak(int32_t device, double *tflops) {
  if (!tflops) return MCI_ERR_INVALID;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return fail(nullptr, MCI_ERR_NO_DEVICE, "no CUDA device %d", device);
  CU(nullptr, cudaSetDevice(device));
  cudaDeviceProp prop;
  CU(nullptr, cudaGetDeviceProperties(&prop, device));
  double *d = nullptr;
  CU(nullptr, cudaMalloc((void **)&d, 8));
  cudaEvent_t e0, e1;
  CU(nullptr, cudaEventCreate(&e0)); CU(nullptr, cudaEventCreate(&e1));
  const int iters = 1 << 14, threads = 256, grid = prop.multiProcessorCount * 8;
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    fp64_peak_kernel<<<grid, threads>>>(d, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    cudaError_t e = cudaEventSynchronize(e1);
    if (e != cudaSuccess) { cudaFree(d); return fail(nullptr, MCI_ERR_CUDA, "fp64 probe: %s", cudaGetErrorString(e)); }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    double fl = 2.0 * 8.0 * iters * (double)threads * grid;
    if (rep > 0) best = std::max(best, fl / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
  *tflops = best;
  return MCI_OK;
}

}  // extern "C"
