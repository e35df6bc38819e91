/* mci_kernels.h — host-visible lookup of the kernel instantiations (each .cu translation unit
 * instantiates a group of templates; the ABI layer launches them through cudaLaunchKernel). */
#pragma once
#include "mci_device.cuh"

namespace mci {

constexpr int POOL_THREADS = 256;
constexpr int POOL_THREADS_STRIPED = 1024; /* HIST_SMEM_STRIPED: one CTA per SM owns the 16-way striped histogram */
constexpr int GENERIC_THREADS = 256;
constexpr int GENERIC_THREADS_BIG = 1024; /* HIST_SMEM_FIXED on layouts whose histograms fill most of an SM's shared memory: one
                                             CTA per SM, so the CTA itself must bring enough warps */
constexpr int GENERIC_THREADS_MID = 512;  /* ... two CTAs per SM */

/* vegas, single Continuous pool: returns NULL when (fid, d) has no pool instantiation for hist_mode
 * (HIST_NONE / HIST_SMEM_ATOMIC / HIST_SMEM_STRIPED / HIST_GLOBAL_RED).  exact: every integrand uses exactly d slots. */
const void *lookup_vegas_pool(int fid, int d, bool exact, int hist_mode, const char **name);
/* vegas, any layout */
const void *lookup_vegas_generic(int fid, const char **name, int threads = GENERIC_THREADS);
const void *lookup_vegas_generic_big(int fid, const char **name, int threads); /* the 512- and 1024-thread instantiations (own translation unit) */
const void *lookup_vegas_trace(int fid);
/* the same probe through the pool kernel's own draw code (NULL: this integrand has no pool instantiation) */
const void *lookup_vegas_pool_trace(int fid, int d);
/* Markov-chain solvers */
const void *lookup_vegasmc(int fid, const char **name);
const void *lookup_mcmc(int fid, const char **name);

}  // namespace mci
