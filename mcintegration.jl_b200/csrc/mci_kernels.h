/* mci_kernels.h — host-visible lookup of the kernel instantiations (each .cu translation unit
 * instantiates a group of templates; the ABI layer launches them through cudaLaunchKernel). */
#pragma once
#include "mci_device.cuh"

namespace mci {

constexpr int POOL_THREADS = 256;
constexpr int GENERIC_THREADS = 256;
constexpr int CHAIN_THREADS = 128;

/* vegas, single Continuous pool: returns NULL when (fid, d) has no pool instantiation */
const void *lookup_vegas_pool(int fid, int d, int hist_mode, const char **name);
/* vegas, any layout */
const void *lookup_vegas_generic(int fid, const char **name);
const void *lookup_vegas_trace(int fid);
/* Markov-chain solvers */
const void *lookup_vegasmc(int fid, const char **name);
const void *lookup_mcmc(int fid, const char **name);

}  // namespace mci
