/* vegas generic-layout kernels + trace kernels for every built-in integrand */
#include "mci_kernels.h"
#include "mci_vegas.cuh"
namespace mci {
const void *lookup_vegas_pool_a(int, int, bool, int, const char **);
const void *lookup_vegas_pool_b(int, int, bool, int, const char **);
const void *lookup_vegas_pool_c(int, int, bool, int, const char **);

const void *lookup_vegas_pool(int fid, int d, bool exact, int hist_mode, const char **name) {
  const void *f = lookup_vegas_pool_a(fid, d, exact, hist_mode, name);
  if (!f) f = lookup_vegas_pool_b(fid, d, exact, hist_mode, name);
  if (!f) f = lookup_vegas_pool_c(fid, d, exact, hist_mode, name);
  return f;
}

const void *lookup_vegas_generic(int fid, const char **name, int threads) {
  if (threads != GENERIC_THREADS) return lookup_vegas_generic_big(fid, name, threads);
  switch (fid) {
#define M(FID) case FID: if (name) *name = "vegas_generic_kernel<" #FID ">"; return (const void *)vegas_generic_kernel<FID, GENERIC_THREADS>;
    MCI_FOR_EACH_FID(M)
#undef M
  }
  return nullptr;
}

/* L0 probe of the pool path (pool_draw), for every integrand that has pool instantiations */
const void *lookup_vegas_pool_trace(int fid, int d) {
  if (d > MCI_MAX_SLOT) return nullptr;
  switch (fid) {
#define MP(FID) case FID: return (const void *)vegas_pool_trace_kernel<FID, MCI_MAX_SLOT>;
    MP(MCI_F_GAUSS_ND) MP(MCI_F_GENZ_CORNER_POOL) MP(MCI_F_PRODUCT_4) MP(MCI_F_LOG_OVER_SQRT) MP(MCI_F_SPHERE1) MP(MCI_F_SPHERE2)
    MP(MCI_F_LATTICE_POOL) MP(MCI_F_HYPERSPHERE3) MP(MCI_F_POLY3) MP(MCI_F_X2_PLUS_Y2)
#undef MP
  }
  return nullptr;
}

const void *lookup_vegas_trace(int fid) {
  switch (fid) {
#define M(FID) case FID: return (const void *)vegas_trace_kernel<FID>;
    MCI_FOR_EACH_FID(M)
#undef M
  }
  return nullptr;
}
}  // namespace mci
