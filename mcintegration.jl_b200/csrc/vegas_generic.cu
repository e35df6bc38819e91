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

const void *lookup_vegas_generic(int fid, const char **name) {
  switch (fid) {
#define M(FID) case FID: if (name) *name = "vegas_generic_kernel<" #FID ">"; return (const void *)vegas_generic_kernel<FID, GENERIC_THREADS>;
    MCI_FOR_EACH_FID(M)
#undef M
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
