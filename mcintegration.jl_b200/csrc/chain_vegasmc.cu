/* vegasmc kernels for every built-in integrand */
#include "mci_kernels.h"
#include "mci_chain.cuh"
namespace mci {
const void *lookup_vegasmc(int fid, const char **name) {
  switch (fid) {
#define M(FID) case FID: if (name) *name = "vegasmc_kernel<" #FID ">"; return (const void *)vegasmc_kernel<FID>;
    MCI_FOR_EACH_FID(M)
#undef M
  }
  return nullptr;
}
}  // namespace mci
