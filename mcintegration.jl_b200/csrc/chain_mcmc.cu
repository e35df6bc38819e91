/* mcmc kernels for every built-in integrand */
#include "mci_kernels.h"
#include "mci_chain.cuh"
namespace mci {
const void *lookup_mcmc(int fid, const char **name) {
  switch (fid) {
#define M(FID) case FID: if (name) *name = "mcmc_kernel<" #FID ">"; return (const void *)mcmc_kernel<FID>;
    MCI_FOR_EACH_FID(M)
    MCI_FOR_EACH_MCMC_ONLY_FID(M)
#undef M
  }
  return nullptr;
}
}  // namespace mci
