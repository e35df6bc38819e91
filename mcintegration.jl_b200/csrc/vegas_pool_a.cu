/* vegas pool kernels, group A: the BASELINE C2 Gaussian */
#include "vegas_pool_inst.cuh"
namespace mci {
const void *lookup_vegas_pool_a(int fid, int d, bool exact, int hist_mode, const char **name) {
  MCI_POOL_CASE_EXACT(MCI_F_GAUSS_ND, 10)
  MCI_POOL_CASE(MCI_F_GAUSS_ND, 4)
  MCI_POOL_CASE(MCI_F_GAUSS_ND, 12)
  MCI_POOL_CASE(MCI_F_GAUSS_ND, 24)
  return nullptr;
}
}  // namespace mci
