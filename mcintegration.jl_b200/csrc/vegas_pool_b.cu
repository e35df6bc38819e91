/* vegas pool kernels, group B: Genz corner peak (C3 layout A) and the 4-component product (C5 integrand) */
#include "vegas_pool_inst.cuh"
namespace mci {
const void *lookup_vegas_pool_b(int fid, int d, bool exact, int hist_mode, const char **name) {
  MCI_POOL_CASE_EXACT(MCI_F_GENZ_CORNER_POOL, 16)
  MCI_POOL_CASE(MCI_F_GENZ_CORNER_POOL, 12)
  MCI_POOL_CASE(MCI_F_GENZ_CORNER_POOL, 24)
  MCI_POOL_CASE_EXACT(MCI_F_PRODUCT_4, 20)
  MCI_POOL_CASE(MCI_F_PRODUCT_4, 12)
  MCI_POOL_CASE(MCI_F_PRODUCT_4, 24)
  return nullptr;
}
}  // namespace mci
