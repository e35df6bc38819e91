/* vegas pool kernels, group C: the reference's test-suite integrands on one Continuous pool */
#include "vegas_pool_inst.cuh"
namespace mci {
const void *lookup_vegas_pool_c(int fid, int d, bool exact, int hist_mode, const char **name) {
  MCI_POOL_CASE(MCI_F_LOG_OVER_SQRT, 4)
  MCI_POOL_CASE(MCI_F_SPHERE1, 4)
  MCI_POOL_CASE(MCI_F_SPHERE2, 4)
  MCI_POOL_CASE(MCI_F_LATTICE_POOL, 4)
  MCI_POOL_CASE(MCI_F_HYPERSPHERE3, 4)
  MCI_POOL_CASE(MCI_F_POLY3, 4)
  MCI_POOL_CASE(MCI_F_X2_PLUS_Y2, 4)
  return nullptr;
}
}  // namespace mci
