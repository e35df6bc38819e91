/* vegas generic-layout kernels with 1024-thread CTAs (one CTA per SM): layouts whose fixed-point histogram cells take
 * most of the shared memory of an SM (BASELINE C3-B: 16 leaves x 999 bins x 8 B = 128 KB) */
#include "mci_kernels.h"
#include "mci_vegas.cuh"
namespace mci {
const void *lookup_vegas_generic_big(int fid, const char **name) {
  switch (fid) {
#define M(FID) case FID: if (name) *name = "vegas_generic_kernel<" #FID ",1024>"; return (const void *)vegas_generic_kernel<FID, GENERIC_THREADS_BIG>;
    MCI_FOR_EACH_FID(M)
#undef M
  }
  return nullptr;
}
}  // namespace mci
