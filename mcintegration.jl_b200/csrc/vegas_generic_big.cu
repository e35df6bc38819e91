/* vegas generic-layout kernels with 512- and 1024-thread CTAs (two / one CTA per SM): layouts whose fixed-point
 * histogram cells (and maps) take a large part of the shared memory of an SM — BASELINE C3-B: 16 leaves x 999 bins x 8 B
 * = 128 KB of cells, one CTA per SM; example/bubble.jl under :vegas: 64 KB of maps + 32 KB of cells, two */
#include "mci_kernels.h"
#include "mci_vegas.cuh"
namespace mci {
const void *lookup_vegas_generic_big(int fid, const char **name, int threads) {
  if (threads == GENERIC_THREADS_BIG) {
    switch (fid) {
#define M(FID) case FID: if (name) *name = "vegas_generic_kernel<" #FID ",1024>"; return (const void *)vegas_generic_kernel<FID, GENERIC_THREADS_BIG>;
      MCI_FOR_EACH_FID(M)
#undef M
    }
  } else if (threads == GENERIC_THREADS_MID) {
    switch (fid) {
#define M(FID) case FID: if (name) *name = "vegas_generic_kernel<" #FID ",512>"; return (const void *)vegas_generic_kernel<FID, GENERIC_THREADS_MID>;
      MCI_FOR_EACH_FID(M)
#undef M
    }
  }
  return nullptr;
}
}  // namespace mci
