/* helper macros to instantiate vegas_pool_kernel<FID, MAXD, HMODE> groups */
#pragma once
#include "mci_kernels.h"
#include "mci_vegas.cuh"

#define MCI_POOL_CASE(FID, MAXD)                                                                              \
  if (fid == FID && d <= MAXD) {                                                                              \
    if (name) *name = "vegas_pool_kernel<" #FID "," #MAXD ">";                                                \
    switch (hist_mode) {                                                                                      \
      case mci::HIST_NONE: return (const void *)mci::vegas_pool_kernel<FID, MAXD, mci::HIST_NONE, mci::POOL_THREADS>;             \
      case mci::HIST_SMEM_ATOMIC: return (const void *)mci::vegas_pool_kernel<FID, MAXD, mci::HIST_SMEM_ATOMIC, mci::POOL_THREADS>; \
      case mci::HIST_GLOBAL_RED: return (const void *)mci::vegas_pool_kernel<FID, MAXD, mci::HIST_GLOBAL_RED, mci::POOL_THREADS>;   \
      case mci::HIST_WARP_PRIVATE: return (const void *)mci::vegas_pool_kernel<FID, MAXD, mci::HIST_WARP_PRIVATE, mci::POOL_THREADS>; \
    }                                                                                                         \
  }
