/* helper macros to instantiate vegas_pool_kernel<FID, MAXD, HMODE, THREADS, EXACT> groups */
#pragma once
#include "mci_kernels.h"
#include "mci_vegas.cuh"

#define MCI_POOL_SWITCH(FID, MAXD, EXACT)                                                                                          \
  switch (hist_mode) {                                                                                                           \
    case mci::HIST_NONE: return (const void *)mci::vegas_pool_kernel<FID, MAXD, mci::HIST_NONE, mci::POOL_THREADS, EXACT>;        \
    case mci::HIST_SMEM_ATOMIC: return (const void *)mci::vegas_pool_kernel<FID, MAXD, mci::HIST_SMEM_ATOMIC, mci::POOL_THREADS, EXACT>; \
    case mci::HIST_GLOBAL_RED: return (const void *)mci::vegas_pool_kernel<FID, MAXD, mci::HIST_GLOBAL_RED, mci::POOL_THREADS, EXACT>;   \
    case mci::HIST_SMEM_STRIPED:                                                                                                 \
      if (MAXD <= 20) return (const void *)mci::vegas_pool_kernel<FID, (MAXD <= 20 ? MAXD : 1), mci::HIST_SMEM_STRIPED, mci::POOL_THREADS_STRIPED, EXACT>; \
      return nullptr; /* 1024 threads x 64 registers cannot hold more than 20 slots without spilling */                           \
    case mci::HIST_SMEM_FIXED:                                                                                                   \
      if (MAXD <= 20) return (const void *)mci::vegas_pool_kernel<FID, (MAXD <= 20 ? MAXD : 1), mci::HIST_SMEM_FIXED, mci::POOL_THREADS_STRIPED, EXACT>; \
      return nullptr;                                                                                                            \
  }

/* run-time dof <= MAXD, predicated */
#define MCI_POOL_CASE(FID, MAXD)                                                                              \
  if (fid == FID && d <= MAXD) {                                                                              \
    if (name) *name = "vegas_pool_kernel<" #FID "," #MAXD ">";                                                \
    MCI_POOL_SWITCH(FID, MAXD, false)                                                                         \
  }
/* every integrand uses exactly D slots: fully unrolled, no predication */
#define MCI_POOL_CASE_EXACT(FID, D)                                                                           \
  if (fid == FID && d == D && exact) {                                                                        \
    if (name) *name = "vegas_pool_kernel<" #FID "," #D ",exact>";                                             \
    MCI_POOL_SWITCH(FID, D, true)                                                                             \
  }
