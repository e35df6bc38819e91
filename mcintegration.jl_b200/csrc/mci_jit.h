/* mci_jit.h — NVRTC specialisation of the fill kernels for one problem (mci_jit.cu) */
#pragma once
#include <string>
#include <vector>

#include "mci_device.cuh"

namespace mci {

/* Compile `name_exprs` (e.g. "mci::vegas_pool_kernel<100, 10, 4, 1024, true>") against the layout of P and, for
 * MCI_F_USER, the caller's integrand source.  On success kernels = one cudaKernel_t (as const void *, launchable with
 * cudaLaunchKernel) per name expression, plus the cudaLibrary_t as the last entry — or, with compile_only, nothing:
 * the program is compiled (and cached) but not loaded, so no device is needed.  Returns 0 or -1 with `err` set. */
int jit_compile(const DevProblem &P, const char *user_source, const std::vector<std::string> &name_exprs,
                std::vector<const void *> &kernels, std::string &err, const std::string &defines = std::string(),
                bool compile_only = false);
void jit_unload(const void *lib);
std::string jit_problem_header(const DevProblem &P);

}  // namespace mci
