/*
 * mci_jit.cu — problem-specialised kernels through NVRTC.
 *
 * The reference accepts ANY Julia closure f(x, c) (src/main.jl:25-33).  A closure cannot run on the device, so the
 * drop-in equivalent is a CUDA C++ source string: mci_problem_t.user_source defines
 *
 *     template <class XA> __device__ void mci_user_integrand(const XA &X, const double *userdata, double *w);
 *
 * (X(leaf, pos) = value of pool slot `pos` (1-based) of leaf variable `leaf` (0-based, top-level order then leaf
 * order); w[0..n_integrand-1] out).  It is compiled at mci_create together with the SAME device headers the
 * ahead-of-time kernels are built from (embedded in the library), with the whole problem layout — variables, leaves,
 * dof table, neighbour graph, observable layout — baked in as compile-time tables (JitProblem), so every loop over the
 * layout unrolls and the per-sample state lives in registers.  Built-in integrands can ask for the same
 * specialisation (mci_problem_t.jit = 1).
 *
 * NVRTC is bound at run time (dlopen), like NCCL: the library loads and the ahead-of-time kernels work without it.
 * The compiled CUBIN is loaded with cudaLibraryLoadData; the cudaKernel_t it yields is launched through the ordinary
 * cudaLaunchKernel path of mci_abi.cu.
 */
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "mci_jit.h"
#include "build_jit_sources.inc"

namespace mci {
namespace {

typedef struct _nvrtcProgram *nvrtcProgram;
struct NvrtcApi {
  void *lib = nullptr;
  int (*CreateProgram)(nvrtcProgram *, const char *, const char *, int, const char *const *, const char *const *) = nullptr;
  int (*DestroyProgram)(nvrtcProgram *) = nullptr;
  int (*CompileProgram)(nvrtcProgram, int, const char *const *) = nullptr;
  int (*GetCUBINSize)(nvrtcProgram, size_t *) = nullptr;
  int (*GetCUBIN)(nvrtcProgram, char *) = nullptr;
  int (*GetProgramLogSize)(nvrtcProgram, size_t *) = nullptr;
  int (*GetProgramLog)(nvrtcProgram, char *) = nullptr;
  int (*AddNameExpression)(nvrtcProgram, const char *) = nullptr;
  int (*GetLoweredName)(nvrtcProgram, const char *, const char **) = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
  bool ok = false;
};
NvrtcApi g_rtc;
std::mutex g_mu;

bool nvrtc_load(std::string &why) {
  if (g_rtc.ok) return true;
  std::vector<std::string> names;
  if (const char *e = getenv("MCIB200_NVRTC")) names.push_back(e);
  names.insert(names.end(), {"libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so"});
  for (const auto &n : names) {
    g_rtc.lib = dlopen(n.c_str(), RTLD_NOW | RTLD_GLOBAL);
    if (g_rtc.lib) break;
  }
  if (!g_rtc.lib) { why = "libnvrtc.so.12 not loadable (set MCIB200_NVRTC)"; return false; }
#define SYM(field, name) *(void **)(&g_rtc.field) = dlsym(g_rtc.lib, name)
  SYM(CreateProgram, "nvrtcCreateProgram"); SYM(DestroyProgram, "nvrtcDestroyProgram"); SYM(CompileProgram, "nvrtcCompileProgram");
  SYM(GetCUBINSize, "nvrtcGetCUBINSize"); SYM(GetCUBIN, "nvrtcGetCUBIN"); SYM(GetProgramLogSize, "nvrtcGetProgramLogSize");
  SYM(GetProgramLog, "nvrtcGetProgramLog"); SYM(AddNameExpression, "nvrtcAddNameExpression");
  SYM(GetLoweredName, "nvrtcGetLoweredName"); SYM(GetErrorString, "nvrtcGetErrorString");
#undef SYM
  g_rtc.ok = g_rtc.CreateProgram && g_rtc.DestroyProgram && g_rtc.CompileProgram && g_rtc.GetCUBINSize && g_rtc.GetCUBIN &&
             g_rtc.GetProgramLogSize && g_rtc.GetProgramLog && g_rtc.AddNameExpression && g_rtc.GetLoweredName;
  if (!g_rtc.ok) why = "libnvrtc lacks a required entry point";
  return g_rtc.ok;
}

/* one int table: module-scope `__device__ const` storage (a load with a compile-time index folds to an immediate, a
 * run-time index is an ordinary cached load — a constexpr array local to the accessor is copied to the stack on every
 * call with a run-time index: measured 10x slower on the 16-leaf composite) + an accessor member of JitProblem */
void itab(std::string &tabs, std::string &members, const char *name, const int *v, int n) {
  char buf[96];
  snprintf(buf, sizeof buf, "__device__ const int %s[%d] = {", name, n);
  tabs += buf;
  for (int i = 0; i < n; ++i) { snprintf(buf, sizeof buf, i ? ", %d" : "%d", v[i]); tabs += buf; }
  tabs += "};\n";
  snprintf(buf, sizeof buf, "  MCI_JIT_ITAB(%s)\n", name);
  members += buf;
}

struct Compiled {
  std::vector<char> cubin;
  std::vector<std::string> lowered;
};
/* program text + name expressions -> cubin: the same problem created again (every rank of a job, every test) compiles once */
std::map<std::string, Compiled> g_cache;

}  // namespace

namespace {
/* Opt-in disk cache (MCIB200_JIT_CACHE=<directory>): the first use of a problem in a process costs an NVRTC build
 * (0.6 s for the BASELINE C5 kernel, measured); with the cache a later process loads the cubin instead.  A file is named by
 * the FNV-1a hash of everything the cubin depends on — the embedded device headers of this library build, the generated
 * problem header, the program text, the kernel names — and holds the lowered names and the cubin. */
uint64_t fnv1a(uint64_t h, const char *p, size_t n) {
  for (size_t i = 0; i < n; ++i) { h ^= (unsigned char)p[i]; h *= 1099511628211ull; }
  return h;
}
std::string disk_cache_path(const std::string &key) {
  const char *dir = getenv("MCIB200_JIT_CACHE");
  if (!dir || !*dir) return "";
  uint64_t h = 1469598103934665603ull;
  for (int i = 0; i < MCI_JIT_NUM_HEADERS; ++i) h = fnv1a(h, MCI_JIT_HEADER_SOURCES[i], strlen(MCI_JIT_HEADER_SOURCES[i]));
  h = fnv1a(h, key.data(), key.size());
  char name[64];
  snprintf(name, sizeof name, "/mcib200_sm100a_%016llx.cubin", (unsigned long long)h);
  return std::string(dir) + name;
}
bool disk_cache_load(const std::string &path, size_t n_names, Compiled &C) {
  FILE *f = fopen(path.c_str(), "rb");
  if (!f) return false;
  bool ok = false;
  uint64_t magic = 0, nn = 0, nc = 0;
  if (fread(&magic, 8, 1, f) == 1 && magic == 0x4d43494a49543031ull && fread(&nn, 8, 1, f) == 1 && nn == n_names) {
    ok = true;
    for (uint64_t i = 0; ok && i < nn; ++i) {
      uint64_t len = 0;
      ok = fread(&len, 8, 1, f) == 1 && len < (1u << 16);
      std::string low(ok ? len : 0, '\0');
      ok = ok && (len == 0 || fread(&low[0], 1, len, f) == len);
      C.lowered.push_back(low);
    }
    ok = ok && fread(&nc, 8, 1, f) == 1 && nc > 0 && nc < (1ull << 30);
    if (ok) { C.cubin.resize(nc); ok = fread(C.cubin.data(), 1, nc, f) == nc; }
  }
  fclose(f);
  if (!ok) { C.lowered.clear(); C.cubin.clear(); }
  return ok;
}
void disk_cache_store(const std::string &path, const Compiled &C) {
  const std::string tmp = path + ".tmp" + std::to_string((long long)getpid());
  FILE *f = fopen(tmp.c_str(), "wb");
  if (!f) return; /* a cache that cannot be written is not an error */
  const uint64_t magic = 0x4d43494a49543031ull, nn = C.lowered.size(), nc = C.cubin.size();
  bool ok = fwrite(&magic, 8, 1, f) == 1 && fwrite(&nn, 8, 1, f) == 1;
  for (const auto &low : C.lowered) {
    const uint64_t len = low.size();
    ok = ok && fwrite(&len, 8, 1, f) == 1 && (len == 0 || fwrite(low.data(), 1, len, f) == len);
  }
  ok = ok && fwrite(&nc, 8, 1, f) == 1 && fwrite(C.cubin.data(), 1, nc, f) == nc;
  ok = (fclose(f) == 0) && ok;
  if (!ok || rename(tmp.c_str(), path.c_str()) != 0) remove(tmp.c_str()); /* rename: readers never see a partial file */
}
}  // namespace

/* "mci_jit_problem.h": every layout field of DevProblem as a compile-time constant / table */
std::string jit_problem_header(const DevProblem &P) {
  std::string tabs, members, o;
  char buf[640];
  const int nvv = P.nv > 0 ? P.nv : 1;
  bool has_fermik = false;
  for (int l = 0; l < P.nleaf; ++l) has_fermik = has_fermik || P.leaf[l].kind == MCI_LEAF_FERMIK;
  int max_nleaf = 1;
  for (int v = 0; v < P.nv; ++v) max_nleaf = P.top_nleaf[v] > max_nleaf ? P.top_nleaf[v] : max_nleaf;
  itab(tabs, members, "top_nleaf", P.top_nleaf, nvv); itab(tabs, members, "top_first", P.top_first, nvv);
  itab(tabs, members, "top_base", P.top_base, nvv); itab(tabs, members, "maxdof", P.maxdof, nvv);
  itab(tabs, members, "top_tslot", P.top_tslot, nvv); itab(tabs, members, "top_skip", P.top_skip, nvv);
  itab(tabs, members, "dof", P.dof, P.nd * nvv);
  itab(tabs, members, "same_dof", P.same_dof, P.ni); itab(tabs, members, "obs_off", P.obs_off, P.ni);
  itab(tabs, members, "obs_len", P.obs_len, P.ni);
  itab(tabs, members, "nb_ptr", P.nb_ptr, P.nd + 1);
  itab(tabs, members, "nb_idx", P.nb_idx, P.nb_ptr[P.nd] > 0 ? P.nb_ptr[P.nd] : 1);
  snprintf(buf, sizeof buf, "__device__ const DevLeaf leaf[%d] = {\n", P.nleaf);
  tabs += buf;
  for (int l = 0; l < P.nleaf; ++l) {
    const DevLeaf &L = P.leaf[l];
    snprintf(buf, sizeof buf, "  {%d, %d, %d, %d, %d, %d, %d, %d, %d, %d, %d, 0, %d.0, %.17g, %.17g},\n", L.kind, L.nbins, L.adapt,
             L.lower_i, L.map_off, L.acc_off, L.hist_off, L.top, L.base, L.stride, L.smem_off, L.nbins, L.p0, L.p1);
    tabs += buf;
  }
  tabs += "};\n";
  o += "/* generated by mci_jit.cu: the problem layout as compile-time tables */\n#pragma once\nnamespace mci {\n";
  o += "namespace jit_tab {\n" + tabs + "}  // namespace jit_tab\n";
  o += "#define MCI_JIT_ITAB(name) struct T_##name { __device__ __forceinline__ int operator[](int i) const { return jit_tab::name[i]; } } name;\n";
  o += "struct JitProblem {\n";
  snprintf(buf, sizeof buf,
           "  static constexpr int solver = %d, fid = %d, measure_kind = %d, ni = %d, nd = %d, nv = %d, nleaf = %d, n_draw = %d;\n"
           "  static constexpr int obs_total = %d, hist_len = %d, pa_dim = %d, measurefreq = %d, need_pad = %d, any_adapt = %d;\n"
           "  static constexpr int map_doubles = %d, n_tslot = %d, wdim = %d, start_idx = %d, max_nleaf = %d;\n  static constexpr bool has_fermik = %s;\n",
           P.solver, P.fid, P.measure_kind, P.ni, P.nd, P.nv, P.nleaf, P.n_draw, P.obs_total, P.hist_len, P.pa_dim, P.measurefreq,
           P.need_pad, P.any_adapt, P.map_doubles, P.n_tslot, P.wdim, P.start_idx, max_nleaf, has_fermik ? "true" : "false");
  o += buf;
  o += members;
  o += "  struct T_leaf { __device__ __forceinline__ const DevLeaf &operator[](int i) const { return jit_tab::leaf[i]; } } leaf;\n";
  o += "  const double *userdata, *fconst, *pool_npow; double thermal_ratio; uint64_t seed;\n"
       "  __device__ __forceinline__ JitProblem(const DevProblem &p) : userdata(p.userdata), fconst(p.fconst), pool_npow(p.pool_npow), "
       "thermal_ratio(p.thermal_ratio), seed(p.seed) {}\n"
       "  __device__ static constexpr int pa_len_() { return 3 * nd * pa_dim; }\n};\n}  // namespace mci\n";
  return o;
}

int jit_compile(const DevProblem &P, const char *user_source, const std::vector<std::string> &name_exprs,
                std::vector<const void *> &kernels, std::string &err, const std::string &defines, bool compile_only) {
  std::lock_guard<std::mutex> lock(g_mu);
  if (!nvrtc_load(err)) return -1;
  const std::string prob = jit_problem_header(P);
  std::string top;
  top += defines; /* e.g. "#define MCI_CHAIN_MAXT 512\n" */
  top += "#define MCI_JIT 1\n#include \"mci_vegas.cuh\"\n#include \"mci_chain.cuh\"\n";
  const bool user_f = P.fid == MCI_F_USER, user_m = P.measure_kind == MCI_MEASURE_USER;
  if (user_f || user_m) {
    if (!user_source || !*user_source) { err = "a user integrand / user measure needs mci_problem_t.user_source"; return -1; }
    top += "#line 1 \"user_source\"\n";
    top += user_source;
    top += "\n#line 1 \"mci_jit_glue\"\n";
  }
  if (user_m)
    top += "namespace mci {\ntemplate <> struct UserMeasureT<0> {\n  template <class XA>\n"
           "  __device__ __forceinline__ static void call(const XA &X, double *obs, const double *relw, const double *userdata) "
           "{ ::mci_user_measure(X, obs, relw, userdata); }\n};\n}\n";
  if (user_f) {
    top += "namespace mci {\ntemplate <> struct Integrand<MCI_F_USER> {\n"
           "  static constexpr int NI = JitProblem::ni;\n"
           "  template <int MAXD, class XA, bool EXACT = false>\n"
           "  __device__ __forceinline__ static void eval(const XA &X, const Problem &P, double *w) { ::mci_user_integrand(X, P.userdata, w); }\n"
           "};\n}\n";
  }
  std::string key = prob + "\x01" + top;
  for (const auto &n : name_exprs) key += "\x01" + n;
  auto it = g_cache.find(key);
  const std::string disk = it == g_cache.end() ? disk_cache_path(key) : std::string();
  if (it == g_cache.end() && !disk.empty()) {
    Compiled C;
    if (disk_cache_load(disk, name_exprs.size(), C)) it = g_cache.emplace(key, std::move(C)).first;
  }
  if (it == g_cache.end()) {
    std::vector<const char *> hsrc, hname;
    for (int i = 0; i < MCI_JIT_NUM_HEADERS; ++i) { hsrc.push_back(MCI_JIT_HEADER_SOURCES[i]); hname.push_back(MCI_JIT_HEADER_NAMES[i]); }
    hsrc.push_back(prob.c_str()); hname.push_back("mci_jit_problem.h");
    nvrtcProgram prog = nullptr;
    int rc = g_rtc.CreateProgram(&prog, top.c_str(), "mci_jit_program.cu", (int)hsrc.size(), hsrc.data(), hname.data());
    if (rc != 0) { err = "nvrtcCreateProgram failed"; return -1; }
    for (const auto &n : name_exprs) g_rtc.AddNameExpression(prog, n.c_str());
    const char *opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo", "-DMCI_JIT=1"};
    rc = g_rtc.CompileProgram(prog, 4, opts);
    size_t logn = 0;
    g_rtc.GetProgramLogSize(prog, &logn);
    std::string log(logn, '\0');
    if (logn > 1) g_rtc.GetProgramLog(prog, &log[0]);
    if (rc != 0) {
      err = std::string("NVRTC: ") + (g_rtc.GetErrorString ? g_rtc.GetErrorString(rc) : "compile error") + "\n" + log;
      g_rtc.DestroyProgram(&prog);
      return -1;
    }
    Compiled C;
    size_t n = 0;
    g_rtc.GetCUBINSize(prog, &n);
    C.cubin.resize(n);
    g_rtc.GetCUBIN(prog, C.cubin.data());
    for (const auto &ne : name_exprs) {
      const char *low = nullptr;
      if (g_rtc.GetLoweredName(prog, ne.c_str(), &low) != 0 || !low) { err = "no lowered name for " + ne; g_rtc.DestroyProgram(&prog); return -1; }
      C.lowered.push_back(low);
    }
    g_rtc.DestroyProgram(&prog);
    if (!disk.empty()) disk_cache_store(disk, C);
    it = g_cache.emplace(key, std::move(C)).first;
  }
  if (const char *dump = getenv("MCIB200_JIT_DUMP")) { /* keep the cubin for cuobjdump -sass / -res-usage (tools/jit_sass.py) */
    if (FILE *f = fopen(dump, "wb")) { fwrite(it->second.cubin.data(), 1, it->second.cubin.size(), f); fclose(f); }
  }
  if (compile_only) { kernels.clear(); return 0; } /* mci_jit_check: compile, do not load (no device needed) */
  /* one cudaLibrary per handle (context-bound); the cubin itself is cached */
  cudaLibrary_t lib = nullptr;
  cudaError_t e = cudaLibraryLoadData(&lib, it->second.cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
  if (e != cudaSuccess) { err = std::string("cudaLibraryLoadData: ") + cudaGetErrorString(e); return -1; }
  kernels.clear();
  for (const auto &low : it->second.lowered) {
    cudaKernel_t k = nullptr;
    e = cudaLibraryGetKernel(&k, lib, low.c_str());
    if (e != cudaSuccess) { err = std::string("cudaLibraryGetKernel(") + low + "): " + cudaGetErrorString(e); cudaLibraryUnload(lib); return -1; }
    kernels.push_back((const void *)k);
  }
  kernels.push_back((const void *)lib); /* last entry: the library handle, for mci_destroy */
  return 0;
}

void jit_unload(const void *lib) {
  if (lib) cudaLibraryUnload((cudaLibrary_t)lib);
}

}  // namespace mci
