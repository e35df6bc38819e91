"""ctypes mirror of include/mcib200.h (the C ABI of libmcib200.so).

The struct layouts here are the single Python-side statement of the ABI; the
Julia side of the same structs is julia/MCIntegrationB200.jl.
"""
from __future__ import annotations

import ctypes as C
import os

ABI_VERSION = 3

MAX_TOPVAR = 16
MAX_LEAF = 32
MAX_INTEGRAND = 16
MAX_SLOT = 24
MAX_USERDATA = 64
MAX_OBS = 64

OK = 0
ERR_INVALID = -1
ERR_NO_DEVICE = -2
ERR_CUDA = -3
ERR_NUMERIC = -4
ERR_UNSUPPORTED = -5
ERR_COMM = -6
ERR_MCMC_START = -7

SOLVER_VEGAS, SOLVER_VEGASMC, SOLVER_MCMC = 0, 1, 2
LEAF_CONTINUOUS, LEAF_DISCRETE, LEAF_FERMIK = 0, 1, 2
MEASURE_DEFAULT, MEASURE_BIN_BY_DISCRETE, MEASURE_USER = 0, 1, 2
JIT_AUTO, JIT_ON, JIT_OFF = 0, 1, 2
WEIGHT_REAL, WEIGHT_COMPLEX = 0, 1
HIST_AUTO, HIST_SMEM_ATOMIC, HIST_GLOBAL_RED, HIST_WARP_PRIVATE, HIST_SMEM_STRIPED, HIST_SMEM_FIXED = 0, 1, 2, 3, 4, 5

# enum mci_integrand_id
F_GAUSS_ND = 1
F_LOG_OVER_SQRT = 2
F_GENZ_CORNER_POOL = 3
F_GENZ_CORNER_COMP = 4
F_BUBBLE = 5
F_PRODUCT_4 = 6
F_SPHERE1 = 7
F_SPHERE2 = 8
F_DISCRETE_X = 9
F_DISCRETE2_CONST = 10
F_LATTICE_POOL = 11
F_LATTICE_COMP = 12
F_HYPERSPHERE3 = 13
F_POLY3 = 14
F_X2_PLUS_Y2 = 15
F_BUBBLE_FERMIK = 16
F_COMPLEX1 = 17
F_COMPLEX2 = 18
F_USER = 100


class Leaf(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("nbins", C.c_int32),
        ("adapt", C.c_int32),
        ("lower_i", C.c_int32),
        ("lower", C.c_double),
        ("upper", C.c_double),
        ("alpha", C.c_double),
    ]


class Problem(C.Structure):
    _fields_ = [
        ("solver", C.c_int32),
        ("integrand_id", C.c_int32),
        ("measure_kind", C.c_int32),
        ("n_integrand", C.c_int32),
        ("n_topvar", C.c_int32),
        ("n_leaf", C.c_int32),
        ("topvar_nleaf", C.POINTER(C.c_int32)),
        ("topvar_offset", C.POINTER(C.c_int32)),
        ("leaf", C.POINTER(Leaf)),
        ("dof", C.POINTER(C.c_int32)),
        ("obs_len", C.POINTER(C.c_int32)),
        ("neighbor_ptr", C.POINTER(C.c_int32)),
        ("neighbor_idx", C.POINTER(C.c_int32)),
        ("userdata", C.POINTER(C.c_double)),
        ("n_userdata", C.c_int32),
        ("n_block", C.c_int32),
        ("neval_per_block", C.c_uint64),
        ("seed", C.c_uint64),
        ("measurefreq", C.c_int32),
        ("walkers_per_block", C.c_int32),
        ("thermal_ratio", C.c_double),
        ("device", C.c_int32),
        ("hist_policy", C.c_int32),
        ("jit", C.c_int32),
        ("reserved_", C.c_int32),
        ("user_source", C.c_char_p),
        ("chain_burnin", C.c_uint64),
        ("mcmc_start_idx", C.c_int32),
        ("weight_kind", C.c_int32),
    ]


class IterationOut(C.Structure):
    _fields_ = [
        ("block_obs", C.POINTER(C.c_double)),
        ("block_norm", C.POINTER(C.c_double)),
        ("hist", C.POINTER(C.c_double)),
        ("visited", C.POINTER(C.c_double)),
        ("propose", C.POINTER(C.c_double)),
        ("accept", C.POINTER(C.c_double)),
        ("neval", C.POINTER(C.c_uint64)),
    ]


# every symbol include/mcib200.h declares (tests check the .so exports all of them)
EXPORTS = [
    "mci_abi_version",
    "mci_create",
    "mci_destroy",
    "mci_last_error",
    "mci_sizes",
    "mci_set_maps",
    "mci_get_maps",
    "mci_run_iteration",
    "mci_launch_iteration",
    "mci_fetch_iteration",
    "mci_refine",
    "mci_reweight",
    "mci_comm_unique_id",
    "mci_comm_init",
    "mci_comm_finalize",
    "mci_set_partition",
    "mci_stats_device_ptr",
    "mci_trace_samples",
    "mci_launch_count",
    "mci_last_timing",
    "mci_fp64_peak",
    "mci_kernel_name",
    "mci_set_async",
    "mci_stats_sync_host",
    "mci_launch_info",
    "mci_timer_start",
    "mci_timer_stop",
    "mci_fill_history",
    "mci_jit_check",
    "mci_chain_info",
    "mci_nonfinite_count",
    "mci_run_iterations",
    "mci_hist_rescues",
]

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
# MCIB200_LIB overrides the in-tree library (used by tools/ to A/B experimental builds; the product default is in-tree)
LIB_PATH = os.environ.get("MCIB200_LIB") or os.path.join(_PKG_DIR, "libmcib200.so")


class LibraryMissing(RuntimeError):
    pass


_lib = None


def declare(lib, prefix="mci_"):
    """Attach argtypes/restypes for the handle-level API shared by libmcib200 (prefix mci_) and,
    in tests only, the oracle (prefix orc_)."""
    p = prefix
    dp = C.POINTER(C.c_double)
    ip = C.POINTER(C.c_int32)
    H = C.c_void_p

    def f(name, res, args):
        fn = getattr(lib, p + name)
        fn.restype = res
        fn.argtypes = args
        return fn

    f("create", C.c_int, [C.POINTER(Problem), C.POINTER(H)])
    f("destroy", None, [H])
    f("last_error", C.c_char_p, [H])
    f("sizes", C.c_int, [H, ip, ip, ip, ip, ip, ip])
    f("set_maps", C.c_int, [H, dp, dp, dp, dp])
    f("get_maps", C.c_int, [H, dp, dp, dp, dp])
    f("fetch_iteration", C.c_int, [H, C.POINTER(IterationOut)])
    f("refine", C.c_int, [H])
    f("reweight", C.c_int, [H, C.c_double, dp])
    f("set_partition", C.c_int, [H, C.c_int32, C.c_int32])
    f("trace_samples", C.c_int, [H, C.c_uint64, C.c_int32, C.c_uint64, C.c_int32, dp, ip, dp, dp])
    if prefix == "mci_":
        f("abi_version", C.c_int, [])
        f("run_iteration", C.c_int, [H, C.c_uint64, C.POINTER(IterationOut)])
        f("launch_iteration", C.c_int, [H, C.c_uint64])
        f("comm_unique_id", C.c_int, [C.c_void_p])
        f("comm_init", C.c_int, [H, C.c_void_p, C.c_int32, C.c_int32])
        f("comm_finalize", C.c_int, [])
        f("stats_device_ptr", C.c_int, [H, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)])
        f("launch_count", C.c_int, [H, C.POINTER(C.c_uint64)])
        f("last_timing", C.c_int, [H, C.POINTER(C.c_float), C.POINTER(C.c_float)])
        f("fp64_peak", C.c_int, [C.c_int32, dp])
        f("kernel_name", C.c_char_p, [H])
        f("set_async", C.c_int, [H, C.c_int32])
        f("timer_start", C.c_int, [H])
        f("timer_stop", C.c_int, [H, C.POINTER(C.c_float)])
        f("stats_sync_host", C.c_int, [H])
        f("fill_history", C.c_int, [H, C.POINTER(C.c_float), C.c_int32, ip])
        f("jit_check", C.c_int, [C.POINTER(Problem), C.c_char_p, C.c_int32])
        f("chain_info", C.c_int, [H, ip, C.POINTER(C.c_uint64), ip])
        f("launch_info", C.c_int, [H, ip, ip, C.POINTER(C.c_int64), ip, ip, ip])
        f("nonfinite_count", C.c_int, [H, C.POINTER(C.c_uint64)])
        f("hist_rescues", C.c_int, [H, C.POINTER(C.c_uint64)])
        f("run_iterations", C.c_int, [H, C.c_uint64, C.c_int32, C.c_double, dp, C.c_int32, C.c_int32, C.POINTER(IterationOut)])
    return lib


def load_library():
    """Load libmcib200.so.  There is no fallback: a missing library is an error."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LibraryMissing(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  mcintegration.jl_b200 has no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    declare(lib, "mci_")
    ver = lib.mci_abi_version()
    if ver != ABI_VERSION:
        raise LibraryMissing(f"{LIB_PATH} has ABI version {ver}, expected {ABI_VERSION}; rebuild")
    _lib = lib
    return lib
