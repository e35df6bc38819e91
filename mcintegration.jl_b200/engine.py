"""ctypes front of the C ABI (include/mcib200.h): one engine = one mci_handle_t on one GPU.

`HandleEngine` is prefix-generic only so that tests can drive the CPU oracle (prefix `orc_`, built
from oracle/) through the very same host logic; the product only ever instantiates `CudaEngine`,
which loads libmcib200.so and fails loudly when the library or the GPU is missing.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _abi
from .configuration import ProblemSpec


class MCIError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"[{code}] {msg}")
        self.code = code


@dataclass
class IterationStats:
    """What one iteration hands to the host (mci_iteration_out_t)."""
    block_obs: np.ndarray   # [n_block, obs_total] un-normalised observables per block
    block_norm: np.ndarray  # [n_block]
    hist: np.ndarray        # [hist_len]
    visited: np.ndarray     # [Nd]
    propose: np.ndarray     # [3, Nd, pa]
    accept: np.ndarray      # [3, Nd, pa]
    neval: int


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


class HandleEngine:
    prefix = "mci_"

    def __init__(self, lib, spec: ProblemSpec):
        self.lib = lib
        self.spec = spec
        self._p, self._keep = spec.to_ctypes()
        h = C.c_void_p()
        rc = self._fn("create")(C.byref(self._p), C.byref(h))
        if rc != 0:
            msg = self._fn("last_error")(None)
            raise MCIError(rc, (msg or b"").decode())
        self.h = h
        s = [C.c_int32() for _ in range(6)]
        self._fn("sizes")(self.h, *[C.byref(x) for x in s])
        self.grid_len, self.disc_len, self.hist_len, self.obs_total, self.nd, self.pa_len = [x.value for x in s]
        self.n_disc_leaf = sum(1 for L in spec.leaves if L["kind"] == _abi.LEAF_DISCRETE)
        self.accum_len = self.disc_len + self.n_disc_leaf
        self.n_block = spec.n_block
        self.pa_dim = self.pa_len // (3 * self.nd)

    def _fn(self, name):
        return getattr(self.lib, self.prefix + name)

    def _check(self, rc):
        if rc != 0:
            raise MCIError(rc, (self._fn("last_error")(self.h) or b"").decode())

    def close(self):
        if getattr(self, "h", None):
            self._fn("destroy")(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- maps ---------------------------------------------------------------
    def set_maps(self, grids=None, dist=None, accum=None, reweight=None):
        conv = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64)
        g, d, a, r = conv(grids), conv(dist), conv(accum), conv(reweight)
        self._check(self._fn("set_maps")(self.h, _dp(g), _dp(d), _dp(a), _dp(r)))

    def get_maps(self):
        g = np.zeros(self.grid_len)
        d = np.zeros(self.disc_len)
        a = np.zeros(self.accum_len)
        r = np.zeros(self.nd)
        self._check(self._fn("get_maps")(self.h, _dp(g) if self.grid_len else None, _dp(d) if self.disc_len else None,
                                         _dp(a) if self.accum_len else None, _dp(r)))
        return g, d, a, r

    # ---- iteration ------------------------------------------------------------
    def _alloc_out(self):
        st = IterationStats(
            block_obs=np.zeros((self.n_block, self.obs_total)), block_norm=np.zeros(self.n_block),
            hist=np.zeros(self.hist_len), visited=np.zeros(self.nd), propose=np.zeros((3, self.nd, self.pa_dim)),
            accept=np.zeros((3, self.nd, self.pa_dim)), neval=0)
        nev = C.c_uint64(0)
        out = _abi.IterationOut(_dp(st.block_obs), _dp(st.block_norm), _dp(st.hist), _dp(st.visited), _dp(st.propose),
                                _dp(st.accept), C.pointer(nev))
        return st, out, nev

    def fetch_iteration(self) -> IterationStats:
        st, out, nev = self._alloc_out()
        self._check(self._fn("fetch_iteration")(self.h, C.byref(out)))
        st.neval = nev.value
        return st

    def refine(self):
        self._check(self._fn("refine")(self.h))

    def reweight(self, gamma=1.0, goal=None):
        g = None if goal is None else np.ascontiguousarray(goal, dtype=np.float64)
        self._check(self._fn("reweight")(self.h, float(gamma), _dp(g)))

    def set_partition(self, rank, world):
        self._check(self._fn("set_partition")(self.h, int(rank), int(world)))

    def n_draw(self):
        nv = len(self.spec.topvar_nleaf)
        dof = self.spec.dof
        return sum(max(d[v] for d in dof) * self.spec.topvar_nleaf[v] for v in range(nv))

    def trace_samples(self, it, block, first, count):
        nd = self.n_draw()
        x = np.zeros((count, nd))
        bins = np.zeros((count, nd), dtype=np.int32)
        jac = np.zeros(count)
        # integrand values: one double per integrand, (re, im) pairs for complex weights (MCI_WEIGHT_COMPLEX)
        w = np.zeros((count, self.spec.n_integrand * (2 if self.spec.weight_kind == _abi.WEIGHT_COMPLEX else 1)))
        self._check(self._fn("trace_samples")(self.h, int(it), int(block), int(first), int(count), _dp(x),
                                              bins.ctypes.data_as(C.POINTER(C.c_int32)), _dp(jac), _dp(w)))
        return x, bins, jac, w


class CudaEngine(HandleEngine):
    """The product engine: libmcib200.so on one B200."""
    prefix = "mci_"

    def __init__(self, spec: ProblemSpec):
        super().__init__(_abi.load_library(), spec)

    def run_iteration(self, it: int) -> IterationStats:
        st, out, nev = self._alloc_out()
        self._check(self.lib.mci_run_iteration(self.h, int(it), C.byref(out)))
        st.neval = nev.value
        return st

    def launch_iteration(self, it: int):
        self._check(self.lib.mci_launch_iteration(self.h, int(it)))

    def comm_init(self, id_bytes: bytes, rank: int, world: int):
        """attach the NCCL communicator named by the 128-byte id; an id this process has used before is JOINED, not
        re-created, so successive engines (successive integrate(comm=...) calls) share one warm communicator"""
        buf = C.create_string_buffer(id_bytes, 128)
        self._check(self.lib.mci_comm_init(self.h, buf, int(rank), int(world)))

    @staticmethod
    def comm_finalize():
        """destroy every NCCL communicator of this process (collective; after the last engine is closed)"""
        _abi.load_library().mci_comm_finalize()

    @staticmethod
    def comm_unique_id() -> bytes:
        lib = _abi.load_library()
        buf = C.create_string_buffer(128)
        rc = lib.mci_comm_unique_id(buf)
        if rc != 0:
            raise MCIError(rc, "ncclGetUniqueId failed")
        return buf.raw

    def stats_device_ptr(self):
        p = C.c_void_p()
        n = C.c_int64()
        self._check(self.lib.mci_stats_device_ptr(self.h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def nonfinite_count(self) -> int:
        """non-finite integrand values met by the last fetched iteration (the reference's debug warnings)"""
        n = C.c_uint64()
        self._check(self.lib.mci_nonfinite_count(self.h, C.byref(n)))
        return n.value

    def hist_rescues(self) -> int:
        """fetched iterations whose fixed-point histogram left its range and were recomputed in fp64 (mci_hist_rescues)"""
        n = C.c_uint64()
        self._check(self.lib.mci_hist_rescues(self.h, C.byref(n)))
        return n.value

    def run_iterations(self, first_iter: int, niter: int, *, gamma=1.0, goal=None, do_reweight=False, do_refine=True):
        """niter whole iterations (fill -> statistics -> doReweight! -> train!) in ONE library call, no host round
        trip in between (mci_run_iterations); returns the IterationStats of every iteration."""
        outs = (_abi.IterationOut * niter)()
        keep = []
        for k in range(niter):
            st, out, nev = self._alloc_out()
            outs[k] = out
            keep.append((st, nev))
        g = None if goal is None else np.ascontiguousarray(goal, dtype=np.float64)
        self._check(self.lib.mci_run_iterations(self.h, int(first_iter), int(niter), float(gamma), _dp(g),
                                                1 if do_reweight else 0, 1 if do_refine else 0, outs))
        for st, nev in keep:
            st.neval = nev.value
        return [st for st, _ in keep]

    def launch_count(self) -> int:
        n = C.c_uint64()
        self._check(self.lib.mci_launch_count(self.h, C.byref(n)))
        return n.value

    def last_timing(self):
        a, b = C.c_float(), C.c_float()
        self._check(self.lib.mci_last_timing(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def kernel_name(self) -> str:
        return (self.lib.mci_kernel_name(self.h) or b"").decode()

    def set_async(self, on: bool):
        self._check(self.lib.mci_set_async(self.h, 1 if on else 0))

    def timer_start(self):
        self._check(self.lib.mci_timer_start(self.h))

    def timer_stop(self) -> float:
        ms = C.c_float()
        self._check(self.lib.mci_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def fill_history(self, n=64):
        """CUDA-event durations (ms) of the fill kernel of the last n launched iterations, oldest first."""
        buf = (C.c_float * n)()
        got = C.c_int32()
        self._check(self.lib.mci_fill_history(self.h, buf, n, C.byref(got)))
        return [buf[i] for i in range(got.value)]

    def chain_info(self) -> dict:
        """chain solvers: chains per block (as given or auto-chosen), steps per chain, CTAs per block"""
        w, c = C.c_int32(), C.c_int32()
        st = C.c_uint64()
        self._check(self.lib.mci_chain_info(self.h, C.byref(w), C.byref(st), C.byref(c)))
        return dict(walkers_per_block=w.value, steps_per_walker=st.value, ctas_per_block=c.value)

    def stats_sync_host(self):
        self._check(self.lib.mci_stats_sync_host(self.h))

    def launch_info(self) -> dict:
        g, t, hm, tpb, spt = (C.c_int32() for _ in range(5))
        sm = C.c_int64()
        self._check(self.lib.mci_launch_info(self.h, C.byref(g), C.byref(t), C.byref(sm), C.byref(hm), C.byref(tpb),
                                             C.byref(spt)))
        return dict(grid=g.value, threads=t.value, smem_bytes=sm.value, hist_mode=hm.value,
                    tiles_per_block=tpb.value, samples_per_thread=spt.value, kernel=self.kernel_name())


def fp64_peak_tflops(device=0) -> float:
    lib = _abi.load_library()
    t = C.c_double()
    rc = lib.mci_fp64_peak(int(device), C.byref(t))
    if rc != 0:
        raise MCIError(rc, "mci_fp64_peak failed")
    return t.value
