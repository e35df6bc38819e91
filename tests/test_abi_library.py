"""CPU tests of the drop-in boundary: libmcib200.so loads without a GPU, exports exactly the symbols
include/mcib200.h declares, keeps its struct layouts in step with the ctypes mirror, and refuses to run
without a CUDA device (no CPU fallback, no route into oracle/)."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

import mcintegration_jl_b200 as mci
from mcintegration_jl_b200 import _abi
from cases import make_spec

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mcib200.h")


def _header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mci_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_abi.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return _abi.load_library()


def test_header_and_python_mirror_agree():
    assert _header_functions() == sorted(_abi.EXPORTS)


def test_library_exports_every_declared_symbol(lib):
    for name in _header_functions():
        assert hasattr(lib, name), name
    out = subprocess.run(["nm", "-D", "--defined-only", _abi.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (mci_[a-z0-9_]+)", out))
    assert exported == set(_header_functions()), exported ^ set(_header_functions())
    assert lib.mci_abi_version() == _abi.ABI_VERSION


def test_constants_match_header():
    src = open(HEADER).read()
    for name, val in re.findall(r"#define (MCI_[A-Z_]+) (-?\d+)", src):
        py = name.replace("MCI_", "")
        if hasattr(_abi, py):
            assert getattr(_abi, py) == int(val), name
    for body in re.findall(r"enum[^{]*\{([^}]*)\}", src):
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        for name, val in re.findall(r"(MCI_[A-Z0-9_]+)\s*=\s*(-?\d+)", body):
            py = name.replace("MCI_", "")
            assert hasattr(_abi, py), f"{name} missing from _abi.py"
            assert getattr(_abi, py) == int(val), name


def test_struct_layouts_match_a_c_compiler(tmp_path):
    """sizeof/offsetof of the three ABI structs as gcc sees them == the ctypes mirror"""
    fields = {"mci_leaf_t": _abi.Leaf, "mci_problem_t": _abi.Problem, "mci_iteration_out_t": _abi.IterationOut}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', "int main(void){"]
    for cname, cls in fields.items():
        lines.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for f, _ in cls._fields_:
            lines.append(f'printf("{cname}.{f} %zu\\n", offsetof({cname}, {f}));')
    lines.append("return 0;}")
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-o", str(exe), str(src)], check=True)
    got = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for cname, cls in fields.items():
        assert int(got[cname]) == C.sizeof(cls), cname
        for f, _ in cls._fields_:
            assert int(got[f"{cname}.{f}"]) == getattr(cls, f).offset, (cname, f)


def test_null_arguments_are_errors_not_crashes(lib):
    assert lib.mci_create(None, None) == _abi.ERR_INVALID
    assert lib.mci_sizes(None, None, None, None, None, None, None) == _abi.ERR_INVALID
    assert lib.mci_refine(None) == _abi.ERR_INVALID
    assert lib.mci_run_iteration(None, 0, None) == _abi.ERR_INVALID
    lib.mci_destroy(None)  # no-op
    assert b"null" in lib.mci_last_error(None)


def test_invalid_problems_are_rejected_before_touching_the_device(lib):
    from mcintegration_jl_b200.engine import MCIError
    spec, _ = make_spec("gauss10", "vegas")
    spec.solver = 7
    with pytest.raises(MCIError) as e:
        mci.CudaEngine(spec)
    assert e.value.code == _abi.ERR_INVALID and "not supported" in str(e.value)  # main.jl:263
    spec, _ = make_spec("gauss10", "vegas")
    spec.integrand_id = 999
    with pytest.raises(MCIError) as e:
        mci.CudaEngine(spec)
    assert e.value.code == _abi.ERR_UNSUPPORTED


def test_round2_validation_rules(lib):
    """describe_problem rejects, before any device is touched: a userdata blob shorter than the integrand reads (it would
    be zero-filled silently), a NULL / oversized blob, a real/complex mismatch, the bin-by-Discrete measure with several
    integrands under a chain solver (bins would be dropped), an mcmc start index that is no integrand."""
    from mcintegration_jl_b200.engine import MCIError

    def err(spec):
        with pytest.raises(MCIError) as e:
            mci.CudaEngine(spec)
        return e.value

    spec, _ = make_spec("bubble", "vegasmc")
    spec.userdata = spec.userdata[:7]
    e = err(spec)
    assert e.code == _abi.ERR_INVALID and "userdata" in str(e)
    spec, _ = make_spec("genz16_pool", "vegas")
    spec.userdata = spec.userdata[:10]
    assert err(spec).code == _abi.ERR_INVALID
    spec, _ = make_spec("product4", "vegas")
    spec.userdata = list(spec.userdata) + [0.0] * 100  # > MCI_MAX_USERDATA
    assert err(spec).code == _abi.ERR_INVALID
    spec, _ = make_spec("gauss10", "vegas")
    spec.weight_kind = _abi.WEIGHT_COMPLEX
    e = err(spec)
    assert e.code == _abi.ERR_INVALID and "real" in str(e)
    spec, _ = make_spec("complex2", "mcmc")
    spec.weight_kind = _abi.WEIGHT_REAL
    assert err(spec).code == _abi.ERR_INVALID
    spec, _ = make_spec("sphere2", "mcmc")
    spec.mcmc_start_idx = 5
    assert err(spec).code == _abi.ERR_INVALID
    # bin-by-Discrete with two integrands: fine under :vegas (no device here, so it fails LATER, with NO_DEVICE or not
    # at all), refused under the chain solvers
    import mcintegration_jl_b200 as M
    cfg = M.Configuration(var=(M.Continuous(0.0, 1.0), M.Discrete(1, 4, adapt=False)), dof=[[2, 1], [3, 1]], obs=[[0.0] * 4, 0.0])
    spec = M.problem_from_config(cfg, M.integrands.sphere2, solver=_abi.SOLVER_VEGASMC, measure_kind=_abi.MEASURE_BIN_BY_DISCRETE,
                                 n_block=2, neval_per_block=100)
    e = err(spec)
    assert e.code == _abi.ERR_UNSUPPORTED and "one integrand" in str(e)


def test_sixteen_integrands_are_accepted(lib):
    """MCI_MAX_INTEGRAND = 16, the reference's own bound (MaxOrder, src/distribution/distribution.jl:59): a 16-integrand
    user problem passes validation and compiles (NVRTC, no device needed); 17 are refused."""
    import ctypes as C
    src = """template <class XA> __device__ void mci_user_integrand(const XA &X, const double *ud, double *w) {
      for (int i = 0; i < %d; ++i) w[i] = X(0, 1) * (i + 1); }"""
    for n, ok in ((16, True), (17, False)):
        cfg = mci.Configuration(var=mci.Continuous(0.0, 1.0), dof=[[1]] * n) if ok else None
        if ok:
            f = mci.integrands.user_integrand(src % n, n_integrand=n)
            spec = mci.problem_from_config(cfg, f, solver=_abi.SOLVER_VEGAS, measure_kind=0, n_block=2, neval_per_block=100)
            p, keep = spec.to_ctypes()
            log = C.create_string_buffer(4096)
            rc = lib.mci_jit_check(C.byref(p), log, 4096)
            assert rc == 0 or b"nvrtc" in log.value.lower(), log.value  # NVRTC may be absent on a CPU-only box
        else:
            spec, _ = make_spec("gauss10", "vegas")
            spec.n_integrand = n
            from mcintegration_jl_b200.engine import MCIError
            with pytest.raises(MCIError):
                mci.CudaEngine(spec)


def test_no_device_means_no_fallback(lib):
    """Without a GPU mci_create must fail with MCI_ERR_NO_DEVICE — never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from mcintegration_jl_b200.engine import MCIError
    spec, _ = make_spec("gauss10", "vegas")
    with pytest.raises(MCIError) as e:
        mci.CudaEngine(spec)
    assert e.value.code == _abi.ERR_NO_DEVICE
    assert "no CPU fallback" in str(e.value)
    with pytest.raises(MCIError):
        mci.integrate(mci.integrands.gauss_nd, var=mci.Continuous(-5.0, 5.0), dof=[[10]], solver="vegas", neval=1e4)


def test_product_never_references_the_oracle():
    """Nothing in the package, the C ABI or its build may import, link or name oracle/."""
    pkg = os.path.join(ROOT, "mcintegration.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        if "build" in dirpath.split(os.sep) or "__pycache__" in dirpath:
            continue
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", "Makefile")):
                txt = open(os.path.join(dirpath, fn)).read()
                for bad in ("mci_oracle", "libmcioracle", "oracle_lib", "orc_create"):
                    assert bad not in txt, (fn, bad)
    needed = subprocess.run(["ldd", _abi.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in needed
    # the tuning and evidence scripts under tools/ run the product only: whatever needs the oracle as a yardstick lives in
    # tests/ (tests/perf_latency_sweep.py) or is a CPU leg of bench.py
    tools = os.path.join(ROOT, "tools")
    for fn in os.listdir(tools):
        if fn.endswith((".py", ".sh")):
            txt = open(os.path.join(tools, fn)).read()
            for bad in ("mci_oracle", "libmcioracle", "oracle_lib", "orc_create", "oracle_integrate", "OracleEngine"):
                assert bad not in txt, (fn, bad)


def test_missing_library_fails_loudly():
    """no CUDA extension -> an error that says so, never a silent CPU path"""
    code = ("import sys; sys.path.insert(0, %r); import mcintegration_jl_b200 as mci\n"
            "from mcintegration_jl_b200 import integrands\n"
            "try:\n"
            "    mci.integrate(integrands.gauss_nd, var=mci.Continuous(-5.0, 5.0), dof=[[10]], solver='vegas', neval=1e4)\n"
            "except Exception as e:\n"
            "    print(type(e).__name__, str(e)); sys.exit(3)\n" % ROOT)
    env = dict(os.environ, MCIB200_LIB="/nonexistent/libmcib200.so")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 3, r.stdout + r.stderr
    assert "LibraryMissing" in r.stdout and "no CPU fallback" in r.stdout


def test_comm_finalize_without_communicators():
    """mci_comm_finalize is safe to call with nothing to end (no GPU, NCCL never loaded)"""
    lib = _abi.load_library()
    assert lib.mci_comm_finalize() == 0
