"""examples/c_driver.c: the reference's iteration loop written in plain C against include/mcib200.h only — the proof
that the boundary is a C ABI any host can bind (no Python, no torch types).  CPU: it compiles with gcc, links the library
and fails loudly without a device.  GPU: it integrates BASELINE C2."""
import math
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "mcintegration.jl_b200")


@pytest.fixture(scope="module")
def driver(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("cdrv") / "c_driver")
    subprocess.run(["gcc", "-O2", "-std=c11", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", "c_driver.c"), "-o", exe, "-L" + LIBDIR, "-lmcib200",
                    "-Wl,-rpath," + LIBDIR, "-lm"], check=True)
    return exe


def test_c_driver_builds_and_has_no_cpu_fallback(driver):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present: covered by the gpu test")
    r = subprocess.run([driver, "65536", "2"], capture_output=True, text=True)
    assert r.returncode != 0 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_c_driver_integrates_c2(driver):
    r = subprocess.run([driver, str(1 << 24), "10"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    m = re.search(r"result\s+([0-9.eE+-]+) \+- ([0-9.eE+-]+)\s+exact ([0-9.]+)\s+kernel (\S+)", r.stdout)
    assert m, r.stdout
    mean, sigma, exact = float(m.group(1)), float(m.group(2)), float(m.group(3))
    assert abs(exact - math.erf(5 / math.sqrt(2)) ** 10) < 1e-9
    assert abs(mean - exact) < 7 * sigma and sigma < 1e-4
    assert "vegas_pool_kernel" in m.group(4)
