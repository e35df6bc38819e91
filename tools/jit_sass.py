#!/usr/bin/env python
"""Compile the specialised (NVRTC) program of one test case here — no GPU needed — and report per-kernel resource usage
and SASS size, optionally keeping the disassembly:

    python tools/jit_sass.py bubble vegasmc [--sass out.sass]

Uses mci_jit_check (compile only) with MCIB200_JIT_DUMP pointing at a scratch cubin, then cuobjdump."""
import argparse
import ctypes as C
import os
import re
import subprocess
import sys
import tempfile

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("case")
    ap.add_argument("solver", choices=["vegas", "vegasmc", "mcmc"])
    ap.add_argument("--sass", default=None)
    ap.add_argument("--walkers", type=int, default=1024)
    a = ap.parse_args()
    with tempfile.TemporaryDirectory() as tmp:
        cubin = os.path.join(tmp, "k.cubin")
        os.environ["MCIB200_JIT_DUMP"] = cubin
        from mcintegration_jl_b200 import _abi
        from cases import make_spec

        spec, _ = make_spec(a.case, a.solver, n_block=1024, neval_per_block=1 << 15, walkers=a.walkers if a.solver != "vegas" else 1)
        spec.jit = 1
        lib = _abi.load_library()
        p, keep = spec.to_ctypes()
        log = C.create_string_buffer(1 << 16)
        rc = lib.mci_jit_check(C.byref(p), log, 1 << 16)
        if rc != 0:
            print(log.value.decode())
            sys.exit(1)
        res = subprocess.run(["cuobjdump", "-res-usage", cubin], capture_output=True, text=True).stdout
        sass = subprocess.run(["cuobjdump", "-sass", cubin], capture_output=True, text=True).stdout
        if a.sass:
            open(a.sass, "w").write(sass)
        fn, cur, counts = None, None, {}
        for line in sass.splitlines():
            m = re.match(r"\s*Function : (\S+)", line)
            if m:
                cur = m.group(1)
                counts[cur] = 0
            elif cur and re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+\S", line):
                counts[cur] += 1
        names = subprocess.run(["cu++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
        for (k, n), name in zip(counts.items(), names):
            m = re.search(re.escape(k) + r".*?\n\s*(REG:\d+.*)", res)
            print(f"{name[:90]:90s} sass={n:6d} {m.group(1) if m else ''}")


if __name__ == "__main__":
    main()
