#!/usr/bin/env python
"""sha256 over the kernel sources (mcintegration.jl_b200/csrc/*.cu, *.cuh, *.h and include/mcib200.h, in name order).
bench.py prints the numbers of a committed ncu capture (profiles/*_ncu.json: DRAM traffic, pipe utilisation) only
while this hash equals the capture's `csrc_sha256`, so a kernel change cannot leave them stale.
Usage: python tools/csrc_hash.py  -> prints the hash."""
import glob
import hashlib
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def csrc_sha256() -> str:
    d = os.path.join(ROOT, "mcintegration.jl_b200", "csrc")
    files = sorted(f for pat in ("*.cu", "*.cuh", "*.h") for f in glob.glob(os.path.join(d, pat)))
    files.append(os.path.join(ROOT, "include", "mcib200.h"))
    h = hashlib.sha256()
    for f in files:
        h.update(os.path.basename(f).encode() + b"\0")
        h.update(open(f, "rb").read())
    return h.hexdigest()


if __name__ == "__main__":
    print(csrc_sha256())
