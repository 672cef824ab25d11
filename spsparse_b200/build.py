"""Builds libspsparse_b200.so (hand-written sm_100a kernels + the C ABI) in-tree with nvcc."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "lib", "libspsparse_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-fmad=false",  # products and sums must round like the reference's scalar code (no FMA contraction)
    "-Xcompiler", "-fPIC", "-shared", "-cudart", "shared",
]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC))


def stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    inc = os.path.join(os.path.dirname(HERE), "include")
    deps = sources() + [os.path.join(inc, "spsparse_b200.h"), os.path.join(inc, "spsparse_b200", "base.hpp")]
    return any(os.path.getmtime(s) > t for s in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    cmd = ["nvcc"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB, os.path.join(CSRC, "spb_api.cu"), os.path.join(CSRC, "host_symbols.cpp")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode:
        raise RuntimeError("nvcc failed building libspsparse_b200.so")
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose="-v" in sys.argv))
