"""Builds libauromat_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m auromat_b200.csrc.build [--force] [--verbose]

The shared object is git-ignored but travels to the GPU box with the repository snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SOURCES = ["amt.cu"]
# every header next to the sources takes part in the staleness test (a stale .so must never ship)
HEADERS = sorted(f for f in os.listdir(HERE) if f.endswith((".cuh", ".h"))) + \
          [os.path.join(ROOT, "include", "auromat_b200.h")]
LIB = os.path.join(HERE, "libauromat_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    # FP64 parity: no implicit mul+add contraction on the device (see amt_math.cuh) nor on the host
    "-fmad=false",
    "-Xcompiler", "-fPIC,-O2,-ffp-contract=off",
    "-shared", "-cudart", "static",
]


def find_nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; the CUDA extension cannot be built")
    return nvcc


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(HERE, s) for s in SOURCES] + [h if os.path.isabs(h) else os.path.join(HERE, h) for h in HEADERS]
    deps.append(os.path.abspath(__file__))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB
    extra = os.environ.get("AMT_NVCC_EXTRA", "").split()
    cmd = [find_nvcc()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + \
          ["-I", os.path.join(ROOT, "include"), "-o", LIB] + [os.path.join(HERE, s) for s in SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode:
        raise RuntimeError("nvcc failed (exit %d): %s" % (r.returncode, " ".join(cmd)))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
