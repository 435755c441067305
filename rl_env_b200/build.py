"""Build libplantos_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m rl_env_b200.build [--force]

The shared library sits next to its sources so it travels with a repo snapshot; it is
git-ignored.  No torch headers are involved: the library exposes the plain C ABI of
include/plantos.h and links the CUDA runtime statically.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
# PLANTOS_LIB points the loader at an alternative build (kernel-variant experiments)
LIB_PATH = os.environ.get("PLANTOS_LIB") or os.path.join(CSRC, "libplantos_b200.so")
SOURCES = ["plantos_abi.cu"]
HEADERS = ["plantos_common.cuh", "plantos_generic.cuh", "plantos_fast.cuh",
           os.path.join(ROOT, "include", "plantos.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    path = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(path):
        raise RuntimeError("nvcc not found; libplantos_b200.so cannot be built")
    return path


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + [h if os.path.isabs(h) else os.path.join(CSRC, h) for h in HEADERS]
    return any(os.path.getmtime(d) > built for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile if the library is missing or older than its sources; return its path."""
    if not force and not needs_build():
        return LIB_PATH
    cmd = [_nvcc(), *NVCC_FLAGS, "-o", LIB_PATH, *[os.path.join(CSRC, s) for s in SOURCES]]
    proc = subprocess.run(cmd, capture_output=True, text=True, cwd=CSRC)
    log = proc.stdout + proc.stderr
    with open(os.path.join(CSRC, "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + log)
    if verbose:
        print(log)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
