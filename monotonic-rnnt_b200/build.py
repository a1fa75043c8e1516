"""Builds libmonotonic_rnnt.so (the sm_100a library behind the C ABI) in-tree with nvcc.

The output lives in monotonic-rnnt_b200/lib/ (git-ignored, but it travels to the GPU box with the
repository snapshot).  nvcc cross-compiles without a GPU, so this also is the "does it build" check.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
INCLUDE_DIR = os.path.join(ROOT, "include")
CSRC_DIR = os.path.join(PKG_DIR, "csrc")
LIB_DIR = os.path.join(PKG_DIR, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libmonotonic_rnnt.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
    "-cudart", "shared",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _host_cxx() -> str:
    # $CXX=/opt/gcc/bin/g++ in this image lacks parts of the toolchain; prefer the system compiler
    return "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else (shutil.which("g++") or "g++")


def sources() -> list[str]:
    deps = [os.path.join(CSRC_DIR, f) for f in sorted(os.listdir(CSRC_DIR)) if f.endswith((".cu", ".cuh", ".h"))]
    for d, _, files in os.walk(INCLUDE_DIR):
        deps += [os.path.join(d, f) for f in files]
    return deps


def up_to_date() -> bool:
    if not os.path.exists(LIB_PATH):
        return False
    t = os.path.getmtime(LIB_PATH)
    return all(os.path.getmtime(p) <= t for p in sources())


def build(force: bool = False, verbose: bool = False, extra_flags: list[str] | None = None) -> str:
    if not force and up_to_date():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = [_nvcc(), *NVCC_FLAGS, "-ccbin", _host_cxx(), "-I", INCLUDE_DIR, *(extra_flags or []),
           "-o", LIB_PATH, os.path.join(CSRC_DIR, "c_api.cu")]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
        print(" ".join(cmd), file=sys.stderr)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed:\n{res.stdout}\n{res.stderr}")
    if verbose:
        print(res.stderr, file=sys.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
