"""Builds tests/cpp/bin/abi_check (C++ harness of the drop-in C++ surface) against libmonotonic_rnnt.so."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
OUT = os.path.join(HERE, "bin", "abi_check")


def build_abi_check() -> str:
    import monotonic_rnnt_b200 as mr

    lib = mr.build.build()
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    libdir = os.path.dirname(lib)
    cmd = [mr.build._nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O2", "-std=c++17",
           "-cudart", "shared", "-ccbin", mr.build._host_cxx(), "-I", mr.build.INCLUDE_DIR, "-o", OUT,
           os.path.join(HERE, "abi_check.cu"), "-L", libdir, "-lmonotonic_rnnt", "-Xlinker", "-rpath",
           "-Xlinker", "$ORIGIN/../../../monotonic-rnnt_b200/lib"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(res.stdout + res.stderr)
    return OUT


if __name__ == "__main__":
    print(build_abi_check())
