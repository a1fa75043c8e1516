"""Drop-in proof, build step: compile the reference's UNMODIFIED PyTorch binding translation unit
(/root/reference/pytorch_binding/monotonic_rnnt.cu, where it lies -- nothing is copied) against THIS
repository's include/ directory, with the flags the reference's own loader uses
(pytorch_binding/monotonic_rnnt_op.py:9-15: -DRNNT_ENABLE_GPU -O2, include path ../include) plus the sm_100a
arch.  The resulting extension lands in tests/dropin/_build/ (git-ignored; it travels to the GPU box), where
tests/test_gpu_dropin.py loads it and drives it with CUDA tensors.

Run here (needs /root/reference and a few minutes of nvcc):  python tests/dropin/build_dropin.py
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("RNNT_REF_DIR", "/root/reference")
BUILD = os.path.join(HERE, "_build")
NAME = "monotonic_rnnt_cpp"   # the module name the reference's monotonic_rnnt_op.py expects


def main() -> int:
    src = os.path.join(REF, "pytorch_binding", "monotonic_rnnt.cu")
    if not os.path.exists(src):
        print(f"no reference at {REF}: keeping whatever is in {BUILD}")
        return 0
    os.environ.setdefault("CXX", "/usr/bin/g++")
    os.environ.setdefault("CC", "/usr/bin/gcc")
    os.environ["TORCH_CUDA_ARCH_LIST"] = "10.0a"
    os.makedirs(BUILD, exist_ok=True)
    from torch.utils.cpp_extension import load

    load(name=NAME, sources=[src], extra_cuda_cflags=["-DRNNT_ENABLE_GPU", "-O2", "-lineinfo"],
         extra_include_paths=[os.path.join(ROOT, "include")], build_directory=BUILD, verbose=True)
    so = os.path.join(BUILD, NAME + ".so")
    assert os.path.exists(so), so
    for junk in ("monotonic_rnnt.cuda.o", "build.ninja", ".ninja_deps", ".ninja_log"):
        p = os.path.join(BUILD, junk)
        if os.path.exists(p):
            os.remove(p)
    print("built", so)
    return 0


if __name__ == "__main__":
    sys.exit(main())
