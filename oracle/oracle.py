"""TEST INFRASTRUCTURE ONLY -- ctypes/numpy front-end to the CPU oracle.

* ``run(...)``     -> oracle/liboracle.so   (plain-C restatement, rnnt_oracle.c)
* ``run_ref(...)`` -> oracle/_ref/libmrnnt_ref.so (the unmodified reference CPU path,
  compiled from /root/reference by oracle/Makefile; present only if it was built)

Only tests/, ``__graft_entry__.smoke()`` and bench.py's cpu_baseline / ``--impl
reference`` legs may import this module.  The product never does.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from dataclasses import dataclass
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liboracle.so")
_REF = os.path.join(_HERE, "_ref", "libmrnnt_ref.so")
_REF_GPU = os.path.join(_HERE, "_ref", "ref_gpu_time")

_c_int_p = ctypes.POINTER(ctypes.c_int)


def build(force: bool = False) -> None:
    """Compile liboracle.so (and _ref/ when /root/reference is present)."""
    if force or not os.path.exists(_LIB) or (
        os.path.getmtime(_LIB) < max(os.path.getmtime(os.path.join(_HERE, f))
                                     for f in ("rnnt_oracle.c", "rnnt_oracle_body.inc"))):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    ref_dir = os.environ.get("RNNT_REF_DIR", "/root/reference")
    if os.path.isdir(os.path.join(ref_dir, "include")) and (force or not os.path.exists(_REF)):
        subprocess.check_call(["make", "-C", _HERE, "ref", f"RNNT_REF_DIR={ref_dir}"], stdout=subprocess.DEVNULL)
    if os.path.isdir(os.path.join(ref_dir, "include")) and (force or not os.path.exists(_REF_GPU)):
        subprocess.call(["make", "-C", _HERE, "ref_gpu", f"RNNT_REF_DIR={ref_dir}"], stdout=subprocess.DEVNULL)


def run_ref_gpu(B: int, T: int, S: int, V: int, timeout: float = 180.0) -> Optional[dict]:
    """Run the reference's own CUDA timing program (tests/test_time.cu, unmodified, compiled for sm_100a) on the
    current GPU: 10 calls of GpuRNNTComputer<float>::cost_and_grad on its own generated inputs, each timed by
    the program itself with the host clock around the (synchronous) call.  Returns None if the binary is absent
    or fails.  The first call carries one-time CUDA initialisation and is reported separately."""
    if not os.path.exists(_REF_GPU):
        return None
    try:
        out = subprocess.run([_REF_GPU, str(B), str(T), str(S), str(V)], capture_output=True, text=True,
                             timeout=timeout)
    except (subprocess.TimeoutExpired, OSError):
        return None
    times = [float(line.rsplit(":", 1)[1].split()[0]) for line in out.stdout.splitlines()
             if line.startswith("compute_rnnt_loss elapsed time")]
    if out.returncode != 0 or len(times) < 2:
        return None
    steady = sorted(times[1:])
    return {"ms_first_call": times[0], "ms_median": steady[len(steady) // 2], "ms_min": steady[0],
            "calls": len(times)}


_lib = None
_ref = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB)
    return _lib


def have_ref() -> bool:
    return os.path.exists(_REF)


def ref_lib() -> ctypes.CDLL:
    global _ref
    if _ref is None:
        if not have_ref():
            raise FileNotFoundError(f"{_REF} not built (needs /root/reference at build time)")
        _ref = ctypes.CDLL(_REF)
    return _ref


@dataclass
class OracleResult:
    costs: np.ndarray
    grads: Optional[np.ndarray] = None
    denom: Optional[np.ndarray] = None
    alphas: Optional[np.ndarray] = None
    betas: Optional[np.ndarray] = None
    ll_backward: Optional[np.ndarray] = None


def _i32(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.int32))


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def num_rows(T, S) -> int:
    T = np.asarray(T, dtype=np.int64)
    S = np.asarray(S, dtype=np.int64)
    return int((T * (S + 1)).sum())


def run(acts, labels, T, S, V: int, blank: int = 0, alignment=None, max_shift: int = 0,
        align_blank: Optional[int] = None, precision: str = "f32", want_grads: bool = True,
        want_lattice: bool = False, num_threads: int = 0) -> OracleResult:
    """Run the C restatement.

    precision: "f32" (mirrors CpuRNNTComputer<float>), "f64" (float64 inputs) or
    "f64_from_f32" (float32 inputs widened once, everything else in double).
    """
    T = _i32(T); S = _i32(S); labels = _i32(labels)
    B = int(T.shape[0])
    if B <= 0 or (T <= 0).any() or (S < 0).any() or (T < S).any():
        raise ValueError("oracle returned status 2")  # cpu_workspace_manager.h:99-107
    rows = num_rows(T, S)
    if precision == "f32":
        dt, fn = np.float32, lib().mrnnt_oracle_f32
        acts = np.ascontiguousarray(acts, dtype=np.float32)
    elif precision == "f64":
        dt, fn = np.float64, lib().mrnnt_oracle_f64
        acts = np.ascontiguousarray(acts, dtype=np.float64)
    elif precision == "f64_from_f32":
        dt, fn = np.float64, lib().mrnnt_oracle_f64_from_f32
        acts = np.ascontiguousarray(acts, dtype=np.float32)
    else:
        raise ValueError(precision)
    assert acts.size == rows * V, (acts.size, rows, V)
    align = None if alignment is None else _i32(alignment)
    if align_blank is None:
        align_blank = blank
    costs = np.empty(B, dtype=dt)
    grads = np.empty(rows * V, dtype=dt) if want_grads else None
    denom = np.empty(rows, dtype=dt) if want_lattice else None
    alphas = np.empty(rows, dtype=dt) if want_lattice else None
    betas = np.empty(rows, dtype=dt) if want_lattice else None
    llb = np.empty(B, dtype=dt) if want_lattice else None
    fn.restype = ctypes.c_int
    rc = fn(_ptr(acts), _ptr(labels), ctypes.c_int(B), _ptr(T), _ptr(S), ctypes.c_int(V), ctypes.c_int(blank),
            _ptr(align), ctypes.c_int(max_shift), ctypes.c_int(align_blank), ctypes.c_int(num_threads),
            _ptr(costs), _ptr(grads), _ptr(denom), _ptr(alphas), _ptr(betas), _ptr(llb))
    if rc != 0:
        raise ValueError(f"oracle returned status {rc}")
    if grads is not None:
        grads = grads.reshape(rows, V)
    return OracleResult(costs, grads, denom, alphas, betas, llb)


def run_ref(acts, labels, T, S, V: int, blank: int = 0, alignment=None, max_shift: int = 0,
            align_blank: Optional[int] = None, precision: str = "f32", want_grads: bool = True,
            num_threads: int = 0) -> OracleResult:
    """Run the unmodified reference CPU implementation (oracle/_ref)."""
    T = _i32(T); S = _i32(S); labels = _i32(labels)
    B = int(T.shape[0])
    rows = num_rows(T, S)
    if precision == "f32":
        dt, fn = np.float32, ref_lib().mrnnt_ref_f32
    elif precision == "f64":
        dt, fn = np.float64, ref_lib().mrnnt_ref_f64
    else:
        raise ValueError(precision)
    acts = np.ascontiguousarray(acts, dtype=dt)
    assert acts.size == rows * V
    align = None if alignment is None else _i32(alignment)
    if align_blank is None:
        align_blank = blank
    costs = np.empty(B, dtype=dt)
    grads = np.empty(rows * V, dtype=dt) if want_grads else None
    fn.restype = ctypes.c_int
    rc = fn(_ptr(acts), _ptr(labels), ctypes.c_int(B), _ptr(T), _ptr(S), ctypes.c_int(V), ctypes.c_int(blank),
            _ptr(align), ctypes.c_int(max_shift), ctypes.c_int(align_blank), ctypes.c_int(num_threads),
            _ptr(costs), _ptr(grads))
    if rc != 0:
        raise ValueError(f"reference returned status {rc}")
    if grads is not None:
        grads = grads.reshape(rows, V)
    return OracleResult(costs, grads)


def ref_gen_acts(n: int) -> np.ndarray:
    """tests/random.cpp genActs (mt19937 seed 0, U[0,1))."""
    out = np.empty(n, dtype=np.float32)
    ref_lib().mrnnt_ref_gen_acts(_ptr(out), ctypes.c_int(n))
    return out


def ref_gen_labels(V: int, S: int) -> np.ndarray:
    """tests/random.cpp genLabels (mt19937 seed 1, forced repeats)."""
    out = np.empty(S, dtype=np.int32)
    ref_lib().mrnnt_ref_gen_labels(ctypes.c_int(V), ctypes.c_int(S), _ptr(out))
    return out


def num_threads(reference: bool = False) -> int:
    if reference:
        return int(ref_lib().mrnnt_ref_num_threads())
    return int(lib().mrnnt_oracle_num_threads())
