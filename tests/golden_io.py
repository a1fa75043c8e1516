"""Loader for tests/golden/*.npz (written by tests/golden/make_golden.py from the reference itself)."""
from __future__ import annotations

import glob
import os

import numpy as np

import fixtures

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def names() -> list[str]:
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load(name: str):
    """Returns (Case, dict of reference outputs)."""
    z = np.load(os.path.join(GOLDEN_DIR, f"{name}.npz"))
    case = fixtures.Case(name, z["acts"], z["labels"], z["T"], z["S"], int(z["V"]), int(z["blank"]),
                         z["alignment"] if "alignment" in z.files else None, int(z["max_shift"]),
                         z["expect_costs"] if "expect_costs" in z.files else None)
    ref = {k: z[k] for k in ("costs_f32", "grads_f32", "costs_f64", "grads_f64")}
    return case, ref
