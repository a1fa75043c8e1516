"""Helpers for the -m gpu parity tests: run a fixtures.Case through the C ABI on cuda:0."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

import monotonic_rnnt_b200 as mr
from monotonic_rnnt_b200 import _lib


@dataclass
class GpuResult:
    costs: np.ndarray
    grads: Optional[np.ndarray]
    handle: Optional[mr.LossHandle] = None


def to_dev(a, dtype):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dtype).cuda()


def run_case(case, want_grads=True, force_generic=False, keep_handle=False, host_lengths=True,
             poison_grads=True) -> GpuResult:
    acts = to_dev(case.acts.reshape(case.rows, case.V), torch.float32)
    labels = to_dev(case.labels, torch.int32)
    T = to_dev(case.T, torch.int32)
    S = to_dev(case.S, torch.int32)
    h = mr.LossHandle(acts, labels, T, S, lengths_host=(case.T, case.S) if host_lengths else None)
    if force_generic:
        h.set_option(_lib.OPT_FORCE_GENERIC, 1)
    if case.alignment is not None:
        h.restrict_to_alignment(to_dev(case.alignment, torch.int32), case.max_shift, case.blank)
    grads = None
    if want_grads:
        # poison: every element must be overwritten exactly once (the TF op does not pre-zero)
        grads = torch.full_like(acts, float("nan")) if poison_grads else torch.empty_like(acts)
    costs = h.cost_and_grad(case.blank, grads)
    res = GpuResult(costs.numpy().copy(), None if grads is None else grads.cpu().numpy(), h if keep_handle else None)
    if not keep_handle:
        h.close()
    return res


def diff_report(name, got, ref32, ref64):
    """Three-way gradient report: new-vs-f32 reference, new-vs-f64 truth, f32-vs-f64 floor."""
    d32 = np.abs(got.astype(np.float64) - ref32.astype(np.float64))
    d64 = np.abs(got.astype(np.float64) - ref64)
    floor = np.abs(ref32.astype(np.float64) - ref64)
    return {"case": name, "max_abs_vs_f32": float(d32.max()), "max_abs_vs_f64": float(d64.max()),
            "f32_vs_f64_floor": float(floor.max()), "n_gt_1e-5_vs_f64": int((d64 > 1e-5).sum())}
