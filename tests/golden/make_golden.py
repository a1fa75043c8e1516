"""Generate tests/golden/*.npz from the UNMODIFIED reference CPU implementation.

Run in the authoring container (needs /root/reference):

    make -C oracle            # builds oracle/_ref/libmrnnt_ref.so from the reference sources
    python tests/golden/make_golden.py

For every case the file stores the inputs and the outputs of
  * CpuRNNTComputer<float>::cost_and_grad   (keys costs_f32, grads_f32)
  * CpuRNNTComputer<double>::cost_and_grad on the widened logits (costs_f64, grads_f64)
so the GPU box (which has no /root/reference) can check both the oracle restatement
and the CUDA path against what the reference itself produced.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import fixtures  # noqa: E402
from oracle import oracle  # noqa: E402


def dump(case: fixtures.Case) -> None:
    kw = dict(blank=case.blank, alignment=case.alignment, max_shift=case.max_shift)
    r32 = oracle.run_ref(case.acts, case.labels, case.T, case.S, case.V, precision="f32", **kw)
    r64 = oracle.run_ref(case.acts.astype(np.float64), case.labels, case.T, case.S, case.V, precision="f64", **kw)
    c32 = oracle.run_ref(case.acts, case.labels, case.T, case.S, case.V, precision="f32", want_grads=False, **kw)
    assert np.array_equal(c32.costs, r32.costs, equal_nan=True), "cost() and cost_and_grad() disagree"
    out = dict(acts=case.acts, labels=case.labels, T=case.T, S=case.S, V=np.int32(case.V),
               blank=np.int32(case.blank), max_shift=np.int32(case.max_shift),
               costs_f32=r32.costs, grads_f32=r32.grads, costs_f64=r64.costs, grads_f64=r64.grads)
    if case.alignment is not None:
        out["alignment"] = case.alignment
    if case.expect_costs is not None:
        out["expect_costs"] = case.expect_costs
    np.savez_compressed(os.path.join(HERE, f"{case.name}.npz"), **out)
    print(f"{case.name:22s} B={case.B} rows={case.rows} V={case.V} costs={r32.costs}")


def main() -> None:
    oracle.build()
    assert oracle.have_ref(), "oracle/_ref/libmrnnt_ref.so missing: run `make -C oracle` with /root/reference present"
    for case in fixtures.literal_cases() + fixtures.golden_random_cases():
        dump(case)
    # infnan_test (tests/test_cpu.cpp:297-333): T=50 S=10 V=15 from the reference's own mt19937 generators
    T, S, V = 50, 10, 15
    acts = oracle.ref_gen_acts(T * (S + 1) * V).reshape(T * (S + 1), V)
    labels = oracle.ref_gen_labels(V, S).reshape(1, S)
    dump(fixtures.Case("infnan", acts, labels, np.array([T], np.int32), np.array([S], np.int32), V))


if __name__ == "__main__":
    main()
