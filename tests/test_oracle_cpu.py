"""Pins oracle/rnnt_oracle.c (the plain-C restatement) against the reference.

1. hand-derived values from the reference's own tests (tests/test_cpu.cpp:57, :163-188, :291-294,
   :399-430, :510-546) at the reference's own tolerances (costs 1e-4, gradients 1e-2);
2. tests/golden/*.npz = outputs of the compiled reference (float and double templates):
   the float restatement must be BIT-IDENTICAL to CpuRNNTComputer<float>, the double one to 1e-12;
3. when oracle/_ref was built here, a fresh seeded ragged batch straight against the reference.
"""
import numpy as np
import pytest

import fixtures
import golden_io
from oracle import oracle


def _run(case, precision="f32", **kw):
    acts = case.acts if precision != "f64" else case.acts.astype(np.float64)
    return oracle.run(acts, case.labels, case.T, case.S, case.V, blank=case.blank, alignment=case.alignment,
                      max_shift=case.max_shift, precision=precision, **kw)


@pytest.mark.parametrize("case", fixtures.literal_cases(), ids=lambda c: c.name)
def test_reference_test_values(case):
    r = _run(case)
    assert np.all(np.abs(r.costs - case.expect_costs) < 1e-4)          # rnnt_helper::is_close
    c_only = _run(case, want_grads=False)
    assert np.array_equal(c_only.costs, r.costs)                        # bwd_test: cost() == cost_and_grad()


def test_readme_grads_two_decimals():
    r = _run(fixtures.readme_case())
    assert np.all(np.abs(r.grads.ravel() - fixtures.README_GRADS_2DP) < 1e-2)
    r = _run(fixtures.multibatch_case())
    exp = np.concatenate([fixtures.MULTIBATCH_B0_GRADS_2DP, fixtures.README_GRADS_2DP])
    assert np.all(np.abs(r.grads.ravel() - exp) < 1e-2)


@pytest.mark.parametrize("name", golden_io.names())
def test_golden_bit_exact_f32(name):
    case, ref = golden_io.load(name)
    r = _run(case)
    assert np.array_equal(r.costs, ref["costs_f32"])
    assert np.array_equal(r.grads, ref["grads_f32"], equal_nan=True)


@pytest.mark.parametrize("name", golden_io.names())
def test_golden_f64(name):
    case, ref = golden_io.load(name)
    r = _run(case, precision="f64")
    np.testing.assert_allclose(r.costs, ref["costs_f64"], rtol=1e-13, atol=0)
    np.testing.assert_allclose(r.grads, ref["grads_f64"], rtol=0, atol=1e-13)
    r2 = _run(case, precision="f64_from_f32")
    assert np.array_equal(r2.costs, r.costs) and np.array_equal(r2.grads, r.grads)


def test_infnan_finite():
    case, _ = golden_io.load("infnan")
    r = _run(case)
    assert np.isfinite(r.costs).all() and np.isfinite(r.grads).all()


def test_lattice_outputs_consistent():
    case, _ = golden_io.load("rand_v17_shift1")
    r = _run(case, precision="f64_from_f32", want_lattice=True)
    # beta(0,0) equals alpha(T-1,S) (cpu_rnnt.h:257-259 warns when they differ by > 0.1)
    np.testing.assert_allclose(r.ll_backward, -r.costs, rtol=1e-12)
    # every live row's gradient sums to zero over v
    g = r.grads.sum(axis=1)
    assert np.abs(g).max() < 1e-12


def test_validation():
    c = fixtures.readme_case()
    for T, S in (([0], [0]), ([2], [3]), ([4], [-1])):
        with pytest.raises(ValueError):
            oracle.run(c.acts, c.labels, T, S, c.V)


@pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("seed,V,shift", [(101, 9, None), (102, 24, 2), (103, 5, 0)])
def test_against_live_reference(seed, V, shift):
    case = fixtures.random_case(f"live{seed}", seed, B=6, V=V, T_range=(2, 30), S_range=(0, 10))
    if shift is not None:
        al = fixtures.random_alignment(np.random.default_rng(seed), case.T, case.S, case.labels)
        case = case.with_alignment(al, shift)
    kw = dict(blank=case.blank, alignment=case.alignment, max_shift=case.max_shift)
    ref = oracle.run_ref(case.acts, case.labels, case.T, case.S, case.V, **kw)
    got = _run(case)
    assert np.array_equal(got.costs, ref.costs)
    assert np.array_equal(got.grads, ref.grads, equal_nan=True)
