"""-m gpu: the numeric envelope of the logits, and the reference's own size-test shapes.

Every other parity test draws logits from U[0,1) or 3*N(0,1).  Real joint networks produce logits of magnitude
10..100 with offsets, and the arithmetic of the three kernels meets them in different places:
  * K1 forms sum_v 2^(x*kLog2e - ML) with the single-float log2(e); the weights multiply their own logit by the
    two-float value: the difference, max * 1.9e-8 per row, does not cancel between rows of different maxima;
  * a single-float gradient coefficient c = log2(alpha beta / Z) + D carries half an ulp of |D| ~ max * 1.44 into
    every gradient element (magnitude 300: 1e-5 relative).
Both are tested here against the double-precision oracle at north_star's tolerances (costs 1e-5 relative, gradients
1e-5 absolute), with the three-way report (new vs f64, new vs the reference's f32 arithmetic, f32 vs f64 = the
reference's own rounding floor) printed.

Shapes of the reference's size tests: tensorflow_binding/test.py:159-176 (run_size_test: finite costs and gradients);
here with full parity against the oracle.
"""
import numpy as np
import pytest
import torch

import fixtures
from oracle import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gu():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import gpu_util
    return gpu_util


def _transform(case, kind, seed):
    """The same lattice with logits of another envelope."""
    rng = np.random.default_rng(seed)
    a = case.acts.astype(np.float64)
    if kind == "x10":
        a = a * 10.0
    elif kind == "x30":
        a = a * 30.0
    elif kind == "plus100":
        a = a + 100.0
    elif kind == "minus100":
        a = a - 100.0
    elif kind == "rowwise":   # every row its own scale (1..30) and its own offset (-100..100)
        scale = rng.uniform(1.0, 30.0, size=(a.shape[0], 1))
        offset = rng.uniform(-100.0, 100.0, size=(a.shape[0], 1))
        a = a * scale + offset
    else:
        raise ValueError(kind)
    return fixtures.Case(f"{case.name}_{kind}", a.astype(np.float32), case.labels, case.T, case.S, case.V, case.blank,
                         case.alignment, case.max_shift, None, dict(case.meta))


def _three_way(gu, case, capsys, cost_rtol=1e-5, grad_atol=1e-5):
    got = gu.run_case(case)
    o64 = oracle.run(case.acts, case.labels, case.T, case.S, case.V, blank=case.blank, alignment=case.alignment,
                     max_shift=case.max_shift, precision="f64_from_f32")
    o32 = oracle.run(case.acts, case.labels, case.T, case.S, case.V, blank=case.blank, alignment=case.alignment,
                     max_shift=case.max_shift, precision="f32")
    rel = float(np.max(np.abs(got.costs - o64.costs) / np.abs(o64.costs)))
    rep = gu.diff_report(case.name, got.grads, o32.grads, o64.grads)
    with capsys.disabled():
        print(f"\n[{case.name}] cost rel vs f64 {rel:.2e}; grads max|d| vs f64 {rep['max_abs_vs_f64']:.2e}, vs f32 "
              f"reference {rep['max_abs_vs_f32']:.2e}, f32 reference vs f64 (its own floor) {rep['f32_vs_f64_floor']:.2e}")
    assert np.isfinite(got.costs).all() and np.isfinite(got.grads).all()
    assert rel <= cost_rtol, rel
    assert rep["max_abs_vs_f64"] <= grad_atol, rep
    return rep


KINDS = ["x10", "x30", "plus100", "minus100", "rowwise"]


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("dist", ["normal3", "uniform"])
def test_c2_shaped_envelope(gu, capsys, kind, dist):
    """B=4 utterances of c2's shape (T=150, S=40, V=1000: the streaming kernels)."""
    base = fixtures.random_case(f"c2ish_{dist}", 401, B=4, V=1000, T_range=(150, 150), S_range=(40, 40), dist=dist)
    _three_way(gu, _transform(base, kind, 402), capsys)


@pytest.mark.parametrize("kind", KINDS)
def test_long_utterance_envelope(gu, capsys, kind):
    """T=800 frames, S=120 (c4's lattice, 4 chain warps per direction): the per-row error has 800 rows to add up in."""
    base = fixtures.random_case("t800", 411, B=2, V=256, T_range=(800, 800), S_range=(120, 120), dist="normal3")
    _three_way(gu, _transform(base, kind, 412), capsys)


@pytest.mark.parametrize("kind", ["x30", "rowwise"])
def test_envelope_with_alignment_band_and_ragged_lengths(gu, capsys, kind):
    base = fixtures.random_case("ragged", 421, B=6, V=512, T_range=(40, 120), S_range=(5, 30), dist="normal3")
    rng = np.random.default_rng(422)
    al = fixtures.random_alignment(rng, base.T, base.S, base.labels)
    _three_way(gu, _transform(base.with_alignment(al, 3, "ragged_shift3"), kind, 423), capsys)


@pytest.mark.parametrize("kind", ["x30", "minus100"])
def test_envelope_generic_kernels(gu, capsys, kind):
    """V % 4 != 0: the other pair of streaming kernels, the same arithmetic."""
    base = fixtures.random_case("v79", 431, B=3, V=79, T_range=(30, 60), S_range=(5, 20), dist="normal3")
    _three_way(gu, _transform(base, kind, 432), capsys)


# ---- the reference's own size-test shapes (tensorflow_binding/test.py:159-176) -----------------------------------------
SIZE_SHAPES = [(1, 150, 20, 50), (1, 150, 20, 5000), (16, 150, 20, 50), (16, 150, 20, 5000), (2, 391, 300, 79)]


@pytest.mark.parametrize("shape", SIZE_SHAPES, ids=lambda s: "B%d_T%d_S%d_V%d" % s)
def test_reference_size_test_shapes(gu, capsys, shape):
    B, T, S, V = shape
    case = fixtures.random_case("size_%d_%d_%d_%d" % shape, 441, B=B, V=V, T_range=(T, T), S_range=(S, S), dist="uniform")
    _three_way(gu, case, capsys)
