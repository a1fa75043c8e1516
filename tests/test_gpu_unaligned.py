"""-m gpu: vocabularies that are not a multiple of 4 (rows of logits that are not whole 16-byte vectors).

The reference takes any V, and its own size tests use V = 50 and V = 79 (tensorflow_binding/test.py:159-176); real
vocabularies are 1025, 5001, ...  Up to round 1 such shapes fell to the generic kernels (a warp per row, scalar loads).
Now the bulk-copy ring carries them: what is copied for a run of rows is the aligned 16-byte window that covers it, and
the consumer warps read -- and the gradient kernel writes -- a row through the aligned vectors that cover it, masking
the elements at the two ends that belong to the neighbouring rows (include/mrnnt_b200/k1_lse.cuh: StreamWindow).

Checked here: parity with the double-precision oracle; every gradient element written exactly once and none of a
neighbouring row clobbered (the buffer is poisoned with NaN first, and every element is compared); agreement with the
generic kernels; the row-register variants' boundaries (a row spans one aligned vector more than V / 4 when it does not
start on a 16-byte boundary); the per-utterance scale of the backward half.
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


def _oracle64(case):
    return oracle.run(case.acts, case.labels, case.T, case.S, case.V, blank=case.blank, alignment=case.alignment,
                      max_shift=case.max_shift, precision="f64_from_f32")


# V: 17..79 (a few vectors per row), around the 32- and 64-register row variants (1024 +- 1, 2048 +- 1), large (5001)
VOCABS = [17, 18, 19, 50, 79, 127, 1001, 1022, 1023, 1025, 2047, 2049, 5001]


@pytest.mark.parametrize("restricted", [False, True], ids=["free", "aligned"])
@pytest.mark.parametrize("V", VOCABS)
def test_unaligned_vocab_parity(gu, V, restricted):
    B = 3 if V > 1500 else 5
    case = fixtures.random_case(f"ua_v{V}", 700 + V, B=B, V=V, T_range=(12, 40), S_range=(0, 14), dist="normal3",
                                blank=V // 3)
    if restricted:
        al = fixtures.random_alignment(np.random.default_rng(701 + V), case.T, case.S, case.labels, blank=case.blank)
        case = case.with_alignment(al, 2)
    ref = _oracle64(case)
    got = gu.run_case(case)                      # gradients poisoned with NaN before the call
    np.testing.assert_allclose(got.costs, ref.costs, rtol=1e-5)
    assert not np.isnan(got.grads).any(), "a gradient element was not written"
    assert np.abs(got.grads - ref.grads).max() <= 1e-5
    generic = gu.run_case(case, force_generic=True)
    np.testing.assert_allclose(got.costs, generic.costs, rtol=1e-6)
    assert np.abs(got.grads - generic.grads).max() <= 3e-6


@pytest.mark.parametrize("V", [50, 1025])
def test_unaligned_vocab_wraps_the_ring(gu, V):
    """Enough rows for every CTA's ring to wrap several times (the windows' slot offsets change from tile to tile)."""
    case = fixtures.random_case(f"ua_big_v{V}", 711, B=6, V=V, T_range=(150, 220), S_range=(30, 60), dist="uniform")
    ref = _oracle64(case)
    got = gu.run_case(case)
    np.testing.assert_allclose(got.costs, ref.costs, rtol=1e-5)
    assert np.abs(got.grads - ref.grads).max() <= 1e-5


def test_unaligned_vocab_backward_half_scales(gu):
    """mrnnt_enqueue_forward + mrnnt_enqueue_backward(scale) on unaligned rows: the SCALED gradient kernel."""
    import monotonic_rnnt_b200 as mr
    case = fixtures.random_case("ua_scaled", 721, B=4, V=1025, T_range=(20, 50), S_range=(3, 15), dist="normal3")
    ref = _oracle64(case)
    acts = gu.to_dev(case.acts.reshape(case.rows, case.V), torch.float32)
    h = mr.LossHandle(acts, gu.to_dev(case.labels, torch.int32), gu.to_dev(case.T, torch.int32),
                      gu.to_dev(case.S, torch.int32), lengths_host=(case.T, case.S))
    h.enqueue_forward(case.blank, want_grads=True)
    scale = torch.tensor([0.5, -2.0, 1.0, 3.0], device="cuda")
    grads = torch.full_like(acts, float("nan"))
    h.enqueue_backward(grads, scale)
    torch.cuda.synchronize()
    rows_b = case.T.astype(np.int64) * (case.S + 1)
    want = ref.grads * np.repeat(scale.cpu().numpy().astype(np.float64), rows_b)[:, None]
    assert np.abs(grads.cpu().numpy() - want).max() <= 3e-5
    h.close()


def test_unaligned_vocab_padded_layout(gu):
    """The padded [B, T, U, V] layout with an unaligned V: dead padding rows between live ones, windows per run."""
    import monotonic_rnnt_b200 as mr
    case = fixtures.random_case("ua_pad", 731, B=3, V=79, T_range=(10, 25), S_range=(2, 9), dist="normal3")
    ref = _oracle64(case)
    T_dim, U = int(case.T.max()) + 2, int(case.S.max()) + 3
    padded = np.full((case.B, T_dim, U, case.V), np.nan, np.float32)
    row = 0
    for b in range(case.B):
        for t in range(int(case.T[b])):
            for s in range(int(case.S[b]) + 1):
                padded[b, t, s] = case.acts[row]
                row += 1
    acts = torch.from_numpy(padded).cuda()
    h = mr.LossHandle(acts, gu.to_dev(case.labels, torch.int32), gu.to_dev(case.T, torch.int32),
                      gu.to_dev(case.S, torch.int32), lengths_host=(case.T, case.S))
    grads = torch.full_like(acts, float("nan"))
    costs = h.cost_and_grad(case.blank, grads).numpy()
    h.close()
    np.testing.assert_allclose(costs, ref.costs, rtol=1e-5)
    g = grads.cpu().numpy()
    assert not np.isnan(g).any()
    row = 0
    for b in range(case.B):
        for t in range(T_dim):
            for s in range(U):
                if t < case.T[b] and s <= case.S[b]:
                    assert np.abs(g[b, t, s] - ref.grads[row]).max() <= 1e-5
                    row += 1
                else:
                    assert (g[b, t, s] == 0).all()
