"""-m gpu: seeded random sweep of the whole path against the double-precision oracle.

Shapes, batch sizes, vocabulary sizes, blank positions, alignment bands, layouts (packed / padded) and element types
(float32 / bfloat16) are drawn at random from ranges that cross every dispatch boundary of the engine: one or several
chain warps per direction (32 states each), 1..8 CTAs per utterance in the lattice kernel, streaming and generic
kernels, tiles with and without live rows, T == S, S == 0, T == 1; who writes the gradient's zero rows rotates too.
"""
import os

import numpy as np
import pytest
import torch

import fixtures
from oracle import oracle

pytestmark = pytest.mark.gpu

GRAD_ATOL = 1e-5
COST_RTOL = 1e-5
VOCABS = (3, 4, 5, 8, 17, 32, 40, 100, 256, 1000)


def _draw(seed):
    rng = np.random.default_rng(seed)
    B = int(rng.choice([1, 2, 3, 5, 8, 19, 40, 160]))
    V = int(rng.choice(VOCABS))
    t_hi = int(rng.choice([1, 4, 20, 70, 140]))
    s_hi = int(rng.choice([0, 1, 5, 31, 32, 33, 64, 100]))
    if B * t_hi * (min(s_hi, t_hi) + 1) * V > 6_000_000:      # keep the oracle in the seconds range
        B = max(1, 6_000_000 // (t_hi * (min(s_hi, t_hi) + 1) * V))
    t_lo = max(1, t_hi // 3)
    case = fixtures.random_case(f"fuzz{seed}", seed, B=B, V=V, T_range=(t_lo, t_hi), S_range=(0, s_hi),
                                dist=str(rng.choice(["uniform", "normal3"])), blank=int(rng.integers(0, V)))
    if rng.random() < 0.4 and int(case.S.max()) > 0:
        al = fixtures.random_alignment(rng, case.T, case.S, case.labels, case.blank)
        case = case.with_alignment(al, int(rng.choice([0, 1, 3, 1000])))
    return case, bool(rng.random() < 0.35), bool(rng.random() < 0.3), int(rng.choice([0, 1, 2, 8]))


@pytest.mark.parametrize("seed", range(int(os.environ.get("MRNNT_FUZZ_CASES", "48"))))
def test_random_sweep(seed):
    import monotonic_rnnt_b200 as mr
    from monotonic_rnnt_b200 import _lib
    case, padded, bf16, parts = _draw(seed)
    acts32 = np.ascontiguousarray(case.acts, dtype=np.float32).reshape(case.rows, case.V)
    if bf16:
        acts32 = torch.from_numpy(acts32).to(torch.bfloat16).to(torch.float32).numpy()
    o = oracle.run(acts32, case.labels, case.T, case.S, case.V, blank=case.blank, alignment=case.alignment,
                   max_shift=case.max_shift, precision="f64_from_f32")
    T64, S64 = case.T.astype(np.int64), case.S.astype(np.int64)
    labels = case.labels
    if padded:
        T_dim, U = int(T64.max()) + seed % 3, int(S64.max()) + 1 + seed % 2
        full = np.full((case.B, T_dim, U, case.V), np.nan, dtype=np.float32)
        off = 0
        for b in range(case.B):
            n = int(T64[b] * (S64[b] + 1))
            full[b, :T64[b], :S64[b] + 1] = acts32[off:off + n].reshape(T64[b], S64[b] + 1, case.V)
            off += n
        dev_acts = torch.from_numpy(full)
        labels = np.concatenate([labels, np.full((case.B, seed % 2), 7, np.int32)], axis=1)
    else:
        dev_acts = torch.from_numpy(acts32)
    dev_acts = dev_acts.to(torch.bfloat16 if bf16 else torch.float32).cuda()
    h = mr.LossHandle(dev_acts, torch.from_numpy(np.ascontiguousarray(labels)).cuda(),
                      torch.from_numpy(case.T).cuda(), torch.from_numpy(case.S).cuda())
    h.set_option(_lib.OPT_K2_PARTS, parts)
    # who zeroes the dead rows: automatic, the gradient kernel's consumers, 1 / 2 warps of the lattice kernel, the
    # gradient kernel's own zero-fill warp
    h.set_option(_lib.OPT_K2_ZERO_FILL, (-1, 0, 1, 2, 32)[seed % 5])
    h.set_option(_lib.OPT_DYNAMIC_TILES, (-1, 0, 1)[(seed // 5) % 3])   # the gradient kernel's tiles by counter: automatic, off, on
    if case.alignment is not None:
        h.restrict_to_alignment(torch.from_numpy(case.alignment).cuda(), case.max_shift, case.blank)
    grads = torch.full_like(dev_acts, float("nan"))
    costs = h.cost_and_grad(case.blank, grads).numpy()
    h.close()
    feasible = np.isfinite(o.costs)
    assert np.array_equal(np.isfinite(costs), feasible)                     # +inf where the band excludes the end state
    np.testing.assert_allclose(costs[feasible], o.costs[feasible], rtol=COST_RTOL, atol=1e-6)
    g = grads.to(torch.float32).cpu().numpy().astype(np.float64)
    want = o.grads.reshape(case.rows, case.V)
    tol = (np.abs(want) * 2.0 ** -8 if bf16 else 0.0) + GRAD_ATOL
    off = 0
    for b in range(case.B):
        n = int(T64[b] * (S64[b] + 1))
        got = g[b, :T64[b], :S64[b] + 1].reshape(n, case.V) if padded else g[off:off + n]
        if feasible[b]:
            assert np.all(np.abs(got - want[off:off + n]) <= (tol[off:off + n] if bf16 else tol)), (seed, b)
        if padded:
            g[b, :T64[b], :S64[b] + 1] = 0.0
        off += n
    if padded:
        assert np.all(g == 0.0)                                              # padding: written, exactly zero
