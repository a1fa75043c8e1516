"""-m gpu: the CUDA path (through the flat C ABI) against the oracle and the reference's golden outputs.

Tolerances (BASELINE.json north_star): per-utterance costs within 1e-5 relative, gradients within 1e-5
absolute, lengths / bands / label indexing bit-exact.  The gradient tolerance is applied against the
double-precision reference (`grads_f64`, CpuRNNTComputer<double>); against the float reference it holds on
the small fixtures and is bounded by that reference's own rounding floor on the large shapes.
"""
import numpy as np
import pytest
import torch

import fixtures
import golden_io
from oracle import oracle

pytestmark = pytest.mark.gpu

COST_RTOL = 1e-5
GRAD_ATOL = 1e-5


@pytest.fixture(scope="module")
def gu():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import gpu_util
    return gpu_util


def _oracle(case, precision):
    return oracle.run(case.acts, case.labels, case.T, case.S, case.V, blank=case.blank, alignment=case.alignment,
                      max_shift=case.max_shift, precision=precision, want_lattice=True)


def _check_costs(got, ref):
    ref = np.asarray(ref, dtype=np.float64)
    fin = np.isfinite(ref)
    assert np.array_equal(np.isfinite(got), fin)
    np.testing.assert_allclose(got[fin], ref[fin], rtol=COST_RTOL, atol=0)
    assert np.all(got[~fin] == ref[~fin])


@pytest.mark.parametrize("generic", [False, True], ids=["stream", "generic"])
@pytest.mark.parametrize("name", golden_io.names())
def test_golden(gu, name, generic):
    case, ref = golden_io.load(name)
    r = gu.run_case(case, force_generic=generic)
    _check_costs(r.costs, ref["costs_f32"])
    _check_costs(r.costs, ref["costs_f64"])
    assert not np.isnan(r.grads).any(), "an element of grads was not written"
    assert np.abs(r.grads - ref["grads_f64"]).max() <= GRAD_ATOL
    assert np.abs(r.grads - ref["grads_f32"]).max() <= GRAD_ATOL + np.abs(ref["grads_f32"] - ref["grads_f64"]).max()
    if case.expect_costs is not None:
        assert np.all(np.abs(r.costs - case.expect_costs) < 1e-4)      # the reference tests' own tolerance
    # cost() (gradients == NULL) gives the same costs (bwd_test, tests/test_gpu.cu)
    c = gu.run_case(case, want_grads=False, force_generic=generic)
    assert np.array_equal(c.costs, r.costs)


def test_readme_grads_two_decimals(gu):
    r = gu.run_case(fixtures.readme_case())
    assert np.all(np.abs(r.grads.ravel() - fixtures.README_GRADS_2DP) < 1e-2)      # tests/test_gpu.cu:247


@pytest.mark.parametrize("name", ["rand_v17_shift1", "rand_v32", "rand_edges", "rand_wide_shift3", "infnan"])
def test_intermediates(gu, name):
    """K1 (denominators, gathered log-probs), the band, row liveness and K2 (alpha, beta) one by one."""
    from monotonic_rnnt_b200 import _lib
    case, _ = golden_io.load(name)
    o = _oracle(case, "f64_from_f32")
    r = gu.run_case(case, keep_handle=True)
    h = r.handle
    T, S = case.T.astype(np.int64), case.S.astype(np.int64)
    row_start = np.concatenate([[0], np.cumsum(T * (S + 1))])
    assert np.array_equal(h.debug(_lib.DBG_ROWSTART), row_start)                       # bit-exact offsets
    # band: restatement of cpu_workspace_manager.h:207-224
    band = h.debug(_lib.DBG_BAND)
    for b in range(case.B):
        lo = np.zeros(T[b], np.int32); hi = np.full(T[b], S[b], np.int32)
        if case.alignment is not None:
            m = np.concatenate([[0], np.cumsum(case.alignment[b, :T[b]] != case.blank)])
            for t in range(T[b]):
                lo[t] = m[max(0, t + 1 - case.max_shift)]
                hi[t] = m[min(T[b], t + 1 + case.max_shift)]
        assert np.array_equal(band[b, :T[b], 0], lo) and np.array_equal(band[b, :T[b], 1], hi)
    meta = h.debug(_lib.DBG_ROWMETA)
    live = meta != -2
    denom, alpha, beta, lp = (h.debug(w) for w in (_lib.DBG_DENOM, _lib.DBG_ALPHA, _lib.DBG_BETA, _lib.DBG_LP))
    np.testing.assert_allclose(denom[live], o.denom[live], rtol=0, atol=5e-7)
    # a row is dead exactly when alpha(t-1, s) lies outside the lattice -> its gradient row is zero
    assert np.all(o.grads[~live] == 0.0)
    S_max = max(int(S.max()), 1)
    acts = case.acts.reshape(case.rows, case.V).astype(np.float64)
    for b in range(case.B):
        for i in range(int(row_start[b]), int(row_start[b + 1])):
            if not live[i]:
                continue
            assert abs(lp[i, 0] - (acts[i, case.blank] + o.denom[i])) < 5e-7
            s = (i - row_start[b]) % (S[b] + 1)
            if s < S[b]:
                assert meta[i] == b * S_max + s                                           # label indexing bit-exact
                assert abs(lp[i, 1] - (acts[i, case.labels[b, s]] + o.denom[i])) < 5e-7
            else:
                assert meta[i] == -1
    for got, ref in ((alpha, o.alphas), (beta, o.betas)):
        assert np.array_equal(np.isneginf(got), np.isneginf(ref))                        # same band, cell for cell
        fin = np.isfinite(ref)
        np.testing.assert_allclose(got[fin], ref[fin], rtol=0, atol=2e-5)
    ll = h.debug(_lib.DBG_LL)
    np.testing.assert_allclose(ll[0], -o.costs, rtol=1e-7)
    np.testing.assert_allclose(ll[1], o.ll_backward, rtol=1e-7)
    h.close()


SHAPES = [
    # name, B, V, T_range, S_range, dist
    ("v4_tiny_rows", 6, 4, (2, 30), (0, 10), "normal3"),
    ("v8", 5, 8, (5, 40), (1, 20), "uniform"),
    ("v64_g32", 7, 64, (10, 60), (3, 30), "normal3"),
    ("v1000", 4, 1000, (30, 60), (5, 20), "uniform"),
    ("v1000_peaky", 3, 1000, (30, 60), (5, 20), "normal3"),
    ("v1022_unaligned", 3, 1022, (20, 30), (4, 9), "uniform"),
    ("v5000_g1", 2, 5000, (20, 30), (4, 9), "uniform"),
    ("v16384_3stage", 1, 16384, (12, 16), (3, 5), "uniform"),
    ("v20000_too_big_for_ring", 1, 20000, (10, 12), (3, 4), "uniform"),
    ("s_gt_32", 3, 16, (70, 90), (33, 60), "normal3"),
    ("s_gt_64", 2, 12, (150, 170), (70, 120), "uniform"),
    ("s_gt_128", 2, 8, (300, 310), (130, 250), "uniform"),
    ("s_gt_256", 1, 8, (530, 540), (260, 500), "uniform"),
    ("s_gt_512_k4", 1, 8, (620, 640), (520, 600), "uniform"),
    ("s_gt_768_wide_kernel", 1, 8, (830, 840), (770, 800), "normal3"),
]


@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: s[0])
@pytest.mark.parametrize("restricted", [False, True], ids=["free", "aligned"])
def test_random_shapes(gu, shape, restricted):
    name, B, V, tr, sr, dist = shape
    case = fixtures.random_case(name, 1000 + len(name), B=B, V=V, T_range=tr, S_range=sr, dist=dist)
    if restricted:
        al = fixtures.random_alignment(np.random.default_rng(7), case.T, case.S, case.labels)
        case = case.with_alignment(al, 2)
    o64 = _oracle(case, "f64_from_f32")
    r = gu.run_case(case)
    _check_costs(r.costs, o64.costs)
    assert not np.isnan(r.grads).any()
    assert np.abs(r.grads - o64.grads).max() <= GRAD_ATOL
    # the generic kernels agree with the streaming ones to rounding
    g = gu.run_case(case, force_generic=True)
    _check_costs(g.costs, o64.costs)
    assert np.abs(g.grads - o64.grads).max() <= GRAD_ATOL


@pytest.mark.parametrize("V", [16, 1000], ids=["generic_v16", "stream_v1000"])
def test_masked_logits_leave_zero_rows_inside_the_lattice(gu, V):
    """Logits of -inf (a masked vocabulary entry) make whole rows INSIDE the lattice come out as exact zeros: the first
    label may only be emitted in frame 0, so beta(t, 0) = 0 for t >= 1.  Those rows are not the plan's dead rows (which
    the lattice kernel zeroes); the gradient kernel has to write them."""
    case = fixtures.random_case("masked", 77, B=3, V=V, T_range=(12, 20), S_range=(2, 5), dist="uniform")
    acts = case.acts.copy()
    row = 0
    for b in range(len(case.T)):
        Tb, Sb = int(case.T[b]), int(case.S[b])
        lab0 = int(case.labels[b, 0])
        if lab0 != case.blank:
            for t in range(1, Tb):
                acts[row + t * (Sb + 1), lab0] = -np.inf
        row += Tb * (Sb + 1)
    case = fixtures.Case("masked", acts, case.labels, case.T, case.S, case.V, case.blank)
    o64 = _oracle(case, "f64_from_f32")
    assert np.isfinite(o64.costs).all()
    zero_rows = np.where((o64.grads == 0).all(axis=1))[0]
    r = gu.run_case(case)
    _check_costs(r.costs, o64.costs)
    assert not np.isnan(r.grads).any(), "a row of grads was not written"
    assert np.abs(r.grads - o64.grads).max() <= GRAD_ATOL
    assert (r.grads[zero_rows] == 0).all()


def test_device_lengths_fetch_and_repeated_restrict(gu):
    """One manager reused across restrict_to_alignment + cost calls (tests/test_cpu.cpp:412-430), lengths
    fetched from the device (no host copies given)."""
    base = fixtures.readme_case()
    import gpu_util
    acts = gpu_util.to_dev(base.acts, torch.float32); labels = gpu_util.to_dev(base.labels, torch.int32)
    T = gpu_util.to_dev(base.T, torch.int32); S = gpu_util.to_dev(base.S, torch.int32)
    import monotonic_rnnt_b200 as mr
    h = mr.LossHandle(acts, labels, T, S)
    al = gpu_util.to_dev(np.array([[0, 1, 0, 2]], np.int32), torch.int32)
    assert abs(h.cost(0).item() + np.log(0.363)) < 1e-4
    for shift, p in ((2, 0.363), (0, 0.072), (1, 0.2958)):
        h.restrict_to_alignment(al, shift, 0)
        assert abs(h.cost(0).item() + np.log(p)) < 1e-4
    h.close()


def test_infeasible_alignment_gives_inf_cost(gu):
    """Alignment with the wrong number of labels: terminal state outside the band -> cost = +inf
    (SURVEY A.4).  Other utterances of the batch are unaffected."""
    base = golden_io.load("align_mb_shift0")[0]
    al = base.alignment.copy()
    al[0] = [0, 1, 0, 0]          # only one label emitted for S=2
    case = base.with_alignment(al, 0)
    o = _oracle(case, "f64_from_f32")
    r = gu.run_case(case)
    assert np.isposinf(o.costs[0]) and np.isposinf(r.costs[0])
    assert abs(r.costs[1] - o.costs[1]) < 1e-5
    rows0 = int(case.T[0] * (case.S[0] + 1))
    assert np.abs(r.grads[rows0:] - o.grads[rows0:]).max() <= GRAD_ATOL
    assert not np.isfinite(r.grads[:rows0]).all()     # as in the reference: no finite gradient exists


def test_validation_and_errors(gu):
    import monotonic_rnnt_b200 as mr
    c = fixtures.readme_case()
    import gpu_util
    acts = gpu_util.to_dev(c.acts, torch.float32); labels = gpu_util.to_dev(c.labels, torch.int32)
    for T, S in (([0], [0]), ([2], [3]), ([4], [-1])):
        with pytest.raises(mr.RNNTError) as e:
            mr.LossHandle(acts, labels, gpu_util.to_dev(np.array(T, np.int32), torch.int32),
                          gpu_util.to_dev(np.array(S, np.int32), torch.int32))
        assert e.value.status == 2
    with pytest.raises(RuntimeError):
        mr.LossHandle(acts.cpu(), labels, gpu_util.to_dev(c.T, torch.int32), gpu_util.to_dev(c.S, torch.int32))
    h = mr.LossHandle(acts, labels, gpu_util.to_dev(c.T, torch.int32), gpu_util.to_dev(c.S, torch.int32))
    with pytest.raises(mr.RNNTError) as e:
        h.cost(blank_label=3)                      # blank outside [0, V)
    assert e.value.status == 2
    h.close()


def test_autograd_op_matches_reference_semantics(gu):
    """monotonic_rnnt_loss: costs on the device, backward scales the saved gradients per utterance."""
    import monotonic_rnnt_b200 as mr
    import gpu_util
    case, ref = golden_io.load("rand_v32")
    acts = gpu_util.to_dev(case.acts, torch.float32).requires_grad_(True)
    labels = gpu_util.to_dev(case.labels, torch.int32)
    T = gpu_util.to_dev(case.T, torch.int32); S = gpu_util.to_dev(case.S, torch.int32)
    costs = mr.monotonic_rnnt_loss(acts, labels, T, S, blank_label=case.blank)
    assert costs.is_cuda and costs.shape == (case.B,)
    w = torch.arange(1, case.B + 1, device="cuda", dtype=torch.float32)
    (costs * w).sum().backward()
    np.testing.assert_allclose(costs.detach().cpu().numpy(), ref["costs_f64"], rtol=COST_RTOL)
    rows = case.T.astype(np.int64) * (case.S + 1)
    scale = np.repeat(np.arange(1, case.B + 1, dtype=np.float64), rows)[:, None]
    assert np.abs(acts.grad.cpu().numpy() - ref["grads_f64"] * scale).max() <= GRAD_ATOL * case.B
    loss_mod = mr.MonotonicRNNTLoss(blank_label=case.blank)
    c2 = loss_mod(acts.detach(), labels, T, S)
    assert torch.equal(c2, costs.detach())


@pytest.mark.parametrize("name", ["rand_v17", "rand_wide_shift3", "rand_v32"])
def test_backward_half_scales_per_utterance(gu, name):
    """enqueue_forward + enqueue_backward(scale): negative, zero and fractional upstream gradients, generic and
    streaming gradient kernels, repeated backward on one forward (the coefficients are not consumed)."""
    import monotonic_rnnt_b200 as mr
    import gpu_util
    case, ref = golden_io.load(name)
    acts = gpu_util.to_dev(case.acts, torch.float32)
    labels = gpu_util.to_dev(case.labels, torch.int32)
    T = gpu_util.to_dev(case.T, torch.int32); S = gpu_util.to_dev(case.S, torch.int32)
    h = mr.LossHandle(acts, labels, T, S)
    if case.alignment is not None:
        h.restrict_to_alignment(gpu_util.to_dev(case.alignment, torch.int32), case.max_shift, case.blank)
    costs = h.enqueue_forward(case.blank, want_grads=True).clone()
    torch.cuda.synchronize()
    np.testing.assert_allclose(costs.cpu().numpy(), ref["costs_f64"], rtol=COST_RTOL)
    rows = case.T.astype(np.int64) * (case.S + 1)
    w = np.array([(-2.0, 0.0, 0.5, 3.0, -0.25)[b % 5] for b in range(case.B)], dtype=np.float32)
    scale = np.repeat(w.astype(np.float64), rows)[:, None]
    g = torch.full_like(acts, float("nan"))
    for _ in range(2):
        h.enqueue_backward(g, torch.from_numpy(w).cuda())
        torch.cuda.synchronize()
        assert np.abs(g.cpu().numpy() - ref["grads_f64"] * scale).max() <= GRAD_ATOL * 3
    h.enqueue_backward(g, None)                                   # no scale: the plain gradient
    torch.cuda.synchronize()
    assert np.abs(g.cpu().numpy() - ref["grads_f64"]).max() <= GRAD_ATOL
    h2 = mr.LossHandle(acts, labels, T, S)
    with pytest.raises(mr.RNNTError):
        h2.enqueue_backward(g, None)                              # backward without a forward
    h2.enqueue_forward(case.blank, want_grads=False)
    with pytest.raises(mr.RNNTError):
        h2.enqueue_backward(g, None)                              # the forward kept no coefficients
    h.close(); h2.close()


def _pad_case(case, extra_T, extra_U, extra_S):
    """The packed case as the joint network would hold it: [B, T_dim, U, V] with NaN in every padded row (nothing
    there may ever be read), labels [B, S_dim] with an out-of-range label in the padding."""
    T, S = case.T.astype(np.int64), case.S.astype(np.int64)
    T_dim, U = int(T.max()) + extra_T, int(S.max()) + 1 + extra_U
    S_dim = max(int(S.max()) + extra_S, 1)
    acts4 = np.full((case.B, T_dim, U, case.V), np.nan, dtype=np.float32)
    packed = case.acts.reshape(case.rows, case.V)
    labels = np.full((case.B, S_dim), 2 ** 30, dtype=np.int32)
    off = 0
    for b in range(case.B):
        n = int(T[b] * (S[b] + 1))
        acts4[b, :T[b], :S[b] + 1] = packed[off:off + n].reshape(T[b], S[b] + 1, case.V)
        labels[b, :S[b]] = case.labels[b, :S[b]]
        off += n
    return acts4, labels


@pytest.mark.parametrize("name", ["rand_v32", "rand_v17_shift1"])
def test_forward_into_gradient_buffer(gu, name):
    """mrnnt_enqueue_forward_into: the lattice kernel zeroes the dead rows of the buffer the backward half will fill.
    Same result as the one-shot call; a backward half given ANOTHER buffer must still write every row of it."""
    import monotonic_rnnt_b200 as mr
    from monotonic_rnnt_b200 import _lib
    case, ref = golden_io.load(name)
    dev = torch.device("cuda", 0)
    V = 1000  # (the streaming kernels; the golden cases' own V is below their minimum)
    rng = np.random.default_rng(5)
    acts_np = rng.random((case.acts.shape[0], V), dtype=np.float32)
    wide = fixtures.Case(name + "_v1000", acts_np, case.labels, case.T, case.S, V, case.blank)
    if case.alignment is not None:
        wide = wide.with_alignment(case.alignment, case.max_shift)
    o64 = _oracle(wide, "f64_from_f32")
    acts = torch.from_numpy(acts_np).to(dev)
    h = mr.LossHandle(acts, torch.from_numpy(wide.labels).to(dev), torch.from_numpy(wide.T).to(dev),
                      torch.from_numpy(wide.S).to(dev))
    if wide.alignment is not None:
        h.restrict_to_alignment(torch.from_numpy(wide.alignment).to(dev), wide.max_shift, wide.blank)
    h.set_option(_lib.OPT_K2_ZERO_FILL, 2)
    g1 = torch.full_like(acts, float("nan"))
    costs = h.enqueue_forward(wide.blank, grads=g1).clone()
    assert h.get_option(_lib.OPT_K2_ZERO_FILL) == 2
    torch.cuda.synchronize()
    dead = h.debug(_lib.DBG_ROWMETA) == -2
    assert dead.any() and bool((g1[torch.from_numpy(dead).to(dev)] == 0).all()), "the forward half zeroes the dead rows"
    assert bool(torch.isnan(g1[torch.from_numpy(~dead).to(dev)]).all()), "... and nothing else"
    h.enqueue_backward(g1)
    torch.cuda.synchronize()
    _check_costs(costs.cpu().numpy(), o64.costs)
    assert not torch.isnan(g1).any()
    assert np.abs(g1.cpu().numpy() - o64.grads).max() <= GRAD_ATOL
    # forward into g1, backward into g2: g2 complete, nothing assumed about it
    g1.fill_(float("nan"))
    g2 = torch.full_like(acts, float("nan"))
    h.enqueue_forward(wide.blank, grads=g1)
    h.enqueue_backward(g2)
    torch.cuda.synchronize()
    assert not torch.isnan(g2).any()
    assert np.abs(g2.cpu().numpy() - o64.grads).max() <= GRAD_ATOL
    h.close()


@pytest.mark.parametrize("dyn", [0, 1], ids=["round_robin", "dynamic_tiles"])
@pytest.mark.parametrize("mode", [0, 1, 2, 32], ids=["k3_consumers", "k2_1warp", "k2_2warps", "k3_zero_warp"])
@pytest.mark.parametrize("restricted", [False, True], ids=["free", "aligned"])
def test_every_way_of_zeroing_the_dead_rows(gu, mode, restricted, dyn):
    """MRNNT_OPT_K2_ZERO_FILL: the same gradients whoever writes the rows that are zero by construction, on a shape
    the streaming kernels take (V = 1000), with and without an alignment band; the backward half on its own as well.
    MRNNT_OPT_DYNAMIC_TILES: the same again with the streaming kernels' tiles handed out through a counter."""
    import monotonic_rnnt_b200 as mr
    from monotonic_rnnt_b200 import _lib
    case = fixtures.random_case("zero_modes", 4242, B=6, V=1000, T_range=(20, 45), S_range=(3, 14), dist="uniform")
    if restricted:
        al = fixtures.random_alignment(np.random.default_rng(9), case.T, case.S, case.labels)
        case = case.with_alignment(al, 1)
    o64 = _oracle(case, "f64_from_f32")
    dev = torch.device("cuda", 0)
    acts = torch.from_numpy(case.acts).to(dev)
    h = mr.LossHandle(acts, torch.from_numpy(case.labels).to(dev), torch.from_numpy(case.T).to(dev),
                      torch.from_numpy(case.S).to(dev))
    if restricted:
        h.restrict_to_alignment(torch.from_numpy(case.alignment).to(dev), case.max_shift, case.blank)
    h.set_option(_lib.OPT_K2_ZERO_FILL, mode)
    h.set_option(_lib.OPT_DYNAMIC_TILES, dyn)
    for _ in range(3):   # (the hand-out counters must come back to zero after every call)
        g = torch.full_like(acts, float("nan"))
        costs = h.cost_and_grad(case.blank, g).numpy()
        assert h.get_option(_lib.OPT_K2_ZERO_FILL) == mode
        _check_costs(costs, o64.costs)
        assert not torch.isnan(g).any()
        assert np.abs(g.cpu().numpy() - o64.grads).max() <= GRAD_ATOL
    # two halves, the backward one with no help from the forward one
    g = torch.full_like(acts, float("nan"))
    h.enqueue_forward(case.blank, want_grads=True)
    h.enqueue_backward(g)
    torch.cuda.synchronize()
    assert not torch.isnan(g).any()
    assert np.abs(g.cpu().numpy() - o64.grads).max() <= GRAD_ATOL
    h.close()


@pytest.mark.parametrize("dyn", [0, 1], ids=["round_robin", "dynamic_tiles"])
@pytest.mark.parametrize("share", [1, 37, 64, 99], ids=lambda x: f"k2_takes_{x}pct")
@pytest.mark.parametrize("restricted", [False, True], ids=["free", "aligned"])
def test_zero_fill_split_between_lattice_and_gradient_kernel(gu, share, restricted, dyn):
    """MRNNT_OPT_K2_FILL_SHARE: the lattice kernel's fill writes the zero rows of the first `share` percent of the batch's
    units of 32 rows, the gradient kernel's consumer warps those of the rest.  Every element written, the bits of the
    unsplit fill, three calls in a row, and the two halves of a training step on their own."""
    import monotonic_rnnt_b200 as mr
    from monotonic_rnnt_b200 import _lib
    case = fixtures.random_case("fill_split", 777, B=7, V=1000, T_range=(25, 60), S_range=(4, 20), dist="uniform")
    if restricted:
        al = fixtures.random_alignment(np.random.default_rng(19), case.T, case.S, case.labels)
        case = case.with_alignment(al, 2)
    o64 = _oracle(case, "f64_from_f32")
    dev = torch.device("cuda", 0)
    acts = torch.from_numpy(case.acts).to(dev)
    h = mr.LossHandle(acts, torch.from_numpy(case.labels).to(dev), torch.from_numpy(case.T).to(dev),
                      torch.from_numpy(case.S).to(dev))
    if restricted:
        h.restrict_to_alignment(torch.from_numpy(case.alignment).to(dev), case.max_shift, case.blank)
    h.set_option(_lib.OPT_K2_ZERO_FILL, 2)
    h.set_option(_lib.OPT_DYNAMIC_TILES, dyn)
    h.set_option(_lib.OPT_K2_FILL_SHARE, 100)
    g0 = torch.full_like(acts, float("nan"))
    c0 = h.cost_and_grad(case.blank, g0).numpy().copy()
    assert h.get_option(_lib.OPT_K2_FILL_SHARE) == 100
    assert not torch.isnan(g0).any()
    assert np.abs(g0.cpu().numpy() - o64.grads).max() <= GRAD_ATOL
    h.set_option(_lib.OPT_K2_FILL_SHARE, share)
    for _ in range(3):
        g = torch.full_like(acts, float("nan"))
        costs = h.cost_and_grad(case.blank, g).numpy()
        assert h.get_option(_lib.OPT_K2_ZERO_FILL) == 2
        assert abs(h.get_option(_lib.OPT_K2_FILL_SHARE) - share) <= 1
        np.testing.assert_array_equal(costs, c0)
        assert torch.equal(g, g0)
    # the two halves: the forward half is told the buffer, the backward half finishes the fill
    g = torch.full_like(acts, float("nan"))
    h.enqueue_forward(case.blank, want_grads=True, grads=g)
    h.enqueue_backward(g)
    torch.cuda.synchronize()
    assert torch.equal(g, g0)
    # ... and a backward half into ANOTHER buffer than the one the forward half was told: all rows are its own
    g2 = torch.full_like(acts, float("nan"))
    h.enqueue_forward(case.blank, want_grads=True, grads=g)
    h.enqueue_backward(g2)
    torch.cuda.synchronize()
    assert torch.equal(g2, g0)
    h.close()


@pytest.mark.parametrize("pct", [60, 100], ids=["60pct_fixed", "all_fixed_but_the_remainder"])
def test_fixed_share_before_the_tile_counter(gu, pct):
    """MRNNT_OPT_DYNAMIC_TILES = 2..100: that percentage of a CTA's round-robin share of the gradient kernel's tiles is
    fixed, the counter hands out the rest -- on a shape with several tiles per CTA, three calls in a row (the counter must
    come back to zero), bit-identical to the round-robin hand-out."""
    import monotonic_rnnt_b200 as mr
    from monotonic_rnnt_b200 import _lib
    case = fixtures.random_case("fixed_share", 515, B=8, V=1000, T_range=(60, 80), S_range=(10, 14), dist="uniform")
    o64 = _oracle(case, "f64_from_f32")
    dev = torch.device("cuda", 0)
    acts = torch.from_numpy(case.acts).to(dev)
    h = mr.LossHandle(acts, torch.from_numpy(case.labels).to(dev), torch.from_numpy(case.T).to(dev),
                      torch.from_numpy(case.S).to(dev))
    h.set_option(_lib.OPT_DYNAMIC_TILES, 0)
    g0 = torch.full_like(acts, float("nan"))
    c0 = h.cost_and_grad(case.blank, g0).numpy().copy()
    h.set_option(_lib.OPT_DYNAMIC_TILES, pct)
    for _ in range(3):
        g = torch.full_like(acts, float("nan"))
        costs = h.cost_and_grad(case.blank, g).numpy()
        np.testing.assert_array_equal(costs, c0)
        assert torch.equal(g, g0)
    _check_costs(c0, o64.costs)
    assert np.abs(g0.cpu().numpy() - o64.grads).max() <= GRAD_ATOL
    h.close()


def test_automatic_choice_of_who_zeroes_the_dead_rows(gu):
    """A tight alignment band (nearly all rows dead): the zero-fill warps of the LSE and gradient kernels; a band that
    restricts nothing, or no band: the lattice kernel's fill (small shapes: its recursions are most of the call)."""
    import monotonic_rnnt_b200 as mr
    from monotonic_rnnt_b200 import _lib
    case = fixtures.random_case("zero_auto", 99, B=4, V=1000, T_range=(30, 40), S_range=(10, 14), dist="uniform")
    al = fixtures.random_alignment(np.random.default_rng(3), case.T, case.S, case.labels)
    dev = torch.device("cuda", 0)
    acts = torch.from_numpy(case.acts).to(dev)
    for shift, want in ((None, 2), (0, 32), (1000, 2)):
        c = case if shift is None else case.with_alignment(al, shift)
        o64 = _oracle(c, "f64_from_f32")
        h = mr.LossHandle(acts, torch.from_numpy(c.labels).to(dev), torch.from_numpy(c.T).to(dev),
                          torch.from_numpy(c.S).to(dev))
        if shift is not None:
            h.restrict_to_alignment(torch.from_numpy(c.alignment).to(dev), shift, c.blank)
        for _ in range(2):
            g = torch.full_like(acts, float("nan"))
            costs = h.cost_and_grad(c.blank, g).numpy()
            assert h.get_option(_lib.OPT_K2_ZERO_FILL) == want
            _check_costs(costs, o64.costs)
            assert not torch.isnan(g).any()
            assert np.abs(g.cpu().numpy() - o64.grads).max() <= GRAD_ATOL
        h.close()


@pytest.mark.parametrize("name,extra", [("rand_v32", (0, 0, 0)), ("rand_v32", (3, 2, 4)), ("rand_v17_shift1", (1, 5, 0)),
                                        ("rand_edges", (2, 1, 1)), ("rand_wide_shift3", (0, 3, 2)), ("readme", (0, 0, 0))])
def test_padded_layout_matches_packed(gu, name, extra):
    """the padded [B,T,U,V] entry gives the packed results, never reads the padding (NaN there) and
    writes exact zeros into the padded gradient rows; forward/backward halves and the autograd op included."""
    import monotonic_rnnt_b200 as mr
    import gpu_util
    case, ref = golden_io.load(name)
    acts4, labels = _pad_case(case, *extra)
    T64, S64 = case.T.astype(np.int64), case.S.astype(np.int64)
    acts = torch.from_numpy(acts4).cuda()
    lab = torch.from_numpy(labels).cuda()
    T = gpu_util.to_dev(case.T, torch.int32); S = gpu_util.to_dev(case.S, torch.int32)
    h = mr.LossHandle(acts, lab, T, S)
    if case.alignment is not None:
        h.restrict_to_alignment(gpu_util.to_dev(case.alignment, torch.int32), case.max_shift, case.blank)
    grads = torch.full_like(acts, float("nan"))
    costs = h.cost_and_grad(case.blank, grads).numpy()
    np.testing.assert_allclose(costs, ref["costs_f64"], rtol=COST_RTOL)
    g = grads.cpu().numpy()
    want = np.zeros_like(g, dtype=np.float64)
    off = 0
    for b in range(case.B):
        n = int(T64[b] * (S64[b] + 1))
        want[b, :T64[b], :S64[b] + 1] = ref["grads_f64"][off:off + n].reshape(T64[b], S64[b] + 1, case.V)
        off += n
    assert np.isfinite(g).all()
    assert np.abs(g - want).max() <= GRAD_ATOL
    pad = np.ones(g.shape[:3], dtype=bool)
    for b in range(case.B):
        pad[b, :T64[b], :S64[b] + 1] = False
    assert np.all(g[pad] == 0.0)                                   # exact zeros, written (the buffer held NaN)
    c2 = h.cost(case.blank).numpy()
    assert np.array_equal(c2, costs)
    h.close()
    if case.alignment is None:
        a = acts.clone().requires_grad_(True)
        w = torch.linspace(-1.0, 2.0, case.B, device="cuda")
        (mr.monotonic_rnnt_loss(a, lab, T, S, blank_label=case.blank) * w).sum().backward()
        assert a.grad.shape == acts.shape
        scale = w.cpu().numpy().astype(np.float64)[:, None, None, None]
        assert np.abs(a.grad.cpu().numpy() - want * scale).max() <= GRAD_ATOL * 2


def test_padded_layout_validation(gu):
    import monotonic_rnnt_b200 as mr
    import gpu_util
    case, _ = golden_io.load("rand_v32")
    acts4, labels = _pad_case(case, 0, 0, 0)
    T = gpu_util.to_dev(case.T, torch.int32); S = gpu_util.to_dev(case.S, torch.int32)
    with pytest.raises(mr.RNNTError):                              # U smaller than max S + 1
        mr.LossHandle(torch.from_numpy(acts4[:, :, :-1].copy()).cuda(), torch.from_numpy(labels).cuda(), T, S)
    with pytest.raises(mr.RNNTError):                              # fewer frames than max T
        mr.LossHandle(torch.from_numpy(acts4[:, :-1].copy()).cuda(), torch.from_numpy(labels).cuda(), T, S)
    if labels.shape[1] > 1:
        with pytest.raises(mr.RNNTError):                          # labels narrower than max S
            mr.LossHandle(torch.from_numpy(acts4).cuda(), torch.from_numpy(labels[:, :-1].copy()).cuda(), T, S)


@pytest.mark.parametrize("shape", [("pad_s_gt_32", 3, 16, (70, 90), (33, 60), "normal3"),
                                   ("pad_s_gt_128", 2, 8, (300, 310), (130, 250), "uniform"),
                                   ("pad_wide_kernel", 1, 8, (830, 840), (770, 800), "uniform")], ids=lambda s: s[0])
def test_padded_layout_multi_warp_rows(gu, shape):
    """Padded layout through the multi-warp lattice rows and the wide fallback kernel."""
    import monotonic_rnnt_b200 as mr
    import gpu_util
    name, B, V, tr, sr, dist = shape
    case = fixtures.random_case(name, 77, B=B, V=V, T_range=tr, S_range=sr, dist=dist)
    o64 = _oracle(case, "f64_from_f32")
    acts4, labels = _pad_case(case, 2, 3, 1)
    acts = torch.from_numpy(acts4).cuda()
    h = mr.LossHandle(acts, torch.from_numpy(labels).cuda(), gpu_util.to_dev(case.T, torch.int32),
                      gpu_util.to_dev(case.S, torch.int32))
    grads = torch.full_like(acts, float("nan"))
    costs = h.cost_and_grad(case.blank, grads).numpy()
    _check_costs(costs, o64.costs)
    g = grads.cpu().numpy()
    T64, S64 = case.T.astype(np.int64), case.S.astype(np.int64)
    off = 0
    for b in range(case.B):
        n = int(T64[b] * (S64[b] + 1))
        blk = g[b, :T64[b], :S64[b] + 1].reshape(n, case.V)
        assert np.abs(blk - o64.grads.reshape(-1, case.V)[off:off + n]).max() <= GRAD_ATOL
        g[b, :T64[b], :S64[b] + 1] = 0.0
        off += n
    assert np.all(g == 0.0)
    h.close()


def _bf16_round(a):
    """float32 array rounded to the nearest bfloat16 (ties to even), still held as float32."""
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(torch.bfloat16).to(torch.float32).numpy()


BF16_SHAPES = [
    # name, B, V, T_range, S_range, dist           streaming kernels need V % 8 == 0 in bfloat16
    ("bf16_v1000", 4, 1000, (30, 60), (5, 20), "uniform"),
    ("bf16_v1000_peaky", 3, 1000, (30, 60), (5, 20), "normal3"),
    ("bf16_v64_g32", 7, 64, (10, 60), (3, 30), "normal3"),
    ("bf16_v2048_64regs", 2, 2048, (12, 20), (3, 8), "uniform"),
    ("bf16_v5000_two_pass", 2, 5000, (20, 30), (4, 9), "uniform"),
    ("bf16_v12_generic", 5, 12, (5, 40), (1, 20), "normal3"),
    ("bf16_s_gt_32", 3, 16, (70, 90), (33, 60), "normal3"),
]


@pytest.mark.parametrize("shape", BF16_SHAPES, ids=lambda s: s[0])
def test_bf16_logits(gu, shape):
    """bfloat16 logits and gradients, float arithmetic.  The oracle runs in double on the SAME
    (bfloat16-representable) inputs: costs to the float tolerance, gradients to half a bfloat16 ulp of the output."""
    import monotonic_rnnt_b200 as mr
    import gpu_util
    name, B, V, tr, sr, dist = shape
    case = fixtures.random_case(name, 4242, B=B, V=V, T_range=tr, S_range=sr, dist=dist)
    acts32 = _bf16_round(case.acts)
    o = oracle.run(acts32, case.labels, case.T, case.S, case.V, blank=case.blank, precision="f64_from_f32")
    acts = torch.from_numpy(acts32.reshape(case.rows, case.V)).to(torch.bfloat16).cuda()
    labels = gpu_util.to_dev(case.labels, torch.int32)
    T = gpu_util.to_dev(case.T, torch.int32); S = gpu_util.to_dev(case.S, torch.int32)
    for generic in (False, True):
        h = mr.LossHandle(acts, labels, T, S)
        if generic:
            h.set_option(mr._lib.OPT_FORCE_GENERIC, 1)
        grads = torch.full_like(acts, float("nan"))
        costs = h.cost_and_grad(case.blank, grads).numpy()
        _check_costs(costs, o.costs)
        g = grads.to(torch.float32).cpu().numpy().astype(np.float64)
        want = o.grads.reshape(case.rows, case.V)
        assert np.isfinite(g).all()
        assert np.all(np.abs(g - want) <= np.abs(want) * 2.0 ** -8 + GRAD_ATOL)
        assert np.all(g[want == 0.0] == 0.0)
        h.close()
    # the autograd op on bfloat16 logits: gradient in bfloat16, scaled per utterance
    a = acts.clone().requires_grad_(True)
    w = torch.linspace(0.5, 2.0, case.B, device="cuda")
    (mr.monotonic_rnnt_loss(a, labels, T, S, blank_label=case.blank) * w).sum().backward()
    assert a.grad.dtype == torch.bfloat16
    rows = case.T.astype(np.int64) * (case.S + 1)
    scale = np.repeat(w.cpu().numpy().astype(np.float64), rows)[:, None]
    ga = a.grad.to(torch.float32).cpu().numpy().astype(np.float64)
    assert np.all(np.abs(ga - want * scale) <= np.abs(want * scale) * 2.0 ** -7 + GRAD_ATOL * 2)


def test_owned_workspace_is_recycled_across_handles(gu):
    """create_workspace / free_workspace per call, as the reference's torch binding does it: the blocks are recycled
    (larger, smaller, more handles than cache slots, two alive at once) and every call still gives the oracle's result."""
    import ctypes
    import monotonic_rnnt_b200 as mr
    from monotonic_rnnt_b200 import _lib
    import gpu_util
    lib = _lib.load()
    names = ["rand_v32", "readme", "rand_wide", "multibatch", "rand_v17", "rand_edges", "rand_v32", "rand_wide"]
    keep = []
    for round_ in range(2):
        for name in names:
            case, ref = golden_io.load(name)
            acts = gpu_util.to_dev(case.acts, torch.float32)
            labels = gpu_util.to_dev(case.labels, torch.int32)
            T = gpu_util.to_dev(case.T, torch.int32); S = gpu_util.to_dev(case.S, torch.int32)
            h = ctypes.c_void_p()
            _lib.check(lib.mrnnt_create(ctypes.byref(h), acts.data_ptr(), labels.data_ptr(), case.B, T.data_ptr(),
                                        S.data_ptr(), case.V, None, None), "create")
            _lib.check(lib.mrnnt_create_workspace(h), "create_workspace")
            grads = torch.full_like(acts, float("nan"))
            costs = torch.empty(case.B, dtype=torch.float32)
            _lib.check(lib.mrnnt_cost_and_grad(h, case.blank, torch.cuda.current_stream().cuda_stream, costs.data_ptr(),
                                               grads.data_ptr()), "cost_and_grad")
            np.testing.assert_allclose(costs.numpy(), ref["costs_f64"], rtol=COST_RTOL)
            assert np.abs(grads.cpu().numpy() - ref["grads_f64"]).max() <= GRAD_ATOL
            keep.append((h, acts, labels, T, S))
            if len(keep) > 2:                      # two handles stay alive: their blocks must not be handed out
                old = keep.pop(0)
                lib.mrnnt_free_workspace(old[0])
                lib.mrnnt_destroy(old[0])
    for old in keep:
        lib.mrnnt_free_workspace(old[0])
        lib.mrnnt_destroy(old[0])


@pytest.mark.parametrize("shape", [
    # name, B, V, T_range, S_range, aligned, max_shift
    ("ragged_free", 9, 16, (5, 120), (1, 40), False, 0),
    ("ragged_aligned", 9, 16, (5, 120), (1, 40), True, 2),
    ("one_utterance_long", 1, 8, (700, 700), (90, 90), True, 5),
    ("many_short", 300, 8, (3, 12), (0, 3), True, 1),
    ("s_zero", 4, 8, (1, 6), (0, 0), False, 0),
], ids=lambda s: s[0])
def test_fused_plan_equals_the_three_kernels(gu, shape):
    """MRNNT_OPT_FUSED_PLAN: row starts, band and row flags from ONE launch (plan.cuh: plan_fused_kernel) are the arrays
    the three set-up kernels produce, bit for bit -- packed and padded, with and without an alignment band, after a
    repeated restrict_to_alignment -- and so are the costs and gradients behind them."""
    import monotonic_rnnt_b200 as mr
    from monotonic_rnnt_b200 import _lib
    name, B, V, tr, sr, aligned, shift = shape
    case = fixtures.random_case("plan_" + name, 4000 + len(name), B=B, V=V, T_range=tr, S_range=sr, dist="uniform")
    if aligned:
        al = fixtures.random_alignment(np.random.default_rng(31), case.T, case.S, case.labels)
        case = case.with_alignment(al, shift)
    dev = torch.device("cuda", 0)
    acts = torch.from_numpy(case.acts.reshape(case.rows, case.V)).to(dev)
    out = {}
    for fused in (0, 1):
        h = mr.LossHandle(acts, torch.from_numpy(case.labels).to(dev), torch.from_numpy(case.T).to(dev),
                          torch.from_numpy(case.S).to(dev))
        h.set_option(_lib.OPT_FUSED_PLAN, fused)
        if aligned:
            ald = torch.from_numpy(case.alignment).to(dev)
            h.restrict_to_alignment(ald, case.max_shift + 3, case.blank)     # (a first band, replaced before any call ...
            h.restrict_to_alignment(ald, case.max_shift, case.blank)
        g = torch.full_like(acts, float("nan"))
        c = h.cost_and_grad(case.blank, g).clone()
        if aligned:                                                            # ... and once more after a call)
            h.restrict_to_alignment(ald, case.max_shift + 1, case.blank)
            h.cost_and_grad(case.blank, torch.empty_like(acts))
            h.restrict_to_alignment(ald, case.max_shift, case.blank)
            g.fill_(float("nan"))
            c2 = h.cost_and_grad(case.blank, g).clone()
            assert torch.equal(c2, c)
        torch.cuda.synchronize()
        out[fused] = (h.debug(_lib.DBG_ROWSTART), h.debug(_lib.DBG_BAND), h.debug(_lib.DBG_ROWMETA), c, g.clone(),
                      np.int64(h.get_option(_lib.OPT_LAUNCH_COUNT)))
        h.close()
    # the three kernels: 3 launches for the first plan, 2 (band, row flags) per later restrict_to_alignment; fused: 1 each
    assert int(out[0][5]) - int(out[1][5]) == (4 if aligned else 2)
    for a, b in zip(out[0][:5], out[1][:5]):
        if isinstance(a, torch.Tensor):
            assert torch.equal(a, b)
        else:
            assert np.array_equal(a, b)
    o64 = _oracle(case, "f64_from_f32")
    _check_costs(out[1][3].numpy(), o64.costs)
    assert np.abs(out[1][4].cpu().numpy() - o64.grads).max() <= GRAD_ATOL
