"""-m gpu: the BASELINE.json configurations at (or near) full size.

* c2 (B=32 T=150 S=40 V=1000) in full against the oracle: costs 1e-5 relative, gradients 1e-5 absolute
  against the double-precision oracle, and the three-way report of SURVEY 7.3-1 (new vs float reference,
  new vs double truth, float reference vs double truth = the float path's own rounding floor).
* c3 / c4 / c5: the WHOLE batch against the double-precision oracle (the oracle takes it in groups of utterances, all
  host threads: about a minute in all), plus the size-independent properties the domain offers: every gradient row sums to zero over the vocabulary
  (it is softmax-folded), rows outside the lattice are exactly zero, forward and backward likelihood agree,
  cost-only equals cost-and-grad, and the result does not depend on the kernel variant.
These sizes wrap the shared-memory ring of the streaming kernels many times per CTA, which the small
fixtures do not.
"""
import numpy as np
import pytest
import torch

from oracle import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import monotonic_rnnt_b200 as mr
    from monotonic_rnnt_b200 import _lib
    return mr, _lib


def _device_workload(mr, _lib, wl, nbatch=None):
    B = nbatch or wl.B
    rows = int((wl.T[:B].astype(np.int64) * (wl.S[:B] + 1)).sum())
    acts = torch.empty((rows, wl.V), dtype=torch.float32, device="cuda")
    _lib.check(_lib.load().mrnnt_synth_uniform(acts.data_ptr(), rows * wl.V, wl.logits_seed, 0,
                                               torch.cuda.current_stream().cuda_stream), "synth")
    T = torch.from_numpy(wl.T[:B].copy()).cuda()
    S = torch.from_numpy(wl.S[:B].copy()).cuda()
    s_max = int(wl.S[:B].max())
    labels_np = np.ascontiguousarray(wl.labels[:B, :s_max])
    labels = torch.from_numpy(labels_np).cuda()
    al_np = None
    if wl.alignment is not None:
        al_np = np.ascontiguousarray(wl.alignment[:B, :int(wl.T[:B].max())])
    return acts, labels, T, S, labels_np, al_np


def _run(mr, _lib, wl, nbatch=None, generic=False, want_grads=True):
    acts, labels, T, S, labels_np, al_np = _device_workload(mr, _lib, wl, nbatch)
    B = int(T.shape[0])
    h = mr.LossHandle(acts, labels, T, S, lengths_host=(wl.T[:B], wl.S[:B]))
    if generic:
        h.set_option(_lib.OPT_FORCE_GENERIC, 1)
    if al_np is not None:
        h.restrict_to_alignment(torch.from_numpy(al_np).cuda(), wl.max_shift, wl.blank)
    grads = torch.full_like(acts, float("nan")) if want_grads else None
    costs = h.cost_and_grad(wl.blank, grads).numpy().copy()
    ll = h.debug(_lib.DBG_LL) if want_grads else None
    meta = h.debug(_lib.DBG_ROWMETA) if want_grads else None
    h.close()
    return acts, grads, costs, ll, meta, labels_np, al_np


def _oracle(wl, acts, labels_np, al_np, B, precision):
    return oracle.run(acts.cpu().numpy(), labels_np, wl.T[:B], wl.S[:B], wl.V, blank=wl.blank, alignment=al_np,
                      max_shift=wl.max_shift, precision=precision)


def _properties(grads, costs, ll, meta):
    assert not torch.isnan(grads).any(), "an element of grads was not written"
    assert np.isfinite(costs).all()
    np.testing.assert_allclose(ll[0], ll[1], rtol=1e-9, atol=1e-6)           # alpha(T-1,S) == beta(0,0)
    np.testing.assert_allclose(costs, -ll[0], rtol=1e-6)
    row_sums = grads.sum(dim=1, dtype=torch.float64).abs().max().item()
    assert row_sums < 2e-4, row_sums                                          # sum_v g = 0 up to float rounding of V terms
    dead = torch.from_numpy(meta == -2).cuda()
    assert (grads[dead] == 0).all()                                           # rows outside the lattice: exact zeros
    assert grads.abs().max().item() <= 1.0 + 1e-5                             # |posterior differences| <= 1


def test_c2_full_parity_three_way(env, capsys):
    mr, _lib = env
    wl = mr.synth.workload("c2")
    acts, grads, costs, ll, meta, labels_np, al_np = _run(mr, _lib, wl)
    _properties(grads, costs, ll, meta)
    o64 = _oracle(wl, acts, labels_np, al_np, wl.B, "f64_from_f32")
    o32 = _oracle(wl, acts, labels_np, al_np, wl.B, "f32")
    g = grads.cpu().numpy().astype(np.float64)
    rel_cost = np.max(np.abs(costs - o64.costs) / np.abs(o64.costs))
    rel_cost32 = np.max(np.abs(costs - o32.costs) / np.abs(o32.costs))
    d64 = np.abs(g - o64.grads).max()
    d32 = np.abs(g - o32.grads).max()
    floor = np.abs(o32.grads.astype(np.float64) - o64.grads).max()
    with capsys.disabled():
        print(f"\n[c2 three-way] cost rel vs f64 {rel_cost:.2e}, vs f32 {rel_cost32:.2e}; grads max|d| vs f64 "
              f"{d64:.2e}, vs f32 reference {d32:.2e}, f32 reference vs f64 (its own floor) {floor:.2e}")
    assert rel_cost <= 1e-5 and rel_cost32 <= 1e-5
    assert d64 <= 1e-5
    assert d32 <= floor + 1e-5


def _full_batch_against_f64_oracle(mr, _lib, wl, capsys, group_bytes=4 << 30):
    """The WHOLE named batch against the double-precision oracle: the GPU runs the batch once; the oracle takes it in
    groups of consecutive utterances of at most `group_bytes` of logits (labels / alignment re-strided to the group's own
    maxima, as the ABI wants them), all host threads, so that the double-precision gradients of a group fit in memory."""
    acts, grads, costs, ll, meta, labels_np, al_np = _run(mr, _lib, wl)
    rows_b = wl.T.astype(np.int64) * (wl.S.astype(np.int64) + 1)
    starts = np.concatenate([[0], np.cumsum(rows_b)])
    worst_cost, worst_grad, b0, groups = 0.0, 0.0, 0, 0
    while b0 < wl.B:
        b1 = b0 + 1
        while b1 < wl.B and (starts[b1 + 1] - starts[b0]) * wl.V * 4 <= group_bytes:
            b1 += 1
        sh = mr.shard.make_shard(wl.T, wl.S, wl.labels, b0, b1, alignment=wl.alignment)
        a = acts[sh.row0:sh.row1].cpu().numpy()
        o64 = oracle.run(a, sh.labels, sh.T, sh.S, wl.V, blank=wl.blank, alignment=sh.alignment, max_shift=wl.max_shift,
                         precision="f64_from_f32")
        worst_cost = max(worst_cost, float(np.max(np.abs(costs[b0:b1] - o64.costs) / np.abs(o64.costs))))
        g = grads[sh.row0:sh.row1].cpu().numpy()
        step = max(1, (1 << 27) // wl.V)                       # compare in slabs: no second full-size temporary
        for r in range(0, g.shape[0], step):
            worst_grad = max(worst_grad, float(np.abs(g[r:r + step] - o64.grads[r:r + step]).max()))
        del o64, a, g
        b0, groups = b1, groups + 1
    with capsys.disabled():
        print(f"\n[{wl.name} full batch, {wl.B} utterances in {groups} oracle groups] cost rel vs f64 {worst_cost:.2e}, "
              f"grads max|d| vs f64 {worst_grad:.2e}")
    assert worst_cost <= 1e-5, worst_cost
    assert worst_grad <= 1e-5, worst_grad
    del acts, grads
    torch.cuda.empty_cache()


@pytest.mark.parametrize("name", ["c3", "c5", "c4"])
def test_full_batch_parity(env, capsys, name):
    """c3 (64 ragged utterances, 5 GB), c5 (alignment band, 4.7 GB) and c4 (8 utterances of 4.8e8 logits each: 3.9e9 in
    all, more than 2^31 -- 64-bit offsets) in FULL against the double-precision oracle: costs 1e-5 relative, every
    gradient element 1e-5 absolute."""
    mr, _lib = env
    _full_batch_against_f64_oracle(mr, _lib, mr.synth.workload(name), capsys)


@pytest.mark.parametrize("name", ["c3", "c4", "c5"])
def test_full_size_properties(env, name):
    mr, _lib = env
    wl = mr.synth.workload(name)
    acts, grads, costs, ll, meta, _, _ = _run(mr, _lib, wl)
    _properties(grads, costs, ll, meta)
    # cost() alone agrees with cost_and_grad()
    _, _, costs_only, _, _, _, _ = _run(mr, _lib, wl, want_grads=False)
    assert np.array_equal(costs_only, costs)
    del acts, grads
    torch.cuda.empty_cache()


def test_c5_generic_equals_stream(env):
    """The two kernel variants (bulk-copy ring vs direct loads) agree on a ring-wrapping, mostly-dead workload."""
    mr, _lib = env
    wl = mr.synth.workload("c5")
    _, g1, c1, _, _, _, _ = _run(mr, _lib, wl, nbatch=8)
    _, g2, c2, _, _, _, _ = _run(mr, _lib, wl, nbatch=8, generic=True)
    np.testing.assert_allclose(c1, c2, rtol=1e-6)
    assert (g1 - g2).abs().max().item() <= 2e-6
