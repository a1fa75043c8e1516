"""-m gpu: a synchronous call with gradients returns when the COSTS are on the host (MRNNT_OPT_RETURN_EARLY, the
default); the gradient kernel completes in stream order.  The reference's contract is costs on the host on return,
gradients in the caller's device buffer (gpu_rnnt.h:229-232); both of its bindings consume the gradients on the stream
the call was made on.  What must hold: the same bits as with the whole wait, costs valid the moment the call returns,
calls back to back on one stream or on alternating streams, a handle destroyed while its gradient kernel runs."""
import numpy as np
import pytest
import torch

import fixtures
from oracle import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda", 0)


def _bind(case, dev):
    import monotonic_rnnt_b200 as mr
    acts = torch.from_numpy(case.acts.reshape(case.rows, case.V)).to(dev)
    h = mr.LossHandle(acts, torch.from_numpy(case.labels).to(dev), torch.from_numpy(case.T).to(dev),
                      torch.from_numpy(case.S).to(dev), lengths_host=(case.T, case.S))
    return h, acts


def _case(seed=11, B=16, V=1000):
    return fixtures.random_case(f"early{seed}", seed, B=B, V=V, T_range=(60, 150), S_range=(10, 40), dist="uniform")


def test_same_bits_as_the_whole_wait(dev):
    from monotonic_rnnt_b200 import _lib
    case = _case()
    h, acts = _bind(case, dev)
    assert h.get_option(_lib.OPT_RETURN_EARLY) == 1
    g_full = torch.full_like(acts, float("nan"))
    h.set_option(_lib.OPT_RETURN_EARLY, 0)
    assert h.get_option(_lib.OPT_RETURN_EARLY) == 0
    c_full = h.cost_and_grad(case.blank, g_full).clone()
    torch.cuda.synchronize()
    h.set_option(_lib.OPT_RETURN_EARLY, 1)
    ref = oracle.run(case.acts, case.labels, case.T, case.S, case.V, blank=case.blank, precision="f64_from_f32")
    np.testing.assert_allclose(c_full.numpy(), ref.costs, rtol=1e-5)
    for it in range(6):
        g = torch.full_like(acts, float("nan"))
        costs = torch.full((case.B,), float("nan")).pin_memory()
        h.cost_and_grad(case.blank, g, costs)
        # the costs are there NOW, whatever the gradient kernel is doing
        assert torch.equal(costs, c_full), it
        assert torch.equal(g, g_full), it      # (a torch op on the same stream: stream order)
    h.close()


def test_returns_while_the_gradient_kernel_runs(dev):
    """On a c2-sized batch the gradient kernel takes ~0.2 ms: the stream must still be busy right after the return
    at least once in a few tries, and never with the whole wait."""
    from monotonic_rnnt_b200 import _lib
    case = fixtures.random_case("early_big", 5, B=32, V=1000, T_range=(150, 150), S_range=(40, 40), dist="uniform")
    h, acts = _bind(case, dev)
    g = torch.empty_like(acts)
    st = torch.cuda.current_stream()
    h.cost_and_grad(case.blank, g)
    torch.cuda.synchronize()
    busy = 0
    for _ in range(10):
        h.cost_and_grad(case.blank, g)
        busy += 0 if st.query() else 1
        torch.cuda.synchronize()
    assert busy >= 1
    h.set_option(_lib.OPT_RETURN_EARLY, 0)
    for _ in range(5):
        h.cost_and_grad(case.blank, g)
        assert st.query()
    h.close()


def test_back_to_back_and_alternating_streams(dev):
    case = _case(seed=12)
    h, acts = _bind(case, dev)
    want_g = torch.empty_like(acts)
    want_c = h.cost_and_grad(case.blank, want_g).clone()
    torch.cuda.synchronize()
    sa, sb = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    bufs = [torch.empty_like(acts) for _ in range(4)]
    for rep in range(5):
        for b in bufs:
            b.fill_(float("nan"))
        torch.cuda.synchronize()
        got = []
        for i, b in enumerate(bufs):   # one handle, one workspace: the second stream's kernels go behind the first's
            with torch.cuda.stream(sa if i % 2 == 0 else sb):
                got.append(h.cost_and_grad(case.blank, b).clone())
        torch.cuda.synchronize()
        for c, b in zip(got, bufs):
            assert torch.equal(c, want_c)
            assert torch.equal(b, want_g)
    h.close()


def test_handle_destroyed_while_its_gradient_kernel_runs(dev):
    case = _case(seed=13)
    h, acts = _bind(case, dev)
    want_g = torch.empty_like(acts)
    want_c = h.cost_and_grad(case.blank, want_g).clone()
    torch.cuda.synchronize()
    h.close()
    for _ in range(5):
        h, acts2 = _bind(case, dev)
        g = torch.full_like(acts2, float("nan"))
        c = h.cost_and_grad(case.blank, g)
        h.close()                         # (waits for what the call left in flight: its workspace goes back to torch)
        del h
        assert torch.equal(c, want_c)
        assert torch.equal(g, want_g)


def test_cost_only_and_generic_kernels(dev):
    """No gradient kernel, or the generic one (V = 37: no streaming variant): the same contract."""
    from monotonic_rnnt_b200 import _lib
    case = fixtures.random_case("early_gen", 21, B=6, V=37, T_range=(12, 30), S_range=(2, 9), dist="uniform")
    ref = oracle.run(case.acts, case.labels, case.T, case.S, case.V, blank=case.blank, precision="f64_from_f32")
    h, acts = _bind(case, dev)
    h.set_option(_lib.OPT_FORCE_GENERIC, 1)
    for _ in range(3):
        g = torch.full_like(acts, float("nan"))
        c = h.cost_and_grad(case.blank, g)
        np.testing.assert_allclose(c.numpy(), ref.costs, rtol=1e-5)
        assert np.abs(g.cpu().numpy().reshape(-1) - ref.grads.reshape(-1)).max() <= 1e-5
        np.testing.assert_allclose(h.cost(case.blank).numpy(), ref.costs, rtol=1e-5)
    h.close()


def test_peer_reduce_keeps_the_whole_wait_unless_asked(dev):
    """With a peer reduce the world's sum is promised on return (pinned host destination): mode 1 waits for it, mode 2
    returns early and the sum is valid behind a synchronisation."""
    import monotonic_rnnt_b200 as mr
    from monotonic_rnnt_b200 import _lib
    case = _case(seed=14, B=5)
    ref = oracle.run(case.acts, case.labels, case.T, case.S, case.V, blank=case.blank, precision="f64_from_f32")
    want = float(np.sum(ref.costs))
    (boards,) = mr.peer.PeerBoards.local(1, dev)
    total = torch.full((1,), float("nan")).pin_memory()
    h, acts = _bind(case, dev)
    h.set_peer_reduce(boards, total)
    g = torch.empty_like(acts)
    for _ in range(3):
        total.fill_(float("nan"))
        h.cost_and_grad(case.blank, g)
        assert abs(float(total.item()) - want) <= 1e-5 * abs(want)      # on return
    h.set_option(_lib.OPT_RETURN_EARLY, 2)
    for _ in range(3):
        total.fill_(float("nan"))
        c = h.cost_and_grad(case.blank, g)
        np.testing.assert_allclose(c.numpy(), ref.costs, rtol=1e-5)
        torch.cuda.synchronize()
        assert abs(float(total.item()) - want) <= 1e-5 * abs(want)      # in stream order
    h.set_peer_reduce(None, None)
    h.close()
    boards.close()


def test_owned_workspace_goes_back_to_the_cache_while_the_gradient_kernel_runs(dev):
    """create_workspace -> cost_and_grad (early return) -> free_workspace -> destroy, as the reference's torch binding does
    it per call: free_workspace does not wait for the gradient kernel, the block goes back to the cache with the event
    behind it, and the NEXT handle -- other inputs, ANOTHER stream -- that takes the block orders its kernels behind that
    event.  Without the ordering the second handle's LSE kernel would overwrite the per-row records the first handle's
    gradient kernel is still reading."""
    import ctypes
    from monotonic_rnnt_b200 import _lib
    lib = _lib.load()
    ca = fixtures.random_case("recycle_a", 71, B=32, V=1000, T_range=(150, 150), S_range=(40, 40), dist="uniform")
    cb = fixtures.random_case("recycle_b", 72, B=32, V=1000, T_range=(150, 150), S_range=(40, 40), dist="normal3")
    sa, sb = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

    def tensors(c):
        return (torch.from_numpy(c.acts.reshape(c.rows, c.V)).to(dev), torch.from_numpy(c.labels).to(dev),
                torch.from_numpy(c.T).to(dev), torch.from_numpy(c.S).to(dev))

    ta, tb = tensors(ca), tensors(cb)

    def one_call(c, t, stream, grads):
        h = ctypes.c_void_p()
        Th = np.ascontiguousarray(c.T, dtype=np.int32); Sh = np.ascontiguousarray(c.S, dtype=np.int32)
        _lib.check(lib.mrnnt_create(ctypes.byref(h), t[0].data_ptr(), t[1].data_ptr(), c.B, t[2].data_ptr(), t[3].data_ptr(),
                                    c.V, Th.ctypes.data, Sh.ctypes.data), "create")      # (host lengths: no blocking fetch)
        _lib.check(lib.mrnnt_create_workspace(h), "create_workspace")
        costs = torch.empty(c.B, dtype=torch.float32)
        _lib.check(lib.mrnnt_cost_and_grad(h, c.blank, stream.cuda_stream, costs.data_ptr(), grads.data_ptr()), "call")
        busy = not stream.query()
        lib.mrnnt_free_workspace(h)
        lib.mrnnt_destroy(h)
        return costs, busy

    torch.cuda.synchronize()
    ga, gb = torch.empty_like(ta[0]), torch.empty_like(tb[0])
    want_ca, _ = one_call(ca, ta, sa, ga)
    torch.cuda.synchronize()
    want_cb, _ = one_call(cb, tb, sb, gb)
    torch.cuda.synchronize()
    want_ga, want_gb = ga.clone(), gb.clone()
    oa = oracle.run(ca.acts, ca.labels, ca.T, ca.S, ca.V, blank=ca.blank, precision="f64_from_f32")
    np.testing.assert_allclose(want_ca.numpy(), oa.costs, rtol=1e-5)
    assert np.abs(want_ga.cpu().numpy().reshape(-1) - oa.grads.reshape(-1)).max() <= 1e-5
    ran_ahead = 0
    for it in range(12):
        ga.fill_(float("nan")); gb.fill_(float("nan"))
        torch.cuda.synchronize()
        c1, busy1 = one_call(ca, ta, sa, ga)          # returns while its gradient kernel runs; its block goes to the cache
        c2, busy2 = one_call(cb, tb, sb, gb)          # takes that block, on another stream
        ran_ahead += int(busy1)
        torch.cuda.synchronize()
        assert torch.equal(c1, want_ca) and torch.equal(c2, want_cb), it
        assert torch.equal(ga, want_ga), it
        assert torch.equal(gb, want_gb), it
    assert ran_ahead >= 1      # (free_workspace really did not wait)
