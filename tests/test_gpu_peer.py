"""-m gpu: the all-GPU sum of the summed cost over peer memory (include/mrnnt_b200/peer_reduce.cuh).

One device suffices for the protocol: the "ranks" are handles of one process whose boards are plain device
allocations (monotonic_rnnt_b200.peer.PeerBoards.local).  The CUDA-IPC mapping between processes is exercised by
`bench.py --gpus N` (it checks the exchanged sum against an NCCL all-reduce of the same costs) and by
test_two_processes below where the box has two GPUs.
"""
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


def _handle(case, dev):
    import monotonic_rnnt_b200 as mr
    acts = torch.from_numpy(case.acts.reshape(case.rows, case.V)).to(dev)
    h = mr.LossHandle(acts, torch.from_numpy(case.labels).to(dev), torch.from_numpy(case.T).to(dev),
                      torch.from_numpy(case.S).to(dev))
    return h, acts


def _oracle_costs(case):
    return oracle.run(case.acts, case.labels, case.T, case.S, case.V, blank=case.blank, precision="f64_from_f32").costs


@pytest.mark.parametrize("V", [1000, 37], ids=["stream", "generic"])
def test_world_of_one(dev, V):
    """world = 1: the total is this rank's own sum -- with gradients (inside the gradient kernel), cost only and as a
    forward half (one-warp launch); the epoch advances by one per exchange; a new handle takes the boards over."""
    import monotonic_rnnt_b200 as mr
    case = fixtures.random_case("peer1", 77, B=5, V=V, T_range=(12, 30), S_range=(2, 9), dist="uniform")
    want = float(np.sum(_oracle_costs(case)))
    (boards,) = mr.peer.PeerBoards.local(1, dev)
    total = torch.full((1,), float("nan"), device=dev)
    h, acts = _handle(case, dev)
    h.set_peer_reduce(boards, total)
    g = torch.empty_like(acts)
    for i in range(3):
        total.fill_(float("nan"))
        costs = h.cost_and_grad(case.blank, g).numpy()
        assert abs(float(total.item()) - want) <= 1e-5 * abs(want)
        assert abs(float(total.item()) - float(costs.sum(dtype=np.float64))) <= 1e-5 * abs(want)
    total.fill_(float("nan"))
    h.cost(case.blank)
    assert abs(float(total.item()) - want) <= 1e-5 * abs(want)
    total.fill_(float("nan"))
    h.enqueue_forward(case.blank, want_grads=True)
    h.enqueue_backward(g)          # (a backward half on its own exchanges nothing)
    torch.cuda.synchronize()
    assert abs(float(total.item()) - want) <= 1e-5 * abs(want)
    h.sync_peer_epoch()
    assert boards.epoch == 5
    h.close()
    h2, acts2 = _handle(case, dev)  # a new handle continues at the boards' epoch
    h2.set_peer_reduce(boards, total)
    total.fill_(float("nan"))
    h2.cost_and_grad(case.blank, torch.empty_like(acts2))
    assert abs(float(total.item()) - want) <= 1e-5 * abs(want)
    h2.sync_peer_epoch()
    assert boards.epoch == 6
    # pinned host memory as the destination (what bench.py uses)
    total_h = torch.full((1,), float("nan")).pin_memory()
    h2.set_peer_reduce(boards, total_h)
    h2.cost_and_grad(case.blank, torch.empty_like(acts2))
    assert abs(float(total_h.item()) - want) <= 1e-5 * abs(want)
    h2.set_peer_reduce(None, None)
    h2.close()
    boards.close()


def test_two_ranks_on_one_device(dev):
    """world = 2 inside one process: rank 1 publishes from a forward half (one-warp launch) on its own stream and
    waits there for rank 0, whose exchange rides in its gradient kernel.  Both see sum_0 + sum_1, bit-identical."""
    import monotonic_rnnt_b200 as mr
    c0 = fixtures.random_case("peer2a", 5, B=4, V=1000, T_range=(20, 40), S_range=(3, 10), dist="uniform")
    c1 = fixtures.random_case("peer2b", 6, B=3, V=1000, T_range=(15, 25), S_range=(2, 8), dist="uniform")
    want = float(np.sum(_oracle_costs(c0)) + np.sum(_oracle_costs(c1)))
    b0, b1 = mr.peer.PeerBoards.local(2, dev)
    t0 = torch.full((1,), float("nan"), device=dev)
    t1 = torch.full((1,), float("nan"), device=dev)
    h0, a0 = _handle(c0, dev)
    h1, a1 = _handle(c1, dev)
    h0.set_peer_reduce(b0, t0)
    h1.set_peer_reduce(b1, t1)
    g0 = torch.empty_like(a0)
    s1 = torch.cuda.Stream(device=dev)
    torch.cuda.synchronize()
    for _ in range(4):   # (both parities of the slots, twice)
        t0.fill_(float("nan"))
        t1.fill_(float("nan"))
        torch.cuda.synchronize()
        with torch.cuda.stream(s1):
            h1.enqueue_forward(c1.blank, want_grads=False)   # publishes, then waits for rank 0 in a one-warp kernel
        # (no pause: the two ranks' lattice kernels may share the device -- no CTA of either waits for a CTA that has
        # not been dispatched, tests/test_gpu_concurrent.py)
        h0.enqueue(c0.blank, g0)                              # publishes at the start of K3, collects at its end
        torch.cuda.synchronize()
        assert abs(float(t0.item()) - want) <= 1e-5 * abs(want)
        assert float(t0.item()) == float(t1.item())
    h0.close()
    h1.close()
    b0.close()
    b1.close()


def test_peer_that_never_shows_up_fails_for_good(dev):
    """world = 2, but rank 1 never runs: rank 0's collect gives up after its time-out (set to 50 ms here; 60 s by
    default), leaves NaN, and from then on the handle refuses to take part -- it must not publish epochs its peer has
    not collected (the slot invariant is gone once a rank has moved on alone)."""
    import monotonic_rnnt_b200 as mr
    from monotonic_rnnt_b200 import _lib
    c0 = fixtures.random_case("peer_to", 7, B=4, V=1000, T_range=(20, 40), S_range=(3, 10), dist="uniform")
    b0, b1 = mr.peer.PeerBoards.local(2, dev)
    total = torch.zeros(1).pin_memory()
    h0, a0 = _handle(c0, dev)
    h0.set_peer_reduce(b0, total)
    h0.set_peer_timeout_ms(50)
    g0 = torch.empty_like(a0)
    with pytest.raises(_lib.RNNTError) as e:
        h0.cost_and_grad(c0.blank, g0)
    assert e.value.status == 3                       # RNNT_STATUS_EXECUTION_FAILED
    assert np.isnan(float(total.item()))
    assert h0.peer_failed()
    # the gradients of the call itself are complete: only the exchange failed
    ref = oracle.run(c0.acts, c0.labels, c0.T, c0.S, c0.V, blank=c0.blank, precision="f64_from_f32")
    assert np.abs(g0.cpu().numpy() - ref.grads).max() <= 1e-5
    with pytest.raises(_lib.RNNTError):              # final: no further exchange through this handle
        h0.cost_and_grad(c0.blank, g0)
    # without the exchange the handle works again
    h0.set_peer_reduce(None, None)
    costs = h0.cost_and_grad(c0.blank, g0)
    np.testing.assert_allclose(costs.numpy(), ref.costs, rtol=1e-5)
    h0.close()
    b0.close()
    b1.close()


def _two_process_worker(rank, port, out):
    import os
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE="2")
    import torch.distributed as dist
    import monotonic_rnnt_b200 as mr
    torch.cuda.set_device(rank)
    d = torch.device("cuda", rank)
    dist.init_process_group("nccl", device_id=d)
    case = fixtures.random_case(f"peer_mp{rank}", 100 + rank, B=4, V=1000, T_range=(20, 40), S_range=(3, 10), dist="uniform")
    boards = mr.peer.PeerBoards(device=d)
    total = torch.full((1,), float("nan")).pin_memory()
    h, acts = _handle(case, d)
    h.set_peer_reduce(boards, total)
    g = torch.empty_like(acts)
    for _ in range(3):
        costs = h.cost_and_grad(case.blank, g)
    want = torch.tensor([float(costs.double().sum())], dtype=torch.float64, device=d)
    dist.all_reduce(want)
    out[rank] = (float(total.item()), float(want.item()))
    h.close()
    boards.close()
    dist.destroy_process_group()


def test_two_processes(dev):
    """The real thing where two GPUs are there: one process per GPU, boards mapped with CUDA IPC."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Manager().dict()
    procs = [ctx.Process(target=_two_process_worker, args=(r, 29641, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    assert out[0][0] == out[1][0]
    assert abs(out[0][0] - out[0][1]) <= 1e-5 * abs(out[0][1])
