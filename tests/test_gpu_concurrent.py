"""-m gpu: two loss calls in flight on one device at the same time (two handles, two streams).

The lattice kernel runs several CTAs per utterance that hand work to each other.  Up to round 1 part 0 of an utterance
waited for all of its helper CTAs, i.e. for CTAs that might not have been dispatched yet; with two such launches on the
device that can deadlock (the test in test_gpu_peer.py kept the kernels apart with a sleep).  Now the
rows of the shared phases are handed out in blocks and nobody waits for a CTA that is not running
(include/mrnnt_b200/k2_lattice.cuh), so the calls below may overlap in any way; their results must be the bits of the
same calls made one after the other.
"""
import numpy as np
import pytest
import torch

import fixtures

pytestmark = pytest.mark.gpu


def _bind(case, dev):
    import monotonic_rnnt_b200 as mr
    acts = torch.from_numpy(case.acts.reshape(case.rows, case.V)).to(dev)
    h = mr.LossHandle(acts, torch.from_numpy(case.labels).to(dev), torch.from_numpy(case.T).to(dev),
                      torch.from_numpy(case.S).to(dev), lengths_host=(case.T, case.S))
    return h, acts, torch.empty_like(acts)


@pytest.mark.parametrize("shape", [
    # (B, T, S, V): `parts` CTAs per utterance = min(8, 148 // B); both grids together want more SMs than there are
    ("b24_c2ish", 24, 150, 40, 1000),
    ("b4_long", 4, 400, 80, 256),
    ("b40_small", 40, 60, 12, 512),
], ids=lambda s: s[0])
def test_two_handles_on_two_streams(shape):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    name, B, T, S, V = shape
    dev = torch.device("cuda", 0)
    ca = fixtures.random_case(name + "_a", 501, B=B, V=V, T_range=(T // 2, T), S_range=(S // 2, S), dist="uniform")
    cb = fixtures.random_case(name + "_b", 502, B=B, V=V, T_range=(T // 2, T), S_range=(S // 2, S), dist="normal3")
    ha, aa, ga = _bind(ca, dev)
    hb, ab, gb = _bind(cb, dev)
    # the same calls, one after the other
    want_ca = ha.enqueue(ca.blank, ga).clone()
    torch.cuda.synchronize()
    want_cb = hb.enqueue(cb.blank, gb).clone()
    torch.cuda.synchronize()
    want_ga, want_gb = ga.clone(), gb.clone()
    sa, sb = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    for it in range(25):
        ga.fill_(float("nan"))
        gb.fill_(float("nan"))
        torch.cuda.synchronize()
        with torch.cuda.stream(sa):
            ca_dev = ha.enqueue(ca.blank, ga)
        with torch.cuda.stream(sb):
            cb_dev = hb.enqueue(cb.blank, gb)
            if it % 3 == 0:   # back to back on one stream as well
                cb_dev = hb.enqueue(cb.blank, gb)
        torch.cuda.synchronize()
        assert torch.equal(ca_dev, want_ca) and torch.equal(cb_dev, want_cb), it
        assert torch.equal(ga, want_ga) and torch.equal(gb, want_gb), it
    ha.close()
    hb.close()


def test_forward_halves_back_to_back():
    """Forward halves (K1 + K2 with coefficients, no gradient kernel in between) queued back to back: under programmatic
    dependent launch the next lattice launch can start before the previous one has finished; the hand-over words of the
    two launches must not meet."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    dev = torch.device("cuda", 0)
    case = fixtures.random_case("fwd_b2b", 511, B=6, V=64, T_range=(20, 60), S_range=(4, 20), dist="normal3")
    h, acts, grads = _bind(case, dev)
    h.enqueue(case.blank, grads)
    torch.cuda.synchronize()
    want = grads.clone()
    for it in range(10):
        for _ in range(4):
            costs = h.enqueue_forward(case.blank, want_grads=True)
        grads.fill_(float("nan"))
        h.enqueue_backward(grads)
        torch.cuda.synchronize()
        assert torch.equal(grads, want), it
    h.close()
