"""-m gpu: the batch-sharded path of BASELINE.json configs[2] on the kernels.

A shard is a batch of its own: its rows gathered (or sliced) out of the global packed tensor, labels and alignment
re-strided to the shard's own maxima (the ABI derives both strides from the lengths it is given, reference
cpu_workspace_manager.h:44,117-135,208).  Run one after another on ONE GPU -- exactly what N ranks do side by side --
the shards must reproduce the whole batch's costs and gradients BIT FOR BIT: nothing in the three kernels depends on
which utterances share a batch (tile boundaries, chunk sizes and the number of CTAs per utterance all change with the
batch; the arithmetic per row and per lattice cell does not).
"""
import numpy as np
import pytest
import torch

import fixtures

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mr():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import monotonic_rnnt_b200
    return monotonic_rnnt_b200


def _run(mr, acts, labels, T, S, blank, alignment=None, max_shift=0):
    dev = acts.device
    h = mr.LossHandle(acts, torch.from_numpy(np.ascontiguousarray(labels)).to(dev), torch.from_numpy(T.copy()).to(dev),
                      torch.from_numpy(S.copy()).to(dev), lengths_host=(T, S))
    if alignment is not None:
        h.restrict_to_alignment(torch.from_numpy(np.ascontiguousarray(alignment)).to(dev), max_shift, blank)
    grads = torch.full_like(acts, float("nan"))
    costs = h.cost_and_grad(blank, grads).clone()
    h.close()
    return costs, grads


def _check_sharded_equals_whole(mr, acts, labels, T, S, blank, alignment, max_shift, worlds, kinds=("contiguous", "lpt"),
                                exact=True):
    whole_c, whole_g = _run(mr, acts, labels, T, S, blank, alignment, max_shift)
    assert not torch.isnan(whole_g).any()
    for world in worlds:
        for kind in kinds:
            if kind == "contiguous":
                parts = [np.arange(a, b) for a, b in mr.shard.partition_contiguous(T, S, world)]
            else:
                parts = mr.shard.partition_lpt(T, S, world)
            costs = torch.full((len(T),), float("nan"))
            grads = torch.full_like(whole_g, float("nan"))
            for idx in parts:
                if len(idx) == 0:
                    continue
                sh = mr.shard.make_shard_indexed(T, S, labels, idx, alignment=alignment)
                local = mr.shard.gather_rows(acts, sh).contiguous()
                c, g = _run(mr, local, sh.labels, sh.T, sh.S, blank, sh.alignment, max_shift)
                costs[torch.from_numpy(np.asarray(idx))] = c
                mr.shard.scatter_rows(g, sh, grads)
            if exact:
                assert torch.equal(costs, whole_c), (world, kind)
                assert torch.equal(grads, whole_g), (world, kind)      # every bit, NaN-free (every row written by a shard)
            else:
                assert not torch.isnan(grads).any()
                torch.testing.assert_close(costs, whole_c, rtol=2e-6, atol=0.0)
                assert (grads - whole_g).abs().max().item() <= 2e-6


@pytest.mark.parametrize("restricted", [False, True], ids=["free", "aligned"])
@pytest.mark.parametrize("shape", [("ragged_v1024", 12, 1024, (60, 120), (10, 40)),
                                   ("ragged_v50_unaligned", 9, 50, (20, 60), (0, 20)),
                                   ("wide_states", 5, 256, (100, 140), (70, 100))], ids=lambda s: s[0])
def test_shards_reproduce_the_whole_batch_bit_for_bit(mr, shape, restricted):
    name, B, V, T_range, S_range = shape
    case = fixtures.random_case(name, 601, B=B, V=V, T_range=T_range, S_range=S_range, dist="normal3")
    al, shift = None, 0
    if restricted:
        al = fixtures.random_alignment(np.random.default_rng(602), case.T, case.S, case.labels)
        shift = 3
    acts = torch.from_numpy(case.acts.reshape(case.rows, case.V)).cuda()
    # Rows that are not whole 16-byte vectors (V % 4 != 0) are summed through the aligned vectors that cover them: which
    # lane adds which logit depends on where the row starts in its 16 bytes, i.e. on the row's position in the tensor,
    # so a shard agrees with the whole batch to float rounding there, not bit for bit.
    _check_sharded_equals_whole(mr, acts, case.labels, case.T, case.S, case.blank, al, shift, worlds=(2, 3, 4),
                                exact=(V % 4 == 0))


def test_c3_full_batch_sharded_over_8(mr):
    """BASELINE.json configs[2] itself: B=64 ragged utterances, V=1024, 5 GB of logits, cut for 2 and 8 ranks."""
    from monotonic_rnnt_b200 import _lib
    wl = mr.synth.workload("c3")
    acts = torch.empty((wl.rows, wl.V), dtype=torch.float32, device="cuda")
    _lib.check(_lib.load().mrnnt_synth_uniform(acts.data_ptr(), wl.elements, wl.logits_seed, 0,
                                               torch.cuda.current_stream().cuda_stream), "synth")
    _check_sharded_equals_whole(mr, acts, wl.labels, wl.T, wl.S, wl.blank, None, 0, worlds=(2, 8), kinds=("lpt",))
    _check_sharded_equals_whole(mr, acts, wl.labels, wl.T, wl.S, wl.blank, None, 0, worlds=(8,), kinds=("contiguous",))
