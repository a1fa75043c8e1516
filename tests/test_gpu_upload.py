"""-m gpu: host -> device upload of the live rows only (mrnnt_upload_acts, plan.cuh::upload_live_rows_kernel).

The rows the plan calls dead are never read by any kernel, so they need not cross the bus: the device array is
poisoned with NaN, filled from pinned host memory through the handle, and the call that follows must give the bits
of a call on a plain device copy of the same logits -- while the dead rows still hold the poison.
"""
import ctypes

import numpy as np
import pytest
import torch

import fixtures
from gpu_util import to_dev

pytestmark = pytest.mark.gpu


def _cases():
    rng = np.random.default_rng(77)
    ragged = fixtures.random_case("up_ragged", 21, B=5, V=1000, T_range=(20, 40), S_range=(3, 12), dist="uniform")
    banded = ragged.with_alignment(fixtures.random_alignment(rng, ragged.T, ragged.S, ragged.labels), 2, "up_banded")
    odd = fixtures.random_case("up_v37", 22, B=3, V=37, T_range=(5, 20), S_range=(0, 6))   # 148-byte rows: 4-byte units
    return [ragged, banded, odd]


def _run(case, acts, dtype, upload_from=None, copy_engine=None, lengths_host=True):
    import monotonic_rnnt_b200 as mr
    h = mr.LossHandle(acts, to_dev(case.labels, torch.int32), to_dev(case.T, torch.int32), to_dev(case.S, torch.int32),
                      lengths_host=(case.T, case.S) if lengths_host else None)
    if case.alignment is not None:
        h.restrict_to_alignment(to_dev(case.alignment, torch.int32), case.max_shift, case.blank)
    if copy_engine is not None:
        h.set_option(mr._lib.OPT_UPLOAD_COPY_ENGINE, copy_engine)
    if upload_from is not None:
        h.upload_acts(upload_from)
    grads = torch.full_like(acts, float("nan"))
    costs = h.cost_and_grad(case.blank, grads).numpy().copy()
    rowmeta = h.debug(mr._lib.DBG_ROWMETA)
    h.close()
    return costs, grads.float().cpu().numpy(), np.asarray(rowmeta)


@pytest.mark.parametrize("case", _cases(), ids=lambda c: c.name)
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
def test_upload_gives_the_bits_of_a_device_copy(case, dtype):
    host = torch.from_numpy(case.acts.reshape(case.rows, case.V)).to(dtype).contiguous().pin_memory()
    want_costs, want_grads, rowmeta = _run(case, host.cuda(), dtype)
    poisoned = torch.full((case.rows, case.V), float("nan"), dtype=dtype, device="cuda")
    got_costs, got_grads, _ = _run(case, poisoned, dtype, upload_from=host)
    np.testing.assert_array_equal(got_costs, want_costs)
    np.testing.assert_array_equal(got_grads, want_grads)
    dead = rowmeta == -2
    assert dead.any() and not dead.all()
    dev = poisoned.float().cpu().numpy()
    assert np.isnan(dev[dead]).all(), "dead rows must not have crossed the bus"
    np.testing.assert_array_equal(dev[~dead], host.float().numpy()[~dead])


@pytest.mark.parametrize("copy_engine", [0, 1, 40000, -1], ids=["kernel_only", "any_block", "blocks_of_40kB", "default_1MiB"])
@pytest.mark.parametrize("lengths_host", [True, False], ids=["host_lengths", "fetched_lengths"])
def test_middle_blocks_through_the_copy_engine(copy_engine, lengths_host):
    """The all-live block in the middle of every utterance goes through the copy engine on a side stream while the
    kernel brings the ragged frames around it (Engine::upload_live_rows, plan.cuh::upload_middle_block): whoever brings
    a row, the result is the bits of a plain device copy, the dead rows never cross, and the call behind the upload
    waits for both.  T = S, T < 2 S (no middle block), S = 0 (all of the utterance is one) and a block big enough for
    the default threshold are all in the batch."""
    case = fixtures.random_case("up_mid", 31, B=6, V=128, dist="uniform",
                                force=[(30, 4), (12, 12), (12, 8), (40, 0), (700, 3), (9, 2)])   # 700 x 4 rows x 512 B: 1.4 MB
    host = torch.from_numpy(case.acts.reshape(case.rows, case.V)).contiguous().pin_memory()
    want_costs, want_grads, rowmeta = _run(case, host.cuda(), torch.float32)
    for _ in range(2):   # (twice: the side stream's events are reused)
        poisoned = torch.full((case.rows, case.V), float("nan"), dtype=torch.float32, device="cuda")
        got_costs, got_grads, _ = _run(case, poisoned, torch.float32, upload_from=host, copy_engine=copy_engine,
                                       lengths_host=lengths_host)
        np.testing.assert_array_equal(got_costs, want_costs)
        np.testing.assert_array_equal(got_grads, want_grads)
        dead = rowmeta == -2
        dev = poisoned.cpu().numpy()
        assert np.isnan(dev[dead]).all(), "dead rows must not have crossed the bus"
        np.testing.assert_array_equal(dev[~dead], host.numpy()[~dead])


def test_pageable_memory_is_refused():
    import monotonic_rnnt_b200 as mr
    case = fixtures.readme_case()
    acts = to_dev(case.acts.reshape(case.rows, case.V), torch.float32)
    h = mr.LossHandle(acts, to_dev(case.labels, torch.int32), to_dev(case.T, torch.int32), to_dev(case.S, torch.int32))
    pageable = np.ascontiguousarray(case.acts.reshape(case.rows, case.V), dtype=np.float32)
    with pytest.raises(TypeError):
        h.upload_acts(torch.from_numpy(pageable))
    lib = mr._lib.load()
    st = lib.mrnnt_upload_acts(h._h, ctypes.c_void_p(pageable.ctypes.data),
                               ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert st == 2, st   # RNNT_STATUS_INVALID_VALUE (status.h)
    assert lib.mrnnt_upload_acts(h._h, None, None) == 2
    # the handle still works
    costs = h.cost_and_grad(case.blank, torch.empty_like(acts)).numpy()
    np.testing.assert_allclose(costs, [1.01335239], rtol=1e-5)
    h.close()
