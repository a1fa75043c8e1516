"""-m gpu: the drop-in boundary, exercised from the reference's side.

1. tests/cpp/bin/abi_check -- a C++ program that calls compute_rnnt_loss / GpuRNNTWorkspaceManager /
   GpuRNNTComputer exactly like the reference's tests/test_gpu.cu and checks the reference's expected values.
2. tests/dropin/_build/monotonic_rnnt_cpp.so -- the reference's UNMODIFIED PyTorch binding translation unit
   (pytorch_binding/monotonic_rnnt.cu) compiled against THIS repository's include/ (tests/dropin/build_dropin.py,
   built in the authoring container where /root/reference exists).  Its gpu_monotonic_rnnt /
   gpu_monotonic_rnnt_align_restrict entry points are called with CUDA tensors the way the reference's
   monotonic_rnnt_op.py:39-65 calls them, on the fixtures of the reference's pytorch_binding/test.py (which itself
   only ever runs them on CPU tensors).
"""
import importlib.util
import os
import subprocess

import numpy as np
import pytest
import torch

import fixtures
import golden_io

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ABI = os.path.join(ROOT, "tests", "cpp", "bin", "abi_check")
DROPIN = os.path.join(ROOT, "tests", "dropin", "_build", "monotonic_rnnt_cpp.so")


@pytest.fixture(scope="module")
def cuda():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")


def test_cpp_abi_harness(cuda):
    if not os.path.exists(ABI):
        pytest.skip("tests/cpp/bin/abi_check not built (run __graft_entry__.build())")
    res = subprocess.run([ABI], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0 and "ABI OK" in res.stdout, res.stdout + res.stderr


@pytest.fixture(scope="module")
def ref_binding(cuda):
    if not os.path.exists(DROPIN):
        pytest.skip("reference binding not prebuilt (python tests/dropin/build_dropin.py, needs /root/reference)")
    spec = importlib.util.spec_from_file_location("monotonic_rnnt_cpp", DROPIN)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _dev(a, dt):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dt).cuda()


@pytest.mark.parametrize("name", ["readme", "multibatch", "rand_v32", "rand_wide", "infnan"])
def test_unmodified_reference_binding_on_our_headers(ref_binding, name):
    case, ref = golden_io.load(name)
    acts = _dev(case.acts, torch.float32)
    grads = torch.full_like(acts, float("nan"))
    costs = torch.zeros(case.B, dtype=torch.float32)          # host tensor, as monotonic_rnnt_op.py:37
    rc = ref_binding.gpu_monotonic_rnnt(acts, _dev(case.labels, torch.int32), _dev(case.T, torch.int32),
                                        _dev(case.S, torch.int32), costs, grads, case.blank, 0)
    assert rc == 0
    np.testing.assert_allclose(costs.numpy(), ref["costs_f64"], rtol=1e-5)
    assert np.abs(grads.cpu().numpy() - ref["grads_f64"]).max() <= 1e-5


@pytest.mark.parametrize("name", ["align_shift0", "align_shift1", "align_mb_shift1", "rand_v17_shift1",
                                  "rand_wide_shift3"])
def test_unmodified_reference_binding_align_restrict(ref_binding, name):
    case, ref = golden_io.load(name)
    acts = _dev(case.acts, torch.float32)
    grads = torch.full_like(acts, float("nan"))
    costs = torch.zeros(case.B, dtype=torch.float32)
    rc = ref_binding.gpu_monotonic_rnnt_align_restrict(
        acts, _dev(case.labels, torch.int32), _dev(case.T, torch.int32), _dev(case.S, torch.int32),
        _dev(case.alignment, torch.int32), case.max_shift, costs, grads, case.blank, 0)
    assert rc == 0
    np.testing.assert_allclose(costs.numpy(), ref["costs_f64"], rtol=1e-5)
    assert np.abs(grads.cpu().numpy() - ref["grads_f64"]).max() <= 1e-5
    if case.expect_costs is not None:                       # pytorch_binding/test.py: 1.22 / 2.7 at 1e-2
        assert np.all(np.abs(costs.numpy() - case.expect_costs) < 1e-2)


def test_reference_binding_cpu_entry_fails_loudly(ref_binding):
    """The binding's cpu_monotonic_rnnt compiles against our name-only CPU shells and raises: no CPU fallback."""
    case = fixtures.readme_case()
    t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(dt)
    with pytest.raises(RuntimeError):
        ref_binding.cpu_monotonic_rnnt(t(case.acts, torch.float32), t(case.labels, torch.int32),
                                       t(case.T, torch.int32), t(case.S, torch.int32), torch.zeros(1),
                                       torch.zeros(12, 3), 0, 0)


def test_c5_through_the_unmodified_reference_binding(ref_binding):
    """BASELINE configs[4] as it names it: the alignment-restricted shape B=32 T=300 S=60 V=2000 "via PyTorch binding" --
    the reference's own gpu_monotonic_rnnt_align_restrict (pytorch_binding/monotonic_rnnt.cu:116-150), compiled unmodified
    against our include/, on the full c5 batch (4.7 GB of logits).  It must give the bits of this repository's own handle
    on the same device tensors (the same engine behind both faces), every gradient element written (NaN-poisoned buffer,
    95 % of the rows dead), and costs that match the double-precision oracle on a slice."""
    import monotonic_rnnt_b200 as mr
    from monotonic_rnnt_b200 import _lib
    from oracle import oracle
    wl = mr.synth.workload("c5")
    dev = torch.device("cuda", 0)
    acts = torch.empty((wl.rows, wl.V), dtype=torch.float32, device=dev)
    _lib.check(_lib.load().mrnnt_synth_uniform(acts.data_ptr(), wl.elements, wl.logits_seed, 0,
                                               torch.cuda.current_stream().cuda_stream), "synth")
    labels, T, S = _dev(wl.labels, torch.int32), _dev(wl.T, torch.int32), _dev(wl.S, torch.int32)
    al = _dev(wl.alignment, torch.int32)
    grads = torch.full_like(acts, float("nan"))
    costs = torch.zeros(wl.B, dtype=torch.float32)
    for _ in range(2):   # (a new manager and workspace per call, as the binding does it; the second call reuses the block)
        rc = ref_binding.gpu_monotonic_rnnt_align_restrict(acts, labels, T, S, al, wl.max_shift, costs, grads, wl.blank, 0)
        assert rc == 0
    torch.cuda.synchronize()
    assert not torch.isnan(grads).any()
    h = mr.LossHandle(acts, labels, T, S, lengths_host=(wl.T, wl.S))
    h.restrict_to_alignment(al, wl.max_shift, wl.blank)
    g2 = torch.full_like(acts, float("nan"))
    c2 = h.cost_and_grad(wl.blank, g2)
    torch.cuda.synchronize()
    assert torch.equal(c2, costs)
    assert torch.equal(g2, grads)
    h.close()
    # the first two utterances against the double-precision oracle
    nb = 2
    r1 = int(np.sum(wl.T[:nb].astype(np.int64) * (wl.S[:nb] + 1)))
    a_h = acts[:r1].cpu().numpy()
    ref = oracle.run(a_h, wl.labels[:nb], wl.T[:nb], wl.S[:nb], wl.V, blank=wl.blank, alignment=wl.alignment[:nb],
                     max_shift=wl.max_shift, precision="f64_from_f32")
    np.testing.assert_allclose(costs.numpy()[:nb], ref.costs, rtol=1e-5)
    assert np.abs(grads[:r1].cpu().numpy() - ref.grads.reshape(r1, wl.V)).max() <= 1e-5
