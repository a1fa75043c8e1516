"""-m "not gpu": host-side logic -- the C-ABI library loads without a GPU and exports every declared symbol,
size/validation queries, the no-CPU-fallback behaviour, sharding, synthetic generators, the gloo path."""
import ctypes
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

import fixtures
import monotonic_rnnt_b200 as mr
from monotonic_rnnt_b200 import _lib
from oracle import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    names = set()
    for header in ("mrnnt_c_api.h", "rnnt_entrypoint.h"):
        text = open(os.path.join(ROOT, "include", header)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        text = re.sub(r"//[^\n]*", "", text)
        names |= set(re.findall(r"\b(mrnnt_\w+|rnnt_loss_grad_gpu|compute_rnnt_loss|get_workspace_size)\s*\(", text))
    return names


def test_library_loads_and_exports_every_declared_symbol():
    lib = _lib.load()
    declared = _declared_functions()
    assert {"compute_rnnt_loss", "mrnnt_cost_and_grad", "rnnt_loss_grad_gpu"} <= declared
    assert declared == set(_lib.EXPORTED_SYMBOLS)
    for name in declared:
        assert getattr(lib, name) is not None
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.lib_path()], capture_output=True, text=True).stdout
    for name in declared:
        assert re.search(rf"\bT {name}\b", out), name
    assert b"sm_100a" in lib.mrnnt_build_info()


def test_cubin_is_sm100a_with_bulk_copies():
    """The library carries sm_100a SASS and the streaming kernels really use the TMA engine (UBLKCP)."""
    res = subprocess.run(["cuobjdump", "-sass", _lib.lib_path()], capture_output=True, text=True)
    if res.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in res.stdout
    assert "UBLKCP" in res.stdout and "SYNCS" in res.stdout     # cp.async.bulk + mbarrier
    assert "REDUX" in res.stdout and "MUFU.EX2" in res.stdout


def test_workspace_size_query_and_validation():
    lib = _lib.load()
    size = mr.workspace_size([4], [2], 3)
    assert size > 12 * 56
    big = mr.workspace_size([150] * 32, [40] * 32, 1000)
    assert big > 32 * 150 * 41 * 56 and big == mr.workspace_size([150] * 32, [40] * 32, 5000)   # V does not enter
    for T, S in (([0], [0]), ([2], [3]), ([4], [-1]), ([], [])):
        with pytest.raises(mr.RNNTError) as e:
            mr.workspace_size(T, S, 3)
        assert e.value.status == 2                                      # RNNT_STATUS_INVALID_VALUE
    assert mr.workspace_size([3], [3], 5) > 0                           # T == S is legal
    out = ctypes.c_size_t(0)
    assert lib.mrnnt_get_workspace_size(None, None, 1, 3, ctypes.byref(out)) == 2
    h = ctypes.c_void_p()
    assert lib.mrnnt_create(ctypes.byref(h), None, None, 0, None, None, 3, None, None) == 2


def test_padded_layout_size_query_and_validation():
    """mrnnt_get_workspace_size_padded: the tensor dimensions must cover every utterance."""
    lib = _lib.load()
    T = np.array([5, 9, 7], np.int32); S = np.array([2, 4, 0], np.int32)
    out = ctypes.c_size_t(0)

    def q(T_dim, U, stride):
        return lib.mrnnt_get_workspace_size_padded(T.ctypes.data, S.ctypes.data, 3, 11, T_dim, U, stride, ctypes.byref(out))

    assert q(9, 5, 4) == 0
    exact = out.value
    assert exact > 3 * 9 * 5 * 56
    assert q(12, 8, 6) == 0 and out.value > exact                   # sized by the padded extent
    assert q(8, 5, 4) == 2 and q(9, 4, 4) == 2 and q(9, 5, 3) == 2  # too few frames / states / label columns
    assert q(9, 5, 0) == 0                                          # label stride 0: the reference rule, max S
    assert q(0, 5, 4) == 2 and q(9, 0, 4) == 2
    h = ctypes.c_void_p()
    assert lib.mrnnt_create_padded(ctypes.byref(h), None, None, 3, None, None, 11, 8, 5, 4, T.ctypes.data,
                                   S.ctypes.data) == 2              # validated at creation when lengths are given


def test_no_cpu_fallback():
    c = fixtures.readme_case()
    t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(dt)
    with pytest.raises(RuntimeError, match="GPU-only"):
        mr.monotonic_rnnt_loss(t(c.acts, torch.float32), t(c.labels, torch.int32), t(c.T, torch.int32),
                               t(c.S, torch.int32))


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(mr.build, "LIB_PATH", str(tmp_path / "nope.so"))
    monkeypatch.setattr(mr.build, "up_to_date", lambda: False)

    def boom(*a, **k):
        raise RuntimeError("nvcc not found")
    monkeypatch.setattr(mr.build, "build", boom)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.load()


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under the package or include/ imports, links or dlopens it."""
    bad = re.compile(r"^\s*(from|import)\s+oracle\b|liboracle|libmrnnt_ref|rnnt_oracle|oracle/_ref", re.M)
    for top in ("monotonic-rnnt_b200", "include"):
        for d, _, files in os.walk(os.path.join(ROOT, top)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h")):
                    assert not bad.search(open(os.path.join(d, f)).read()), os.path.join(d, f)


# ---- synthetic generators -----------------------------------------------------------------------------
def test_synth_generator():
    a = mr.synth.uniform_logits(1000, seed=0)
    assert a.dtype == np.float32 and a.min() >= 0.0 and a.max() < 1.0
    assert np.array_equal(mr.synth.uniform_logits(300, 0, 700), a[700:])       # counter based: offsets compose
    assert not np.array_equal(mr.synth.uniform_logits(1000, seed=1), a)
    assert abs(float(a.mean()) - 0.5) < 0.05
    lab = mr.synth.labels_for(8, 40, 1000)
    assert lab.min() >= 1 and lab.max() <= 999                                  # never blank
    for name, rows in (("c2", 196800), ("c4", 774400), ("c5", 585600)):
        wl = mr.synth.workload(name)
        assert wl.rows == rows and wl.algorithmic_bytes == 12 * rows * wl.V
    wl = mr.synth.workload("c3")
    assert wl.T[0] == 400 and wl.S[0] == 80 and (wl.T >= wl.S).all() and wl.T.max() <= 400
    wl = mr.synth.workload("c5")
    assert ((wl.alignment != 0).sum(axis=1) == wl.S).all()                      # a valid alignment per utterance


# ---- sharding -----------------------------------------------------------------------------------------
@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_partition_covers_and_balances(world):
    wl = mr.synth.workload("c3")
    parts = mr.shard.partition_contiguous(wl.T, wl.S, world)
    assert len(parts) == world and parts[0][0] == 0 and parts[-1][1] == wl.B
    assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
    w = wl.T.astype(np.int64) * (wl.S + 1)
    loads = [int(w[a:b].sum()) for a, b in parts]
    assert min(loads) > 0 and max(loads) <= 1.10 * (sum(loads) / world)           # the best contiguous cut (c3 at 8: 1.068)
    # ... and it IS the best one: no contiguous cut into `world` ranges has a smaller maximum (brute force at world <= 3)
    if world in (2, 3):
        import itertools
        cum = np.concatenate([[0], np.cumsum(w)])
        best = min(max(int(cum[b] - cum[a]) for a, b in zip((0,) + c, c + (wl.B,)))
                   for c in itertools.combinations(range(1, wl.B), world - 1))
        assert max(loads) == best
    fixed = mr.shard.partition_contiguous([150] * 32, [40] * 32, world)
    if 32 % world == 0:
        assert all(b - a == 32 // world for a, b in fixed)
    few = mr.shard.partition_contiguous([5, 6], [1, 2], 4)                      # B < world: empty tails
    assert sum(b - a for a, b in few) == 2


@pytest.mark.parametrize("world", [2, 4, 8])
def test_lpt_assignment_balances_better_than_any_contiguous_cut(world):
    wl = mr.synth.workload("c3")
    parts = mr.shard.partition_lpt(wl.T, wl.S, world)
    assert sorted(np.concatenate(parts).tolist()) == list(range(wl.B))            # every utterance exactly once
    assert all((np.diff(p) > 0).all() for p in parts)                             # ascending inside a rank
    lpt = mr.shard.imbalance(wl.T, wl.S, parts)
    contiguous = mr.shard.imbalance(wl.T, wl.S, mr.shard.partition_contiguous(wl.T, wl.S, world))
    assert lpt <= 1.01 and lpt <= contiguous
    assert mr.shard.partition_lpt([5, 6], [1, 2], 4)[2].size == 0                 # B < world: empty ranks


def test_indexed_shards_equal_whole_batch():
    """The same for shards that are NOT contiguous (LPT assignment): gather the utterances' rows, run, scatter back."""
    case = fixtures.random_case("shardlpt", 78, B=9, V=13, T_range=(4, 25), S_range=(0, 9))
    al = fixtures.random_alignment(np.random.default_rng(4), case.T, case.S, case.labels)
    whole = oracle.run(case.acts, case.labels, case.T, case.S, case.V, alignment=al, max_shift=2)
    for world in (2, 4):
        costs = np.full(case.B, np.nan, np.float32)
        grads = np.full_like(whole.grads, np.nan)
        for idx in mr.shard.partition_lpt(case.T, case.S, world):
            sh = mr.shard.make_shard_indexed(case.T, case.S, case.labels, idx, alignment=al)
            local = mr.shard.gather_rows(case.acts, sh)
            assert local.shape[0] == sh.rows
            r = oracle.run(local, sh.labels, sh.T, sh.S, case.V, alignment=sh.alignment, max_shift=2)
            costs[idx] = r.costs
            mr.shard.scatter_rows(r.grads, sh, grads)
        assert np.array_equal(costs, whole.costs) and np.array_equal(grads, whole.grads)


def test_sharded_results_equal_whole_batch():
    """Utterances are independent: running each shard on its own (labels / alignment re-strided to the shard's
    own maxima, as the ABI requires) reproduces the whole-batch costs and gradients bit for bit."""
    case = fixtures.random_case("shardme", 77, B=7, V=11, T_range=(4, 25), S_range=(0, 9))
    al = fixtures.random_alignment(np.random.default_rng(3), case.T, case.S, case.labels)
    whole = oracle.run(case.acts, case.labels, case.T, case.S, case.V, alignment=al, max_shift=1)
    for world in (2, 3):
        costs, grads = [], []
        for b0, b1 in mr.shard.partition_contiguous(case.T, case.S, world):
            sh = mr.shard.make_shard(case.T, case.S, case.labels, b0, b1, alignment=al)
            r = oracle.run(case.acts[sh.row0:sh.row1], sh.labels, sh.T, sh.S, case.V, alignment=sh.alignment,
                           max_shift=1)
            costs.append(r.costs)
            grads.append(r.grads)
        assert np.array_equal(np.concatenate(costs), whole.costs)
        assert np.array_equal(np.concatenate(grads), whole.grads)


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    case = fixtures.random_case("gloo", 5, B=6, V=9, T_range=(3, 15), S_range=(0, 6))
    b0, b1 = mr.shard.partition_contiguous(case.T, case.S, world)[rank]
    sh = mr.shard.make_shard(case.T, case.S, case.labels, b0, b1)
    r = oracle.run(case.acts[sh.row0:sh.row1], sh.labels, sh.T, sh.S, case.V)
    total = mr.shard.allreduce_cost_sum(torch.from_numpy(r.costs))        # the path's ONE collective
    q.put((rank, float(total), b0, b1))
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world_size_2_cost_allreduce():
    import torch.multiprocessing as tmp
    ctx = tmp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 1000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    case = fixtures.random_case("gloo", 5, B=6, V=9, T_range=(3, 15), S_range=(0, 6))
    whole = oracle.run(case.acts, case.labels, case.T, case.S, case.V)
    for rank, total, b0, b1 in got:
        assert abs(total - float(whole.costs.sum(dtype=np.float64))) < 1e-3
    assert sorted((b0, b1) for _, _, b0, b1 in got)[0][0] == 0


def test_bench_reference_arm_schema():
    env = dict(os.environ, MRNNT_BENCH_TEST_B="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "utt/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["higher_is_better"] is True
    # other ranks of a torchrun launch exit 0 without work
    env["RANK"] = "1"
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1"],
                         capture_output=True, text=True, env=env, timeout=60)
    assert out.returncode == 0 and out.stdout.strip() == ""


@pytest.mark.parametrize("name", ["c2", "c3", "c4", "c5"])
def test_bench_cpu_sample_is_a_bounded_prefix(name):
    """bench.py's host legs (CPU baseline, checker, reference arm) run on a prefix of the batch: at most 4 GiB of
    logits (so the reference's int indexing stays in range), labels / alignment re-strided to the prefix's
    own maxima (the reference derives both strides from the lengths it is given)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    import monotonic_rnnt_b200 as mr
    wl = mr.synth.workload(name)
    k, rows_k, labels, align = bench._cpu_sample(wl)
    assert 1 <= k <= wl.B
    assert rows_k == int((wl.T[:k].astype(np.int64) * (wl.S[:k].astype(np.int64) + 1)).sum())
    assert rows_k * wl.V < 2 ** 31 and (rows_k * wl.V * 4 <= bench.CPU_SAMPLE_BYTES or k == 1)
    assert labels.shape == (k, int(wl.S[:k].max())) and labels.flags["C_CONTIGUOUS"]
    np.testing.assert_array_equal(labels, wl.labels[:k, : labels.shape[1]])
    if wl.alignment is None:
        assert align is None
    else:
        assert align.shape == (k, int(wl.T[:k].max())) and align.flags["C_CONTIGUOUS"]
    if name == "c2":
        assert k == wl.B          # the headline shape is checked whole
    k1, rows_1, _, _ = bench._cpu_sample(wl, 1)
    assert k1 == 1 and rows_1 == int(wl.T[0]) * (int(wl.S[0]) + 1)


def test_option_numbers_of_the_python_mirror_match_the_header():
    """monotonic_rnnt_b200/_lib.py repeats the MRNNT_OPT_* / MRNNT_DBG_* numbers of include/mrnnt_c_api.h (ctypes cannot read a
    header): every constant of the mirror must carry the header's value, and every option of the header must be there."""
    import re
    from monotonic_rnnt_b200 import _lib
    text = open(os.path.join(ROOT, "include", "mrnnt_c_api.h")).read()
    header = {m.group(1): int(m.group(2)) for m in re.finditer(r"\bMRNNT_((?:OPT|DBG)_[A-Z0-9_]+)\s*=\s*(\d+)", text)}
    assert len([k for k in header if k.startswith("OPT_")]) >= 15 and len([k for k in header if k.startswith("DBG_")]) >= 8
    for name, value in header.items():
        assert getattr(_lib, name, None) == value, name
    for name in dir(_lib):
        if name.startswith(("OPT_", "DBG_")):
            assert header.get(name) == getattr(_lib, name), name
