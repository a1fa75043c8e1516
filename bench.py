#!/usr/bin/env python
"""bench.py -- loss+grad throughput of the monotonic RNN-T hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # our CUDA path (one process per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU implementation, host cores

One "step" = one pass of the hot path over one batch of synthetic input: the configuration BASELINE.json
quotes the metric on (configs[1]: B=32 T=150 S=40 V=1000 fp32 logits, fixed lengths) per GPU.  Weak
scaling (default): every rank runs that batch on its own logits.  --scaling strong (BASELINE.json configs[2]: "batch-
sharded at 1/2/4/8 B200"): ONE batch is cut over the ranks by utterance (monotonic_rnnt_b200/shard.py: LPT assignment,
or --partition contiguous), every rank regenerates its utterances' logits from the global element index, runs the same
call on its shard, and value = the batch's utterances / the slowest rank's step time; rank 0 then also runs the whole
batch alone and asserts the gathered per-utterance costs equal it bit for bit.  Either way the only collective is the
sum of the costs (inside the gradient kernel, over peer memory).  Rank 0 prints ONE JSON line.

  value     whole-job utterances/s with inputs resident in HBM; the timed region is K calls of the C-ABI
            entry (mrnnt_cost_and_grad: K1 -> K2 -> K3, costs on the host on return, one stream sync)
            bracketed by barrier + synchronize, CUDA events on the launch stream, max over ranks.  The K-step
            region is run --blocks times (default 5); ms_per_step is the MEDIAN block (all blocks and the per-rank
            times are in the line), so that one host hiccup in a 7 ms region cannot move the number.
  e2e       same metric through the public host API with HOST buffers: every step brings the step's
            logits/labels/lengths from pinned host memory to the device, builds the handle, runs, reads costs
            back.  Gradients stay on the device, as in the reference's contract (its GPU path leaves them there,
            gpu_rnnt.h:229): the leg measures the upload + call.  Two ways (e2e_paths): the whole tensor through the
            copy engine, and the rows the lattice reads only (mrnnt_upload_acts); e2e is the faster.
  roofline  dominant kernel (K3, gradient) timed per launch with CUDA events recorded around it on the same stream
            in a second, instrumented pass of K steps.  achieved = the bytes the kernel's job requires (read the logits
            of live rows once + write the gradient rows it owns once) / duration; traffic = DRAM bytes per launch from
            the committed ncu capture of this workload (profiles/traffic.json); dram_frac = traffic / duration / peak.
            (The 12-bytes-per-logit convention of SURVEY 8d, which charges dead rows as read, is under call_roofline.)
  cpu_baseline  the reference's own CPU implementation (oracle/_ref, compiled from the unmodified
            reference sources) on the box's host cores, same inputs, bounded sample; also used here as the
            checker of the GPU result (costs vs it, gradients vs the double-precision oracle: asserted), which is
            the only reason bench.py touches oracle/.
Inputs (0.79 GB logits + 0.79 GB gradients per step) exceed the 126 MB L2, so no L2 flush is needed
between iterations (config.l2: "inputs>L2").
--workload c3|c4|c5 puts the other shapes BASELINE.json names through the same contract (the driver's line is c2,
the default); their host legs run on a prefix of the batch of at most 4 GiB of logits.  --workload c2v1025 / c4v5001:
the same shapes with a vocabulary that is not a multiple of 4 (unaligned rows).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "loss_grad_utterances_per_sec"
UNIT = "utt/s"
WORKLOAD = "c2"  # BASELINE.json configs[1]; --workload c3 / c4 / c5 reports the other named shapes the same way
SCALING = "weak"
WORKLOAD_NAMES = {
    "c2v1025": "synthetic B={B} T={T} S={S} V={V} fp32 logits, fixed lengths: c2 with rows that are not whole 16-byte vectors, per GPU",
    "c4v5001": "large-vocab B={B} T={T} S={S} V={V} fp32 logits: c4 with rows that are not whole 16-byte vectors, per GPU",
    "c2": "synthetic B={B} T={T} S={S} V={V} fp32 logits, fixed lengths (BASELINE.json configs[1]) per GPU",
    "c3": "synthetic B={B} T<={T} S<={S} V={V} fp32 logits, random per-utterance T_b/S_b, packed "
          "(BASELINE.json configs[2]) per GPU",
    "c4": "large-vocab B={B} T={T} S={S} V={V} fp32 logits (BASELINE.json configs[3]) per GPU",
    "c5": "alignment-restricted B={B} T={T} S={S} V={V} fp32 logits, max distance 5 (BASELINE.json configs[4]) per GPU",
}
CPU_SAMPLE_BYTES = 4 << 30   # the host baseline / checker runs on a prefix of the batch of at most this many logit bytes


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def _traffic(kernel: str):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full capture of this workload, if any."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f).get(WORKLOAD, {}).get(kernel)
    return None


def _config(wl, extra=None):
    cfg = {"workload": f"{wl.name}: " + WORKLOAD_NAMES[wl.name].format(B=wl.B, T=int(wl.T.max()), S=int(wl.S.max()), V=wl.V),
           "batch_per_gpu": wl.B, "rows_per_gpu": wl.rows, "logit_bytes_per_gpu": wl.elements * 4,
           "algorithmic_bytes_per_step_per_gpu": wl.algorithmic_bytes, "l2": "inputs>L2 (no flush needed)",
           "parallelism": "utterance-sharded, one sum of the costs over all GPUs"}
    if extra:
        cfg.update(extra)
    return cfg


# ----------------------------------------------------------------------------------------------------
# clocks: NVML sampled from a thread while the timed regions run
# ----------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index: int):
        self.samples, self.reasons, self.power = [], set(), []
        self.sm_max = None
        self._stop = threading.Event()
        self._active = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = int(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        except Exception as exc:  # NVML missing: report that, never fake numbers
            self._nv = None
            self.error = repr(exc)

    def _run(self):
        nv = self._nv
        while not self._stop.is_set():
            if self._active.is_set():
                try:
                    self.samples.append(int(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                    self.power.append(nv.nvmlDeviceGetPowerUsage(self._h) / 1000.0)
                    get = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                        nv.nvmlDeviceGetCurrentClocksThrottleReasons
                    mask = int(get(self._h))
                    for bit, name in self.REASONS.items():
                        if mask & bit and name != "gpu_idle":
                            self.reasons.add(name)
                except Exception:
                    pass
            time.sleep(0.002)

    def start(self):
        self._active.set()

    def pause(self):
        self._active.clear()

    def result(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=1.0)
        if self._nv is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "error": getattr(self, "error", "nvml")}
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.sm_max,
                "reasons": sorted(self.reasons), "samples": len(self.samples),
                "power_w_max": max(self.power) if self.power else None}


# ----------------------------------------------------------------------------------------------------
# reference arm: the reference's CPU implementation on the host cores
# ----------------------------------------------------------------------------------------------------
def _host_threads() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def _cpu_sample(wl, max_utts=None, max_bytes=None):
    """A prefix of the batch for the host legs: (k, rows_k, labels, alignment) with the labels / the alignment re-strided
    to the prefix's own maxima (the reference derives both strides from the lengths it is given,
    cpu_workspace_manager.h:44,122) and at most CPU_SAMPLE_BYTES of logits -- which also keeps the reference's int
    indexing below 2^31 elements (c4 as a whole overflows it)."""
    rows_b = wl.T.astype(np.int64) * (wl.S.astype(np.int64) + 1)
    cum = np.cumsum(rows_b) * wl.V * 4
    k = int(np.searchsorted(cum, max_bytes or CPU_SAMPLE_BYTES, side="right"))
    k = max(1, min(wl.B, k, max_utts or wl.B))
    s_max = max(1, int(wl.S[:k].max()))
    labels = np.ascontiguousarray(wl.labels[:k, :s_max])
    align = None if wl.alignment is None else np.ascontiguousarray(wl.alignment[:k, : int(wl.T[:k].max())])
    return k, int(rows_b[:k].sum()), labels, align


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # under torchrun only rank 0 measures the host baseline
    import monotonic_rnnt_b200 as mr
    from oracle import oracle

    oracle.build()
    wl = mr.synth.workload(WORKLOAD, B=int(os.environ["MRNNT_BENCH_TEST_B"]) if "MRNNT_BENCH_TEST_B" in os.environ
                           else None)  # the override exists for tests/test_host_cpu.py only
    kind = "reference" if oracle.have_ref() else "port"
    runner = oracle.run_ref if kind == "reference" else oracle.run
    # all the host threads this process may use, named explicitly: torchrun exports OMP_NUM_THREADS=1 to its workers,
    # and the reference's own default (num_threads = 0 -> omp_get_max_threads()) would then time ONE thread
    cores = _host_threads()
    k_max, rows_max, _, _ = _cpu_sample(wl)
    all_acts = mr.synth.uniform_logits(rows_max * wl.V, wl.logits_seed, 0)  # generated once, outside the timing

    def step(B):
        k, rows_k, labels, align = _cpu_sample(wl, B)
        acts = all_acts[: rows_k * wl.V]
        t0 = time.perf_counter()
        res = runner(acts, labels, wl.T[:k], wl.S[:k], wl.V, blank=wl.blank, alignment=align, max_shift=wl.max_shift,
                     precision="f32", want_grads=True, num_threads=cores)
        dt = time.perf_counter() - t0
        assert np.isfinite(res.costs).all()
        return dt

    # bounded sample: as many utterances of the workload as keep the whole run within ~2 minutes
    probe_B = min(k_max, max(1, cores))
    t_probe = step(probe_B)
    per_utt_s = t_probe / probe_B
    budget = 120.0 / max(1, args.steps + args.warmup)
    B = int(max(1, min(k_max, budget / per_utt_s)))
    if B >= cores:
        B = B // cores * cores
    for _ in range(args.warmup):
        step(B)
    t = [step(B) for _ in range(args.steps)]
    ms = 1000.0 * sum(t) / len(t)
    value = B / (ms / 1000.0)
    sample = (f"the first {B} of the {wl.B} utterances of {wl.name} per step (full T<={int(wl.T.max())} "
              f"S<={int(wl.S.max())} V={wl.V}"
              + (f", alignment band +-{wl.max_shift}" if wl.alignment is not None else "")
              + f"), CpuRNNTComputer<float>::cost_and_grad, -O2 -fopenmp, {cores} threads (the reference runs one "
              f"utterance per thread)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": _config(wl, {"sample_batch": B}),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    _emit(args.out_fd, line)


# ----------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------
def _bind_to_gpu_numa_node(index: int) -> None:
    """Pin this process to the CPUs next to its GPU (NVML's ideal affinity), so that the pinned host buffers of the
    e2e leg are allocated on the GPU's own NUMA node and its H2D copies do not cross the socket interconnect."""
    try:
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(index))
    except Exception:
        pass  # no NVML / no permission: run unbound


def _local_workload(mr, wl, args, rank, world):
    """Weak scaling: the named batch on every rank.  Strong scaling: this rank's utterances of the ONE batch."""
    import dataclasses
    if SCALING != "strong" or world == 1:
        return wl, None, None
    if args.partition == "contiguous":
        parts = [np.arange(a, b) for a, b in mr.shard.partition_contiguous(wl.T, wl.S, world)]
    else:
        parts = mr.shard.partition_lpt(wl.T, wl.S, world)
    if min(len(p) for p in parts) == 0:
        raise SystemExit(f"--scaling strong: {wl.B} utterances cannot feed {world} ranks")
    sh = mr.shard.make_shard_indexed(wl.T, wl.S, wl.labels, parts[rank], alignment=wl.alignment)
    local = dataclasses.replace(wl, B=len(parts[rank]), T=sh.T, S=sh.S, labels=sh.labels, alignment=sh.alignment)
    return local, sh, parts


def run_b200(args) -> None:
    import ctypes

    import torch
    import torch.distributed as dist

    import monotonic_rnnt_b200 as mr
    from monotonic_rnnt_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with: python -m torch.distributed.run --nproc-per-node N bench.py --gpus N ...")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path "
                         "(use --impl reference for the host baseline)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    _bind_to_gpu_numa_node(local_rank)  # before any pinned allocation: first touch decides where the pages live
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def gather_floats(x):
        """One float per rank -> the list of all ranks' values (on every rank)."""
        t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
        if world == 1:
            return [float(x)]
        out = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return [float(o.item()) for o in out]

    lib = _lib.load()
    wl_global = mr.synth.workload(WORKLOAD)
    wl, shard, parts = _local_workload(mr, wl_global, args, rank, world)
    strong = shard is not None
    n = wl.elements
    stream = torch.cuda.current_stream()

    # ---- inputs resident in HBM ------------------------------------------------------------------
    acts = torch.empty((wl.rows, wl.V), dtype=torch.float32, device=dev)
    if strong:
        # every utterance's logits from its GLOBAL element index: the same bits as in the one-GPU run of the whole batch
        off = 0
        for r0, r1 in shard.row_ranges:
            ne = (r1 - r0) * wl.V
            _lib.check(lib.mrnnt_synth_uniform(acts.data_ptr() + 4 * off, ne, wl.logits_seed, r0 * wl.V,
                                               stream.cuda_stream), "synth")
            off += ne
    else:
        _lib.check(lib.mrnnt_synth_uniform(acts.data_ptr(), n, wl.logits_seed, rank * n, stream.cuda_stream), "synth")
    labels = torch.from_numpy(wl.labels).to(dev)
    T = torch.from_numpy(wl.T).to(dev)
    S = torch.from_numpy(wl.S).to(dev)
    grads = torch.empty_like(acts)
    costs_host = torch.empty(wl.B, dtype=torch.float32).pin_memory()
    handle = mr.LossHandle(acts, labels, T, S, lengths_host=(wl.T, wl.S))
    if wl.alignment is not None:
        handle.restrict_to_alignment(torch.from_numpy(wl.alignment).to(dev), wl.max_shift, wl.blank)

    cost_sum_host = torch.zeros((), dtype=torch.float32).pin_memory()

    # the C-ABI entry itself, bound once (handle, blank, stream, costs on the host, gradients on the device): the
    # Python mirror adds argument checks and attribute look-ups that are not part of the path being measured
    abi_call = lib.mrnnt_cost_and_grad
    abi_args = (handle._h, ctypes.c_int(wl.blank), ctypes.c_void_p(stream.cuda_stream),
                ctypes.c_void_p(costs_host.data_ptr()), ctypes.c_void_p(grads.data_ptr()))

    # N > 1: the path's only collective, the all-GPU sum of the summed cost (4 bytes), is done by the kernels themselves
    # over peer memory (monotonic_rnnt_b200/peer.py, include/mrnnt_b200/peer_reduce.cuh): the step is the SAME call as
    # at N = 1 and returns with the costs and the world's sum on the host.  If the boards cannot be mapped (no CUDA IPC
    # between the ranks) the step falls back to an NCCL all-reduce on a side stream; the JSON line says which.
    boards, collective = None, "none (1 GPU)"
    if world > 1:
        collective = "fused: peer-memory stores from the gradient kernel (NVLink), no collective kernel"
        try:
            if args.collective == "nccl":
                raise _lib.RNNTError(0, "--collective nccl")
            boards = mr.peer.PeerBoards(device=dev)
            handle.set_peer_reduce(boards, cost_sum_host)
        except _lib.RNNTError as exc:
            boards = None
            collective = f"nccl all-reduce on a side stream ({exc})"
        flag = torch.tensor([1 if boards is not None else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)     # all ranks take the same path
        if int(flag.item()) == 0 and boards is not None:
            handle.set_peer_reduce(None, None)
            boards = None
            collective = "nccl all-reduce on a side stream (a peer could not map the boards)"

    def one_step():
        if world == 1 or boards is not None:
            st = abi_call(*abi_args)  # costs on the host on return; gradients (and the world's sum) in stream order
            if st != 0:
                _lib.check(st, "mrnnt_cost_and_grad")
            return
        # Fallback at N > 1: the same three kernels without a host round trip in between.  The costs are final after
        # K2, so the all-reduce and the copies of the costs to the host run on a side stream / NCCL's stream next to K3
        # on the compute stream; ONE host synchronisation at the end of the step.
        st = fwd_call(*fwd_args)                       # K1, K2
        if st != 0:
            _lib.check(st, "mrnnt_enqueue_forward_into")
        k2_done.record(stream)
        st = bwd_call(*bwd_args)                       # K3 (compute stream)
        if st != 0:
            _lib.check(st, "mrnnt_enqueue_backward")
        with torch.cuda.stream(side):
            side.wait_event(k2_done)
            costs_host.copy_(dev_costs, non_blocking=True)
            total = dev_costs.sum()
            dist.all_reduce(total)                     # NCCL's stream, ordered after `side` up to here
            cost_sum_host.copy_(total, non_blocking=True)
        side.synchronize()
        stream.synchronize()

    side = torch.cuda.Stream(device=dev) if world > 1 else None
    k2_done = torch.cuda.Event() if world > 1 else None
    dev_costs = handle.device_costs()
    # (the forward half is told which buffer the backward half will fill: the lattice kernel zeroes its dead rows)
    fwd_call, bwd_call = lib.mrnnt_enqueue_forward_into, lib.mrnnt_enqueue_backward
    fwd_args = (handle._h, ctypes.c_int(wl.blank), ctypes.c_void_p(stream.cuda_stream), ctypes.c_void_p(grads.data_ptr()))
    bwd_args = (handle._h, ctypes.c_void_p(stream.cuda_stream), ctypes.c_void_p(grads.data_ptr()), None)

    if world > 1:
        handle.set_option(_lib.OPT_RESERVED_SMS, args.reserve_sms)
    # The call returns when the costs are on the host; the gradient kernel completes in stream order (the default at
    # N = 1, MRNNT_OPT_RETURN_EARLY).  With the fused exchange the world's sum arrives at the END of the gradient kernel,
    # so there the same behaviour has to be asked for (value 2: the sum is valid in stream order as well -- the bench
    # reads it behind the block's synchronisation, as a training loop reads its logged loss).
    if boards is not None and not args.full_wait:
        handle.set_option(_lib.OPT_RETURN_EARLY, 2)
    if args.full_wait:
        handle.set_option(_lib.OPT_RETURN_EARLY, 0)
    call_semantics = ("mrnnt_cost_and_grad returns only when everything it launched has completed (--full-wait)" if args.full_wait else
                      "mrnnt_cost_and_grad returns as soon as the costs are on the host; the gradient kernel completes in "
                      "stream order (MRNNT_OPT_RETURN_EARLY) and every timed block ends with a device synchronisation")
    clocks = ClockSampler(local_rank) if rank == 0 else None
    for _ in range(max(args.warmup, 3)):
        one_step()
    if boards is not None:
        # the exchange must have worked on EVERY rank before anything is timed through it (a peer whose stores do not
        # arrive leaves NaN after the kernel's time-out): otherwise all ranks switch to the NCCL step, and say so
        want = torch.tensor([float(costs_host.double().sum())], dtype=torch.float64, device=dev)
        dist.all_reduce(want)
        torch.cuda.synchronize()                        # (the sum lands when the gradient kernel ends)
        got = float(cost_sum_host.item())
        ok = np.isfinite(got) and abs(got - float(want.item())) <= 1e-5 * abs(float(want.item()))
        flag = torch.tensor([1 if ok else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            handle.set_peer_reduce(None, None)
            boards = None
            collective = "nccl all-reduce on a side stream (the peer-memory exchange failed its check in warm-up)"
            for _ in range(3):
                one_step()

    # ---- timed region: value.  `blocks` times EXACTLY K steps, each block bracketed by barrier + synchronize; a block's
    #      time is the max over ranks of the CUDA-event time on the launch stream; the MEDIAN block is reported --------
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    block_rank_ms, block_launches, wall_ms = [], [], []
    for _ in range(max(1, args.blocks)):
        barrier()
        if clocks:
            clocks.start()
        launches0 = handle.get_option(_lib.OPT_LAUNCH_COUNT)
        t0 = time.perf_counter()
        ev0.record(stream)
        for _ in range(args.steps):
            one_step()
        ev1.record(stream)
        barrier()
        wall_ms.append(1000.0 * (time.perf_counter() - t0))
        if clocks:
            clocks.pause()
        block_launches.append(handle.get_option(_lib.OPT_LAUNCH_COUNT) - launches0)
        block_rank_ms.append(gather_floats(ev0.elapsed_time(ev1)))
    block_ms = [max(b) for b in block_rank_ms]                       # max over ranks, per block
    order = sorted(range(len(block_ms)), key=lambda i: block_ms[i])
    mid = order[(len(order) - 1) // 2]                               # the median block (lower median: a block that ran)
    ms_per_step = block_ms[mid] / args.steps
    B_job = wl_global.B if strong else world * wl.B
    value = B_job / (ms_per_step / 1000.0)
    gpu_launches = int(block_launches[mid])
    # the world's sum as the kernels exchanged it, against a library all-reduce of the same costs (outside the timing)
    collective_check = None
    if world > 1:
        fused_total = float(cost_sum_host.item())
        want = torch.tensor([float(costs_host.double().sum())], dtype=torch.float64, device=dev)
        dist.all_reduce(want)
        collective_check = {"world_sum": fused_total, "nccl_f64_sum_of_the_same_costs": float(want.item()),
                            "rel_diff": abs(fused_total - float(want.item())) / abs(float(want.item()))}
        assert collective_check["rel_diff"] < 1e-5, collective_check

    # ---- strong scaling: the sharded result IS the whole batch's (outside the timing) ------------------------------
    shard_info = None
    if strong:
        rows_per_rank = [int(sum((wl_global.T[p].astype(np.int64) * (wl_global.S[p] + 1)))) for p in parts]
        all_costs = torch.full((wl_global.B,), float("nan"), dtype=torch.float32, device=dev)
        all_costs[torch.from_numpy(np.asarray(parts[rank])).to(dev)] = costs_host.to(dev)
        gathered = [torch.empty_like(all_costs) for _ in range(world)]
        dist.all_gather(gathered, all_costs)
        merged = torch.stack(gathered).nan_to_num(nan=0.0).sum(dim=0).cpu()    # (every utterance is owned by one rank)
        whole_equal = None
        if rank == 0:
            g = wl_global
            acts_w = torch.empty((g.rows, g.V), dtype=torch.float32, device=dev)
            _lib.check(lib.mrnnt_synth_uniform(acts_w.data_ptr(), g.elements, g.logits_seed, 0, stream.cuda_stream), "synth")
            hw = mr.LossHandle(acts_w, torch.from_numpy(g.labels).to(dev), torch.from_numpy(g.T).to(dev),
                               torch.from_numpy(g.S).to(dev), lengths_host=(g.T, g.S))
            if g.alignment is not None:
                hw.restrict_to_alignment(torch.from_numpy(g.alignment).to(dev), g.max_shift, g.blank)
            whole = hw.cost_and_grad(g.blank, torch.empty_like(acts_w)).clone()
            hw.close()
            del acts_w
            torch.cuda.empty_cache()
            whole_equal = bool(torch.equal(whole, merged))
            assert whole_equal, "sharded per-utterance costs differ from the whole batch on one GPU"
        shard_info = {"partition": args.partition, "utterances_per_rank": [int(len(p)) for p in parts],
                      "rows_per_rank": rows_per_rank,
                      "imbalance_max_over_mean": max(rows_per_rank) / (sum(rows_per_rank) / world),
                      "max_T_per_rank": [int(wl_global.T[p].max()) for p in parts],
                      "costs_bit_identical_to_the_whole_batch_on_one_gpu": whole_equal}

    # ---- the same K steps through the asynchronous entry (mrnnt_enqueue: costs stay on the device, no host round trip
    #      per step, one synchronisation at the end): what a training loop that never looks at the costs sees ----
    async_ms = None
    if world == 1:
        enq = lib.mrnnt_enqueue
        enq_args = (handle._h, ctypes.c_int(wl.blank), ctypes.c_void_p(stream.cuda_stream), ctypes.c_void_p(grads.data_ptr()))
        for _ in range(3):
            enq(*enq_args)
        barrier()
        ev0.record(stream)
        for _ in range(args.steps):
            enq(*enq_args)
        ev1.record(stream)
        barrier()
        async_ms = ev0.elapsed_time(ev1) / args.steps

    # ---- the same K steps with the whole wait inside every call (MRNNT_OPT_RETURN_EARLY 0): what the early return buys
    full_wait_ms = None
    if world == 1 and not args.full_wait:
        handle.set_option(_lib.OPT_RETURN_EARLY, 0)
        for _ in range(3):
            one_step()
        barrier()
        ev0.record(stream)
        for _ in range(args.steps):
            one_step()
        ev1.record(stream)
        barrier()
        full_wait_ms = ev0.elapsed_time(ev1) / args.steps
        handle.set_option(_lib.OPT_RETURN_EARLY, 1)

    # ---- the call as the reference's torch binding pays for it (pytorch_binding/monotonic_rnnt.cu:99-111): a new manager,
    #      cudaMalloc of the workspace, cost_and_grad, cudaFree -- every call ----
    alloc_ms = None
    if world == 1:
        hp = ctypes.c_void_p()
        Th = np.ascontiguousarray(wl.T, dtype=np.int32); Sh = np.ascontiguousarray(wl.S, dtype=np.int32)

        def alloc_step():
            _lib.check(lib.mrnnt_create(ctypes.byref(hp), acts.data_ptr(), labels.data_ptr(), wl.B, T.data_ptr(), S.data_ptr(),
                                        wl.V, Th.ctypes.data, Sh.ctypes.data), "mrnnt_create")
            _lib.check(lib.mrnnt_create_workspace(hp), "mrnnt_create_workspace")
            st = lib.mrnnt_cost_and_grad(hp, wl.blank, stream.cuda_stream, costs_host.data_ptr(), grads.data_ptr())
            lib.mrnnt_free_workspace(hp)
            lib.mrnnt_destroy(hp)
            _lib.check(st, "mrnnt_cost_and_grad")

        if wl.alignment is None:
            for _ in range(3):
                alloc_step()
            n_alloc = max(3, min(args.steps, 20))
            barrier()
            t_a = time.perf_counter()
            for _ in range(n_alloc):
                alloc_step()
            torch.cuda.synchronize()
            alloc_ms = 1000.0 * (time.perf_counter() - t_a) / n_alloc

    # ---- instrumented pass: per-kernel durations (CUDA events around K1, K2, K3 on the launch stream) -
    handle.set_option(_lib.OPT_TIMING, 1)
    handle.cost_and_grad(wl.blank, grads, costs_host)
    acc = np.zeros(3)
    if clocks:
        clocks.start()
    for _ in range(args.steps):
        handle.cost_and_grad(wl.blank, grads, costs_host)
        acc += np.array(handle.last_timings())
    if clocks:
        clocks.pause()
    handle.set_option(_lib.OPT_TIMING, 0)
    k_ms = acc / args.steps
    # who writes the gradient's zero rows (rows whose alpha(t-1, s) lies outside the lattice): the lattice kernel,
    # while its recursions run, or the gradient kernel.  The roofline below charges each kernel with what it moves.
    zero_fill_warps = handle.get_option(_lib.OPT_K2_ZERO_FILL)
    dead_rows = int((handle.debug(_lib.DBG_ROWMETA) == -2).sum())
    costs_gpu = costs_host.clone().numpy()

    # ---- e2e: host buffers, copies inside the timed region --------------------------------------------
    if args.no_e2e:                                      # (no host copy of the logits at all: 15.5 GB per rank on c4)
        want_host = rank == 0 and world == 1 and not args.no_cpu_baseline   # (the checker reads a plain host copy)
        acts_h = acts.cpu() if want_host else torch.empty((0, wl.V), dtype=torch.float32)
    else:
        acts_h = torch.empty((wl.rows, wl.V), dtype=torch.float32).pin_memory()
        acts_h.copy_(acts)
    labels_h = torch.from_numpy(wl.labels).pin_memory()
    T_h = torch.from_numpy(wl.T).pin_memory()
    S_h = torch.from_numpy(wl.S).pin_memory()
    align_h = None if wl.alignment is None else torch.from_numpy(wl.alignment).pin_memory()
    align_d = None if align_h is None else torch.empty_like(align_h, device=dev)
    h2d = acts_h.numel() * 4 + labels_h.numel() * 4 + T_h.numel() * 4 + S_h.numel() * 4 \
        + (0 if align_h is None else align_h.numel() * 4)
    d2h = wl.B * 4

    live_bytes = (wl.rows - dead_rows) * wl.V * 4

    def e2e_step(mode):
        live_rows_only = mode != "whole_tensor"
        labels.copy_(labels_h, non_blocking=True)
        T.copy_(T_h, non_blocking=True)
        S.copy_(S_h, non_blocking=True)
        h = mr.LossHandle(acts, labels, T, S, lengths_host=(wl.T, wl.S))
        if align_h is not None:
            align_d.copy_(align_h, non_blocking=True)
            h.restrict_to_alignment(align_d, wl.max_shift, wl.blank)
        if live_rows_only:
            if mode == "live_rows_kernel_only":
                h.set_option(_lib.OPT_UPLOAD_COPY_ENGINE, 0)
            h.upload_acts(acts_h)                        # only the rows the lattice reads cross the bus (mrnnt_upload_acts)
        else:
            acts.copy_(acts_h, non_blocking=True)        # the whole tensor through the copy engine
        if boards is not None:
            h.set_peer_reduce(boards, cost_sum_host)     # (a new handle takes the boards over at their epoch)
        h.cost_and_grad(wl.blank, grads, costs_host)     # returns with the costs (and the world's sum) on the host
        h.sync_peer_epoch()
        h.close()

    e2e_steps = 0 if args.no_e2e else max(3, min(args.steps, 20 if h2d < (4 << 30) else 3))
    if boards is not None:
        handle.sync_peer_epoch()
    e2e_paths = {}
    for mode in (() if args.no_e2e else ("whole_tensor", "live_rows_kernel_only", "live_rows")):
        live_rows_only = mode != "whole_tensor"
        if live_rows_only:
            acts.fill_(float("nan"))                     # what the upload does not bring must not matter
        e2e_step(mode)
        barrier()
        if clocks:
            clocks.start()
        ev0.record(stream)
        for _ in range(e2e_steps):
            e2e_step(mode)
        ev1.record(stream)
        barrier()
        if clocks:
            clocks.pause()
        rank_ms = gather_floats(ev0.elapsed_time(ev1) / e2e_steps)
        ms = max(rank_ms)
        assert np.allclose(costs_host.numpy(), costs_gpu, rtol=1e-6)
        bytes_step = (h2d - acts_h.numel() * 4 + live_bytes) if live_rows_only else h2d
        host_read = gather_floats(bytes_step)
        e2e_paths[mode] = {
            "value": B_job / (ms / 1000.0), "unit": UNIT, "ms_per_step": ms, "steps": e2e_steps,
            "h2d_bytes_per_step": bytes_step,
            "d2h_bytes_per_step": d2h,
            "ms_per_step_per_rank": rank_ms,
            "host_read_GBps_all_ranks": sum(host_read) / (ms * 1e-3) / 1e9,
            "gradients": "stay on the device (the reference's GPU contract, gpu_rnnt.h:229: costs to the host, gradients "
                         "to the caller's device buffer): this leg is the upload of the logits plus the call",
            "path": {"live_rows": "pinned host -> LossHandle(...) -> mrnnt_upload_acts (the all-live block in the middle of every "
                                  "utterance through the copy engine on a side stream, the ragged frames around it by a kernel that "
                                  "reads pinned host memory over PCIe; dead rows never cross) -> mrnnt_cost_and_grad -> costs on host",
                     "live_rows_kernel_only": "pinned host -> LossHandle(...) -> mrnnt_upload_acts with MRNNT_OPT_UPLOAD_COPY_ENGINE 0 (a "
                                              "kernel reads every live row from host memory over PCIe) -> mrnnt_cost_and_grad -> costs on host",
                     "whole_tensor": "pinned host -> H2D copy of the whole tensor -> LossHandle(...) -> mrnnt_cost_and_grad -> costs "
                                     "on host"}[mode]}
    e2e_best = max(e2e_paths.values(), key=lambda e: e["value"]) if e2e_paths else None  # (None: --no-e2e, profiling runs)
    if not args.no_e2e:
        acts.copy_(acts_h)                               # (the checker below reads the device copy's gradients)
    torch.cuda.synchronize()
    assert np.allclose(costs_host.numpy(), costs_gpu, rtol=1e-6)

    # ---- CPU baseline beside it (rank 0, N == 1 only), doubling as the checker ---------------------------
    cpu_baseline, parity = None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import oracle
        oracle.build()
        kind = "reference" if oracle.have_ref() else "port"
        runner = oracle.run_ref if kind == "reference" else oracle.run
        cores = _host_threads()
        k, rows_k, labels_k, align_k = _cpu_sample(wl)
        acts_np = acts_h.numpy().reshape(-1)[: rows_k * wl.V]
        times, res = [], None
        t_budget = time.perf_counter()
        for i in range(3):
            t1 = time.perf_counter()
            res = runner(acts_np, labels_k, wl.T[:k], wl.S[:k], wl.V, blank=wl.blank, alignment=align_k,
                         max_shift=wl.max_shift, precision="f32", want_grads=True, num_threads=cores)
            times.append(time.perf_counter() - t1)
            if time.perf_counter() - t_budget > 20.0:
                break
        best = min(times)
        which = f"the full {wl.name} batch ({wl.B} utterances)" if k == wl.B else \
            f"the first {k} of the {wl.B} utterances of {wl.name} ({rows_k * wl.V * 4 / 1e9:.2f} GB of logits)"
        cpu_baseline = {"value": k / best, "unit": UNIT, "cores": cores, "kind": kind,
                        "sample": f"{which}, best of {len(times)} runs of "
                                  f"CpuRNNTComputer<float>::cost_and_grad (-O2 -fopenmp), {best * 1000:.0f} ms"}
        g = grads[:rows_k].cpu().numpy()
        # gradients: against the double-precision oracle (the float reference is its own rounding noise away from exact
        # arithmetic, 3e-4 .. 3e-3 on these shapes) on a prefix of at most 1 GiB of logits (>= 1 utterance)
        k64, rows64, labels64, align64 = _cpu_sample(wl, max_bytes=1 << 30)
        o64 = oracle.run(acts_h.numpy().reshape(-1)[: rows64 * wl.V], labels64, wl.T[:k64], wl.S[:k64], wl.V, blank=wl.blank,
                         alignment=align64, max_shift=wl.max_shift, precision="f64_from_f32", want_grads=True,
                         num_threads=cores)
        parity = {"checker": f"oracle/{'_ref' if kind == 'reference' else 'liboracle'} f32 on the same inputs ({which}) for the "
                             f"costs; the double-precision oracle on the first {k64} utterances for the gradients",
                  "cost_max_rel_vs_f32_reference": float(np.max(np.abs(costs_gpu[:k] - res.costs) / np.abs(res.costs))),
                  "cost_max_rel_vs_f64_oracle": float(np.max(np.abs(costs_gpu[:k64] - o64.costs) / np.abs(o64.costs))),
                  "grad_max_abs_vs_f64_oracle": float(np.abs(g[:rows64] - o64.grads).max()),
                  "grad_max_abs_vs_f32_reference": float(np.abs(g - res.grads).max()),
                  "f32_reference_vs_f64_oracle": float(np.abs(res.grads[:rows64] - o64.grads).max()),
                  "tolerance": "costs 1e-5 relative, gradients 1e-5 absolute (north_star), asserted"}
        assert parity["cost_max_rel_vs_f32_reference"] <= 1e-5, parity
        assert parity["cost_max_rel_vs_f64_oracle"] <= 1e-5, parity
        assert parity["grad_max_abs_vs_f64_oracle"] <= 1e-5, parity

    # ---- secondary: the reference's own CUDA path (unmodified tests/test_time.cu for sm_100a) on this GPU -----
    ref_cuda = None
    ref_cuda_runs = (wl.alignment is None and wl.elements < 2 ** 31 and int(wl.T.min()) == int(wl.T.max())
                     and int(wl.S.min()) == int(wl.S.max()))   # its program takes one (T, S) and indexes with int
    if rank == 0 and world == 1 and not args.no_cpu_baseline and ref_cuda_runs:
        from oracle import oracle
        del acts_h
        r = oracle.run_ref_gpu(wl.B, int(wl.T[0]), int(wl.S[0]), wl.V)
        if r is not None:
            ref_cuda = {"value": wl.B / (r["ms_median"] * 1e-3), "unit": UNIT, "ms_per_call": r["ms_median"],
                        "ms_min": r["ms_min"], "ms_first_call": r["ms_first_call"], "calls": r["calls"],
                        "what": "reference GpuRNNTComputer<float>::cost_and_grad via its own tests/test_time.cu "
                                "(unmodified, nvcc -O2 sm_100a), host clock around each synchronous call, median of "
                                "calls 2..10, its own generated inputs of the same shape"}

    clock_info = clocks.result() if clocks else None
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return

    peak, peak_src = _peaks()
    # K3's job, in bytes per launch: read the logits of the live rows once, write the gradient rows it owns once (all
    # of them, less the zero rows the lattice kernel's fill has already written -- those are that kernel's)
    live_rows = wl.rows - dead_rows
    zero_note = None
    if zero_fill_warps > 0 and zero_fill_warps < 32:
        k3_written = live_rows
        zero_note = "the dead rows' zeros are written by the lattice kernel's fill warps while its recursions run"
    elif zero_fill_warps >= 32 or (wl.alignment is not None and zero_fill_warps == 0):
        # tight alignment band: the zero rows are written by the zero-fill warps of K1, K2 and K3 out of ONE counter
        # (DESIGN 3.2); the split is not known on the host, K3 is charged with none of them (a lower bound)
        k3_written = live_rows
        zero_note = "zero rows written by the fill warps of K1, K2 and K3 from one counter; K3 charged with none of them"
    else:
        k3_written = wl.rows
    k3_bytes = 4 * wl.V * (live_rows + k3_written)
    k3_traffic = _traffic("k3_grad_tma_kernel")   # DRAM bytes per launch from the committed ncu capture of this workload
    if (zero_fill_warps >= 32 or (wl.alignment is not None and zero_fill_warps == 0)) and k3_traffic is not None:
        # how many of the zero rows K3's fill warp took is only known from the counters: the capture's write bytes
        k3_bytes = max(k3_bytes, int(k3_traffic))
        zero_note = ("zero rows written by the fill warps of K1, K2 and K3 from one counter; K3's share taken from the ncu capture "
                     "(profiles/traffic.json): its job is what it moved there")
    k3_gbs = k3_bytes / (k_ms[2] * 1e-3) / 1e9
    k3_dram_gbs = None if k3_traffic is None else k3_traffic / (k_ms[2] * 1e-3) / 1e9
    call_gbs = wl.algorithmic_bytes / (ms_per_step * 1e-3) / 1e9
    # what the call has to move whoever moves it: the live rows' logits twice (K1, K3), every gradient row once
    call_required = 4 * wl.V * (2 * live_rows + wl.rows)
    call_required_gbs = call_required / (ms_per_step * 1e-3) / 1e9
    call_traffic = None
    tj = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tj):
        with open(tj) as f:
            e = json.load(f).get(WORKLOAD, {})
        if all(kname in e for kname in ("k1_lse_tma_kernel", "k2_lattice_kernel", "k3_grad_tma_kernel")):
            call_traffic = e["k1_lse_tma_kernel"] + e["k2_lattice_kernel"] + e["k3_grad_tma_kernel"]
    # (the captures are of the whole named batch on one GPU: not comparable with a shard's step time)
    call_dram_gbs = None if (call_traffic is None or strong) else call_traffic / (ms_per_step * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": SCALING,
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": _config(wl_global if strong else wl, {"collective": collective, "scaling": SCALING,
                                                        "batch_total": B_job, "call": call_semantics}),
        "timing": {"blocks": len(block_ms), "steps_per_block": args.steps, "reported": "median block, max over ranks",
                   "ms_per_step_blocks": [b / args.steps for b in block_ms],
                   "ms_per_step_min_block": min(block_ms) / args.steps, "ms_per_step_max_block": max(block_ms) / args.steps,
                   "ms_per_step_per_rank_median_block": [x / args.steps for x in block_rank_ms[mid]],
                   "wall_ms_blocks": wall_ms},
        "shards": shard_info,
        "collective_check": collective_check,
        "roofline": {"kernel": "k3_grad_tma_kernel", "bound": "hbm", "achieved": k3_gbs, "peak": peak, "unit": "GB/s",
                     "frac": k3_gbs / peak, "traffic": k3_traffic, "peak_source": peak_src,
                     "dram_achieved": k3_dram_gbs, "dram_frac": None if k3_dram_gbs is None else k3_dram_gbs / peak,
                     "note": "achieved = the bytes the kernel's job requires per launch (4 V (live rows read + gradient rows "
                             "written by this kernel)) / this run's launch duration; traffic = DRAM bytes ncu counted per "
                             "launch of this workload (profiles/traffic.json, isolated cold-cache run); dram_frac = traffic / "
                             "this run's duration / peak.  Rank 0's shard under --scaling strong.",
                     "algorithmic_bytes_per_launch": k3_bytes, "ms_per_launch": float(k_ms[2]),
                     "live_rows": live_rows, "dead_rows": dead_rows, "gradient_rows_written_by_this_kernel": k3_written,
                     "k2_zero_fill_warps": zero_fill_warps, "zero_rows_note": zero_note,
                     "survey_8d_convention": {"bytes": 2 * 4 * n, "GBps": 2 * 4 * n / (k_ms[2] * 1e-3) / 1e9,
                                              "note": "8 bytes per logit whether the row is ever read or not"}},
        "kernels_ms": {"k1_lse_gather": float(k_ms[0]), "k2_lattice": float(k_ms[1]), "k3_grad": float(k_ms[2]),
                       "sum": float(k_ms.sum()), "k1_GBps_of_live_logits": 4 * wl.V * live_rows / (k_ms[0] * 1e-3) / 1e9},
        "call_roofline": {"algorithmic_bytes_survey_8d": wl.algorithmic_bytes, "achieved_GBps_survey_8d": call_gbs,
                          "frac_of_measured_peak_survey_8d": call_gbs / peak, "frac_of_8TBps_nominal_survey_8d": call_gbs / 8000.0,
                          "required_bytes": call_required, "required_GBps": call_required_gbs,
                          "required_frac_of_measured_peak": call_required_gbs / peak,
                          "required_note": "4 V (2 live rows read + all gradient rows written): dead rows are never read; "
                                           "the whole synchronous call, the lattice kernel's time included",
                          "dram_traffic_bytes": call_traffic,
                          "dram_GBps": call_dram_gbs,
                          "dram_frac_of_measured_peak": None if call_dram_gbs is None else call_dram_gbs / peak,
                          "per_gpu": True},
        "e2e": e2e_best, "e2e_paths": e2e_paths,
        "async_enqueue": None if async_ms is None else {
            "value": wl.B / (async_ms / 1000.0), "unit": UNIT, "ms_per_step": async_ms,
            "what": "the same steps through mrnnt_enqueue (no host synchronisation per step, one at the end)"},
        "full_wait": None if full_wait_ms is None else {
            "value": wl.B / (full_wait_ms / 1000.0), "unit": UNIT, "ms_per_step": full_wait_ms,
            "what": "the same steps with MRNNT_OPT_RETURN_EARLY 0: every call waits for its gradient kernel before it returns"},
        "per_call_workspace": None if alloc_ms is None else {
            "value": wl.B / (alloc_ms / 1000.0), "unit": UNIT, "ms_per_step": alloc_ms,
            "what": "new handle + cudaMalloc of the workspace + set-up kernels + cost_and_grad + cudaFree per call (host clock), "
                    "the way the reference's torch binding drives its manager"},
        "gpu_launches": gpu_launches,
        "gpu_launches_note": "kernel launches counted by the engine (MRNNT_OPT_LAUNCH_COUNT) inside the reported block of K steps, rank 0",
        "clocks": clock_info, "wall_ms_timed_region": wall_ms[mid],
        "cpu_baseline": cpu_baseline, "reference_cuda_same_gpu": ref_cuda, "parity": parity,
        "build": lib.mrnnt_build_info().decode(),
    }
    _emit(args.out_fd, line)


def _claim_stdout() -> int:
    """Rank 0 prints ONE JSON line on stdout; libraries (NCCL prints its version there) get stderr instead."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    return real


def _emit(fd: int, line: dict) -> None:
    sys.stdout.flush()
    os.write(fd, (json.dumps(line) + "\n").encode())


def main() -> None:
    global WORKLOAD
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--workload", choices=sorted(WORKLOAD_NAMES), default=WORKLOAD,
                    help="the named shape (BASELINE.json configs[1..4]); the contract's bench line is c2, the default")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the host-CPU baseline/checker leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (profiling runs under ncu only)")
    ap.add_argument("--full-wait", action="store_true",
                    help="every call waits for all of its kernels before it returns (MRNNT_OPT_RETURN_EARLY 0)")
    ap.add_argument("--collective", choices=["fused", "nccl"], default="fused",
                    help="N > 1: the sum of the costs over peer memory inside the gradient kernel, or NCCL on a side stream")
    ap.add_argument("--reserve-sms", type=int, default=0,
                    help="N > 1: SMs the gradient kernel leaves to the concurrent all-reduce")
    ap.add_argument("--scaling", choices=["weak", "strong"], default="weak",
                    help="weak: the named batch on every GPU; strong: ONE batch sharded over the GPUs by utterance")
    ap.add_argument("--partition", choices=["lpt", "contiguous"], default="lpt",
                    help="--scaling strong: longest-processing-time assignment of single utterances, or contiguous ranges")
    ap.add_argument("--blocks", type=int, default=5, help="how often the K-step timed region is run (median reported)")
    args = ap.parse_args()
    global SCALING
    SCALING = args.scaling
    WORKLOAD = args.workload
    args.out_fd = _claim_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
