"""Development aid: the whole call in the stream (no per-kernel events, programmatic dependent launches on, early
return) for a list of MRNNT_OPT_K2_FILL_SHARE values -- how much of the zero fill the lattice kernel should keep.
    python tools/share_sweep.py c2 c3 --shares 100,80,65,50 [--shard 0/8] [--steps 100]
"""
from __future__ import annotations

import argparse
import dataclasses
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import monotonic_rnnt_b200 as mr  # noqa: E402
from monotonic_rnnt_b200 import _lib  # noqa: E402


def run(name, shares, steps, shard, zero, reps, opts=()):
    wl = mr.synth.workload(name)
    if shard:
        r, n = (int(x) for x in shard.split("/"))
        idx = mr.shard.partition_lpt(wl.T, wl.S, n)[r]
        sh = mr.shard.make_shard_indexed(wl.T, wl.S, wl.labels, idx, alignment=wl.alignment)
        wl = dataclasses.replace(wl, name=f"{wl.name}[{shard}]", B=len(idx), T=sh.T, S=sh.S, labels=sh.labels, alignment=sh.alignment)
    dev = torch.device("cuda", 0)
    lib = _lib.load()
    acts = torch.empty((wl.rows, wl.V), dtype=torch.float32, device=dev)
    _lib.check(lib.mrnnt_synth_uniform(acts.data_ptr(), wl.elements, wl.logits_seed, 0, torch.cuda.current_stream().cuda_stream), "synth")
    h = mr.LossHandle(acts, torch.from_numpy(wl.labels).to(dev), torch.from_numpy(wl.T).to(dev), torch.from_numpy(wl.S).to(dev),
                      lengths_host=(wl.T, wl.S))
    if wl.alignment is not None:
        h.restrict_to_alignment(torch.from_numpy(wl.alignment).to(dev), wl.max_shift, wl.blank)
    grads = torch.empty_like(acts)
    costs = torch.empty(wl.B, dtype=torch.float32).pin_memory()
    h.set_option(_lib.OPT_K2_ZERO_FILL, zero)
    for k, v in opts:
        h.set_option(k, v)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ref = None
    for share in shares:
        h.set_option(_lib.OPT_K2_FILL_SHARE, share)
        for _ in range(5):
            h.cost_and_grad(wl.blank, grads, costs)
        torch.cuda.synchronize()
        if ref is None:
            ref = (costs.clone(), grads.clone())
        else:
            assert torch.equal(costs, ref[0]) and torch.equal(grads, ref[1]), share
        ts = []
        for _ in range(reps):
            torch.cuda.synchronize()
            ev0.record()
            for _ in range(steps):
                h.cost_and_grad(wl.blank, grads, costs)
            ev1.record()
            torch.cuda.synchronize()
            ts.append(ev0.elapsed_time(ev1) / steps * 1e3)
        print(f"{wl.name} B={wl.B} zero={zero} opts={list(opts)} share={share:4d} (used {h.get_option(_lib.OPT_K2_FILL_SHARE):3d}, fill warps "
              f"{h.get_option(_lib.OPT_K2_ZERO_FILL)}): call {np.median(ts):8.1f} us (min {min(ts):8.1f})  {wl.B / np.median(ts) * 1e6:9.0f} utt/s", flush=True)
    h.close()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("workloads", nargs="*", default=["c2"])
    ap.add_argument("--shares", default="100,80,65,50")
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--shard", default="")
    ap.add_argument("--zero", type=int, default=-1)
    ap.add_argument("--opt", action="append", default=[], help="ID=VALUE: mrnnt_set_option before the sweep (repeatable)")
    a = ap.parse_args()
    for name in a.workloads:
        run(name, [int(x) for x in a.shares.split(",")], a.steps, a.shard, a.zero, a.reps,
            [tuple(int(x) for x in o.split("=")) for o in a.opt])
