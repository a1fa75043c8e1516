"""Per-kernel device times (CUDA events inside the engine) for the named workloads and tuning knobs.

    python tools/kernel_times.py [c2 c3 c5 ...] [--iters 20]

Prints one line per (workload, K1 warps, K3 warps).  Development aid; bench.py is the contract.
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import monotonic_rnnt_b200 as mr  # noqa: E402
from monotonic_rnnt_b200 import _lib  # noqa: E402


def run(name: str, iters: int, combos, padded: bool = False, bf16: bool = False, compact: int = -1,
        zeros=(-1,), dyn: int = -1, shard: str = "", parts: int = 0, share: int = -1) -> None:
    wl = mr.synth.workload(name)
    if shard:   # rank r's utterances of the batch cut over N ranks (LPT, as bench.py --scaling strong does)
        import dataclasses
        r, n = (int(x) for x in shard.split("/"))
        idx = mr.shard.partition_lpt(wl.T, wl.S, n)[r]
        sh = mr.shard.make_shard_indexed(wl.T, wl.S, wl.labels, idx, alignment=wl.alignment)
        wl = dataclasses.replace(wl, name=f"{wl.name}[{shard}]", B=len(idx), T=sh.T, S=sh.S, labels=sh.labels, alignment=sh.alignment)
    dev = torch.device("cuda", 0)
    lib = _lib.load()
    if padded:   # the joint network's own [B, T_max, S_max+1, V] tensor; the padding holds ordinary numbers too
        shape = (wl.B, int(wl.T.max()), int(wl.S.max()) + 1, wl.V)
        acts = torch.empty(shape, dtype=torch.float32, device=dev)
        _lib.check(lib.mrnnt_synth_uniform(acts.data_ptr(), acts.numel(), 0, 0,
                                           torch.cuda.current_stream().cuda_stream), "s")
    else:
        acts = torch.empty((wl.rows, wl.V), dtype=torch.float32, device=dev)
        _lib.check(lib.mrnnt_synth_uniform(acts.data_ptr(), wl.elements, 0, 0,
                                           torch.cuda.current_stream().cuda_stream), "s")
    if bf16:
        acts = acts.to(torch.bfloat16)
    labels = torch.from_numpy(wl.labels).to(dev)
    T = torch.from_numpy(wl.T).to(dev)
    S = torch.from_numpy(wl.S).to(dev)
    grads = torch.empty_like(acts)
    h = mr.LossHandle(acts, labels, T, S, lengths_host=(wl.T, wl.S))
    if wl.alignment is not None:
        h.restrict_to_alignment(torch.from_numpy(wl.alignment).to(dev), wl.max_shift, wl.blank)
    h.set_option(_lib.OPT_TIMING, 1)
    h.set_option(_lib.OPT_K1_COMPACT, compact)
    h.set_option(_lib.OPT_DYNAMIC_TILES, dyn)
    h.set_option(_lib.OPT_K2_PARTS, parts)
    h.set_option(_lib.OPT_K2_FILL_SHARE, share)
    costs = torch.empty(wl.B, dtype=torch.float32).pin_memory()
    n4 = wl.elements * (2 if bf16 else 4)
    for k1w, k3w, zf in [(a, b, z) for a, b in combos for z in zeros]:
        h.set_option(_lib.OPT_K2_ZERO_FILL, zf)
        h.set_option(_lib.OPT_K1_WARPS, k1w)
        h.set_option(_lib.OPT_K3_WARPS, k3w)
        ts, wall = [], []
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for i in range(iters + 3):
            ev0.record()
            h.cost_and_grad(wl.blank, grads, costs)
            ev1.record()
            torch.cuda.synchronize()
            if i >= 3:
                ts.append(h.last_timings())
                wall.append(ev0.elapsed_time(ev1))
        k = np.median(np.array(ts), axis=0)
        w = float(np.median(wall))
        print(f"{wl.name}{' BF16' if bf16 else ''}{' PADDED rows=' + str(acts.numel() // wl.V) if padded else ''} B={wl.B} V={wl.V} rows={wl.rows} "
              f"k1w={k1w} k3w={k3w} zero={zf} dyn={dyn} share={share}: "
              f"K1 {k[0]*1e3:7.1f} us ({n4/k[0]/1e6:6.0f} GB/s of 1xN)  K2 {k[1]*1e3:7.1f} us  "
              f"K3 {k[2]*1e3:7.1f} us ({2*n4/k[2]/1e6:6.0f} GB/s of 2xN)  call {w*1e3:7.1f} us "
              f"({3*n4/w/1e6:6.0f} GB/s of 3xN, {wl.B/w*1e3:8.0f} utt/s)", flush=True)
    # cost-only call (K1 + alpha pass, no beta / coefficients / K3)
    ts = []
    for i in range(8):
        h.cost(wl.blank, costs)
        ts.append(h.last_timings())
    k = np.median(np.array(ts[3:]), axis=0)
    print(f"{wl.name} cost-only: K1 {k[0]*1e3:7.1f} us  K2(alpha only) {k[1]*1e3:7.1f} us", flush=True)
    h.close()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("workloads", nargs="*", default=["c2"])
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--combos", default="24:24")
    ap.add_argument("--bf16", action="store_true", help="bfloat16 logits and gradients")
    ap.add_argument("--compact", type=int, default=-1, help="K1 dead-tile compaction: 1 / 0 forced, -1 automatic")
    ap.add_argument("--padded", action="store_true", help="feed the padded [B,T,S+1,V] tensor instead of packed rows")
    ap.add_argument("--dyn", default="-1", help="comma list of MRNNT_OPT_DYNAMIC_TILES values (0 / 1, -1 automatic)")
    ap.add_argument("--zero", default="-1", help="comma list of MRNNT_OPT_K2_ZERO_FILL values to compare")
    ap.add_argument("--shard", default="", help="r/N: rank r's utterances of the batch cut over N ranks")
    ap.add_argument("--parts", type=int, default=0, help="MRNNT_OPT_K2_PARTS")
    ap.add_argument("--share", default="-1", help="comma list of MRNNT_OPT_K2_FILL_SHARE values")
    a = ap.parse_args()
    combos = [tuple(int(x) for x in c.split(":")) for c in a.combos.split(",")]
    for name in a.workloads:
        for d in [int(x) for x in a.dyn.split(',')]:
            for sh in [int(x) for x in a.share.split(',')]:
                run(name, a.iters, combos, a.padded, a.bf16, a.compact, [int(z) for z in a.zero.split(',')], d, a.shard, a.parts, sh)
