"""Batch sharding for the multi-GPU path (one process per GPU).

Utterances are independent and the packed layout makes an utterance a contiguous slice of rows, so the path
partitions with NO data-path collective; the only exchange is one all-reduce of the summed cost
(north_star).  What is balanced is sum_b T_b*(S_b+1) -- the streamed bytes per V -- not utterance counts.

Two assignments:
  * ``partition_contiguous``: contiguous utterance ranges [b0, b1) with the smallest possible largest range (a shard
    is then ONE slice of the caller's packed tensor: ``acts[row0:row1]``, no gather);
  * ``partition_lpt``: longest-processing-time-first over single utterances, for callers that build each rank's
    packed tensor themselves anyway (a data loader, the synthetic bench): the largest shard is within one utterance's
    weight / world of the mean, where a contiguous cut of 64 ragged utterances over 8 ranks is typically 10-20 % off.

Either way a shard carries its OWN label / alignment strides: the ABI indexes labels with max_b S_b and alignments
with max_b T_b of the arrays it is given (reference cpu_workspace_manager.h:44,117-135,208).
Per-utterance results do not depend on which utterances share a batch: the kernels' arithmetic per row / per lattice
cell is the same whatever the batch (tests/test_gpu_shard.py: sharded == whole batch, bit for bit).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np


def _weights(T: Sequence[int], S: Sequence[int]) -> np.ndarray:
    T = np.asarray(T, dtype=np.int64)
    S = np.asarray(S, dtype=np.int64)
    return T * (S + 1)


def partition_contiguous(T: Sequence[int], S: Sequence[int], world_size: int) -> List[Tuple[int, int]]:
    """Contiguous [b0, b1) ranges, one per rank, with the smallest possible largest row count (exact: binary search on
    the capacity with a greedy feasibility check).

    Every rank gets at least one utterance when B >= world_size; with B < world_size the tail ranks get
    empty ranges.
    """
    w = _weights(T, S)
    B = int(w.shape[0])
    if world_size <= 1 or B == 0:
        return [(0, B)] + [(B, B)] * (max(world_size, 1) - 1)

    def cuts_for(cap: int) -> Optional[List[int]]:
        bounds, load = [0], 0
        for b in range(B):
            if w[b] > cap:
                return None
            if load + w[b] > cap:
                bounds.append(b)
                load = 0
            load += int(w[b])
        bounds.append(B)
        return bounds if len(bounds) - 1 <= world_size else None

    lo, hi = int(w.max()), int(w.sum())
    while lo < hi:
        mid = (lo + hi) // 2
        if cuts_for(mid) is not None:
            hi = mid
        else:
            lo = mid + 1
    bounds = cuts_for(lo)
    assert bounds is not None
    # fewer ranges than ranks: split the heaviest range that still has more than one utterance (never raises the maximum)
    while len(bounds) - 1 < min(world_size, B):
        loads = [(int(w[bounds[i]:bounds[i + 1]].sum()), i) for i in range(len(bounds) - 1) if bounds[i + 1] - bounds[i] > 1]
        _, i = max(loads)
        b0, b1 = bounds[i], bounds[i + 1]
        cum = np.cumsum(w[b0:b1])
        k = int(np.searchsorted(cum, cum[-1] / 2.0, side="left"))
        k = min(max(k + 1, 1), b1 - b0 - 1)
        bounds.insert(i + 1, b0 + k)
    ranges = [(bounds[i], bounds[i + 1]) for i in range(len(bounds) - 1)]
    return ranges + [(B, B)] * (world_size - len(ranges))


def partition_lpt(T: Sequence[int], S: Sequence[int], world_size: int) -> List[np.ndarray]:
    """Longest-processing-time-first: utterances by decreasing row count, each to the least loaded rank.  Returns one
    ascending index array per rank (possibly empty when B < world_size).  Deterministic (ties: lower index first)."""
    w = _weights(T, S)
    order = sorted(range(int(w.shape[0])), key=lambda b: (-int(w[b]), b))
    loads = [0] * world_size
    out: List[List[int]] = [[] for _ in range(world_size)]
    for b in order:
        r = min(range(world_size), key=lambda i: (loads[i], i))
        out[r].append(b)
        loads[r] += int(w[b])
    return [np.array(sorted(x), dtype=np.int64) for x in out]


def imbalance(T: Sequence[int], S: Sequence[int], parts) -> float:
    """Largest shard's row count over the mean (1.0 = perfect).  `parts`: ranges or index arrays."""
    w = _weights(T, S)
    loads = []
    for p in parts:
        idx = np.arange(p[0], p[1]) if isinstance(p, tuple) else np.asarray(p, dtype=np.int64)
        loads.append(int(w[idx].sum()) if len(idx) else 0)
    mean = sum(loads) / max(1, len(loads))
    return max(loads) / mean if mean > 0 else 1.0


@dataclass
class Shard:
    b0: int              # contiguous shards: the range; indexed shards: first / one-past-last utterance index
    b1: int
    row0: int            # contiguous shards: the slice acts[row0:row1] of the global packed tensor
    row1: int
    T: np.ndarray        # int32 [n]
    S: np.ndarray
    labels: np.ndarray   # int32 [n, S_max(shard)]  (the ABI derives the stride from max(S) it is given)
    alignment: Optional[np.ndarray]  # int32 [n, T_max(shard)]
    index: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int64))   # global utterance ids, ascending
    row_ranges: List[Tuple[int, int]] = field(default_factory=list)            # [row0, row1) of each utterance, globally

    @property
    def contiguous(self) -> bool:
        return len(self.index) == 0 or bool((np.diff(self.index) == 1).all())

    @property
    def rows(self) -> int:
        return int(sum(r1 - r0 for r0, r1 in self.row_ranges))


def make_shard_indexed(T, S, labels, index, alignment=None) -> Shard:
    """Host-side metadata of the utterances `index` (ascending global ids) as one batch of their own: lengths, labels
    and alignment re-strided to the shard's maxima, and the global row range of every utterance (what to gather)."""
    T = np.asarray(T, dtype=np.int32)
    S = np.asarray(S, dtype=np.int32)
    labels = np.asarray(labels, dtype=np.int32)
    index = np.asarray(index, dtype=np.int64)
    rows = T.astype(np.int64) * (S.astype(np.int64) + 1)
    cum = np.concatenate([[0], np.cumsum(rows)])
    Ts, Ss = T[index].copy(), S[index].copy()
    s_max = max(int(Ss.max()) if len(Ss) else 0, 1)
    lab = np.zeros((len(index), s_max), dtype=np.int32)
    w = min(s_max, labels.shape[1])
    lab[:, :w] = labels[index, :w]
    al = None
    if alignment is not None:
        alignment = np.asarray(alignment, dtype=np.int32)
        t_max = int(Ts.max()) if len(Ts) else 0
        al = np.ascontiguousarray(alignment[index, :t_max])
    ranges = [(int(cum[b]), int(cum[b + 1])) for b in index]
    b0 = int(index[0]) if len(index) else 0
    b1 = int(index[-1]) + 1 if len(index) else 0
    contiguous = len(index) == 0 or bool((np.diff(index) == 1).all())
    row0 = ranges[0][0] if ranges else 0
    row1 = ranges[-1][1] if (ranges and contiguous) else row0 + sum(r1 - r0 for r0, r1 in ranges)
    return Shard(b0, b1, row0, row1, Ts, Ss, np.ascontiguousarray(lab), al, index, ranges)


def make_shard(T, S, labels, b0: int, b1: int, alignment=None) -> Shard:
    """Slice host-side metadata for utterances [b0, b1) and re-stride labels/alignment to the shard's maxima.

    The reference ABI indexes labels with stride max_b S_b and alignments with stride max_b T_b OF THE
    ARRAYS IT IS GIVEN (cpu_workspace_manager.h:44,122,208), so a shard must carry its own strides.
    """
    return make_shard_indexed(T, S, labels, np.arange(b0, b1, dtype=np.int64), alignment)


def gather_rows(global_rows, shard: Shard):
    """The shard's packed tensor out of the global one ([rows, V] numpy array or torch tensor; a view when contiguous)."""
    if shard.contiguous:
        return global_rows[shard.row0:shard.row1]
    pieces = [global_rows[r0:r1] for r0, r1 in shard.row_ranges]
    if isinstance(global_rows, np.ndarray):
        return np.concatenate(pieces, axis=0)
    import torch
    return torch.cat(pieces, dim=0)


def scatter_rows(local_rows, shard: Shard, global_out) -> None:
    """Inverse of gather_rows: the shard's gradient rows back into the global packed array."""
    off = 0
    for r0, r1 in shard.row_ranges:
        global_out[r0:r1] = local_rows[off:off + (r1 - r0)]
        off += r1 - r0


def allreduce_cost_sum(costs_dev):
    """Sum of the per-utterance costs over all ranks: ONE all-reduce of one float (NCCL on GPUs, gloo in
    the CPU tests).  Returns a 0-d tensor on costs_dev.device."""
    import torch
    import torch.distributed as dist

    total = costs_dev.sum(dtype=torch.float32)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(total, op=dist.ReduceOp.SUM)
    return total
