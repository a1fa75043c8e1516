"""Batch sharding for the multi-GPU path (one process per GPU).

Utterances are independent and the packed layout makes a shard a contiguous slice, so the path
partitions with NO data-path collective; the only exchange is one all-reduce of the summed cost
(north_star).  Shards are contiguous utterance ranges balancing sum_b T_b*(S_b+1)*V
(the streamed bytes), not utterance counts.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np


def partition_contiguous(T: Sequence[int], S: Sequence[int], world_size: int) -> List[Tuple[int, int]]:
    """Contiguous [b0, b1) ranges, one per rank, minimising the largest range's row count greedily.

    Every rank gets at least one utterance when B >= world_size; with B < world_size the tail ranks get
    empty ranges.
    """
    T = np.asarray(T, dtype=np.int64)
    S = np.asarray(S, dtype=np.int64)
    B = int(T.shape[0])
    w = T * (S + 1)
    cum = np.concatenate([[0], np.cumsum(w)])
    total = int(cum[-1])
    bounds = [0]
    for r in range(1, world_size):
        target = total * r / world_size
        b = int(np.searchsorted(cum, target, side="left"))
        # choose the closer of the two neighbouring cut points
        if b > 0 and abs(cum[b - 1] - target) <= abs(cum[min(b, B)] - target):
            b -= 1
        lo = bounds[-1] + 1 if B - bounds[-1] > world_size - r else bounds[-1]
        hi = B - (world_size - r) if B >= world_size else B
        b = max(min(b, hi), min(lo, B))
        bounds.append(b)
    bounds.append(B)
    return [(bounds[i], bounds[i + 1]) for i in range(world_size)]


@dataclass
class Shard:
    b0: int
    b1: int
    row0: int            # first packed row of the shard in the global acts
    row1: int
    T: np.ndarray        # int32 [b1-b0]
    S: np.ndarray
    labels: np.ndarray   # int32 [b1-b0, S_max(shard)]  (the ABI derives the stride from max(S) it is given)
    alignment: Optional[np.ndarray]  # int32 [b1-b0, T_max(shard)]


def make_shard(T, S, labels, b0: int, b1: int, alignment=None) -> Shard:
    """Slice host-side metadata for utterances [b0, b1) and re-stride labels/alignment to the shard's maxima.

    The reference ABI indexes labels with stride max_b S_b and alignments with stride max_b T_b OF THE
    ARRAYS IT IS GIVEN (cpu_workspace_manager.h:44,122,208), so a shard must carry its own strides.
    """
    T = np.asarray(T, dtype=np.int32)
    S = np.asarray(S, dtype=np.int32)
    labels = np.asarray(labels, dtype=np.int32)
    rows = T.astype(np.int64) * (S.astype(np.int64) + 1)
    cum = np.concatenate([[0], np.cumsum(rows)])
    Ts, Ss = T[b0:b1].copy(), S[b0:b1].copy()
    s_max = max(int(Ss.max()) if len(Ss) else 0, 1)
    lab = np.zeros((b1 - b0, s_max), dtype=np.int32)
    w = min(s_max, labels.shape[1])
    lab[:, :w] = labels[b0:b1, :w]
    al = None
    if alignment is not None:
        alignment = np.asarray(alignment, dtype=np.int32)
        t_max = int(Ts.max()) if len(Ts) else 0
        al = np.ascontiguousarray(alignment[b0:b1, :t_max])
    return Shard(b0, b1, int(cum[b0]), int(cum[b1]), Ts, Ss, np.ascontiguousarray(lab), al)


def allreduce_cost_sum(costs_dev):
    """Sum of the per-utterance costs over all ranks: ONE all-reduce of one float (NCCL on GPUs, gloo in
    the CPU tests).  Returns a 0-d tensor on costs_dev.device."""
    import torch
    import torch.distributed as dist

    total = costs_dev.sum(dtype=torch.float32)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(total, op=dist.ReduceOp.SUM)
    return total
