"""Synthetic workloads: the named configurations of BASELINE.json and the counter-based generator.

Logits: x[i] = (splitmix64(seed ^ i) >> 40) * 2^-24, i = global element index -- U[0,1) like the
reference's genActs (tests/random.cpp:4-20) but reproducible per element, so a GPU shard, the host
oracle and a test can all regenerate the same bits without moving data.  The device twin is
mrnnt_synth_uniform in csrc/c_api.cu.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = (x.astype(np.uint64) + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


def uniform_logits(n: int, seed: int = 0, index_offset: int = 0) -> np.ndarray:
    """Host twin of mrnnt_synth_uniform (bit-identical)."""
    idx = np.arange(index_offset, index_offset + n, dtype=np.uint64)
    u = splitmix64(np.uint64(seed) ^ idx)
    return ((u >> np.uint64(40)).astype(np.float32) * np.float32(1.0 / 16777216.0)).astype(np.float32)


def labels_for(B: int, S_max: int, V: int, seed: int = 1) -> np.ndarray:
    """Never-blank labels 1 + splitmix64(seed ^ (b*S_max+s)) % (V-1), like genLabels (random.cpp:22-30)."""
    idx = np.arange(B * max(S_max, 1), dtype=np.uint64)
    lab = 1 + (splitmix64(np.uint64(seed) ^ idx) % np.uint64(max(V - 1, 1))).astype(np.int64)
    return lab.reshape(B, max(S_max, 1)).astype(np.int32)


@dataclass
class Workload:
    name: str
    B: int
    V: int
    T: np.ndarray            # int32 [B]
    S: np.ndarray            # int32 [B]
    labels: np.ndarray       # int32 [B, S_max]
    alignment: Optional[np.ndarray] = None   # int32 [B, T_max]
    max_shift: int = 0
    blank: int = 0
    logits_seed: int = 0

    @property
    def rows(self) -> int:
        return int((self.T.astype(np.int64) * (self.S.astype(np.int64) + 1)).sum())

    @property
    def elements(self) -> int:
        return self.rows * self.V

    @property
    def algorithmic_bytes(self) -> int:
        """2 x read logits + 1 x write grads (BASELINE.json / SURVEY 8d)."""
        return 3 * 4 * self.elements

    def describe(self) -> str:
        return (f"{self.name}: B={self.B} T<={int(self.T.max())} S<={int(self.S.max())} V={self.V} "
                f"rows={self.rows} logits={self.elements * 4 / 1e9:.3f} GB")


def _alignment(T: np.ndarray, S: np.ndarray, labels: np.ndarray, seed: int, blank: int = 0) -> np.ndarray:
    rng = np.random.default_rng(seed)
    B, T_max = len(T), int(T.max())
    al = np.full((B, T_max), blank, dtype=np.int32)
    for b in range(B):
        frames = np.sort(rng.choice(int(T[b]), size=int(S[b]), replace=False))
        al[b, frames] = labels[b, : int(S[b])]
    return al


def workload(name: str, B: Optional[int] = None) -> Workload:
    """The configurations named in BASELINE.json (`configs`), optionally with a different batch size."""
    if name == "c2":      # synthetic B=32 T=150 S=40 V=1000, fixed lengths (the headline shape)
        B = B or 32
        T = np.full(B, 150, np.int32); S = np.full(B, 40, np.int32); V = 1000
        return Workload("c2", B, V, T, S, labels_for(B, 40, V))
    if name == "c3":      # B=64 T<=400 S<=80 V=1024, random per-utterance lengths, packed
        B = B or 64
        rng = np.random.default_rng(1234)
        T = rng.integers(200, 401, size=B).astype(np.int32)
        S = rng.integers(40, 81, size=B).astype(np.int32)
        T[0], S[0] = 400, 80
        V = 1024
        return Workload("c3", B, V, T, S, labels_for(B, int(S.max()), V))
    if name == "c4":      # large vocabulary B=8 T=800 S=120 V=5000 (> 2^31 logits: 64-bit offsets)
        B = B or 8
        T = np.full(B, 800, np.int32); S = np.full(B, 120, np.int32); V = 5000
        return Workload("c4", B, V, T, S, labels_for(B, 120, V))
    if name == "c5":      # alignment-restricted B=32 T=300 S=60 V=2000, max distance 5
        B = B or 32
        T = np.full(B, 300, np.int32); S = np.full(B, 60, np.int32); V = 2000
        labels = labels_for(B, 60, V)
        return Workload("c5", B, V, T, S, labels, _alignment(T, S, labels, seed=7), max_shift=5)
    if name == "c2v1025":  # c2 with a vocabulary that is not a multiple of 4: rows are not whole 16-byte vectors
        B = B or 32
        T = np.full(B, 150, np.int32); S = np.full(B, 40, np.int32); V = 1025
        return Workload("c2v1025", B, V, T, S, labels_for(B, 40, V))
    if name == "c4v5001":  # c4 likewise (real vocabularies: 1025, 5001, ...)
        B = B or 8
        T = np.full(B, 800, np.int32); S = np.full(B, 120, np.int32); V = 5001
        return Workload("c4v5001", B, V, T, S, labels_for(B, 120, V))
    if name == "dense":   # tools only: the size of c2 with (almost) no dead rows, to separate streaming efficiency
        B = B or 32       # from the effect of skipping rows
        T = np.full(B, 1500, np.int32); S = np.full(B, 3, np.int32); V = 1000
        return Workload("dense", B, V, T, S, labels_for(B, 3, V))
    raise KeyError(name)
