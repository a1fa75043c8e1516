"""Shared test inputs.

The literal cases restate the fixtures of the reference's own tests
(/root/reference/tests/test_cpu.cpp: fwd/bwd/grads_test :10-192, multibatch_test
:194-295, align_restrict_test :335-438, align_restrict_multibatch_test :440-552)
so that the parity tests read like the reference's.  Random cases use seeded numpy
generators and are stored, with the reference's outputs, under tests/golden/.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

import numpy as np

README_PROBS = np.array([
    # t = 0
    0.6, 0.3, 0.1,  0.7, 0.1, 0.2,  0.5, 0.1, 0.4,
    # t = 1
    0.5, 0.4, 0.1,  0.5, 0.1, 0.4,  0.8, 0.1, 0.1,
    # t = 2
    0.4, 0.3, 0.3,  0.5, 0.1, 0.4,  0.7, 0.2, 0.1,
    # t = 3
    0.8, 0.1, 0.1,  0.3, 0.1, 0.6,  0.8, 0.1, 0.1,
], dtype=np.float32)

# the utterance b=0 of multibatch_test (T=2, S=1)
MULTIBATCH_B0_PROBS = np.array([
    0.6, 0.3, 0.1,  0.7, 0.1, 0.2,
    0.5, 0.4, 0.1,  0.5, 0.1, 0.4,
], dtype=np.float32)

# tests/test_cpu.cpp:163-183 (2-decimal README gradients, tolerance 1e-2)
README_GRADS_2DP = np.array([
    0.04, -0.14, 0.1,   0.0, 0.0, 0.0,     0.0, 0.0, 0.0,
    0.13, -0.19, 0.06,  -0.04, 0.04, -0.01, 0.0, 0.0, 0.0,
    0.06, -0.1, 0.04,   0.01, 0.07, -0.08, -0.06, 0.04, 0.02,
    0.0, 0.0, 0.0,      0.14, 0.05, -0.19, -0.11, 0.05, 0.05,
], dtype=np.float32)

# tests/test_cpu.cpp:253-264 (utterance 0 of multibatch_test)
MULTIBATCH_B0_GRADS_2DP = np.array([
    -0.02, -0.08, 0.1,  0.0, 0.0, 0.0,
    0.31, -0.37, 0.06,  -0.19, 0.04, 0.15,
], dtype=np.float32)


def logf(p: np.ndarray) -> np.ndarray:
    """std::log on floats, as the reference tests build their logits (test_cpu.cpp:47-48)."""
    return np.log(p.astype(np.float32)).astype(np.float32)


@dataclass
class Case:
    name: str
    acts: np.ndarray                 # float32 [rows, V] packed
    labels: np.ndarray               # int32 [B, S_max]
    T: np.ndarray                    # int32 [B]
    S: np.ndarray                    # int32 [B]
    V: int
    blank: int = 0
    alignment: Optional[np.ndarray] = None   # int32 [B, T_max]
    max_shift: int = 0
    expect_costs: Optional[np.ndarray] = None    # hand-derived values from the reference tests
    meta: dict = field(default_factory=dict)

    @property
    def B(self) -> int:
        return int(self.T.shape[0])

    @property
    def rows(self) -> int:
        return int((self.T.astype(np.int64) * (self.S.astype(np.int64) + 1)).sum())

    def with_alignment(self, alignment, max_shift: int, name: Optional[str] = None, expect_costs=None) -> "Case":
        return Case(name or f"{self.name}_shift{max_shift}", self.acts, self.labels, self.T, self.S, self.V,
                    self.blank, np.asarray(alignment, dtype=np.int32), int(max_shift), expect_costs, dict(self.meta))


def readme_case() -> Case:
    return Case("readme", logf(README_PROBS).reshape(12, 3), np.array([[1, 2]], np.int32),
                np.array([4], np.int32), np.array([2], np.int32), 3,
                expect_costs=np.array([-np.log(0.363)], np.float32))


def multibatch_case() -> Case:
    acts = np.concatenate([logf(MULTIBATCH_B0_PROBS), logf(README_PROBS)]).reshape(16, 3)
    return Case("multibatch", acts, np.array([[1, 0], [1, 2]], np.int32), np.array([2, 4], np.int32),
                np.array([1, 2], np.int32), 3,
                expect_costs=np.array([-np.log(0.39), -np.log(0.363)], np.float32))


def align_cases() -> list[Case]:
    base = readme_case()
    al = [[0, 1, 0, 2]]
    return [
        base.with_alignment(al, 2, "align_shift2", np.array([-np.log(0.363)], np.float32)),
        base.with_alignment(al, 0, "align_shift0", np.array([-np.log(0.072)], np.float32)),
        base.with_alignment(al, 1, "align_shift1", np.array([-np.log(0.2958)], np.float32)),
    ]


def align_multibatch_cases() -> list[Case]:
    acts = np.concatenate([logf(README_PROBS), logf(README_PROBS)]).reshape(24, 3)
    base = Case("align_mb", acts, np.array([[1, 2], [1, 2]], np.int32), np.array([4, 4], np.int32),
                np.array([2, 2], np.int32), 3)
    al = [[0, 1, 0, 2], [1, 2, 0, 0]]
    return [
        base.with_alignment(al, 3, "align_mb_shift3", np.array([-np.log(0.363)] * 2, np.float32)),
        base.with_alignment(al, 0, "align_mb_shift0", np.array([-np.log(0.072), -np.log(0.0672)], np.float32)),
        base.with_alignment(al, 1, "align_mb_shift1", np.array([-np.log(0.2958), -np.log(0.192)], np.float32)),
    ]


def random_alignment(rng: np.random.Generator, T: np.ndarray, S: np.ndarray, labels: np.ndarray,
                     blank: int = 0) -> np.ndarray:
    """A valid alignment per utterance: S_b distinct emission frames carrying the labels, blank elsewhere."""
    B, T_max = len(T), int(T.max())
    al = np.full((B, T_max), blank, dtype=np.int32)
    for b in range(B):
        frames = np.sort(rng.choice(int(T[b]), size=int(S[b]), replace=False))
        for k, f in enumerate(frames):
            lab = int(labels[b, k])
            al[b, f] = lab if lab != blank else (blank + 1)
    return al


def random_case(name: str, seed: int, B: int, V: int, T_range=(3, 20), S_range=(0, 8), dist: str = "normal3",
                blank: int = 0, force=None) -> Case:
    """Seeded ragged batch.  `force` optionally pins (T_b, S_b) of the leading utterances."""
    rng = np.random.default_rng(seed)
    T = rng.integers(T_range[0], T_range[1] + 1, size=B).astype(np.int32)
    S = np.array([rng.integers(S_range[0], min(S_range[1], int(t)) + 1) for t in T], dtype=np.int32)
    if force:
        for b, (t, s) in enumerate(force):
            T[b], S[b] = t, s
    S_max = max(int(S.max()), 1)
    labels = rng.integers(0, V, size=(B, S_max)).astype(np.int32)
    # mostly non-blank labels, as in genLabels (tests/random.cpp:22-30), but keep a few blanks:
    # the reference's branch order (blank before label, cpu_rnnt.h:224-232) must be honoured.
    mask = rng.random((B, S_max)) < 0.9
    nonblank = rng.integers(1, V, size=(B, S_max)).astype(np.int32)
    nonblank = np.where(nonblank == blank, (blank + 1) % V, nonblank)
    labels = np.where(mask, nonblank, labels).astype(np.int32)
    rows = int((T.astype(np.int64) * (S + 1)).sum())
    if dist == "normal3":
        acts = (3.0 * rng.standard_normal((rows, V))).astype(np.float32)
    elif dist == "uniform":
        acts = rng.random((rows, V), dtype=np.float32)
    else:
        raise ValueError(dist)
    return Case(name, acts, labels, T, S, V, blank, meta={"seed": seed, "dist": dist})


def golden_random_cases() -> list[Case]:
    """The seeded cases whose reference outputs are stored in tests/golden/."""
    cases = []
    c = random_case("rand_v17", 11, B=4, V=17)
    cases.append(c)
    rng = np.random.default_rng(111)
    al = random_alignment(rng, c.T, c.S, c.labels)
    cases.append(c.with_alignment(al, 1, "rand_v17_shift1"))
    cases.append(c.with_alignment(al, 0, "rand_v17_shift0"))
    cases.append(c.with_alignment(al, 4, "rand_v17_shift4"))
    c = random_case("rand_v32", 12, B=5, V=32, T_range=(8, 40), S_range=(0, 12), dist="uniform")
    cases.append(c)
    # edge shapes: T==S (single path), S==0 (blank only), T==1
    c = random_case("rand_edges", 13, B=5, V=8, force=[(6, 6), (5, 0), (1, 0), (1, 1), (9, 3)])
    cases.append(c)
    # blank not at index 0
    c = random_case("rand_blank5", 14, B=3, V=12, blank=5)
    cases.append(c)
    # wider than one warp of label states (S+1 > 32) and V % 4 == 0
    c = random_case("rand_wide", 15, B=2, V=20, T_range=(60, 70), S_range=(40, 50))
    cases.append(c)
    rng = np.random.default_rng(151)
    cases.append(c.with_alignment(random_alignment(rng, c.T, c.S, c.labels), 3, "rand_wide_shift3"))
    return cases


def literal_cases() -> list[Case]:
    return [readme_case(), multibatch_case()] + align_cases() + align_multibatch_cases()
