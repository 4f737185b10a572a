"""Synthetic workloads of BASELINE.json's configs: frozen seeds, numpy
default_rng(seed).integers(0, 4) -> ACGT, one sequence per row, rows end in '\\n' (seq_t layout).
Shared by bench.py and the tests; not part of the product."""
from __future__ import annotations

import numpy as np

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def _rows(rng, count, length):
    rows = np.empty((count, length + 1), dtype=np.uint8)
    rows[:, :length] = ACGT[rng.integers(0, 4, size=(count, length), dtype=np.uint8)]
    rows[:, length] = 10
    return rows


def _mutated(rng, base, count, max_subs):
    length = base.shape[0]
    rows = np.empty((count, length + 1), dtype=np.uint8)
    rows[:, :length] = base
    rows[:, length] = 10
    k = rng.integers(0, max_subs + 1, size=count)
    for j in range(max_subs):
        sel = np.nonzero(k > j)[0]
        pos = rng.integers(0, length, size=sel.shape[0])
        rows[sel, pos] = ACGT[rng.integers(0, 4, size=sel.shape[0])]
    return rows


CONFIGS = {
    # name: (algo name, query_len, subject_len, full subject count, seed, extra)
    "C2": dict(algo="bitpal_packed", qlen=150, slen=150, count=1_000_000, seed=12345),
    "C3": dict(algo="banded", qlen=100, slen=100, count=10_000_000, seed=777, threshold=5),
    # C3 with its two halves interleaved at random: the same subjects, order-independence check of the banded early exit
    "C3s": dict(algo="banded", qlen=100, slen=100, count=10_000_000, seed=777, threshold=5),
    "C4": dict(algo="myers_semiglobal", qlen=1000, slen=1000, count=1_000_000, seed=4),
    "C5": dict(algo="bitpal_packed", qlen=5000, slen=5000, count=1_000_000, seed=5),
}


def shuffled(subjects, name: str = "C3s"):
    """The rows of `subjects` in the (seeded) random order of config `name`."""
    return np.ascontiguousarray(subjects[np.random.default_rng(CONFIGS[name]["seed"] + 1).permutation(subjects.shape[0])])


def make(name: str, count: int | None = None, shard: int = 0):
    """Returns (query rows [1, qlen+1], subject rows [count, slen+1]) for a config; `count`
    truncates the subject set (the generator is sequential, so the first `count` subjects of the
    full set are reproduced only for the iid configs C2/C4/C5; C3 keeps its 50/50 mix at any size).
    shard > 0 draws the subjects from a different stream (the query stays the config's)."""
    cfg = CONFIGS[name]
    n = cfg["count"] if count is None else count
    rng = np.random.default_rng(cfg["seed"])
    query = _rows(rng, 1, cfg["qlen"])
    if shard:      # another shard of the same workload (multi-GPU weak scaling): same query, same recipe, fresh subjects
        rng = np.random.default_rng(cfg["seed"] * 1000 + shard)
    if name in ("C3", "C3s"):
        half = n // 2
        similar = _mutated(rng, query[0, : cfg["qlen"]], half, 8)
        rest = _rows(rng, n - half, cfg["slen"])
        subjects = np.concatenate([similar, rest])
        if name == "C3s":
            subjects = shuffled(subjects, name)
    elif name == "C4":
        subjects = _rows(rng, n, cfg["slen"])
        # 1 % planted query substrings so that the semi-global minima vary
        planted = rng.choice(n, size=max(1, n // 100), replace=False)
        for idx in planted:
            a = int(rng.integers(0, cfg["qlen"] - 200))
            ln = int(rng.integers(100, 200))
            b = int(rng.integers(0, cfg["slen"] - ln))
            subjects[idx, b:b + ln] = query[0, a:a + ln]
    else:
        subjects = _rows(rng, n, cfg["slen"])
    return query, subjects
