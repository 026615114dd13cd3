"""Shared helpers for the parity tests: build a DB with the CPU oracle, query it with both sides."""
from __future__ import annotations

import os
from typing import List, Sequence, Tuple

import numpy as np


def oracle_build_db(oracle, genomes: Sequence[Tuple[str, bytes]], k: int, directory: str, fpr: float = 0.001,
                    largest: int = 10_000, seeds=(0x5EED0001, 0x5EED0002), rot: int = 26):
    t = oracle.Tree(k, fpr, largest, seeds[0], seeds[1], rot=rot)
    for gid, seq in genomes:
        t.insert(gid, seq)
    t.save(directory)
    return t


def gpu_query(tree, reads: Sequence[bytes], threshold: float):
    """Returns (list of frozenset of DFS leaf indices per read) through pf_query_block."""
    from phagefilter_b200.query import PackedReads, query_packed
    p = PackedReads(list(reads))
    try:
        off, leaf = query_packed(tree, p, threshold, want_hits=True)
    finally:
        p.close()
    return [frozenset(int(x) for x in leaf[int(off[i]):int(off[i + 1])]) for i in range(len(reads))]


def random_genomes(rng: np.random.Generator, n: int, lo: int, hi: int, alphabet: bytes = b"ACGT") -> List[Tuple[str, bytes]]:
    a = np.frombuffer(alphabet, dtype=np.uint8)
    return [(f"g{i}", a[rng.integers(0, len(a), size=int(rng.integers(lo, hi + 1)))].tobytes()) for i in range(n)]


def sample_reads(rng: np.random.Generator, genomes, n: int, length: int, err: float) -> List[bytes]:
    out = []
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    for _ in range(n):
        g = genomes[int(rng.integers(0, len(genomes)))][1]
        if len(g) < length:
            out.append(g)
            continue
        s = int(rng.integers(0, len(g) - length + 1))
        r = np.frombuffer(g[s:s + length], dtype=np.uint8).copy()
        m = rng.random(length) < err
        r[m] = acgt[rng.integers(0, 4, size=int(m.sum()))]
        out.append(r.tobytes())
    return out
