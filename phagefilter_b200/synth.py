"""Deterministic synthetic phage genomes and simulated reads.

Genomes: `n_families` ancestors of i.i.d. uniform ACGT with length U[len_lo, len_hi]; every family has
`family_size` members, each the ancestor with `divergence` substitutions per base.
Reads follow the reference's simulator model (benchmarking/bench/simulate_reads.py:28-48): uniform start,
fixed length, substitution-only errors, replacement base uniform over ACGT (may equal the original).
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def make_genomes(n_families: int, family_size: int, seed: int, len_lo: int = 40_000, len_hi: int = 60_000,
                 divergence: float = 0.05) -> List[Tuple[str, bytes]]:
    rng = np.random.Generator(np.random.PCG64(seed))
    out: List[Tuple[str, bytes]] = []
    for f in range(n_families):
        n = int(rng.integers(len_lo, len_hi + 1))
        anc = rng.integers(0, 4, size=n, dtype=np.uint8)
        for m in range(family_size):
            g = anc.copy()
            if m > 0 or family_size == 1:
                mask = rng.random(n) < divergence
                g[mask] = rng.integers(0, 4, size=int(mask.sum()), dtype=np.uint8)
            out.append((f"SYN_{f:05d}_{m:02d}", _ACGT[g].tobytes()))
    return out


def simulate_reads(genomes: List[Tuple[str, bytes]], n_reads: int, read_len: int, seed: int,
                   error_rates: Tuple[float, ...] = (0.0, 0.01), background_frac: float = 0.0,
                   chunk: int = 100_000) -> Tuple[np.ndarray, np.ndarray]:
    """Returns (reads uint8 [n_reads, read_len] of ASCII bytes, source genome index or -1 for background).
    Read i uses error_rates[i % len(error_rates)]."""
    rng = np.random.Generator(np.random.PCG64(seed))
    cat = np.frombuffer(b"".join(g for _, g in genomes), dtype=np.uint8)
    lens = np.array([len(g) for _, g in genomes], dtype=np.int64)
    starts = np.concatenate([[0], np.cumsum(lens)[:-1]])
    ok = np.nonzero(lens >= read_len)[0]
    reads = np.empty((n_reads, read_len), dtype=np.uint8)
    src = np.empty(n_reads, dtype=np.int64)
    rates = np.array(error_rates, dtype=np.float64)
    ar = np.arange(read_len, dtype=np.int64)
    for lo in range(0, n_reads, chunk):
        hi = min(lo + chunk, n_reads)
        n = hi - lo
        gi = ok[rng.integers(0, len(ok), size=n)]
        st = (rng.random(n) * (lens[gi] - read_len + 1)).astype(np.int64)
        blk = cat[(starts[gi] + st)[:, None] + ar[None, :]].copy()
        er = rates[(np.arange(lo, hi) % len(rates))]
        mask = rng.random((n, read_len)) < er[:, None]
        blk[mask] = _ACGT[rng.integers(0, 4, size=int(mask.sum()), dtype=np.uint8)]
        bg = rng.random(n) < background_frac
        if bg.any():
            blk[bg] = _ACGT[rng.integers(0, 4, size=(int(bg.sum()), read_len), dtype=np.uint8)]
            gi = gi.copy()
            gi[bg] = -1
        reads[lo:hi] = blk
        src[lo:hi] = gi
    return reads, src


def reads_to_concat(reads: np.ndarray) -> Tuple[bytes, np.ndarray]:
    n, L = reads.shape
    offs = np.arange(n + 1, dtype=np.uint64) * np.uint64(L)
    return reads.tobytes(), offs


def simulate_reads_fast(genomes: List[Tuple[str, bytes]], n_reads: int, read_len: int, seed: int,
                        error_rates: Tuple[float, ...] = (0.0, 0.01), background_frac: float = 0.0,
                        chunk: int = 250_000) -> Tuple[np.ndarray, np.ndarray]:
    """Same model and return values as simulate_reads, drawn in a cheaper order (a different random stream): which reads
    are background is decided first, background reads are filled directly with uniform bases, and only the reads that
    come from a genome pay for the gather and the substitution mask.  For large spike-in workloads (BASELINE configs
    3-5: 90 % background) this is about 5x faster."""
    rng = np.random.Generator(np.random.PCG64(seed))
    cat = np.frombuffer(b"".join(g for _, g in genomes), dtype=np.uint8)
    lens = np.array([len(g) for _, g in genomes], dtype=np.int64)
    starts = np.concatenate([[0], np.cumsum(lens)[:-1]])
    ok = np.nonzero(lens >= read_len)[0]
    reads = np.empty((n_reads, read_len), dtype=np.uint8)
    src = np.full(n_reads, -1, dtype=np.int64)
    rates = np.array(error_rates, dtype=np.float64)
    ar = np.arange(read_len, dtype=np.int64)
    for lo in range(0, n_reads, chunk):
        hi = min(lo + chunk, n_reads)
        n = hi - lo
        bg = rng.random(n) < background_frac
        blk = reads[lo:hi]
        n_bg = int(bg.sum())
        if n_bg:
            blk[bg] = _ACGT[rng.integers(0, 4, size=(n_bg, read_len), dtype=np.uint8)]
        fg = np.nonzero(~bg)[0]
        if len(fg):
            gi = ok[rng.integers(0, len(ok), size=len(fg))]
            st = (rng.random(len(fg)) * (lens[gi] - read_len + 1)).astype(np.int64)
            sub = cat[(starts[gi] + st)[:, None] + ar[None, :]].copy()
            er = rates[(lo + fg) % len(rates)]
            mask = rng.random((len(fg), read_len)) < er[:, None]
            sub[mask] = _ACGT[rng.integers(0, 4, size=int(mask.sum()), dtype=np.uint8)]
            blk[fg] = sub
            src[lo + fg] = gi
    return reads, src
