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


def reference_query_outputs(oracle, db_dir, records, theta, block_size, pos, neg, search_depth=None):
    """The reference's query driver (src/main.rs:249-376) restated with the oracle as engine.
    records: list of DNASequence (id, sequence, quality).  Returns (classification_csv, pos_records, neg_records)
    where pos/neg records are lists of (header_id, genome_set, seq, qual) / (id, seq, qual)."""
    tree = oracle.Tree.load(db_dir)
    if search_depth is not None:
        tree.prune_tree(search_depth)
    ids = tree.leaf_ids()
    filtering = pos or neg
    pos_out, neg_out = [], []
    for lo in range(0, len(records), block_size):
        blk = records[lo:lo + block_size]
        res = tree.query_batch([r.sequence for r in blk], theta, want_hits=True)
        if not filtering:
            continue
        result_map = {}  # read id -> set of genome ids (ResultMap, keyed by id, cleared per block)
        for r, l in res.hits:
            result_map.setdefault(blk[int(r)].id, set()).add(ids[int(l)])
        for rec in blk:
            seq = rec.sequence.upper()
            if rec.id in result_map:
                if pos:
                    pos_out.append((rec.id, frozenset(result_map[rec.id]), seq, rec.quality))
            elif neg:
                neg_out.append((rec.id, seq, rec.quality))
    return tree.classification_csv(), pos_out, neg_out


def parse_filter_file(path):
    """POS/NEG_FILTERING.{fa,fq} -> list of (id, genome frozenset or None, seq, qual)."""
    out = []
    with open(path, "rb") as f:
        lines = f.read().split(b"\n")
    i = 0
    while i < len(lines):
        h = lines[i]
        if not h:
            i += 1
            continue
        if h.startswith(b"@"):
            seq, qual = lines[i + 1], lines[i + 3]
            i += 4
        else:
            seq, qual = lines[i + 1], None
            i += 2
        head = h[1:].decode()
        if " |" in head:
            rid, g = head.split(" |", 1)
            out.append((rid, frozenset(x for x in g.split(",") if x), seq, qual))
        else:
            out.append((head, None, seq, qual))
    return out
