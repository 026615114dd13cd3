"""A second, independent restatement of the reference's build + query path in plain Python (test infrastructure).

Written from the reference's sources only -- not from oracle/pf_oracle.c -- so that the two restatements check each
other (tests/test_pyref_differential_cpu.py): canonical k-mers (file_parser.rs:114-148 with bio's complement table),
FxHasher over Vec<u8> (hasher.rs:12-21, hash_iter.rs:13-45; rustc-hash 2.x per SURVEY App. A), filter geometry in f32
(bloom_filter.rs:342-357), insert / contains / union / distance (bloom_filter.rs:142-150,275-332), the greedy tree
insert (bloom_tree.rs:128-246), prune_tree (:302-330), the recursive descent (query.rs:38-158) and the leaf report
(query.rs:173-218).  Pure Python: only for small cases.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Set, Tuple

import numpy as np

from tests.test_oracle_cpu import py_fx_hash

MASK = (1 << 64) - 1
_PAIRS = "AT CG YR WW SS KM DH VB NN"  # bio::alphabets::dna::complement: IUPAC pairs, case preserved, the rest maps to itself
COMP = list(range(256))
for _a, _b in (p for p in _PAIRS.split()):
    for x, y in ((_a, _b), (_b, _a), (_a.lower(), _b.lower()), (_b.lower(), _a.lower())):
        COMP[ord(x)] = ord(y)


def get_lex_less(kmer: bytes) -> bytes:
    rc = bytes(COMP[c] for c in reversed(kmer))
    return rc if rc < kmer else kmer  # Ordering::Greater => revcomp; Less and Equal => forward (file_parser.rs:116-120)


def get_kmers(seq: bytes, k: int) -> List[bytes]:
    if k > len(seq) or k == 0:
        return []
    return [get_lex_less(seq[i:i + k]) for i in range(len(seq) - k + 1)]


def needed_bits(fpr: float, n: int) -> int:
    f = np.float32
    ln2 = f(0.6931471805599453)
    ln22 = f(ln2 * ln2)
    return int(np.rint(f(f(n) * f(np.log(f(f(1.0) / f(fpr))) / ln22))))  # f32::round is half away from zero; no ties here


def optimal_num_hashes(bits: int, n: int) -> int:
    f = np.float32
    v = int(np.floor(f(f(f(bits) / f(n)) * f(0.6931471805599453)) + f(0.5)))
    return min(max(v, 2), 200)


_H12: Dict[tuple, Tuple[int, int]] = {}


class Filter:
    def __init__(self, m: int, k_hashes: int, seeds: Tuple[int, int], rot: int):
        self.m, self.K, self.seeds, self.rot = m, k_hashes, seeds, rot
        self.bits = bytearray(m)  # one byte per bit: simple beats fast here

    def _indices(self, item: bytes):
        key = (self.seeds, self.rot, item)
        hh = _H12.get(key)
        if hh is None:  # the two hashes depend on the item only: computed once per distinct k-mer (speed, not semantics)
            hh = _H12[key] = (py_fx_hash(self.seeds[0], item, self.rot), py_fx_hash(self.seeds[1], item, self.rot))
        h1, h2 = hh
        for i in range(self.K):
            g = h1 if i == 0 else h2 if i == 1 else (((h1 + i) & MASK) * h2) & MASK
            yield g % self.m

    def insert(self, item: bytes) -> None:
        for idx in self._indices(item):
            self.bits[idx] = 1

    def contains(self, item: bytes) -> bool:
        return all(self.bits[idx] for idx in self._indices(item))

    def union(self, other: "Filter") -> None:
        self.bits = bytearray(a | b for a, b in zip(self.bits, other.bits))

    def distance(self, other: "Filter") -> int:
        return sum(a ^ b for a, b in zip(self.bits, other.bits))


class Node:
    def __init__(self, tax_id: str, flt: Filter):
        self.tax_id, self.filter = tax_id, flt
        self.left: Optional[Node] = None
        self.right: Optional[Node] = None
        self.mapped_reads = 0

    def is_leaf(self) -> bool:
        return self.left is None and self.right is None


class Tree:
    def __init__(self, k: int, fpr: float, largest: int, seed1: int, seed2: int, rot: int = 26):
        self.k, self.seeds, self.rot = k, (seed1, seed2), rot
        self.m = needed_bits(fpr, largest)
        self.K = optimal_num_hashes(self.m, largest)
        self.root: Optional[Node] = None
        self._n_internal = 0

    def _new_filter(self) -> Filter:
        return Filter(self.m, self.K, self.seeds, self.rot)

    def insert(self, genome_id: str, seq: bytes) -> None:  # bloom_tree.rs:128-163
        leaf = Node(genome_id, self._new_filter())
        for kmer in get_kmers(seq, self.k):
            leaf.filter.insert(kmer)
        self.root = leaf if self.root is None else self._add(self.root, leaf)

    def _add(self, cur: Node, node: Node) -> Node:  # add_to_tree, bloom_tree.rs:187-214
        if cur.left is not None and cur.right is not None:
            cur.filter.union(node.filter)
            right_d, left_d = cur.right.filter.distance(node.filter), cur.left.filter.distance(node.filter)
            if right_d < left_d:
                cur.right = self._add(cur.right, node)
            else:
                cur.left = self._add(cur.left, node)
            return cur
        assert cur.is_leaf(), "Node with only one child encountered"
        self._n_internal += 1  # init_internal_node, :226-246 (the name is random there; it does not matter here)
        parent = Node(f"Internal_Node_{self._n_internal}", self._new_filter())
        parent.filter.union(node.filter)
        parent.filter.union(cur.filter)
        parent.left, parent.right = cur, node
        return parent

    def prune_tree(self, depth: int) -> None:  # bloom_tree.rs:302-330: nodes at `depth` lose their children
        def rec(n: Optional[Node], d: int):
            if n is None:
                return
            if d >= depth:
                n.left = n.right = None
            else:
                rec(n.left, d + 1)
                rec(n.right, d + 1)
        rec(self.root, 0)

    def leaves(self) -> List[Node]:  # left-first DFS (query.rs:197-218)
        out: List[Node] = []

        def rec(n: Optional[Node]):
            if n is None:
                return
            if n.is_leaf():
                out.append(n)
            rec(n.left)
            rec(n.right)
        rec(self.root)
        return out

    def query_batch(self, reads: Sequence[bytes], threshold: float) -> Set[Tuple[int, int]]:
        """Returns {(read index, DFS leaf index)}; leaf counters accumulate in the nodes (query.rs:99-158)."""
        kmers = [get_kmers(r, self.k) for r in reads]
        leaf_index = {id(n): i for i, n in enumerate(self.leaves())}
        hits: Set[Tuple[int, int]] = set()
        th = np.float32(threshold)

        def passes(flt: Filter, i: int) -> bool:  # query_passes, query.rs:38-49
            n = sum(1 for km in kmers[i] if flt.contains(km))
            return n >= int(np.ceil(np.float32(th * np.float32(len(kmers[i])))))

        def rec(node: Node, subset: List[int]):
            ok = [i for i in subset if passes(node.filter, i)]
            if not node.is_leaf():
                if ok:
                    if node.left is not None:
                        rec(node.left, ok)
                    if node.right is not None:
                        rec(node.right, ok)
            else:
                node.mapped_reads += len(ok)
                hits.update((i, leaf_index[id(node)]) for i in ok)
        if self.root is not None:
            rec(self.root, list(range(len(reads))))
        return hits

    def classification_csv(self) -> str:  # save_leaf_counts, query.rs:173-183
        return "".join(f"{n.tax_id},{n.mapped_reads}\n" for n in self.leaves() if n.mapped_reads > 0)

    def preorder(self) -> List[Tuple[bool, int]]:
        """(is leaf, depth) in pre-order: the topology, names left out."""
        out: List[Tuple[bool, int]] = []

        def rec(n: Optional[Node], d: int):
            if n is None:
                return
            out.append((n.is_leaf(), d))
            rec(n.left, d + 1)
            rec(n.right, d + 1)
        rec(self.root, 0)
        return out
