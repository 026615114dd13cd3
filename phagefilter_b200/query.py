"""query_batch / get_leaf_counts / save_leaf_counts (reference: src/query.rs:66-218) on the GPU."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from .bloom_tree import BloomTree
from .file_parser import DNASequence
from .result_map import ResultMap


class PackedReads:
    """Host batch in the 2-bit layout of pf_read_batch (pinned when a GPU is present)."""

    def __init__(self, reads: Sequence[bytes]):
        self.n_reads = len(reads)
        offs = np.zeros(self.n_reads + 1, dtype=np.uint64)
        if self.n_reads:
            offs[1:] = np.cumsum(np.fromiter((len(r) for r in reads), dtype=np.uint64, count=self.n_reads))
        blob = b"".join(reads)
        self._h = C.c_void_p()
        _lib.check(_lib.lib().pf_pack_reads(blob, offs.ctypes.data_as(C.POINTER(C.c_uint64)), self.n_reads,
                                            C.byref(self._h)))
        self.batch = _lib.lib().pf_packed_batch(self._h)

    @classmethod
    def from_concat(cls, blob: bytes, offs: np.ndarray) -> "PackedReads":
        self = cls.__new__(cls)
        self.n_reads = len(offs) - 1
        offs = np.ascontiguousarray(offs, dtype=np.uint64)
        self._h = C.c_void_p()
        _lib.check(_lib.lib().pf_pack_reads(blob, offs.ctypes.data_as(C.POINTER(C.c_uint64)), self.n_reads,
                                            C.byref(self._h)))
        self.batch = _lib.lib().pf_packed_batch(self._h)
        return self

    def nbytes(self) -> int:
        b = self.batch.contents
        n = b.n_reads * 12 + b.n_words * 4
        if b.n_exc:
            n += b.n_reads * 4 + (b.n_exc + 1) * 8 + int(b.exc_off[b.n_exc])
        return int(n)

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            _lib.lib().pf_packed_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _hits_to_numpy(h: _lib.Hits, n_reads: int) -> Tuple[np.ndarray, np.ndarray]:
    off = np.ctypeslib.as_array(h.read_off, shape=(n_reads + 1,)).copy() if n_reads >= 0 and h.read_off else \
        np.zeros(n_reads + 1, dtype=np.uint64)
    leaf = np.ctypeslib.as_array(h.leaf, shape=(int(h.n_hits),)).copy() if h.n_hits else np.zeros(0, dtype=np.uint32)
    return off, leaf


def query_packed(tree: BloomTree, packed: PackedReads, threshold: float, want_hits: bool = True
                 ) -> Tuple[np.ndarray, np.ndarray]:
    """One pf_query_block call: returns the CSR (read_off, leaf) of matched DFS leaf indices."""
    h = _lib.Hits()
    _lib.check(_lib.lib().pf_query_block(tree._h, packed.batch, C.c_float(threshold), int(want_hits), C.byref(h)))
    return _hits_to_numpy(h, packed.n_reads)


def query_sharded(tree: BloomTree, packed: PackedReads, threshold: float, want_hits: bool = True
                  ) -> Tuple[np.ndarray, np.ndarray]:
    """pf_query_sharded: collective over the ranks of a subtree-sharded tree; every rank passes its own block
    (possibly empty) and gets the CSR of its own reads."""
    h = _lib.Hits()
    _lib.check(_lib.lib().pf_query_sharded(tree._h, packed.batch, C.c_float(threshold), int(want_hits), C.byref(h)))
    return _hits_to_numpy(h, packed.n_reads)


def query_batch(bloom_tree: BloomTree, read_set: Sequence[DNASequence], threshold: float,
                result_map: Optional[ResultMap]) -> BloomTree:
    """query.rs:66-82.  Leaf counters accumulate inside the tree across calls; result_map receives
    (read id -> genome id) for every leaf a read passes, only when the block carries sequences
    (`first_read.sequence.is_some()`, query.rs:147-153 -- here: result_map given)."""
    if not read_set:
        return bloom_tree
    seqs = [r.sequence or b"" for r in read_set]
    packed = PackedReads(seqs)
    try:
        want = result_map is not None
        off, leaf = query_packed(bloom_tree, packed, threshold, want_hits=want)
        if want:
            ids = bloom_tree.leaf_ids()
            for r, read in enumerate(read_set):
                for j in range(int(off[r]), int(off[r + 1])):
                    result_map.add_read_map(read.id, ids[int(leaf[j])])
    finally:
        packed.close()
    return bloom_tree


def get_leaf_counts(bloom_tree: BloomTree) -> List[Tuple[str, int]]:
    """query.rs:197-218: (tax id, mapped reads) in left-first DFS leaf order."""
    return list(zip(bloom_tree.leaf_ids(), (int(c) for c in bloom_tree.leaf_counts())))


def save_leaf_counts(bloom_tree: BloomTree, output_path: str) -> None:
    """query.rs:173-183: "{id},{count}\\n" for leaves with count > 0."""
    _lib.check(_lib.lib().pf_save_leaf_counts(bloom_tree._h, output_path.encode()))
