"""ctypes binding of the CPU ORACLE (oracle/pf_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
PARITY UNPINNED at the bitvec-serde / bincode (file format) boundaries; the rustc-hash arithmetic is pinned against
a real rustc-hash 2.x build by tests/test_hash_pin_cpu.py (see pf_oracle.h).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Iterable, List, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libpforacle.so")
DEFAULT_ROT = 26


def build(force: bool = False) -> str:
    """Compile oracle/libpforacle.so with the committed Makefile."""
    src = [os.path.join(_HERE, f) for f in ("pf_oracle.c", "pf_oracle.h", "Makefile")]
    stale = force or not os.path.exists(_LIB_PATH) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in src
    )
    if stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "libpforacle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


class _QueryResult(C.Structure):
    _fields_ = [
        ("n_hits", C.c_uint64),
        ("hit_read", C.POINTER(C.c_uint32)),
        ("hit_leaf", C.POINTER(C.c_uint32)),
        ("pairs", C.c_uint64),
        ("probes_ref", C.c_uint64),
        ("probes_sched", C.c_uint64),
    ]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_LIB_PATH)
    u8p, u64p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint64)
    L.pfo_hash_bytes.restype = C.c_uint64
    L.pfo_hash_bytes.argtypes = [C.c_char_p, C.c_size_t]
    L.pfo_fx_hash.restype = C.c_uint64
    L.pfo_fx_hash.argtypes = [C.c_uint64, C.c_char_p, C.c_size_t, C.c_int]
    L.pfo_hash_iter.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, u64p]
    L.pfo_complement.restype = C.c_uint8
    L.pfo_complement.argtypes = [C.c_uint8]
    L.pfo_get_lex_less.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p]
    L.pfo_revcomp.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p]
    L.pfo_num_kmers.restype = C.c_size_t
    L.pfo_num_kmers.argtypes = [C.c_size_t, C.c_size_t]
    L.pfo_get_kmers.argtypes = [C.c_char_p, C.c_size_t, C.c_size_t, C.c_char_p]
    L.pfo_needed_bits.restype = C.c_uint64
    L.pfo_needed_bits.argtypes = [C.c_float, C.c_uint32]
    L.pfo_optimal_num_hashes.restype = C.c_uint32
    L.pfo_optimal_num_hashes.argtypes = [C.c_uint64, C.c_uint32]
    L.pfo_filter_new.restype = C.c_void_p
    L.pfo_filter_new.argtypes = [C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint64]
    L.pfo_filter_free.argtypes = [C.c_void_p]
    L.pfo_filter_insert.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.c_int]
    L.pfo_filter_contains.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.c_int, u64p]
    L.pfo_filter_union.argtypes = [C.c_void_p, C.c_void_p]
    L.pfo_filter_distance.restype = C.c_uint64
    L.pfo_filter_distance.argtypes = [C.c_void_p, C.c_void_p]
    L.pfo_filter_save.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
    L.pfo_filter_load.restype = C.c_void_p
    L.pfo_filter_load.argtypes = [C.c_char_p]
    L.pfo_tree_new.restype = C.c_void_p
    L.pfo_tree_new.argtypes = [C.c_uint64, C.c_float, C.c_uint32, C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_uint64]
    L.pfo_tree_free.argtypes = [C.c_void_p]
    L.pfo_tree_insert.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_size_t]
    L.pfo_tree_save.argtypes = [C.c_void_p, C.c_char_p]
    L.pfo_tree_load.restype = C.c_void_p
    L.pfo_tree_load.argtypes = [C.c_char_p, C.c_int]
    L.pfo_tree_prune.argtypes = [C.c_void_p, C.c_uint64]
    L.pfo_last_error.restype = C.c_char_p
    for name in ("pfo_tree_num_nodes", "pfo_tree_num_leaves", "pfo_tree_kmer_size", "pfo_tree_num_bits"):
        getattr(L, name).restype = C.c_uint64
        getattr(L, name).argtypes = [C.c_void_p]
    L.pfo_tree_num_hashes.restype = C.c_uint32
    L.pfo_tree_num_hashes.argtypes = [C.c_void_p]
    L.pfo_tree_seeds.argtypes = [C.c_void_p, u64p, u64p]
    L.pfo_tree_leaf_id.restype = C.c_char_p
    L.pfo_tree_leaf_id.argtypes = [C.c_void_p, C.c_uint64]
    L.pfo_tree_leaf_count.restype = C.c_uint64
    L.pfo_tree_leaf_count.argtypes = [C.c_void_p, C.c_uint64]
    L.pfo_tree_reset_counts.argtypes = [C.c_void_p]
    L.pfo_tree_preorder.restype = C.c_uint64
    L.pfo_tree_preorder.argtypes = [C.c_void_p, u8p, C.POINTER(C.c_uint32), C.c_uint64]
    L.pfo_tree_node_name.restype = C.c_char_p
    L.pfo_tree_node_name.argtypes = [C.c_void_p, C.c_uint64]
    L.pfo_query_batch.argtypes = [C.c_void_p, C.c_char_p, u64p, C.c_uint32, C.c_float, C.c_int, C.c_int,
                                  C.POINTER(_QueryResult)]
    L.pfo_query_result_free.argtypes = [C.POINTER(_QueryResult)]
    L.pfo_query_blocks.argtypes = [C.c_void_p, C.c_char_p, u64p, C.c_uint32, C.c_float, C.c_int, C.c_int, C.c_uint32,
                                   C.POINTER(_QueryResult)]
    L.pfo_tree_load_lazy.restype = C.c_void_p
    L.pfo_tree_load_lazy.argtypes = [C.c_char_p, C.c_int, C.c_int]
    L.pfo_tree_cache_stats.argtypes = [C.c_void_p, u64p, u64p, u64p]
    L.pfo_query_batch_sched.argtypes = [C.c_void_p, C.c_char_p, u64p, C.c_uint32, C.c_float, C.c_int, C.c_int,
                                        C.c_int, C.c_uint32, C.POINTER(_QueryResult)]
    L.pfo_need.restype = C.c_uint64
    L.pfo_need.argtypes = [C.c_float, C.c_uint64]
    L.pfo_classification_csv.restype = C.c_uint64
    L.pfo_classification_csv.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64]
    _lib = L
    return L


# ---- thin functional wrappers -------------------------------------------------------------
def hash_bytes(b: bytes) -> int:
    return lib().pfo_hash_bytes(b, len(b))


def fx_hash(seed: int, item: bytes, rot: int = DEFAULT_ROT) -> int:
    return lib().pfo_fx_hash(seed, item, len(item), rot)


def hash_iter(h1: int, h2: int, count: int) -> List[int]:
    out = (C.c_uint64 * max(count, 1))()
    lib().pfo_hash_iter(h1, h2, count, out)
    return [int(out[i]) for i in range(count)]


def get_lex_less(kmer: bytes) -> bytes:
    out = C.create_string_buffer(len(kmer))
    lib().pfo_get_lex_less(kmer, len(kmer), out)
    return out.raw


def revcomp(kmer: bytes) -> bytes:
    out = C.create_string_buffer(len(kmer))
    lib().pfo_revcomp(kmer, len(kmer), out)
    return out.raw


def get_kmers(seq: bytes, k: int) -> List[bytes]:
    n = lib().pfo_num_kmers(len(seq), k)
    out = C.create_string_buffer(max(n * k, 1))
    lib().pfo_get_kmers(seq, len(seq), k, out)
    return [out.raw[i * k:(i + 1) * k] for i in range(n)]


def needed_bits(fpr: float, n: int) -> int:
    return lib().pfo_needed_bits(fpr, n)


def optimal_num_hashes(bits: int, n: int) -> int:
    return lib().pfo_optimal_num_hashes(bits, n)


def need(threshold: float, n_kmers: int) -> int:
    return lib().pfo_need(threshold, n_kmers)


class Filter:
    """BloomFilter (bloom_filter.rs:84-93)."""

    def __init__(self, m: int = 0, K: int = 0, seed1: int = 0, seed2: int = 0, _ptr=None, rot: int = DEFAULT_ROT):
        self.rot = rot
        self._p = _ptr if _ptr is not None else lib().pfo_filter_new(m, K, seed1, seed2)
        if not self._p:
            raise MemoryError("pfo_filter_new")

    @classmethod
    def load(cls, path: str, rot: int = DEFAULT_ROT) -> "Filter":
        p = lib().pfo_filter_load(path.encode())
        if not p:
            raise RuntimeError(lib().pfo_last_error().decode())
        return cls(_ptr=p, rot=rot)

    def _s(self):
        class S(C.Structure):
            _fields_ = [("m", C.c_uint64), ("nwords", C.c_uint64), ("words", C.POINTER(C.c_uint64)),
                        ("K", C.c_uint32), ("seed1", C.c_uint64), ("seed2", C.c_uint64)]
        return C.cast(self._p, C.POINTER(S)).contents

    @property
    def m(self):
        return int(self._s().m)

    @property
    def K(self):
        return int(self._s().K)

    @property
    def seeds(self):
        s = self._s()
        return int(s.seed1), int(s.seed2)

    def words(self) -> np.ndarray:
        s = self._s()
        return np.ctypeslib.as_array(s.words, shape=(int(s.nwords),)).copy()

    def insert(self, item: bytes) -> bool:
        return bool(lib().pfo_filter_insert(self._p, item, len(item), self.rot))

    def contains(self, item: bytes) -> bool:
        return bool(lib().pfo_filter_contains(self._p, item, len(item), self.rot, None))

    def union(self, other: "Filter") -> None:
        lib().pfo_filter_union(self._p, other._p)

    def distance(self, other: "Filter") -> int:
        return int(lib().pfo_filter_distance(self._p, other._p))

    def save(self, path: str, recorded_path: str | None = None) -> None:
        if lib().pfo_filter_save(self._p, path.encode(), (recorded_path or path).encode()):
            raise RuntimeError(lib().pfo_last_error().decode())

    def __del__(self):
        if getattr(self, "_p", None):
            lib().pfo_filter_free(self._p)
            self._p = None


def concat_reads(reads: Sequence[bytes]) -> Tuple[bytes, np.ndarray]:
    offs = np.zeros(len(reads) + 1, dtype=np.uint64)
    if len(reads):
        offs[1:] = np.cumsum([len(r) for r in reads], dtype=np.uint64)
    return b"".join(reads), offs


class QueryResult:
    def __init__(self, hits: np.ndarray, pairs: int, probes_ref: int, probes_sched: int):
        self.hits = hits  # (n,2) uint32: read index, DFS leaf index
        self.pairs = pairs
        self.probes_ref = probes_ref
        self.probes_sched = probes_sched

    def hit_sets(self, n_reads: int) -> List[frozenset]:
        out = [set() for _ in range(n_reads)]
        for r, l in self.hits:
            out[int(r)].add(int(l))
        return [frozenset(s) for s in out]


class Tree:
    """BloomTree (bloom_tree.rs:28-48) with every filter resident in host memory."""

    def __init__(self, kmer_size: int, fpr: float = 0.001, largest_genome: int = 1_000_000, seed1: int = 0x5EED0001,
                 seed2: int = 0x5EED0002, rot: int = DEFAULT_ROT, name_mode: int = 0, name_seed: int = 0, _ptr=None):
        self._p = _ptr if _ptr is not None else lib().pfo_tree_new(kmer_size, fpr, largest_genome, seed1, seed2, rot,
                                                                   name_mode, name_seed)
        if not self._p:
            raise MemoryError("pfo_tree_new")

    @classmethod
    def load(cls, directory: str, rot: int = DEFAULT_ROT) -> "Tree":
        p = lib().pfo_tree_load(directory.encode(), rot)
        if not p:
            raise RuntimeError(lib().pfo_last_error().decode())
        return cls(0, _ptr=p)

    @classmethod
    def load_lazy(cls, directory: str, cache_size: int = 10, rot: int = DEFAULT_ROT) -> "Tree":
        """Reference-faithful filter store: only tree.bin is read, filters come through an LRU of `cache_size` entries
        and every miss re-reads the .bf file (BFLruCache, cache.rs:55-88; --cache-size default 10, main.rs:119-122)."""
        p = lib().pfo_tree_load_lazy(directory.encode(), rot, cache_size)
        if not p:
            raise RuntimeError(lib().pfo_last_error().decode())
        return cls(0, _ptr=p)

    def cache_stats(self) -> Tuple[int, int, int]:
        """(filter files loaded, cache hits, filter bytes decoded) of a load_lazy tree."""
        a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
        lib().pfo_tree_cache_stats(self._p, C.byref(a), C.byref(b), C.byref(c))
        return int(a.value), int(b.value), int(c.value)

    def query_blocks(self, reads: Sequence[bytes], threshold: float, block_size: int = 100, threads: int = 0,
                     want_hits: bool = True, concat: Tuple[bytes, np.ndarray] | None = None) -> QueryResult:
        """The reference's driver loop (main.rs:334-368): query_batch once per block of `block_size` reads
        (--block-size-reads, default 100); hit read indices are relative to the whole set."""
        seqs, offs = concat if concat is not None else concat_reads(reads)
        n = len(offs) - 1
        res = _QueryResult()
        rc = lib().pfo_query_blocks(self._p, seqs, offs.ctypes.data_as(C.POINTER(C.c_uint64)), n, C.c_float(threshold),
                                    threads, int(want_hits), block_size, C.byref(res))
        if rc:
            raise RuntimeError(lib().pfo_last_error().decode())
        nh = int(res.n_hits)
        hits = np.zeros((nh, 2), dtype=np.uint32)
        if nh:
            hits[:, 0] = np.ctypeslib.as_array(res.hit_read, shape=(nh,))
            hits[:, 1] = np.ctypeslib.as_array(res.hit_leaf, shape=(nh,))
        out = QueryResult(hits, int(res.pairs), int(res.probes_ref), int(res.probes_sched))
        lib().pfo_query_result_free(C.byref(res))
        return out

    def insert(self, genome_id: str, seq: bytes) -> None:
        lib().pfo_tree_insert(self._p, genome_id.encode(), seq, len(seq))

    def save(self, directory: str) -> None:
        if lib().pfo_tree_save(self._p, directory.encode()):
            raise RuntimeError(lib().pfo_last_error().decode())

    def prune_tree(self, search_depth: int) -> None:
        lib().pfo_tree_prune(self._p, search_depth)

    @property
    def kmer_size(self) -> int:
        return int(lib().pfo_tree_kmer_size(self._p))

    @property
    def num_bits(self) -> int:
        return int(lib().pfo_tree_num_bits(self._p))

    @property
    def num_hashes(self) -> int:
        return int(lib().pfo_tree_num_hashes(self._p))

    @property
    def seeds(self) -> Tuple[int, int]:
        a, b = C.c_uint64(), C.c_uint64()
        lib().pfo_tree_seeds(self._p, C.byref(a), C.byref(b))
        return int(a.value), int(b.value)

    @property
    def num_nodes(self) -> int:
        return int(lib().pfo_tree_num_nodes(self._p))

    @property
    def num_leaves(self) -> int:
        return int(lib().pfo_tree_num_leaves(self._p))

    def leaf_ids(self) -> List[str]:
        return [lib().pfo_tree_leaf_id(self._p, i).decode() for i in range(self.num_leaves)]

    def leaf_counts(self) -> List[Tuple[str, int]]:
        """get_leaf_counts (query.rs:197-218): DFS leaf order."""
        return [(lib().pfo_tree_leaf_id(self._p, i).decode(), int(lib().pfo_tree_leaf_count(self._p, i)))
                for i in range(self.num_leaves)]

    def reset_counts(self) -> None:
        lib().pfo_tree_reset_counts(self._p)

    def preorder(self) -> List[Tuple[str, bool, int]]:
        n = self.num_nodes
        leaf = (C.c_uint8 * max(n, 1))()
        depth = (C.c_uint32 * max(n, 1))()
        lib().pfo_tree_preorder(self._p, leaf, depth, n)
        return [(lib().pfo_tree_node_name(self._p, i).decode(), bool(leaf[i]), int(depth[i])) for i in range(n)]

    def query_batch(self, reads: Sequence[bytes], threshold: float, threads: int = 0, want_hits: bool = True,
                    concat: Tuple[bytes, np.ndarray] | None = None) -> QueryResult:
        """query::query_batch (query.rs:66-82) on one block; leaf counters accumulate in the tree."""
        seqs, offs = concat if concat is not None else concat_reads(reads)
        n = len(offs) - 1
        res = _QueryResult()
        rc = lib().pfo_query_batch(self._p, seqs, offs.ctypes.data_as(C.POINTER(C.c_uint64)), n,
                                   C.c_float(threshold), threads, int(want_hits), C.byref(res))
        if rc:
            raise RuntimeError(lib().pfo_last_error().decode())
        nh = int(res.n_hits)
        hits = np.zeros((nh, 2), dtype=np.uint32)
        if nh:
            hits[:, 0] = np.ctypeslib.as_array(res.hit_read, shape=(nh,))
            hits[:, 1] = np.ctypeslib.as_array(res.hit_leaf, shape=(nh,))
        out = QueryResult(hits, int(res.pairs), int(res.probes_ref), int(res.probes_sched))
        lib().pfo_query_result_free(C.byref(res))
        return out

    def query_sched(self, reads: Sequence[bytes], threshold: float, lazy: bool = True, threads: int = 0,
                    concat: Tuple[bytes, np.ndarray] | None = None, group_rounds: int | None = None) -> QueryResult:
        """The GPU kernel's schedule restated on the CPU (same decisions as query_batch, different work).
        group_rounds: rounds of 32 k-mers a lane owns at once; default = the kernel's rule
        clamp(ceil(max k-mers per read / 32), 1, 8)."""
        seqs, offs = concat if concat is not None else concat_reads(reads)
        n = len(offs) - 1
        if group_rounds is None:
            k = self.kmer_size
            lens = np.diff(offs.astype(np.int64)) if n else np.zeros(0, dtype=np.int64)
            max_nk = int(max(0, (lens.max() - k + 1) if n and k and lens.max() >= k else 0))
            group_rounds = min(8, max(1, -(-max_nk // 32)))
        res = _QueryResult()
        rc = lib().pfo_query_batch_sched(self._p, seqs, offs.ctypes.data_as(C.POINTER(C.c_uint64)), n,
                                         C.c_float(threshold), threads, 1, int(lazy), group_rounds, C.byref(res))
        if rc:
            raise RuntimeError(lib().pfo_last_error().decode())
        nh = int(res.n_hits)
        hits = np.zeros((nh, 2), dtype=np.uint32)
        if nh:
            hits[:, 0] = np.ctypeslib.as_array(res.hit_read, shape=(nh,))
            hits[:, 1] = np.ctypeslib.as_array(res.hit_leaf, shape=(nh,))
        out = QueryResult(hits, int(res.pairs), int(res.probes_ref), int(res.probes_sched))
        lib().pfo_query_result_free(C.byref(res))
        return out

    def classification_csv(self) -> str:
        """save_leaf_counts (query.rs:173-183)."""
        n = lib().pfo_classification_csv(self._p, None, 0)
        buf = C.create_string_buffer(int(n) + 1)
        lib().pfo_classification_csv(self._p, buf, n)
        return buf.raw[:n].decode()

    def __del__(self):
        if getattr(self, "_p", None):
            lib().pfo_tree_free(self._p)
            self._p = None

