"""BloomTree -- the flattened, device-resident gSBT behind the reference's BloomTree interface
(src/bloom_tree.rs:28-61, 302-330, 364-386) plus the GPU builder for `build` (:100-299, 339-355)."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Tuple

import numpy as np

from . import _lib


class BloomTree:
    """Query-side tree.  `load` mirrors BloomTree::load(directory, bf_cache); the cache argument is
    accepted and ignored because every filter is resident in HBM (cache.rs is replaced)."""

    def __init__(self, directory: str, device: int = 0, search_depth: Optional[int] = None):
        self.directory = directory
        self.device = device
        self._h = C.c_void_p()
        _lib.check(_lib.lib().pf_db_open(directory.encode(), device, -1 if search_depth is None else search_depth,
                                         C.byref(self._h)))
        self._info = _lib.DbInfo()
        _lib.check(_lib.lib().pf_db_info(self._h, C.byref(self._info)))

    # -- reference interface --------------------------------------------------------------------
    @classmethod
    def load(cls, directory: str, bf_cache=None, device: int = 0) -> "BloomTree":
        """bloom_tree.rs:364-386"""
        return cls(directory, device)

    @classmethod
    def open_sharded(cls, directory: str, device: int, rank: int, nranks: int, nccl_id: bytes,
                     search_depth: Optional[int] = None, cut_level: Optional[int] = None) -> "BloomTree":
        """Subtree-sharded tree for databases larger than one GPU's HBM (pf_db_open_sharded): collective over
        `nranks` processes; levels above the cut are replicated, each subtree below it lives on one rank."""
        self = cls.__new__(cls)
        self.directory, self.device = directory, device
        self._h = C.c_void_p()
        _lib.check(_lib.lib().pf_db_open_sharded(directory.encode(), device, -1 if search_depth is None else search_depth,
                                                 nranks, rank, nccl_id, -1 if cut_level is None else cut_level,
                                                 C.byref(self._h)))
        self._info = _lib.DbInfo()
        _lib.check(_lib.lib().pf_db_info(self._h, C.byref(self._info)))
        return self

    def shard_info(self) -> _lib.ShardInfo:
        o = _lib.ShardInfo()
        _lib.check(_lib.lib().pf_shard_info(self._h, C.byref(o)))
        return o

    def shard_stats(self) -> _lib.ShardStats:
        o = _lib.ShardStats()
        _lib.check(_lib.lib().pf_shard_stats(self._h, C.byref(o)))
        return o

    def allreduce_counts(self) -> None:
        """ONE ncclAllReduce(sum) of the per-leaf counters over the handle's communicator."""
        _lib.check(_lib.lib().pf_allreduce_counts(self._h))

    def prune_tree(self, search_depth: int) -> None:
        """bloom_tree.rs:302-330: nodes at depth >= search_depth become leaves.  Counters restart,
        as in the reference where pruning happens before the first query (main.rs:293-299)."""
        self.close()
        self.__init__(self.directory, self.device, search_depth)

    @property
    def kmer_size(self) -> int:
        return int(self._info.kmer_size)

    # -- introspection ----------------------------------------------------------------------------
    @property
    def info(self) -> _lib.DbInfo:
        return self._info

    def leaf_ids(self) -> List[str]:
        L = _lib.lib()
        return [L.pf_db_leaf_id(self._h, i).decode() for i in range(int(self._info.n_leaves))]

    def leaf_counts(self) -> np.ndarray:
        out = np.zeros(max(int(self._info.n_leaves), 1), dtype=np.uint64)
        _lib.check(_lib.lib().pf_leaf_counts(self._h, out.ctypes.data_as(C.POINTER(C.c_uint64))))
        return out[: int(self._info.n_leaves)]

    def reset_counts(self) -> None:
        _lib.check(_lib.lib().pf_reset_counts(self._h))

    def set_hash_rot(self, rot: int) -> None:
        _lib.check(_lib.lib().pf_db_set_hash_rot(self._h, rot))
        _lib.check(_lib.lib().pf_db_info(self._h, C.byref(self._info)))

    def detect_hash_rot(self, dfs_leaf: int, genome: bytes) -> int:
        rot = C.c_int(-1)
        _lib.check(_lib.lib().pf_db_detect_hash_rot(self._h, dfs_leaf, genome, len(genome), C.byref(rot)))
        return rot.value

    def set_exhaustive(self, on: bool) -> None:
        _lib.check(_lib.lib().pf_db_set_exhaustive(self._h, int(on)))

    def set_hash_cache_bytes(self, nbytes: int) -> None:
        _lib.check(_lib.lib().pf_db_set_hash_cache_bytes(self._h, nbytes))

    def set_memo(self, on: bool, budget_bytes: int = 0) -> None:
        """k-mer memo at exact nodes (pf_db_set_memo): same results, fewer probes on deep-coverage batches."""
        _lib.check(_lib.lib().pf_db_set_memo(self._h, int(on), budget_bytes))

    MODE_AUTO, MODE_PAIR, MODE_SLICED = 0, 1, 2

    def set_mode(self, mode: int) -> None:
        """pf_db_set_mode: 0 = cost model decides per (threshold, read length), 1 = node-at-a-time descent,
        2 = bit-sliced tiles.  Results are identical in every mode."""
        _lib.check(_lib.lib().pf_db_set_mode(self._h, int(mode)))

    def set_handover(self, handover: int) -> None:
        """pf_db_set_handover: sliced path below the cut -- 0 tiles all the way down, 1 hand-over to the node-at-a-time
        descent, -1 tiles if they fit the free HBM."""
        _lib.check(_lib.lib().pf_db_set_handover(self._h, int(handover)))

    def set_tile_cols(self, cols: int) -> None:
        """pf_db_set_tile_cols: most columns per tile of the sliced path (32, 64, 128 or 256)."""
        _lib.check(_lib.lib().pf_db_set_tile_cols(self._h, int(cols)))

    def set_frontier_cap(self, pairs: int) -> None:
        """pf_db_set_frontier_cap: chunks of reads whose frontier outgrows this many pairs are cut in half and redone."""
        _lib.check(_lib.lib().pf_db_set_frontier_cap(self._h, pairs))

    def set_lazy(self, on: bool) -> None:
        _lib.check(_lib.lib().pf_db_set_lazy(self._h, int(on)))

    def node_steps(self, threshold: float, nominal_kmers: int = 131) -> np.ndarray:
        out = np.zeros(int(self._info.n_nodes), dtype=np.uint32)
        _lib.check(_lib.lib().pf_db_node_steps(self._h, C.c_float(threshold), nominal_kmers,
                                               out.ctypes.data_as(C.POINTER(C.c_uint32))))
        return out

    def node_plan(self, threshold: float, nominal_kmers: int = 131) -> Tuple[np.ndarray, np.ndarray]:
        """(probe steps, k-mer sampling stride) per node in level order."""
        steps = np.zeros(int(self._info.n_nodes), dtype=np.uint32)
        strides = np.zeros(int(self._info.n_nodes), dtype=np.uint32)
        _lib.check(_lib.lib().pf_db_node_plan(self._h, C.c_float(threshold), nominal_kmers,
                                              steps.ctypes.data_as(C.POINTER(C.c_uint32)),
                                              strides.ctypes.data_as(C.POINTER(C.c_uint32))))
        return steps, strides

    def stats(self) -> _lib.Stats:
        s = _lib.Stats()
        _lib.check(_lib.lib().pf_get_stats(self._h, C.byref(s)))
        return s

    def reset_stats(self) -> None:
        _lib.check(_lib.lib().pf_reset_stats(self._h))

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            _lib.lib().pf_db_close(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class BloomTreeBuilder:
    """Build side: BloomTree::new + insert + save (bloom_tree.rs:100-145, 339-355) on the GPU,
    writing the reference's on-disk format."""

    def __init__(self, kmer_size: int, false_pos_rate: float = 0.001, largest_expected_genome: int = 1_000_000,
                 hash_states: Tuple[int, int] = (0x5EED0001, 0x5EED0002), device: int = 0, name_mode: int = 0,
                 name_seed: int = 0):
        self._h = C.c_void_p()
        _lib.check(_lib.lib().pf_builder_create(kmer_size, C.c_float(false_pos_rate), largest_expected_genome,
                                                hash_states[0], hash_states[1], device, name_mode, name_seed,
                                                C.byref(self._h)))

    def insert(self, genome_id: str, sequence: bytes) -> None:
        _lib.check(_lib.lib().pf_builder_insert(self._h, genome_id.encode(), sequence, len(sequence)))

    def save(self, directory: str) -> None:
        _lib.check(_lib.lib().pf_builder_save(self._h, directory.encode()))

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            _lib.lib().pf_builder_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
