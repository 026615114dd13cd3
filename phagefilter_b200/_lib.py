"""ctypes binding of libpfgpu.so (the C ABI declared in include/pfgpu.h).

There is no CPU fallback: if the CUDA library is missing this module raises at import of the
first symbol, and every compute entry point fails loudly without a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpfgpu.so")

PF_OK = 0
STATUS_NAMES = {0: "PF_OK", 1: "PF_ERR_ARG", 2: "PF_ERR_IO", 3: "PF_ERR_FORMAT", 4: "PF_ERR_CUDA",
                5: "PF_ERR_NOMEM", 6: "PF_ERR_NCCL", 7: "PF_ERR_STATE"}


class PfError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"{STATUS_NAMES.get(status, status)}: {message}")
        self.status = status
        self.message = message


class DbInfo(C.Structure):
    _fields_ = [
        ("kmer_size", C.c_uint64), ("num_bits", C.c_uint64), ("words_per_filter", C.c_uint64),
        ("n_nodes", C.c_uint64), ("n_leaves", C.c_uint64), ("n_filters", C.c_uint64), ("n_levels", C.c_uint64),
        ("filter_bytes", C.c_uint64), ("seed1", C.c_uint64), ("seed2", C.c_uint64),
        ("num_hashes", C.c_uint32), ("largest_genome", C.c_uint32), ("false_pos_rate", C.c_float),
        ("device", C.c_int32), ("hash_rot", C.c_int32), ("fast_path", C.c_int32),
        ("n_internal", C.c_uint64), ("n_monotone", C.c_uint64),
    ]


class ReadBatch(C.Structure):
    _fields_ = [
        ("n_reads", C.c_uint32), ("n_exc", C.c_uint32),
        ("lengths", C.POINTER(C.c_uint32)), ("word_off", C.POINTER(C.c_uint64)),
        ("packed", C.POINTER(C.c_uint32)), ("n_words", C.c_uint64),
        ("exc_index", C.POINTER(C.c_uint32)), ("exc_off", C.POINTER(C.c_uint64)),
        ("exc_bytes", C.POINTER(C.c_uint8)),
        ("max_length", C.c_uint32), ("total_bases", C.c_uint64),
    ]


class Hits(C.Structure):
    _fields_ = [("n_hits", C.c_uint64), ("read_off", C.POINTER(C.c_uint64)), ("leaf", C.POINTER(C.c_uint32))]


class Stats(C.Structure):
    _fields_ = [
        ("blocks", C.c_uint64), ("reads", C.c_uint64), ("pairs", C.c_uint64), ("probes_issued", C.c_uint64),
        ("levels", C.c_uint64), ("probe_launches", C.c_uint64), ("other_launches", C.c_uint64),
        ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64),
        ("probe_kernel_ms", C.c_double), ("device_ms", C.c_double), ("group_rounds", C.c_uint64),
        ("memo_hits", C.c_uint64), ("memo_lookups", C.c_uint64),
        ("sliced_blocks", C.c_uint64), ("sliced_tiles", C.c_uint64), ("sliced_table_bytes", C.c_uint64),
        ("chunk_splits", C.c_uint64), ("sector_loads", C.c_uint64), ("sliced_pairs", C.c_uint64),
        ("sliced_kernel_ms", C.c_double), ("sliced_launches", C.c_uint64), ("line_loads", C.c_uint64), ("entry_kernel_ms", C.c_double),
    ]


class ShardInfo(C.Structure):
    _fields_ = [("sharded", C.c_int32), ("rank", C.c_int32), ("nranks", C.c_int32), ("cut_level", C.c_uint32),
                ("top_nodes", C.c_uint64), ("owned_nodes", C.c_uint64), ("resident_filters", C.c_uint64),
                ("resident_bytes", C.c_uint64)]


class ShardStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("queries", "collectives", "reads_gathered", "pairs_top", "pairs_subtrees",
                                          "pairs_sent", "pairs_received", "hits_sent", "bytes_sent", "bytes_received")]


# every symbol include/pfgpu.h declares: (restype, argtypes)
_VP = C.c_void_p
SYMBOLS = {
    "pf_last_error": (C.c_char_p, []),
    "pf_version": (C.c_char_p, []),
    "pf_db_open": (C.c_int, [C.c_char_p, C.c_int, C.c_int64, C.POINTER(_VP)]),
    "pf_db_info": (C.c_int, [_VP, C.POINTER(DbInfo)]),
    "pf_db_leaf_id": (C.c_char_p, [_VP, C.c_uint64]),
    "pf_db_set_hash_rot": (C.c_int, [_VP, C.c_int]),
    "pf_db_detect_hash_rot": (C.c_int, [_VP, C.c_uint64, C.c_char_p, C.c_uint64, C.POINTER(C.c_int)]),
    "pf_db_close": (None, [_VP]),
    "pf_pack_reads": (C.c_int, [C.c_char_p, C.POINTER(C.c_uint64), C.c_uint32, C.POINTER(_VP)]),
    "pf_pack_reads_ptrs": (C.c_int, [C.POINTER(C.c_char_p), C.POINTER(C.c_uint32), C.c_uint32, C.POINTER(_VP)]),
    "pf_packed_batch": (C.POINTER(ReadBatch), [_VP]),
    "pf_packed_free": (None, [_VP]),
    "pf_packed_reserve_like": (C.c_int, [C.POINTER(_VP), _VP]),
    "pf_alloc_pinned": (_VP, [C.c_size_t]),
    "pf_free_pinned": (None, [_VP]),
    "pf_thread_set_device": (C.c_int, [C.c_int]),
    "pf_query_block": (C.c_int, [_VP, C.POINTER(ReadBatch), C.c_float, C.c_int, C.POINTER(Hits)]),
    "pf_batch_upload": (C.c_int, [_VP, C.POINTER(ReadBatch), C.POINTER(_VP)]),
    "pf_query_device": (C.c_int, [_VP, _VP, C.c_float, C.c_int, C.POINTER(Hits)]),
    "pf_batch_upload_async": (C.c_int, [_VP, C.POINTER(ReadBatch), C.POINTER(_VP)]),
    "pf_batch_free": (None, [_VP, _VP]),
    "pf_leaf_counts": (C.c_int, [_VP, C.POINTER(C.c_uint64)]),
    "pf_reset_counts": (C.c_int, [_VP]),
    "pf_save_leaf_counts": (C.c_int, [_VP, C.c_char_p]),
    "pf_get_stats": (C.c_int, [_VP, C.POINTER(Stats)]),
    "pf_reset_stats": (C.c_int, [_VP]),
    "pf_db_stream": (_VP, [_VP]),
    "pf_db_set_exhaustive": (C.c_int, [_VP, C.c_int]),
    "pf_db_set_lazy": (C.c_int, [_VP, C.c_int]),
    "pf_db_set_mode": (C.c_int, [_VP, C.c_int]),
    "pf_db_set_handover": (C.c_int, [_VP, C.c_int]),
    "pf_db_set_tile_cols": (C.c_int, [_VP, C.c_int]),
    "pf_db_set_frontier_cap": (C.c_int, [_VP, C.c_uint64]),
    "pf_db_set_memo": (C.c_int, [_VP, C.c_int, C.c_uint64]),
    "pf_db_set_hash_cache_bytes": (C.c_int, [_VP, C.c_uint64]),
    "pf_db_node_steps": (C.c_int, [_VP, C.c_float, C.c_uint64, C.POINTER(C.c_uint32)]),
    "pf_db_node_plan": (C.c_int, [_VP, C.c_float, C.c_uint64, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "pf_nccl_unique_id": (C.c_int, [C.c_char_p]),
    "pf_comm_init": (C.c_int, [_VP, C.c_int, C.c_int, C.c_char_p]),
    "pf_allreduce_counts": (C.c_int, [_VP]),
    "pf_db_open_sharded": (C.c_int, [C.c_char_p, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_char_p, C.c_int64, C.POINTER(_VP)]),
    "pf_shard_info": (C.c_int, [_VP, C.POINTER(ShardInfo)]),
    "pf_shard_stats": (C.c_int, [_VP, C.POINTER(ShardStats)]),
    "pf_shard_plan": (C.c_int, [C.c_char_p, C.c_int64, C.c_int, C.c_int64, C.POINTER(C.c_uint32), C.POINTER(C.c_int32),
                                C.c_uint64, C.POINTER(C.c_uint64)]),
    "pf_query_sharded": (C.c_int, [_VP, C.POINTER(ReadBatch), C.c_float, C.c_int, C.POINTER(Hits)]),
    "pf_query_sharded_device": (C.c_int, [_VP, _VP, C.c_float, C.c_int, C.POINTER(Hits)]),
    "pf_builder_create": (C.c_int, [C.c_uint64, C.c_float, C.c_uint32, C.c_uint64, C.c_uint64, C.c_int, C.c_int,
                                    C.c_uint64, C.POINTER(_VP)]),
    "pf_builder_open": (C.c_int, [C.c_char_p, C.c_int, C.POINTER(_VP)]),
    "pf_builder_set_hash_rot": (C.c_int, [_VP, C.c_int]),
    "pf_builder_insert": (C.c_int, [_VP, C.c_char_p, C.c_char_p, C.c_uint64]),
    "pf_builder_save": (C.c_int, [_VP, C.c_char_p]),
    "pf_builder_free": (None, [_VP]),
    "pf_needed_bits": (C.c_uint64, [C.c_float, C.c_uint32]),
    "pf_optimal_num_hashes": (C.c_uint32, [C.c_uint64, C.c_uint32]),
    "pf_microbench_sectors": (C.c_int, [C.c_int, C.c_uint64, C.c_int, C.POINTER(C.c_double)]),
    "pf_plan_tiles": (C.c_int, [C.c_uint64, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_int32),
                                C.POINTER(C.c_uint64), C.POINTER(C.c_uint8), C.c_uint64, C.c_uint32, C.c_float, C.c_uint64,
                                C.c_int, C.POINTER(C.c_uint8), C.POINTER(C.c_int32), C.POINTER(C.c_uint32),
                                C.POINTER(C.c_int32), C.POINTER(C.c_uint32), C.c_uint64, C.POINTER(C.c_uint64),
                                C.POINTER(C.c_uint64)]),
    "pf_plan_tile_layout": (C.c_int, [C.c_uint64, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_int32),
                                      C.POINTER(C.c_uint64), C.POINTER(C.c_uint8), C.c_uint64, C.c_uint32, C.c_float, C.c_uint64,
                                      C.c_int, C.c_int, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32),
                                      C.POINTER(C.c_uint32), C.POINTER(C.c_int32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint8),
                                      C.POINTER(C.c_uint32), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64),
                                      C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
}

_lib = None


def lib() -> C.CDLL:
    """Load libpfgpu.so; raise if it has not been built (no fallback of any kind)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `make -C phagefilter_b200/csrc` "
                "(or `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)  # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(status: int) -> None:
    if status != PF_OK:
        raise PfError(status, (lib().pf_last_error() or b"").decode(errors="replace"))
