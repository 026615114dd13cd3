"""phagefilter_b200 -- B200-native implementation of PhageFilter's `query` hot path.

Host-side mirror of the reference's interface for that path (same names and argument meaning as
src/bloom_tree.rs, src/query.rs, src/result_map.rs, src/file_parser.rs) on top of the C ABI of
libpfgpu.so (include/pfgpu.h).  All compute runs in hand-written CUDA kernels for sm_100a; there is
no CPU fallback.
"""
from . import _lib
from .bloom_tree import BloomTree
from .file_parser import DNASequence, ReadQueue
from .query import get_leaf_counts, query_batch, save_leaf_counts
from .result_map import ResultMap

__all__ = ["BloomTree", "DNASequence", "ReadQueue", "ResultMap", "query_batch", "get_leaf_counts",
           "save_leaf_counts", "_lib"]
