"""Multi-GPU host logic: reads shard by rank, the tree is replicated, per-leaf counts are summed once.

The query path has no data-path collective (reads are independent); the only exchange is one
all-reduce(sum) of u64[n_leaves].  On GPUs that is pf_allreduce_counts (NCCL inside libpfgpu); this module
holds the host-side pieces that are backend independent so they can be tested with gloo on CPU.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


def shard_blocks(n_blocks: int, rank: int, world_size: int) -> List[int]:
    """Blocks are dealt round-robin: block i belongs to rank i mod N (main.rs:334-368 reads blocks serially)."""
    return list(range(rank, n_blocks, world_size))


def block_ranges(n_reads: int, block_size: int) -> List[Tuple[int, int]]:
    return [(lo, min(lo + block_size, n_reads)) for lo in range(0, n_reads, block_size)]


def combine_counts(local_counts: np.ndarray, group=None) -> np.ndarray:
    """Sum per-leaf counters over all ranks with torch.distributed (gloo on CPU, NCCL on GPU tensors)."""
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(np.ascontiguousarray(local_counts.astype(np.int64)))
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.numpy().astype(np.uint64)


def shard_plan(db_dir: str, nranks: int, search_depth=None, cut_level=None) -> Tuple[int, np.ndarray]:
    """Partition of a tree into a replicated top and per-rank subtrees (pf_shard_plan; host only, reads tree.bin).
    Returns (cut level, owner per node in level order: -1 = replicated, else rank)."""
    import ctypes as C
    from . import _lib
    L = _lib.lib()
    n = C.c_uint64(0)
    cut = C.c_uint32(0)
    sd = -1 if search_depth is None else search_depth
    cl = -1 if cut_level is None else cut_level
    _lib.check(L.pf_shard_plan(db_dir.encode(), sd, nranks, cl, C.byref(cut), None, 0, C.byref(n)))
    owner = np.zeros(int(n.value), dtype=np.int32)
    _lib.check(L.pf_shard_plan(db_dir.encode(), sd, nranks, cl, C.byref(cut), owner.ctypes.data_as(C.POINTER(C.c_int32)),
                               int(n.value), C.byref(n)))
    return int(cut.value), owner


def exchange_nccl_id(rank: int, group=None) -> bytes:
    """The 128-byte ncclUniqueId libpfgpu's own communicator starts from: made on rank 0 (pf_nccl_unique_id) and
    handed to the other ranks with torch.distributed (any backend)."""
    import ctypes as C
    import torch
    import torch.distributed as dist
    from . import _lib
    buf = C.create_string_buffer(128)
    if rank == 0:
        _lib.check(_lib.lib().pf_nccl_unique_id(buf))
    t = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        if dist.get_backend(group) == "nccl":
            t = t.cuda()
        dist.broadcast(t, 0, group=group)
    return bytes(t.cpu().numpy().tobytes())
