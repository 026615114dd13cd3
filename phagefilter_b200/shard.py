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
