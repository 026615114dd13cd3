"""world_size-2 gloo test of the N>1 host logic on CPU: shard reads by rank, classify each shard (with the
oracle standing in for the per-rank GPU), sum the per-leaf counters with one all-reduce, compare with the
single-rank result."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, db_dir, reads, theta, block, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from oracle import pf_oracle
    from phagefilter_b200.shard import block_ranges, combine_counts, shard_blocks
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    tree = pf_oracle.Tree.load(db_dir)
    ranges = block_ranges(len(reads), block)
    for b in shard_blocks(len(ranges), rank, world):
        lo, hi = ranges[b]
        tree.query_batch(reads[lo:hi], theta, threads=1)
    local = np.array([c for _, c in tree.leaf_counts()], dtype=np.uint64)
    total = combine_counts(local)
    np.save(os.path.join(out_dir, f"counts_{rank}.npy"), total)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_counts_equal_single_rank(oracle, tmp_path):
    from phagefilter_b200.shard import block_ranges, shard_blocks
    from tests.util import oracle_build_db, random_genomes, sample_reads
    rng = np.random.default_rng(42)
    genomes = random_genomes(rng, 8, 1500, 2500)
    d = str(tmp_path / "db")
    t = oracle_build_db(oracle, genomes, 20, d, largest=4000)
    reads = sample_reads(rng, genomes, 500, 100, 0.0) + sample_reads(rng, genomes, 250, 100, 0.03) + [b"ACGT", b""]
    t.query_batch(reads, 0.8)
    want = np.array([c for _, c in t.leaf_counts()], dtype=np.uint64)
    # every block is owned by exactly one rank
    ranges = block_ranges(len(reads), 64)
    owned = sorted(b for r in range(2) for b in shard_blocks(len(ranges), r, 2))
    assert owned == list(range(len(ranges)))
    port = _free_port()
    mp.spawn(_worker, args=(2, port, d, reads, 0.8, 64, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        got = np.load(str(tmp_path / f"counts_{r}.npy"))
        assert (got == want).all()
