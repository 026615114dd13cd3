"""One rank of a subtree-sharded query (launched by tests/test_gpu_sharded.py under torch.distributed.run).

Every rank opens its shard of the database, queries its own slice of the reads block by block through
pf_query_sharded (the last block of the shorter slice is empty on purpose) and dumps what it saw; the parent
test compares the union with the CPU oracle."""
import argparse
import os
import pickle
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--db", required=True)
    ap.add_argument("--reads", required=True, help="pickle: list of bytes")
    ap.add_argument("--out", required=True)
    ap.add_argument("--theta", type=float, default=1.0)
    ap.add_argument("--cut", type=int, default=-1)
    ap.add_argument("--blocks", type=int, default=2)
    ap.add_argument("--want-hits", type=int, default=1)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("gloo")
    from phagefilter_b200 import BloomTree
    from phagefilter_b200.query import PackedReads, query_sharded
    from phagefilter_b200.shard import exchange_nccl_id
    reads = pickle.load(open(args.reads, "rb"))
    # uneven slices: rank r takes reads r, r + world + (r odd), ... -> simply contiguous uneven cuts
    cuts = [0] + [int(len(reads) * (i + 1) / world * (0.8 if i + 1 < world else 1.0)) for i in range(world)]
    mine = reads[cuts[rank]:cuts[rank + 1]]
    tree = BloomTree.open_sharded(args.db, local_rank, rank, world, exchange_nccl_id(rank),
                                  cut_level=None if args.cut < 0 else args.cut)
    per_block = (max(len(reads[cuts[r]:cuts[r + 1]]) for r in range(world)) + args.blocks - 1) // args.blocks
    hit_sets = []
    for b in range(args.blocks):
        blk = mine[b * per_block:(b + 1) * per_block]  # may be empty: the call is collective all the same
        p = PackedReads(blk)
        off, leaf = query_sharded(tree, p, args.theta, want_hits=bool(args.want_hits))
        p.close()
        hit_sets += [sorted(int(x) for x in leaf[int(off[i]):int(off[i + 1])]) for i in range(len(blk))]
    partial = tree.leaf_counts().copy()
    tree.allreduce_counts()
    total = tree.leaf_counts().copy()
    si, ss, st = tree.shard_info(), tree.shard_stats(), tree.stats()
    out = {
        "rank": rank, "lo": cuts[rank], "hi": cuts[rank + 1], "hit_sets": hit_sets, "partial": partial, "total": total,
        "info": {f: getattr(si, f) for f, _ in si._fields_}, "shard_stats": {f: getattr(ss, f) for f, _ in ss._fields_},
        "pairs": int(st.pairs), "probes": int(st.probes_issued), "n_filters": int(tree.info.n_filters),
        "n_nodes": int(tree.info.n_nodes),
    }
    pickle.dump(out, open(os.path.join(args.out, f"rank{rank}.pkl"), "wb"))
    tree.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
