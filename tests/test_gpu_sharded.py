"""Subtree-sharded query (pf_db_open_sharded / pf_query_sharded) against the CPU oracle.

Single process: one rank owning every subtree exercises the two-phase descent, the slice exchange with itself
and the hit routing on any GPU box.  Multi process (needs >= 2 GPUs): the worker script under torchrun, reads
split unevenly over the ranks, results compared with the oracle's single tree."""
import os
import pickle
import subprocess
import sys

import numpy as np
import pytest

from tests.util import oracle_build_db

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _reads(genomes, n, seed):
    from phagefilter_b200.synth import simulate_reads
    reads, _ = simulate_reads(genomes, n, 150, seed=seed, error_rates=(0.0, 0.01, 0.05), background_frac=0.3)
    rl = [r.tobytes() for r in reads]
    g0 = genomes[0][1]
    rl += [b"", b"ACGT", g0[:20], g0[100:160].lower(), g0[200:290].replace(b"A", b"N", 1), g0[300:420], g0[300:420]]
    return rl


@pytest.fixture(scope="module")
def db(oracle, tmp_path_factory):
    from phagefilter_b200.synth import make_genomes
    genomes = make_genomes(6, 4, seed=314, len_lo=2500, len_hi=3500)
    d = str(tmp_path_factory.mktemp("shard") / "db")
    ot = oracle_build_db(oracle, genomes, 20, d, largest=5000)
    return genomes, d, ot


@pytest.mark.parametrize("cut", [None, 1, 2, 4, 99])
@pytest.mark.parametrize("theta", [1.0, 0.7])
def test_single_rank_shard_equals_oracle(oracle, db, cut, theta):
    import ctypes as C
    from phagefilter_b200 import BloomTree, _lib
    from phagefilter_b200.query import PackedReads, get_leaf_counts, query_packed, query_sharded
    genomes, d, ot = db
    reads = _reads(genomes, 3000, 5)
    idbuf = C.create_string_buffer(128)
    _lib.check(_lib.lib().pf_nccl_unique_id(idbuf))
    tree = BloomTree.open_sharded(d, 0, 0, 1, idbuf.raw, cut_level=cut)
    si = tree.shard_info()
    assert si.sharded == 1 and si.nranks == 1 and si.top_nodes + si.owned_nodes == tree.info.n_nodes
    if cut is not None:
        assert si.cut_level == min(cut, tree.info.n_levels)
    ot.reset_counts()
    want = ot.query_batch(reads, theta)
    p = PackedReads(reads)
    with pytest.raises(_lib.PfError):  # the replicated-tree entry point refuses a sharded handle
        query_packed(tree, p, theta)
    for rep in range(2):  # counters accumulate over calls like mapped_reads
        off, leaf = query_sharded(tree, p, theta)
        got = [frozenset(int(x) for x in leaf[int(off[i]):int(off[i + 1])]) for i in range(len(reads))]
        assert got == want.hit_sets(len(reads))
    assert [(i, c) for i, c in get_leaf_counts(tree)] == [(i, 2 * c) for i, c in ot.leaf_counts()]
    # counts only
    tree.reset_counts()
    off, leaf = query_sharded(tree, p, theta, want_hits=False)
    assert len(leaf) == 0 and get_leaf_counts(tree) == ot.leaf_counts()
    # the same work as the replicated tree does (same plan, same pairs)
    rep_tree = BloomTree.load(d)
    rep_tree.set_mode(1)      # the sharded descent is node-at-a-time
    rep_tree.set_memo(False)  # deterministic work counts
    tree.set_memo(False)
    query_packed(rep_tree, p, theta)
    tree.reset_stats()
    query_sharded(tree, p, theta)
    a, b = rep_tree.stats(), tree.stats()
    assert (a.pairs, a.probes_issued) == (b.pairs, b.probes_issued)
    ss = tree.shard_stats()
    assert ss.pairs_top + ss.pairs_subtrees >= b.pairs and ss.pairs_sent == 0
    p.close()
    rep_tree.close()
    tree.close()


def _n_gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world,cut,theta,want_hits", [(2, -1, 1.0, 1), (2, 2, 0.7, 1), (2, 3, 1.0, 0), (4, -1, 0.7, 1)])
def test_multi_rank_shard_equals_oracle(oracle, db, tmp_path, world, cut, theta, want_hits):
    if _n_gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    genomes, d, ot = db
    reads = _reads(genomes, 4000, 6)
    rp = str(tmp_path / "reads.pkl")
    pickle.dump(reads, open(rp, "wb"))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
           "127.0.0.1", "--master-port", str(29600 + world * 10 + (cut if cut > 0 else 0)),
           os.path.join(ROOT, "tests", "shard_worker.py"), "--db", d, "--reads", rp, "--out", str(tmp_path), "--theta",
           str(theta), "--cut", str(cut), "--blocks", "2", "--want-hits", str(want_hits)]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    ot.reset_counts()
    want = ot.query_batch(reads, theta)
    want_sets = [sorted(s) for s in want.hit_sets(len(reads))]
    want_counts = np.array([c for _, c in ot.leaf_counts()], dtype=np.uint64)
    outs = [pickle.load(open(str(tmp_path / f"rank{r}.pkl"), "rb")) for r in range(world)]
    partial_sum = np.zeros_like(want_counts)
    resident = 0
    for o in outs:
        if want_hits:
            assert o["hit_sets"] == want_sets[o["lo"]:o["hi"]]
        assert (o["total"] == want_counts).all()
        partial_sum += o["partial"]
        resident += o["n_filters"]
        assert o["info"]["nranks"] == world and o["info"]["rank"] == o["rank"]
    assert (partial_sum == want_counts).all()
    n_nodes, top = outs[0]["n_nodes"], outs[0]["info"]["top_nodes"]
    assert resident == n_nodes + (world - 1) * top  # every subtree lives on exactly one rank
    if outs[0]["info"]["cut_level"] > 1 and theta >= 1.0:
        assert sum(o["shard_stats"]["pairs_sent"] for o in outs) > 0  # the frontier really crossed ranks
    assert sum(o["shard_stats"]["pairs_sent"] for o in outs) == sum(o["shard_stats"]["pairs_received"] for o in outs)
