"""Scaled-down versions of BASELINE.json configs 3-5 as parity cases (the full sizes are bench/perf material):
wide frontiers at -f 0.8 with 90 % background reads, 10 kb long reads at -f 0.9 (multi-group k-mer path),
and a larger, deeper tree.  GPU (through the C ABI) vs the oracle, bit-exact, including the kernel's work."""
import numpy as np
import pytest

from tests.util import gpu_query, oracle_build_db

pytestmark = pytest.mark.gpu


def _check(oracle, ot, gt, reads, theta):
    from phagefilter_b200.query import get_leaf_counts
    ot.reset_counts()
    gt.reset_counts()
    gt.reset_stats()
    want = ot.query_batch(reads, theta)
    sched = ot.query_sched(reads, theta, lazy=True)
    # bit-sliced tiles, then whatever the cost model picks
    for mode, handover in ((2, 0), (2, 1), (0, -1)):
        gt.set_mode(mode)
        gt.set_handover(handover)
        gt.reset_counts()
        assert gpu_query(gt, reads, theta) == want.hit_sets(len(reads)), (mode, handover)
        assert get_leaf_counts(gt) == ot.leaf_counts(), (mode, handover)
    gt.set_mode(1)  # node-at-a-time descent: work counts are predicted by the oracle's restatement of its schedule
    gt.reset_counts()
    gt.reset_stats()
    # default: k-mer memo on -- same results and pairs (how many probes are issued depends on timing)
    got = gpu_query(gt, reads, theta)
    assert got == want.hit_sets(len(reads))
    assert get_leaf_counts(gt) == ot.leaf_counts()
    st = gt.stats()
    assert st.pairs == sched.pairs
    # memo off: the kernel's work is exactly its restatement's
    gt.set_memo(False)
    gt.reset_counts()
    gt.reset_stats()
    got = gpu_query(gt, reads, theta)
    gt.set_memo(True)
    assert got == want.hit_sets(len(reads))
    assert get_leaf_counts(gt) == ot.leaf_counts()
    st = gt.stats()
    assert (st.pairs, st.probes_issued) == (sched.pairs, sched.probes_sched)
    gt.set_mode(0)
    return want, st


@pytest.fixture(scope="module")
def db300(oracle, tmp_path_factory):
    from phagefilter_b200 import BloomTree
    from phagefilter_b200.synth import make_genomes
    genomes = make_genomes(30, 10, seed=1003, len_lo=2500, len_hi=3500)
    d = str(tmp_path_factory.mktemp("cfg3") / "db")
    ot = oracle_build_db(oracle, genomes, 20, d, largest=5000)
    gt = BloomTree.load(d)
    yield genomes, ot, gt
    gt.close()


def test_cfg3_like_wide_frontier(oracle, db300):
    """10 % phage spike-in (1 % substitutions) + 90 % random background, -f 0.8."""
    from phagefilter_b200.synth import simulate_reads
    genomes, ot, gt = db300
    assert gt.info.n_leaves == 300 and gt.info.n_levels > 9
    reads, src = simulate_reads(genomes, 20_000, 150, seed=2003, error_rates=(0.01,), background_frac=0.9)
    want, st = _check(oracle, ot, gt, [r.tobytes() for r in reads], 0.8)
    spiked = int((src >= 0).sum())
    assert spiked > 1000 and len(want.hits) >= spiked * 0.5  # most spike-in reads classify
    assert oracle.need(0.8, 131) == 105


def test_cfg4_like_long_reads(oracle, tmp_path):
    """10 kb reads (9,981 k-mers: 39 groups of 8 rounds), 10 % from genomes with 0.1 % substitutions, -f 0.9."""
    from phagefilter_b200 import BloomTree
    from phagefilter_b200.synth import make_genomes, simulate_reads
    genomes = make_genomes(8, 5, seed=1004, len_lo=11_000, len_hi=14_000)
    d = str(tmp_path / "db")
    ot = oracle_build_db(oracle, genomes, 20, d, largest=16_000)
    gt = BloomTree.load(d)
    reads, src = simulate_reads(genomes, 400, 10_000, seed=2004, error_rates=(0.001,), background_frac=0.9)
    rl = [r.tobytes() for r in reads] + [genomes[3][1][:9_990], genomes[3][1][100:10_119]]
    want, st = _check(oracle, ot, gt, rl, 0.9)
    assert st.group_rounds == 8
    assert oracle.need(0.9, 9981) == 8983
    assert len(want.hits) > 20
    # the same batch in tiny hash-cache chunks
    gt.set_hash_cache_bytes(8 * 9981 * 3)
    _check(oracle, ot, gt, rl, 0.9)
    gt.close()


def test_deep_unbalanced_tree(oracle, tmp_path):
    """Unrelated genomes inserted greedily give a deep, unbalanced tree; reads with mixed lengths."""
    from phagefilter_b200 import BloomTree
    from phagefilter_b200.synth import make_genomes, simulate_reads
    genomes = make_genomes(120, 1, seed=77, len_lo=800, len_hi=1500, divergence=0.0)
    d = str(tmp_path / "db")
    ot = oracle_build_db(oracle, genomes, 21, d, largest=2000)
    gt = BloomTree.load(d)
    depth = max(dd for _, _, dd in ot.preorder())
    assert gt.info.n_levels == depth + 1 and depth >= 8
    r1, _ = simulate_reads(genomes, 1500, 100, seed=5, error_rates=(0.0, 0.02), background_frac=0.3)
    r2, _ = simulate_reads(genomes, 300, 250, seed=6, error_rates=(0.0,))
    rl = [r.tobytes() for r in r1] + [r.tobytes() for r in r2]
    for theta in (1.0, 0.5):
        _check(oracle, ot, gt, rl, theta)
    gt.close()
