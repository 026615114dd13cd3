"""Randomised differential test: random geometry (fpr, largest genome => m, K), k, tree shape, thresholds,
read lengths (0 .. several groups), alphabets with exception bytes -- GPU through the C ABI vs the oracle,
bit-exact in results and in the kernel's work (pairs, probes) in default, exact and exhaustive modes."""
import numpy as np
import pytest

from tests.util import gpu_query, oracle_build_db

pytestmark = pytest.mark.gpu


def _case(seed):
    rng = np.random.default_rng(seed)
    k = int(rng.choice([4, 9, 15, 17, 19, 20, 23, 27, 31, 32, 35]))
    fpr = float(rng.choice([0.2, 0.05, 0.01, 0.001, 1e-4]))
    largest = int(rng.choice([300, 1000, 4000, 20000]))
    n_gen = int(rng.integers(1, 24))
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    genomes = []
    base = acgt[rng.integers(0, 4, size=int(rng.integers(300, 2500)))]
    for i in range(n_gen):
        if rng.random() < 0.5:  # relative of the base genome
            g = base.copy()
            m = rng.random(len(g)) < rng.choice([0.0, 0.01, 0.05, 0.3])
            g[m] = acgt[rng.integers(0, 4, size=int(m.sum()))]
        else:
            g = acgt[rng.integers(0, 4, size=int(rng.integers(50, 3000)))]
        g = bytearray(g.tobytes())
        if rng.random() < 0.3 and len(g) > 40:
            p = int(rng.integers(0, len(g) - 10))
            g[p:p + 5] = rng.choice([b"NNNNN", b"acgtn", b"RYKMS"])
        genomes.append((f"g{i}", bytes(g)))
    reads = []
    for _ in range(int(rng.integers(20, 400))):
        gi = genomes[int(rng.integers(0, n_gen))][1]
        L = int(rng.choice([0, 1, k - 1, k, k + 1, 40, 100, 150, 300, 1200]))
        if rng.random() < 0.25 or len(gi) < L:
            r = acgt[rng.integers(0, 4, size=L)].tobytes()
        else:
            s = int(rng.integers(0, len(gi) - L + 1))
            r = bytearray(gi[s:s + L])
            for p in np.nonzero(rng.random(L) < rng.choice([0.0, 0.01, 0.05]))[0]:
                r[int(p)] = b"ACGTN"[int(rng.integers(0, 5))]
            r = bytes(r)
        reads.append(r)
    thetas = [float(x) for x in rng.choice([0.0, 0.05, 0.3, 0.5, 0.8, 0.9, 0.99, 1.0, 1.2], size=3, replace=False)]
    depth = int(rng.integers(0, 5)) if rng.random() < 0.25 else None
    return k, fpr, largest, genomes, reads, thetas, depth


@pytest.mark.parametrize("seed", range(24))
def test_fuzz(oracle, tmp_path, seed):
    from phagefilter_b200 import BloomTree
    from phagefilter_b200.query import get_leaf_counts
    k, fpr, largest, genomes, reads, thetas, depth = _case(1000 + seed)
    d = str(tmp_path / "db")
    ot = oracle_build_db(oracle, genomes, k, d, fpr=fpr, largest=largest)
    gt = BloomTree(d, 0, depth)
    if depth is not None:
        ot.prune_tree(depth)
    assert gt.leaf_ids() == ot.leaf_ids()
    assert (gt.info.num_bits, gt.info.num_hashes) == (ot.num_bits, ot.num_hashes)
    for theta in thetas:
        ot.reset_counts()
        ref = ot.query_batch(reads, theta)
        want = ref.hit_sets(len(reads))
        for mode, handover in ((2, 0), (2, 1), (0, -1)):  # tiles all the way down; tiles + hand-over; the cost model's choice
            gt.set_mode(mode)
            gt.set_handover(handover)
            gt.reset_counts()
            assert gpu_query(gt, reads, theta) == want, (seed, theta, "mode", mode, handover)
            assert get_leaf_counts(gt) == ot.leaf_counts(), (seed, theta, "mode", mode, handover)
        gt.set_mode(1)  # node-at-a-time descent: the work counts below are its schedule's
        for lazy in (True, False):
            sched = ot.query_sched(reads, theta, lazy=lazy)
            for memo in (True, False):  # the k-mer memo never changes results; without it the work is deterministic
                gt.reset_counts()
                gt.reset_stats()
                gt.set_lazy(lazy)
                gt.set_memo(memo)
                assert gpu_query(gt, reads, theta) == want, (seed, theta, lazy, memo)
                assert get_leaf_counts(gt) == ot.leaf_counts()
                st = gt.stats()
                if memo:
                    assert st.pairs == sched.pairs, (seed, theta, lazy)
                else:
                    assert (st.pairs, st.probes_issued) == (sched.pairs, sched.probes_sched), (seed, theta, lazy)
        gt.set_lazy(True)
        gt.set_memo(True)
        gt.set_exhaustive(True)
        gt.reset_counts()
        gt.reset_stats()
        assert gpu_query(gt, reads, theta) == want
        st = gt.stats()
        assert (st.pairs, st.probes_issued) == (ref.pairs, ref.probes_ref)
        gt.set_exhaustive(False)
    gt.close()
