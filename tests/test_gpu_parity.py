"""GPU parity: libpfgpu (through the C ABI) against the CPU oracle on the same seeded inputs.

Bit-exact bar: identical matched-leaf sets per read, identical leaf counters, identical
CLASSIFICATION.csv, identical probe counts (exhaustive mode == reference semantics; default mode ==
the oracle's restatement of the kernel schedule).
"""
import os

import numpy as np
import pytest

from tests.util import gpu_query, oracle_build_db, random_genomes, sample_reads

pytestmark = pytest.mark.gpu

FOUR = [("baseline", b"ATCAG"), ("diff", b"TTTAG"), ("onediff_first", b"CTCAG"), ("onediff_mid", b"ATTAG")]


def _open(directory, depth=None):
    from phagefilter_b200 import BloomTree
    return BloomTree(directory, 0, depth)


def _compare(oracle, otree, gtree, reads, theta, check_probes=True):
    """Run both sides on one block; compare hits, counters, csv and probe counts."""
    from phagefilter_b200.query import get_leaf_counts
    otree.reset_counts()
    ores = otree.query_batch(reads, theta)  # the reference's semantics
    want = ores.hit_sets(len(reads))
    ghits = None
    # bit-sliced tiles (pf_sliced.cu): every node evaluated exactly, 256 at a time
    gtree.set_mode(2)
    gtree.reset_counts()
    gtree.reset_stats()
    assert gpu_query(gtree, reads, theta) == want, "sliced"
    assert get_leaf_counts(gtree) == otree.leaf_counts(), "sliced"
    assert gtree.stats().sliced_blocks == 1
    # the same with tiles for the cut only and the survivors handed to the node-at-a-time descent
    gtree.set_handover(1)
    gtree.reset_counts()
    assert gpu_query(gtree, reads, theta) == want, "sliced + hand-over"
    assert get_leaf_counts(gtree) == otree.leaf_counts(), "sliced + hand-over"
    gtree.set_handover(-1)
    # whatever the cost model picks
    gtree.set_mode(0)
    gtree.reset_counts()
    assert gpu_query(gtree, reads, theta) == want, "auto"
    assert get_leaf_counts(gtree) == otree.leaf_counts(), "auto"
    gtree.set_mode(1)  # node-at-a-time descent: its work counts are predicted exactly by the oracle's restatement
    for lazy, memo in ((True, True), (True, False), (False, True), (False, False)):
        gtree.reset_counts()
        gtree.reset_stats()
        gtree.set_exhaustive(False)
        gtree.set_lazy(lazy)
        gtree.set_memo(memo)
        ghits = gpu_query(gtree, reads, theta)
        assert ghits == want, f"lazy={lazy} memo={memo}"
        assert get_leaf_counts(gtree) == otree.leaf_counts()
        if check_probes:
            # the oracle's restatement of the kernel schedule predicts the kernel's work exactly (with the k-mer
            # memo the probes actually issued depend on timing; unknown k-mers are probed three steps at a time)
            sched = otree.query_sched(reads, theta, lazy=lazy)
            assert sched.hit_sets(len(reads)) == want
            st = gtree.stats()
            if memo:
                assert st.pairs == sched.pairs, f"lazy={lazy}"
            else:
                assert (st.pairs, st.probes_issued) == (sched.pairs, sched.probes_sched), f"lazy={lazy}"
    gtree.set_memo(True)
    if check_probes:
        # reference-faithful probing: every k-mer of every pair until its first clear bit
        gtree.reset_counts()
        gtree.reset_stats()
        gtree.set_exhaustive(True)
        assert gpu_query(gtree, reads, theta) == want
        st = gtree.stats()
        assert (st.pairs, st.probes_issued) == (ores.pairs, ores.probes_ref)
        gtree.set_exhaustive(False)
    gtree.set_lazy(True)
    gtree.set_mode(0)


@pytest.mark.parametrize("k", [3, 4, 5])
def test_toy_tree_reference_cases(oracle, tmp_path, k):
    """The reference's own query tests (src/query.rs:268-380) on the 4-genome toy tree."""
    d = str(tmp_path / "db")
    ot = oracle_build_db(oracle, FOUR, k, d, largest=1000)
    gt = _open(d)
    assert gt.info.num_bits == 14378 and gt.info.num_hashes == 10
    reads = [b"ATCAG", b"TCAG", b"ATCA", b"AAAA", b"TTTAG", b"CTCAG", b"ATTAG", b"AT", b""]
    for theta in (0.0, 0.1, 0.51, 1.0):
        _compare(oracle, ot, gt, reads, theta)
    # semantics the reference asserts: identical read maps to its genome; theta 0 passes everything
    from phagefilter_b200.query import get_leaf_counts
    gt.reset_counts()
    hits = gpu_query(gt, [b"ATCAG"], 1.0)
    ids = gt.leaf_ids()
    assert "baseline" in {ids[i] for i in hits[0]}
    assert "diff" not in {ids[i] for i in hits[0]}
    gt.reset_counts()
    hits = gpu_query(gt, [b"GGGGG"], 0.0)
    assert hits[0] == frozenset(range(4))
    gt.close()


def test_counts_accumulate_across_calls(oracle, tmp_path):
    """query.rs:357-380: counters accumulate over successive query_batch calls."""
    from phagefilter_b200 import DNASequence, ResultMap, query_batch
    from phagefilter_b200.query import get_leaf_counts
    d = str(tmp_path / "db")
    ot = oracle_build_db(oracle, FOUR, 4, d, largest=1000)
    gt = _open(d)
    rm = ResultMap()
    gt = query_batch(gt, [DNASequence(b"TCAG", None, "read_tcag")], 0.1, rm)
    gt = query_batch(gt, [DNASequence(b"ATCA", None, "read_atca")], 0.1, rm)
    ot.query_batch([b"TCAG"], 0.1)
    ot.query_batch([b"ATCA"], 0.1)
    counts = dict(get_leaf_counts(gt))
    assert counts == dict(ot.leaf_counts())
    assert counts["baseline"] >= 2 and counts["diff"] == 0
    assert rm.read_mapped("read_tcag") and "baseline" in rm.get_ext_id("read_tcag")
    gt.close()


@pytest.mark.parametrize("k", [3, 8, 16, 17, 20, 21, 25, 31, 32, 33])
def test_random_db_all_k(oracle, tmp_path, k):
    """All three hash_bytes length branches, the 2-bit register path (17..32) and the byte path."""
    rng = np.random.default_rng(100 + k)
    genomes = random_genomes(rng, 9, 1500, 3000)
    # a family of near-duplicates so internal nodes are shared by several leaves
    base = genomes[0][1]
    for j in range(3):
        b = bytearray(base)
        for p in rng.integers(0, len(b), size=20):
            b[int(p)] = b"ACGT"[int(rng.integers(0, 4))]
        genomes.append((f"fam{j}", bytes(b)))
    d = str(tmp_path / "db")
    ot = oracle_build_db(oracle, genomes, k, d, largest=4000)
    gt = _open(d)
    assert gt.info.fast_path == (1 if 17 <= k <= 32 else 0)
    reads = sample_reads(rng, genomes, 300, 150, 0.0) + sample_reads(rng, genomes, 300, 150, 0.02)
    reads += [bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=150)) for _ in range(100)]
    reads += sample_reads(rng, genomes, 40, 37, 0.0) + sample_reads(rng, genomes, 20, 700, 0.01)
    for theta in (0.0, 0.3, 0.8, 0.9, 1.0):
        _compare(oracle, ot, gt, reads, theta)
    gt.close()


def test_adversarial_reads(oracle, tmp_path):
    """Short reads, length == k, N / lower-case / IUPAC bytes, palindromes, duplicates."""
    k = 20
    rng = np.random.default_rng(7)
    genomes = random_genomes(rng, 6, 1200, 2000)
    g0 = bytearray(genomes[0][1])
    g0[100:110] = b"NNNNNNNNNN"            # N run inside a genome
    g0[300:340] = bytes(g0[300:340]).lower()  # soft-masked stretch
    g0[500] = ord("R")
    genomes[0] = ("g0", bytes(g0))
    pal = b"ACGTACGTACGTACGTACGT"  # its own reverse complement
    genomes.append(("pal", pal * 4 + b"GATTACA" * 10))
    d = str(tmp_path / "db")
    ot = oracle_build_db(oracle, genomes, k, d, largest=4000)
    gt = _open(d)
    g0b = genomes[0][1]
    reads = [
        b"", b"A", b"ACGT" * 4 + b"ACG",      # shorter than k: zero k-mers, pass everywhere
        g0b[50:70],                            # exactly k
        g0b[90:240],                           # spans the N run
        g0b[280:430],                          # spans lower-case bases
        g0b[450:600],                          # spans the IUPAC base
        g0b[280:430].upper(),                  # upper-cased copy must NOT match the soft-masked k-mers
        pal * 3, pal,                          # palindromic k-mers
        g0b[600:750], g0b[600:750],            # duplicate reads
        b"N" * 150, b"acgt" * 40,
        bytes(range(33, 33 + 90)),             # arbitrary printable bytes
        g0b[700:850].replace(b"A", b"a", 1),
    ]
    for theta in (0.0, 0.3, 0.8, 1.0):
        _compare(oracle, ot, gt, reads, theta)
    gt.close()


@pytest.mark.parametrize("depth", [0, 1, 3])
def test_search_depth(oracle, tmp_path, depth):
    """prune_tree (bloom_tree.rs:302-330): internal nodes at the cut become leaves named Internal_Node_*."""
    rng = np.random.default_rng(11)
    genomes = random_genomes(rng, 10, 1500, 2500)
    d = str(tmp_path / "db")
    ot = oracle_build_db(oracle, genomes, 20, d, largest=4000)
    ot.prune_tree(depth)
    gt = _open(d)
    gt.prune_tree(depth)
    assert gt.leaf_ids() == ot.leaf_ids()
    reads = sample_reads(rng, genomes, 200, 100, 0.0) + sample_reads(rng, genomes, 200, 100, 0.05)
    for theta in (0.5, 1.0):
        _compare(oracle, ot, gt, reads, theta)
    gt.close()


def test_classification_csv_and_single_leaf(oracle, tmp_path):
    """save_leaf_counts (query.rs:173-183) and a one-genome DB whose root is a leaf."""
    from phagefilter_b200.query import save_leaf_counts
    rng = np.random.default_rng(3)
    genomes = random_genomes(rng, 1, 2000, 2000)
    d = str(tmp_path / "db1")
    ot = oracle_build_db(oracle, genomes, 20, d, largest=4000)
    gt = _open(d)
    reads = sample_reads(rng, genomes, 50, 100, 0.0) + [b"ACGT" * 30]
    _compare(oracle, ot, gt, reads, 1.0)
    out = str(tmp_path / "CLASSIFICATION.csv")
    save_leaf_counts(gt, out)
    assert open(out).read() == ot.classification_csv()
    gt.close()


def test_hash_rot_detection(oracle, tmp_path):
    """Self-certifying KAT: rebuilding a leaf from its genome reproduces the stored bits only with the
    rotate the DB was built with."""
    rng = np.random.default_rng(5)
    genomes = random_genomes(rng, 3, 1500, 2000)
    for rot in (26, 20):
        d = str(tmp_path / f"db{rot}")
        ot = oracle_build_db(oracle, genomes, 20, d, largest=4000, rot=rot)
        gt = _open(d)
        ids = gt.leaf_ids()
        li = ids.index("g1")
        assert gt.detect_hash_rot(li, genomes[1][1]) == rot
        assert gt.detect_hash_rot(li, genomes[2][1]) == -1
        gt.set_hash_rot(rot)
        reads = sample_reads(rng, genomes, 100, 120, 0.01)
        _compare(oracle, ot, gt, reads, 0.9)
        gt.close()


def test_gpu_builder_writes_identical_db(oracle, tmp_path):
    """The GPU builder (pf_builder_*) must write byte-identical tree.bin and .bf files to the oracle's
    restatement of BloomTree::insert / save."""
    from phagefilter_b200.bloom_tree import BloomTreeBuilder
    rng = np.random.default_rng(21)
    genomes = random_genomes(rng, 14, 1000, 3000)
    g = bytearray(genomes[3][1])
    g[10:14] = b"NNnn"
    genomes[3] = ("g3", bytes(g))
    genomes.append(("dup_of_0", genomes[0][1]))  # identical genome joins its twin (bloom_tree.rs:662-734)
    for k in (20, 5):
        d1, d2 = str(tmp_path / f"o{k}"), str(tmp_path / f"g{k}")
        oracle_build_db(oracle, genomes, k, d1, largest=4000)
        b = BloomTreeBuilder(k, 0.001, 4000)
        for gid, seq in genomes:
            b.insert(gid, seq)
        b.save(d2)
        b.close()
        f1, f2 = sorted(os.listdir(d1)), sorted(os.listdir(d2))
        assert f1 == f2
        for name in f1:
            a, c = open(os.path.join(d1, name), "rb").read(), open(os.path.join(d2, name), "rb").read()
            if name.endswith(".bf"):
                # the recorded file_path embeds the directory; compare everything before it
                cut = len(a) - (8 + len(os.path.join(d1, name)))
                assert a[:cut] == c[:cut], name
            else:
                assert a == c, name


def test_non_monotone_tree_stays_exact(oracle, tmp_path):
    """The reference names interior nodes with a random u16 (bloom_tree.rs:232-234); colliding names share
    one .bf, so an interior filter need not contain its children.  Such nodes must be evaluated exactly:
    emulate a collision by overwriting one interior node's file with a sibling subtree's filter."""
    import shutil
    rng = np.random.default_rng(77)
    genomes = random_genomes(rng, 10, 1500, 2500)
    d = str(tmp_path / "db")
    ot = oracle_build_db(oracle, genomes, 20, d, largest=4000)
    pre = ot.preorder()
    internals = [n for n, leaf, depth in pre if not leaf and depth >= 1]
    assert len(internals) >= 2
    shutil.copyfile(os.path.join(d, internals[-1] + ".bf"), os.path.join(d, internals[0] + ".bf"))
    shutil.copyfile(os.path.join(d, genomes[0][0] + ".bf"), os.path.join(d, pre[0][0] + ".bf"))  # root too
    ot = oracle.Tree.load(d)
    gt = _open(d)
    assert gt.info.n_monotone < gt.info.n_internal
    steps = gt.node_steps(1.0)
    assert steps[0] == gt.info.num_hashes  # the broken root is exact
    reads = sample_reads(rng, genomes, 300, 120, 0.0) + sample_reads(rng, genomes, 200, 120, 0.03)
    for theta in (0.5, 1.0):
        _compare(oracle, ot, gt, reads, theta)
    gt.close()


def test_lazy_steps_table(oracle, tmp_path):
    """Leaves are always exact; verified interior nodes use fewer steps at high thresholds."""
    rng = np.random.default_rng(78)
    genomes = random_genomes(rng, 16, 2000, 3000)
    d = str(tmp_path / "db")
    oracle_build_db(oracle, genomes, 20, d, largest=4000)
    gt = _open(d)
    K = gt.info.num_hashes
    assert gt.info.n_monotone == gt.info.n_internal == 15
    s1 = gt.node_steps(1.0)
    gt.set_lazy(False)
    assert (gt.node_steps(1.0) == K).all()
    gt.set_lazy(True)
    s3 = gt.node_steps(0.3)
    assert s1.max() == K and s1.min() < K and s3.max() == K  # leaves exact, interior nodes cheaper or skipped
    assert int((s1 == K).sum()) >= gt.info.n_leaves
    gt.set_exhaustive(True)
    assert (gt.node_steps(1.0) == K).all()
    gt.close()


def test_chunked_hash_cache(oracle, tmp_path):
    """A batch whose cached k-mer hashes exceed the budget is processed in several chunks of reads;
    hits, counters and work counts must not depend on the chunking."""
    rng = np.random.default_rng(79)
    genomes = random_genomes(rng, 8, 1500, 2500)
    d = str(tmp_path / "db")
    ot = oracle_build_db(oracle, genomes, 20, d, largest=4000)
    gt = _open(d)
    reads = sample_reads(rng, genomes, 300, 150, 0.0) + sample_reads(rng, genomes, 300, 150, 0.02) + [b"", b"ACGT"]
    reads += sample_reads(rng, genomes, 5, 1200, 0.0)
    for budget in (8 * 131 * 7, 8, 1 << 30):  # ~7 reads per chunk; one read per chunk; everything at once
        gt.set_hash_cache_bytes(budget)
        _compare(oracle, ot, gt, reads, 0.8)
    gt.close()


def test_kmer_memo_deep_coverage(oracle, tmp_path):
    """30x coverage of a few genomes: k-mers at the leaves are answered by the memo; results are unchanged,
    with a tiny memo (evictions) as with the default one, at -f 1.0 and 0.8, with reads holding non-ACGT bytes."""
    from phagefilter_b200 import BloomTree
    from phagefilter_b200.query import get_leaf_counts
    from phagefilter_b200.synth import make_genomes, simulate_reads
    genomes = make_genomes(3, 4, seed=99, len_lo=4000, len_hi=6000)
    d = str(tmp_path / "db")
    ot = oracle_build_db(oracle, genomes, 20, d, largest=8000)
    gt = BloomTree.load(d)
    gt.set_mode(1)  # the memo belongs to the node-at-a-time descent (with 3.7 MB tables the cost model would pick tiles here)
    reads, _ = simulate_reads(genomes, 12000, 150, seed=7, error_rates=(0.0, 0.01))
    rl = [r.tobytes() for r in reads] + [genomes[0][1][10:160].lower(), genomes[1][1][:150].replace(b"C", b"N", 2)]
    for theta in (1.0, 0.8):
        ot.reset_counts()
        want = ot.query_batch(rl, theta)
        sched = ot.query_sched(rl, theta, lazy=True)
        for budget in (0, 12 * 4096 * 8):  # default budget, then the smallest regions
            gt.set_memo(True, budget)
            gt.reset_counts()
            gt.reset_stats()
            assert gpu_query(gt, rl, theta) == want.hit_sets(len(rl))
            assert get_leaf_counts(gt) == ot.leaf_counts()
            st = gt.stats()
            assert st.pairs == sched.pairs
            # how many k-mers the memo answers depends on timing (a pair only profits from pairs that finished
            # before it); with 12,000 reads nearly all of them are in flight at once, so only demand some
            assert st.memo_hits > 0.05 * 131 * len(want.hits)
            assert st.probes_issued < 0.9 * sched.probes_sched
    gt.close()


def test_frontier_cap_splits_chunks(oracle, tmp_path):
    """A block whose (read, node) frontier outgrows the pair index is worked through in smaller chunks of reads instead of
    being refused (round-1 behaviour: PF_ERR_NOMEM).  Forced here with a tiny cap; results must not change."""
    from phagefilter_b200.query import get_leaf_counts
    rng = np.random.default_rng(31)
    genomes = random_genomes(rng, 24, 1500, 2500)
    d = str(tmp_path / "db")
    ot = oracle_build_db(oracle, genomes, 20, d, largest=3000)
    reads = sample_reads(rng, genomes, 3000, 120, 0.01) + [b"ACGTACGTAC", b""]
    gt = _open(d)
    for theta in (1.0, 0.6):
        ot.reset_counts()
        want = ot.query_batch(reads, theta)
        for mode, handover in ((1, -1), (2, 0), (2, 1)):
            gt.set_mode(mode)
            gt.set_handover(handover)
            gt.set_frontier_cap(0xFF000000)
            gt.reset_counts()
            assert gpu_query(gt, reads, theta) == want.hit_sets(len(reads))
            gt.set_frontier_cap(1500)  # fewer pairs than reads: the block must be cut several times
            gt.reset_counts()
            gt.reset_stats()
            assert gpu_query(gt, reads, theta) == want.hit_sets(len(reads)), (theta, mode)
            assert get_leaf_counts(gt) == ot.leaf_counts(), (theta, mode)
            st = gt.stats()
            if mode == 1:
                assert st.chunk_splits >= 1, theta  # a chunk outgrew the cap half-way down and was redone in halves
            else:
                assert st.probe_launches >= 3, theta  # (read, entry tile) pairs are bounded up front: several chunks
    gt.close()
