"""Full-size checks at BASELINE.json configs[1] (100 genomes x ~50 kb, default geometry, 1,000,000 x 150 bp
reads, -f 1.0) through size-independent properties, plus an oracle slice."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N_READS = 1_000_000


@pytest.fixture(scope="module")
def cfg2(tmp_path_factory):
    from phagefilter_b200 import BloomTree
    from phagefilter_b200.bloom_tree import BloomTreeBuilder
    from phagefilter_b200.synth import make_genomes, reads_to_concat, simulate_reads
    genomes = make_genomes(10, 10, 1001)
    d = str(tmp_path_factory.mktemp("cfg2") / "db")
    b = BloomTreeBuilder(20, 0.001, 1_000_000)
    for gid, seq in genomes:
        b.insert(gid, seq)
    b.save(d)
    b.close()
    reads, src = simulate_reads(genomes, N_READS, 150, 2001, error_rates=(0.0, 0.01))
    tree = BloomTree.load(d)
    yield genomes, d, reads, src, tree
    tree.close()


def _query(tree, reads):
    from phagefilter_b200.query import PackedReads, query_packed
    from phagefilter_b200.synth import reads_to_concat
    blob, offs = reads_to_concat(reads)
    p = PackedReads.from_concat(blob, offs)
    try:
        return query_packed(tree, p, 1.0, want_hits=True)
    finally:
        p.close()


def test_fullsize_properties(cfg2, oracle):
    genomes, d, reads, src, tree = cfg2
    info = tree.info
    assert (info.num_bits, info.num_hashes, info.n_leaves, info.n_nodes) == (14_377_587, 10, 100, 199)
    assert info.n_monotone == info.n_internal == 99
    ids = tree.leaf_ids()
    leaf_of = {gid: i for i, gid in enumerate(ids)}
    tree.reset_counts()
    off, leaf = _query(tree, reads)
    counts = tree.leaf_counts().astype(np.int64)
    # (1) histogram == hit list: every hit counted once
    assert counts.sum() == len(leaf) == int(off[-1])
    assert (np.bincount(leaf, minlength=100) == counts).all()
    # (2) no false negatives: an error-free read always reaches the leaf of its source genome
    n_hits = np.diff(off.astype(np.int64))
    src_leaf = np.array([leaf_of[genomes[g][0]] for g in src])
    exact = np.arange(N_READS) % 2 == 0  # error_rates=(0.0, 0.01): even reads are error free
    first = np.minimum(off[:-1].astype(np.int64), max(len(leaf) - 1, 0))
    has_src = np.zeros(N_READS, dtype=bool)
    for j in range(int(n_hits.max())):
        sel = n_hits > j
        has_src[sel] |= leaf[first[sel] + j] == src_leaf[sel]
    assert has_src[exact].all()
    # (3) leaves ascending within a read
    multi = np.nonzero(n_hits > 1)[0][:2000]
    for r in multi:
        seg = leaf[int(off[r]):int(off[r + 1])]
        assert (np.diff(seg.astype(np.int64)) > 0).all()
    # (4) idempotence / accumulation: a second identical block doubles every counter
    off2, leaf2 = _query(tree, reads)
    assert (off2 == off).all() and (leaf2 == leaf).all()
    assert (tree.leaf_counts().astype(np.int64) == 2 * counts).all()
    # (5) permutation invariance: shuffling the block permutes the hit lists and keeps the counters
    perm = np.random.default_rng(5).permutation(N_READS)
    tree.reset_counts()
    off3, leaf3 = _query(tree, reads[perm])
    assert (tree.leaf_counts().astype(np.int64) == counts).all()
    assert (np.diff(off3.astype(np.int64)) == n_hits[perm]).all()
    # (6) exact evaluation of every node and chunked hash cache give the same answer
    tree.reset_counts()
    tree.set_lazy(False)
    off4, leaf4 = _query(tree, reads)
    tree.set_lazy(True)
    assert (off4 == off).all() and (leaf4 == leaf).all()
    tree.set_hash_cache_bytes(64 << 20)
    tree.reset_counts()
    off5, leaf5 = _query(tree, reads)
    tree.set_hash_cache_bytes(16 << 30)
    assert (off5 == off).all() and (leaf5 == leaf).all()
    # (7) oracle slice: 3000 reads spread over the block
    sel = np.linspace(0, N_READS - 1, 3000).astype(np.int64)
    ot = oracle.Tree.load(d)
    res = ot.query_batch([reads[i].tobytes() for i in sel], 1.0)
    want = res.hit_sets(len(sel))
    for j, r in enumerate(sel):
        assert frozenset(int(x) for x in leaf[int(off[r]):int(off[r + 1])]) == want[j]


def test_fullsize_cfg2_sliced_equals_node_at_a_time(cfg2):
    """The bit-sliced tiles on the full config-2 block: identical CSR to the node-at-a-time descent (both variants)."""
    genomes, d, reads, src, tree = cfg2
    tree.set_mode(1)
    tree.reset_counts()
    off, leaf = _query(tree, reads)
    counts = tree.leaf_counts().copy()
    for handover in (0, 1):
        tree.set_mode(2)
        tree.set_handover(handover)
        tree.reset_counts()
        tree.reset_stats()
        off2, leaf2 = _query(tree, reads)
        assert tree.stats().sliced_blocks == 1
        assert (off2 == off).all() and (leaf2 == leaf).all(), handover
        assert (tree.leaf_counts() == counts).all(), handover
    tree.set_mode(0)
    tree.set_handover(-1)


@pytest.fixture(scope="module")
def cfg3s(tmp_path_factory):
    """BASELINE configs 3 and 4 at a tenth of the genomes (1,000 genomes, 1,999 nodes, 3.6 GB of filters at the default
    geometry: the same filter and tile-table sizes as the 10,000-genome tree, built on the GPU in seconds)."""
    from phagefilter_b200 import BloomTree
    from phagefilter_b200.bloom_tree import BloomTreeBuilder
    from phagefilter_b200.synth import make_genomes
    genomes = make_genomes(100, 10, 1003)
    d = str(tmp_path_factory.mktemp("cfg3s") / "db")
    b = BloomTreeBuilder(20, 0.001, 1_000_000)
    for gid, seq in genomes:
        b.insert(gid, seq)
    b.save(d)
    b.close()
    tree = BloomTree.load(d)
    yield genomes, d, tree
    tree.close()


def _query_theta(tree, reads, theta):
    from phagefilter_b200.query import PackedReads, query_packed
    from phagefilter_b200.synth import reads_to_concat
    blob, offs = reads_to_concat(reads)
    p = PackedReads.from_concat(blob, offs)
    try:
        return query_packed(tree, p, theta, want_hits=True)
    finally:
        p.close()


@pytest.mark.parametrize("shape", ["cfg3", "cfg4"])
def test_cfg3_cfg4_shapes_at_default_geometry(cfg3s, oracle, shape):
    """cfg3: 400,000 x 150 bp, 10 % spike-in with 1 % substitutions, -f 0.8.  cfg4: 4,000 x 10 kb, 0.1 % substitutions,
    -f 0.9.  Every evaluation path gives the same CSR and counters; an oracle slice against the 3.6 GB tree agrees."""
    from phagefilter_b200.synth import simulate_reads_fast
    genomes, d, tree = cfg3s
    assert tree.info.n_leaves == 1000 and tree.info.num_bits == 14_377_587
    if shape == "cfg3":
        n, L, theta, err = 400_000, 150, 0.8, (0.01,)
    else:
        n, L, theta, err = 4_000, 10_000, 0.9, (0.001,)
    reads, src = simulate_reads_fast(genomes, n, L, 2003, error_rates=err, background_frac=0.9)
    results = {}
    for name, mode, handover in (("node", 1, -1), ("tiles", 2, 0), ("tiles+handover", 2, 1), ("auto", 0, -1)):
        tree.set_mode(mode)
        tree.set_handover(handover)
        tree.reset_counts()
        off, leaf = _query_theta(tree, reads, theta)
        results[name] = (off, leaf, tree.leaf_counts().copy())
    tree.set_mode(0)
    tree.set_handover(-1)
    off, leaf, counts = results["node"]
    for name, (o, l, c) in results.items():
        assert (o == off).all() and (l == leaf).all() and (c == counts).all(), name
    assert int(counts.sum()) == len(leaf) and len(leaf) > 0.03 * n  # most spike-in reads classify
    # background reads never hit
    n_hits = np.diff(off.astype(np.int64))
    assert n_hits[src < 0].sum() == 0
    sel = np.linspace(0, n - 1, 300 if shape == "cfg3" else 40).astype(np.int64)
    sel = np.unique(np.concatenate([sel, np.nonzero(src >= 0)[0][:100 if shape == "cfg3" else 10]]))
    ot = oracle.Tree.load(d)
    want = ot.query_batch([reads[i].tobytes() for i in sel], theta).hit_sets(len(sel))
    for j, r in enumerate(sel):
        assert frozenset(int(x) for x in leaf[int(off[r]):int(off[r + 1])]) == want[j], (shape, int(r))
