"""Two independent restatements of the reference agree: the C oracle (oracle/pf_oracle.c) and the plain-Python one
(tests/pyref.py, written from the reference's sources).  Random small databases and reads: geometry, tree topology,
filter bits of every node, hit sets, leaf counters accumulated over two blocks, CLASSIFICATION.csv, pruning --
k below/inside/above the 2-bit fast-path range, both FxHasher rotates, IUPAC / lower-case / non-nucleotide bytes,
reads shorter than k, thresholds 0 .. 1."""
import numpy as np
import pytest

from tests import pyref


def _rand_seq(rng, n, alphabet):
    a = np.frombuffer(alphabet, dtype=np.uint8)
    return a[rng.integers(0, len(a), size=n)].tobytes()


@pytest.mark.parametrize("case", range(8))
def test_oracle_and_python_restatement_agree(oracle, tmp_path, case):
    rng = np.random.default_rng(100 + case)
    k = [5, 16, 17, 20, 21, 31, 32, 33][case]
    rot = 20 if case == 3 else 26
    fpr, largest = [(0.01, 400), (0.001, 300), (0.05, 500), (0.01, 350)][case % 4]
    seeds = (int(rng.integers(1, 2**63)), int(rng.integers(1, 2**63)))
    alphabet = b"ACGT" if case % 2 == 0 else b"ACGTACGTACGTNacgtRYKM"
    n_genomes = int(rng.integers(1, 8))
    # families of related genomes so that the greedy insert has real choices; one genome shorter than k
    genomes, base = [], _rand_seq(rng, 260, alphabet)
    for g in range(n_genomes):
        if g % 3 == 0:
            base = _rand_seq(rng, int(rng.integers(k + 1, 300)), alphabet)
        s = bytearray(base)
        for _ in range(len(s) // 20):
            s[int(rng.integers(0, len(s)))] = alphabet[int(rng.integers(0, len(alphabet)))]
        genomes.append((f"g{g}", bytes(s)))
    if case == 5:
        genomes.append(("tiny", b"ACGT"))  # no k-mers: an empty leaf

    o = oracle.Tree(k, fpr, largest, seeds[0], seeds[1], rot=rot)
    p = pyref.Tree(k, fpr, largest, seeds[0], seeds[1], rot=rot)
    for gid, seq in genomes:
        o.insert(gid, seq)
        p.insert(gid, seq)
    assert (o.num_bits, o.num_hashes) == (p.m, p.K)
    assert [(leaf, d) for _, leaf, d in o.preorder()] == p.preorder()
    assert o.leaf_ids() == [n.tax_id for n in p.leaves()]

    # filter bits of every node (pre-order) through the oracle's saved files
    d = str(tmp_path / "db")
    o.save(d)
    nodes = []

    def walk(n):
        if n is None:
            return
        nodes.append(n)
        walk(n.left)
        walk(n.right)
    walk(p.root)
    for (name, _, _), node in zip(o.preorder(), nodes):
        words = oracle.Filter.load(f"{d}/{name}.bf", rot=rot).words()
        bits = np.unpackbits(words.view(np.uint8), bitorder="little")[:p.m]
        assert (bits == np.frombuffer(bytes(node.filter.bits), dtype=np.uint8)).all(), name

    # reads: error-free, mutated, unrelated, shorter than k, exactly k
    reads = []
    for i in range(40):
        gid, seq = genomes[int(rng.integers(0, len(genomes)))]
        if len(seq) <= k or i % 7 == 0:
            reads.append(_rand_seq(rng, int(rng.integers(0, 2 * k + 3)), alphabet))
            continue
        L = int(rng.integers(k, min(len(seq), 3 * k) + 1))
        s0 = int(rng.integers(0, len(seq) - L + 1))
        r = bytearray(seq[s0:s0 + L])
        if i % 3 == 0:
            r[int(rng.integers(0, L))] = alphabet[int(rng.integers(0, len(alphabet)))]
        reads.append(bytes(r))
    for theta in (1.0, 0.8, 0.33, 0.0):
        o.reset_counts()
        for n in p.leaves():
            n.mapped_reads = 0
        for lo in (0, 25):  # two blocks: counters accumulate (query.rs:143)
            blk = reads[lo:lo + 25] if lo == 0 else reads[25:]
            got = {(int(r), int(l)) for r, l in o.query_batch(blk, theta).hits}
            assert got == p.query_batch(blk, theta), (theta, lo)
        assert o.classification_csv() == p.classification_csv()
        assert [c for _, c in o.leaf_counts()] == [n.mapped_reads for n in p.leaves()]

    # --search-depth: nodes at the cut become leaves labelled with their own id (bloom_tree.rs:302-330)
    for depth in (0, 1, 2):
        o2 = oracle.Tree.load(d, rot=rot)
        p2 = pyref.Tree(k, fpr, largest, seeds[0], seeds[1], rot=rot)
        for gid, seq in genomes:
            p2.insert(gid, seq)
        o2.prune_tree(depth)
        p2.prune_tree(depth)
        assert [(leaf, dd) for _, leaf, dd in o2.preorder()] == p2.preorder()
        got = {(int(r), int(l)) for r, l in o2.query_batch(reads, 0.5).hits}
        assert got == p2.query_batch(reads, 0.5), depth
        assert [c for _, c in o2.leaf_counts()] == [n.mapped_reads for n in p2.leaves()]


def test_python_restatement_reproduces_golden_fixture():
    """tests/golden/expected.json is oracle-generated; the plain-Python restatement, built from the same genomes,
    reproduces its geometry, topology, leaf order, hit sets and CLASSIFICATION.csv (theta 1.0 / 0.8 and --search-depth 2)."""
    import json
    import os
    from phagefilter_b200.file_parser import read_records
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    exp = json.load(open(os.path.join(gold, "expected.json")))
    genomes = [(r.id, r.sequence) for r in read_records(os.path.join(gold, "genomes.fa"))]
    reads = [r.sequence for r in read_records(os.path.join(gold, "reads.fq"))]

    def build():
        t = pyref.Tree(exp["k"], exp["fpr"], exp["largest"], 0x5EED0001, 0x5EED0002)
        for gid, seq in genomes:
            t.insert(gid, seq)
        return t
    t = build()
    assert (t.m, t.K) == (exp["num_bits"], exp["num_hashes"])
    assert [n.tax_id for n in t.leaves()] == exp["leaf_ids"]
    assert t.preorder() == [(bool(leaf), d) for _, leaf, d in exp["preorder"]]
    for theta in ("1.0", "0.8"):
        for n in t.leaves():
            n.mapped_reads = 0
        hits = t.query_batch(reads, float(theta))
        got = [sorted(l for r, l in hits if r == i) for i in range(len(reads))]
        assert got == exp["cases"][theta]["hits"], theta
        assert t.classification_csv() == exp["cases"][theta]["csv"]
    t2 = build()
    t2.prune_tree(2)
    hits = t2.query_batch(reads, 0.8)
    case = exp["cases"]["depth2"]
    assert [sorted(l for r, l in hits if r == i) for i in range(len(reads))] == case["hits"]
    assert len(t2.leaves()) == len(case["leaf_ids"]) and t2.classification_csv().count("\n") == case["csv"].count("\n")
    # pruned leaves are interior nodes: their names are arbitrary (random u16 in the reference), the counts are not
    assert [c.split(",")[1] for c in t2.classification_csv().split()] == [c.split(",")[1] for c in case["csv"].split()]
