"""CPU tests of the oracle against every known-answer the reference's own tests hold for this path
(SURVEY.md section 4/8c), plus an independent pure-Python restatement of the hash.

Hash VALUES are not pinned by the reference (no golden vector exists there): "parity unpinned" at the
rustc-hash boundary.  What is pinned here: canonicalisation, k-mer windows, the HashIter derivation,
filter geometry, Hamming distance, tree topology, query semantics, get_ext_id strings.
"""
import os

import numpy as np
import pytest

MASK = (1 << 64) - 1
FX_K = 0xF1357AEA2E62A9C5
S1, S2, PREV = 0x243F6A8885A308D3, 0x13198A2E03707344, 0xA4093822299F31D0


def py_mulmix(x, y):
    p = x * y
    return (p & MASK) ^ (p >> 64)


def py_hash_bytes(b: bytes) -> int:
    """Independent restatement of rustc-hash 2.x hash_bytes with Python ints."""
    n = len(b)
    le = lambda o, w: int.from_bytes(b[o:o + w], "little")
    s0, s1 = S1, S2
    if n <= 16:
        if n >= 8:
            s0 ^= le(0, 8)
            s1 ^= le(n - 8, 8)
        elif n >= 4:
            s0 ^= le(0, 4)
            s1 ^= le(n - 4, 4)
        elif n > 0:
            s0 ^= b[0]
            s1 ^= (b[n - 1] << 8) | b[n // 2]
    else:
        off = 0
        while off < n - 16:
            t = py_mulmix(s0 ^ le(off, 8), PREV ^ le(off + 8, 8))
            s0, s1 = s1, t
            off += 16
        s0 ^= le(n - 16, 8)
        s1 ^= le(n - 8, 8)
    return py_mulmix(s0, s1) ^ n


def py_fx_hash(seed: int, item: bytes, rot: int = 26) -> int:
    s = 0
    for x in (seed, len(item), py_hash_bytes(item)):
        s = ((s + x) * FX_K) & MASK
    return ((s << rot) | (s >> (64 - rot))) & MASK


def test_get_lex_less_reference_vectors(oracle):
    """file_parser.rs:396-407"""
    assert oracle.get_lex_less(b"ACGT") == b"ACGT"  # palindrome -> forward
    assert oracle.get_lex_less(b"AATG") == b"AATG"  # revcomp CATT is greater
    assert oracle.get_lex_less(b"GTAG") == b"CTAC"  # revcomp is smaller


def test_get_kmers_reference_vectors(oracle):
    """file_parser.rs:380-393: windows and canonical form; k > len and k == 0 give nothing."""
    kmers = oracle.get_kmers(b"ATCAG", 3)
    assert len(kmers) == 3
    assert kmers == [oracle.get_lex_less(b"ATC"), oracle.get_lex_less(b"TCA"), oracle.get_lex_less(b"CAG")]
    assert oracle.get_kmers(b"ATC", 4) == []
    assert oracle.get_kmers(b"ATC", 0) == []
    assert oracle.get_kmers(b"ATC", 3) == [oracle.get_lex_less(b"ATC")]


def test_complement_table(oracle):
    """bio::alphabets::dna::complement: IUPAC pairs, case preserved, everything else identity."""
    pairs = dict(zip(b"AGCTYRWSKMDVHBN", b"TCGARYWSMKHBDVN"))
    for v in range(256):
        c = oracle.lib().pfo_complement(v)
        if v in pairs:
            assert c == pairs[v]
        elif v - 32 in pairs:
            assert c == pairs[v - 32] + 32
        else:
            assert c == v
    assert oracle.revcomp(b"ACGTNacgtnRY") == b"RYnacgtNACGT"


def test_hash_iter_derivation(oracle):
    """hash_iter.rs:67-99: item0 = h1, item1 = h2, item i>=2 = (h1 + i).wrapping_mul(h2)."""
    h1, h2 = 0xFFFFFFFFFFFFFFF0, 0x123456789ABCDEF1
    out = oracle.hash_iter(h1, h2, 6)
    assert out[0] == h1 and out[1] == h2
    for i in range(2, 6):
        assert out[i] == ((h1 + i) * h2) & MASK
    for count in (0, 1, 2, 5):
        assert len(oracle.hash_iter(h1, h2, count)) == count


def test_different_seeds_different_hashes(oracle):
    """hasher.rs:36-48"""
    assert oracle.fx_hash(5, b"Hello world!") != oracle.fx_hash(10, b"Hello world!")


def test_hash_matches_independent_python_restatement(oracle):
    rng = np.random.default_rng(0)
    for n in list(range(0, 70)) + [100, 257]:
        b = rng.integers(0, 256, size=n, dtype=np.uint8).tobytes()
        assert oracle.hash_bytes(b) == py_hash_bytes(b), n
        for rot in (26, 20):
            assert oracle.fx_hash(0xDEADBEEF12345678, b, rot) == py_fx_hash(0xDEADBEEF12345678, b, rot)


def test_filter_geometry(oracle):
    """bloom_filter.rs:342-357 evaluated in f32."""
    assert oracle.needed_bits(0.001, 1_000_000) == 14_377_587
    assert oracle.optimal_num_hashes(14_377_587, 1_000_000) == 10
    assert oracle.needed_bits(0.001, 1000) == 14_378
    assert oracle.optimal_num_hashes(14_378, 1000) == 10
    b = oracle.needed_bits(1e-5, 500_000)
    assert b == 11_981_322 and oracle.optimal_num_hashes(b, 500_000) == 17
    assert oracle.needed_bits(0.01, 1000) > 1000  # test_with_rate_sizing, bloom_filter.rs:467


def test_filter_insert_contains_union_distance(oracle, tmp_path):
    """bloom_filter.rs:378-465"""
    a = oracle.Filter(14378, 10, 1, 2)
    b = oracle.Filter(14378, 10, 1, 2)
    assert a.insert(b"ACGTA") is True
    assert a.insert(b"ACGTA") is False
    assert a.contains(b"ACGTA") and not b.contains(b"ACGTA")
    b.insert(b"TTTTT")
    d0 = a.distance(b)
    assert d0 == int(np.unpackbits((a.words() ^ b.words()).view(np.uint8)).sum())
    b.union(a)
    assert b.contains(b"ACGTA") and b.contains(b"TTTTT")
    # Hamming distance 3 for 0b00101101 vs 0b10100111 (bloom_filter.rs:378-391)
    assert bin(0b00101101 ^ 0b10100111).count("1") == 3
    p = str(tmp_path / "x.bf")
    b.save(p)
    c = oracle.Filter.load(p)
    assert (c.words() == b.words()).all() and c.m == 14378 and c.K == 10 and c.seeds == (1, 2)
    assert os.path.getsize(p) == 45 + 8 * 225 + 4 + 16 + 1 + 8 + len(p)  # SURVEY App. B layout


def _four(oracle, k):
    t = oracle.Tree(k, 0.001, 1000)
    for gid, seq in [("baseline", b"ATCAG"), ("diff", b"TTTAG"), ("onediff_first", b"CTCAG"), ("onediff_mid", b"ATTAG")]:
        t.insert(gid, seq)
    return t


def test_query_reference_cases(oracle):
    """query.rs:268-380 restated against the oracle."""
    t = oracle.Tree(3, 0.001, 1000)
    t.insert("genome", b"ATCGCA")
    assert t.query_batch([b"ATCG"], 1.0).hit_sets(1) == [frozenset({0})]
    assert t.query_batch([b"AAAA"], 1.0).hit_sets(1) == [frozenset()]
    assert t.query_batch([b"ATCG", b"AAAA"], 0.0).hit_sets(2) == [frozenset({0}), frozenset({0})]

    t = _four(oracle, 5)
    t.query_batch([b"ATCAG"], 0.1)
    c = dict(t.leaf_counts())
    assert c["baseline"] >= 1 and c["diff"] == 0

    t = _four(oracle, 4)
    t.query_batch([b"TCAG"], 0.1)
    c = dict(t.leaf_counts())
    assert c["baseline"] >= 1 and c["diff"] == 0

    t = _four(oracle, 4)
    t.query_batch([b"TCAG", b"ATCA"], 0.51)
    c = dict(t.leaf_counts())
    assert c["baseline"] >= 1 and c["diff"] == 0

    t = _four(oracle, 4)
    t.query_batch([b"TCAG"], 0.1)
    t.query_batch([b"ATCA"], 0.1)
    c = dict(t.leaf_counts())
    assert c["baseline"] >= 2 and c["diff"] == 0


def test_tree_topology(oracle, tmp_path):
    """bloom_tree.rs:458-734: left = existing, right = new; an identical genome joins its twin; save/load."""
    t = oracle.Tree(3, 0.001, 1000)
    t.insert("g1", b"ATCAGTTT")
    assert t.preorder() == [("g1", True, 0)]
    t.insert("g2", b"GGGCCCAA")
    assert [(n, l) for n, l, _ in t.preorder()] == [("Internal_Node_0", False), ("g1", True), ("g2", True)]
    t.insert("g3", b"ATCAGTTT")  # same as g1 -> descends to g1's side (distance 0, ties go left)
    pre = t.preorder()
    assert [n for n, _, _ in pre] == ["Internal_Node_0", "Internal_Node_1", "g1", "g3", "g2"]
    assert [d for _, _, d in pre] == [0, 1, 2, 2, 1]
    t.insert("g4", b"GGGCCCAA")  # same as g2 -> goes right
    assert [n for n, _, _ in t.preorder()] == ["Internal_Node_0", "Internal_Node_1", "g1", "g3", "Internal_Node_2", "g2", "g4"]
    d = str(tmp_path / "db")
    t.save(d)
    t2 = oracle.Tree.load(d)
    assert t2.preorder() == t.preorder() and t2.kmer_size == 3
    assert t2.seeds == t.seeds and t2.num_bits == t.num_bits and t2.num_hashes == t.num_hashes
    t2.prune_tree(1)
    assert t2.leaf_ids() == ["Internal_Node_1", "Internal_Node_2"]
    t2.prune_tree(0)
    assert t2.leaf_ids() == ["Internal_Node_0"]


def test_need_rounding(oracle):
    """(threshold * n_k as f32).ceil() as usize in f32 (query.rs:48)."""
    f32 = np.float32
    for th in (0.0, 0.1, 0.3, 0.51, 0.8, 0.9, 1.0, 1.5, -0.5):
        for n in (0, 1, 3, 81, 131, 9981, 1 << 25):
            want = int(max(0.0, np.ceil(f32(th) * f32(n))))
            assert oracle.need(th, n) == want, (th, n)
    assert oracle.need(0.9, 9981) == 8983
    assert oracle.need(0.8, 131) == 105


def test_probe_counts_consistent(oracle):
    """The kernel schedule never issues more probes than the reference semantics and decides identically."""
    rng = np.random.default_rng(1)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    t = oracle.Tree(20, 0.001, 3000)
    gs = [acgt[rng.integers(0, 4, size=2500)].tobytes() for _ in range(6)]
    for i, g in enumerate(gs):
        t.insert(f"g{i}", g)
    reads = [g[s:s + 150] for g in gs for s in (0, 500, 1000)] + [acgt[rng.integers(0, 4, size=150)].tobytes() for _ in range(20)]
    for th in (0.3, 0.8, 1.0):
        r = t.query_batch(reads, th)
        want = r.hit_sets(len(reads))
        exact = t.query_sched(reads, th, lazy=False)
        lazy = t.query_sched(reads, th, lazy=True)
        assert exact.hit_sets(len(reads)) == want and lazy.hit_sets(len(reads)) == want
        assert exact.pairs == r.pairs and lazy.pairs > 0  # the plan may skip saturated top nodes entirely
        assert 0 < exact.probes_sched <= r.probes_ref and 0 < lazy.probes_sched


def test_reference_faithful_lru_mode_equals_resident_mode(oracle, tmp_path):
    """BFLruCache (cache.rs:55-88) + the block loop of main.rs:334-368: filters re-read from disk through an LRU of
    --cache-size entries, query_batch per --block-size-reads block.  Same hits, counters and work as one resident pass."""
    from tests.util import oracle_build_db, random_genomes, sample_reads
    rng = np.random.default_rng(5)
    genomes = random_genomes(rng, 14, 1200, 2500)
    d = str(tmp_path / "db")
    oracle_build_db(oracle, genomes, 20, d, largest=3000)
    reads = sample_reads(rng, genomes, 350, 100, 0.01) + [b"ACGT", b"N" * 40]
    full = oracle.Tree.load(d)
    want = full.query_batch(reads, 0.8)
    for cache, block in ((1, 100), (3, 7), (10, 100), (64, 1000)):
        lazy = oracle.Tree.load_lazy(d, cache_size=cache)
        got = lazy.query_blocks(reads, 0.8, block_size=block, threads=4)
        assert sorted(map(tuple, got.hits.tolist())) == sorted(map(tuple, want.hits.tolist()))
        assert lazy.leaf_counts() == full.leaf_counts()
        assert got.probes_ref == want.probes_ref
        loads, hits, nbytes = lazy.cache_stats()
        n_blocks = -(-len(reads) // block)
        assert loads >= 1 and nbytes == loads * ((full.num_bits + 63) // 64 * 8)
        if cache == 1:
            assert loads >= n_blocks  # a one-entry cache reloads at least the root's sibling chain every block
        with pytest.raises(RuntimeError):
            lazy.query_sched(reads, 0.8)
