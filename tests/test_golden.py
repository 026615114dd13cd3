"""Golden fixture (tests/golden, made by tests/golden/make_golden.py with the oracle).
CPU: the oracle still reproduces the committed vectors.  GPU: the product reproduces them through the C ABI
without the oracle in the loop."""
import json
import os

import pytest

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
EXP = json.load(open(os.path.join(HERE, "expected.json")))
DB = os.path.join(HERE, "db")


def _reads():
    from phagefilter_b200.file_parser import read_records
    return list(read_records(os.path.join(HERE, "reads.fq")))


def test_oracle_reproduces_golden(oracle):
    reads = [r.sequence for r in _reads()]
    assert len(reads) == len(EXP["cases"]["1.0"]["hits"])
    t = oracle.Tree.load(DB)
    assert t.leaf_ids() == EXP["leaf_ids"]
    assert [[n, int(l), d] for n, l, d in t.preorder()] == EXP["preorder"]
    assert (t.num_bits, t.num_hashes, t.kmer_size) == (EXP["num_bits"], EXP["num_hashes"], EXP["k"])
    for th in ("1.0", "0.8", "0.3"):
        t.reset_counts()
        res = t.query_batch(reads, float(th))
        c = EXP["cases"][th]
        assert [sorted(s) for s in res.hit_sets(len(reads))] == c["hits"]
        assert t.classification_csv() == c["csv"]
        assert (res.pairs, res.probes_ref) == (c["pairs"], c["probes_ref"])
    for depth in (0, 2):
        t2 = oracle.Tree.load(DB)
        t2.prune_tree(depth)
        res = t2.query_batch(reads, 0.8)
        c = EXP["cases"][f"depth{depth}"]
        assert t2.leaf_ids() == c["leaf_ids"] and t2.classification_csv() == c["csv"]


def test_oracle_rebuilds_golden_db_bytes(oracle, tmp_path):
    """Rebuilding the DB from genomes.fa reproduces the committed tree.bin and filter payloads."""
    from phagefilter_b200.file_parser import read_records
    t = oracle.Tree(EXP["k"], EXP["fpr"], EXP["largest"], 0x5EED0001, 0x5EED0002)
    for rec in read_records(os.path.join(HERE, "genomes.fa")):
        t.insert(rec.id, rec.sequence)
    d = str(tmp_path / "db")
    t.save(d)
    assert open(os.path.join(d, "tree.bin"), "rb").read() == open(os.path.join(DB, "tree.bin"), "rb").read()
    for name in os.listdir(DB):
        if name.endswith(".bf"):
            a = oracle.Filter.load(os.path.join(DB, name))
            b = oracle.Filter.load(os.path.join(d, name))
            assert (a.words() == b.words()).all() and a.K == b.K and a.seeds == b.seeds


@pytest.mark.gpu
def test_gpu_reproduces_golden(tmp_path):
    from phagefilter_b200 import BloomTree, ResultMap, query_batch, save_leaf_counts
    recs = _reads()
    for th in ("1.0", "0.8", "0.3"):
        tree = BloomTree.load(DB)
        assert tree.leaf_ids() == EXP["leaf_ids"]
        rm = ResultMap()
        # the reference's block loop (main.rs:334-368): blocks of 100 reads
        got = {}
        for lo in range(0, len(recs), 100):
            blk = recs[lo:lo + 100]
            query_batch(tree, blk, float(th), rm)
            for r in blk:
                got[r.id] = sorted(rm.read_map.get(r.id, ()))
            rm.empty_read_map()
        c = EXP["cases"][th]
        ids = EXP["leaf_ids"]
        for i, r in enumerate(recs):
            assert got[r.id] == sorted(ids[j] for j in c["hits"][i]), (th, r.id)
        out = str(tmp_path / f"CLASSIFICATION_{th}.csv")
        save_leaf_counts(tree, out)
        assert open(out).read() == c["csv"]
        tree.close()
    for depth in (0, 2):
        tree = BloomTree.load(DB)
        tree.prune_tree(depth)
        c = EXP["cases"][f"depth{depth}"]
        assert tree.leaf_ids() == c["leaf_ids"]
        query_batch(tree, recs, 0.8, None)
        out = str(tmp_path / f"CLASSIFICATION_d{depth}.csv")
        save_leaf_counts(tree, out)
        assert open(out).read() == c["csv"]
        tree.close()


@pytest.mark.gpu
def test_gpu_builder_reproduces_golden_db(tmp_path):
    from phagefilter_b200.bloom_tree import BloomTreeBuilder
    from phagefilter_b200.file_parser import read_records
    b = BloomTreeBuilder(EXP["k"], EXP["fpr"], EXP["largest"])
    for rec in read_records(os.path.join(HERE, "genomes.fa")):
        b.insert(rec.id, rec.sequence)
    d = str(tmp_path / "db")
    b.save(d)
    b.close()
    assert open(os.path.join(d, "tree.bin"), "rb").read() == open(os.path.join(DB, "tree.bin"), "rb").read()
    for name in os.listdir(DB):
        if name.endswith(".bf"):
            a, c = open(os.path.join(DB, name), "rb").read(), open(os.path.join(d, name), "rb").read()
            payload = 45 + 8 * ((EXP["num_bits"] + 63) // 64) + 4 + 16  # everything before the recorded path
            assert a[:payload] == c[:payload], name
