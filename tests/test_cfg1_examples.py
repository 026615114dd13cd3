"""BASELINE.json config #1: build from the reference's examples/genomes/viral_genome_dir (107 genomes, default
geometry) and query examples/test_reads (9 x 10,000 x 100 bp), k default, -f 1.0.

Inputs are reference data: read from /root/reference/examples when present (this container) or from
tests/_cfg1_data (git-ignored copy made by tests/golden/fetch_cfg1.py; travels to the GPU box).  The expected
CLASSIFICATION.csv is committed (tests/golden/cfg1_classification.csv, oracle-generated with fixed seeds)."""
import os
import subprocess
from collections import Counter

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "cfg1_classification.csv")
BIN = os.path.join(ROOT, "phagefilter_b200", "bin", "phage_filter")
SEEDS = (0x5EED0001, 0x5EED0002)


def _data():
    for base in ("/root/reference/examples", os.path.join(ROOT, "tests", "_cfg1_data")):
        g = os.path.join(base, "genomes", "viral_genome_dir") if base.endswith("examples") else os.path.join(base, "viral_genome_dir")
        r = os.path.join(base, "test_reads")
        if os.path.isdir(g) and os.path.isdir(r):
            return g, r
    pytest.skip("config #1 inputs not available (run tests/golden/fetch_cfg1.py where /root/reference exists)")


def _genome_records(gdir):
    """Records in the order the host driver inserts them: files popped from the end of the sorted listing."""
    from phagefilter_b200.file_parser import has_supported_extension, read_records
    files = sorted(os.path.join(gdir, n) for n in os.listdir(gdir) if has_supported_extension(os.path.join(gdir, n)))
    for f in reversed(files):
        yield from read_records(f)


def _read_files(rdir):
    from phagefilter_b200.file_parser import has_supported_extension
    return list(reversed(sorted(os.path.join(rdir, n) for n in os.listdir(rdir) if has_supported_extension(os.path.join(rdir, n)))))


@pytest.fixture(scope="module")
def oracle_cfg1(oracle, tmp_path_factory):
    gdir, rdir = _data()
    t = oracle.Tree(20, 0.001, 1_000_000, *SEEDS)
    n = 0
    for rec in _genome_records(gdir):
        t.insert(rec.id, rec.sequence)
        n += 1
    assert n == 107 and t.num_leaves == 107 and t.num_nodes == 213
    assert (t.num_bits, t.num_hashes) == (14_377_587, 10)
    return t, gdir, rdir


def test_cfg1_oracle_matches_committed_classification(oracle_cfg1):
    from phagefilter_b200.file_parser import read_records
    t, gdir, rdir = oracle_cfg1
    t.reset_counts()
    total = 0
    for f in _read_files(rdir):
        recs = list(read_records(f))
        total += len(recs)
        for lo in range(0, len(recs), 10_000):
            t.query_batch([r.sequence for r in recs[lo:lo + 10_000]], 1.0, want_hits=False)
    assert total == 90_000
    csv = t.classification_csv()
    if not os.path.exists(GOLD):  # first run in the container that has the reference: create the golden file
        open(GOLD, "w").write(csv)
    assert csv == open(GOLD).read()
    # sanity on the data itself: reads are named <genome>_<n> (bench/utils.py:194-212); error-free files map back
    counts = dict(line.rsplit(",", 1) for line in csv.strip().split("\n"))
    assert sum(int(v) for v in counts.values()) >= 30_000


@pytest.mark.gpu
def test_cfg1_gpu_cli_end_to_end(oracle, tmp_path):
    """phage_filter build + query on config #1: CLASSIFICATION.csv byte-identical to the committed golden,
    POS/NEG files of one read file identical (as record multisets) to the reference driver restated with the oracle."""
    from phagefilter_b200.file_parser import read_records
    from tests.util import parse_filter_file, reference_query_outputs
    gdir, rdir = _data()
    db, out = str(tmp_path / "tree"), str(tmp_path / "output")
    p = subprocess.run([BIN, "build", "--genomes", gdir, "--db-path", db, "--seed-one", str(SEEDS[0]), "--seed-two", str(SEEDS[1]),
                        "--node-names", "counter"], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    assert len([f for f in os.listdir(db) if f.endswith(".bf")]) == 213
    p = subprocess.run([BIN, "query", "-r", rdir, "-o", out, "-d", db, "-f", "1.0"], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    assert open(os.path.join(out, "CLASSIFICATION.csv")).read() == open(GOLD).read()
    one = _read_files(rdir)[0]
    out2 = str(tmp_path / "output2")
    p = subprocess.run([BIN, "query", "-r", one, "-o", out2, "-d", db, "-f", "1.0", "--pos-filter", "--neg-filter", "-b", "1000"],
                       capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    recs = list(read_records(one))
    csv, pos, neg = reference_query_outputs(oracle, db, recs, 1.0, 1000, True, True)
    assert open(os.path.join(out2, "CLASSIFICATION.csv")).read() == csv
    assert Counter(parse_filter_file(os.path.join(out2, "POS_FILTERING.fq"))) == Counter(pos)
    assert Counter((i, s, q) for i, _, s, q in parse_filter_file(os.path.join(out2, "NEG_FILTERING.fq"))) == Counter(neg)
    # precision of the classification on error-free reads: the source genome is always among the matches
    truth_hits = sum(1 for rid, gs, _, _ in pos if rid.rsplit("_", 1)[0] in gs)
    assert truth_hits == len(pos) or "e0.0" not in one
