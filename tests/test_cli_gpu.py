"""The C++ host driver `phage_filter` (reference command line over the C ABI) against the reference's driver
loop (src/main.rs:249-376) restated with the oracle: CLASSIFICATION.csv byte-identical; POS/NEG files equal
as multisets of records with genome annotations compared as sets (the reference's own order is unspecified,
SURVEY.md section 0)."""
import gzip
import os
import shutil
import subprocess
from collections import Counter

import pytest

from tests.util import parse_filter_file, reference_query_outputs

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "phagefilter_b200", "bin", "phage_filter")
GOLD = os.path.join(ROOT, "tests", "golden")


def run(*args):
    p = subprocess.run([BIN, *args], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    return p.stdout


def _records(path):
    from phagefilter_b200.file_parser import read_records
    return list(read_records(path))


@pytest.mark.parametrize("theta,block,depth", [(1.0, 100, None), (0.8, 7, None), (0.3, 1000, None), (0.8, 64, 2)])
def test_query_outputs(oracle, tmp_path, theta, block, depth):
    out = str(tmp_path / "out")
    os.makedirs(out)
    open(os.path.join(out, "stale.txt"), "w").write("x")  # --out is deleted and re-created (main.rs:380-391)
    args = ["query", "-r", os.path.join(GOLD, "reads.fq"), "-o", out, "-d", os.path.join(GOLD, "db"), "-f", str(theta),
            "-b", str(block), "--pos-filter", "--neg-filter", "-t", "4", "-c", "10"]
    if depth is not None:
        args += ["--search-depth", str(depth)]
    stdout = run(*args)
    assert "Querying reads..." in stdout and "Filtering settings: positive=true; negative=true" in stdout
    assert stdout.rstrip().endswith("Finished.")
    assert sorted(os.listdir(out)) == ["CLASSIFICATION.csv", "NEG_FILTERING.fq", "POS_FILTERING.fq"]
    recs = _records(os.path.join(GOLD, "reads.fq"))
    csv, pos, neg = reference_query_outputs(oracle, os.path.join(GOLD, "db"), recs, theta, block, True, True, depth)
    assert open(os.path.join(out, "CLASSIFICATION.csv")).read() == csv
    got_pos = parse_filter_file(os.path.join(out, "POS_FILTERING.fq"))
    got_neg = parse_filter_file(os.path.join(out, "NEG_FILTERING.fq"))
    assert Counter(got_pos) == Counter(pos)
    assert Counter((i, s, q) for i, _, s, q in got_neg) == Counter(neg)
    assert len(got_pos) + len(got_neg) == len(recs)


def test_query_without_filter_flags_fasta_gz_and_dir(oracle, tmp_path):
    """No POS/NEG files without the flags; FASTA input gives .fa outputs; gzip and directories are read."""
    reads_dir = tmp_path / "reads"
    reads_dir.mkdir()
    recs = _records(os.path.join(GOLD, "reads.fq"))
    with gzip.open(reads_dir / "a.fasta.gz", "wb") as f:
        for r in recs[:120]:
            f.write(b">%s extra words\n%s\n" % (r.id.encode(), r.sequence))
    with open(reads_dir / "b.fa", "wb") as f:
        for r in recs[120:]:
            s = r.sequence
            f.write(b">%s\n%s\n%s\n" % (r.id.encode(), s[:40], s[40:]))  # multi-line FASTA
    (reads_dir / "ignored.txt").write_text("not a sequence file")
    out = str(tmp_path / "out")
    stdout = run("query", "--reads", str(reads_dir), "--out", out, "--db-path", os.path.join(GOLD, "db"))
    assert "positive=false; negative=false" in stdout
    assert os.listdir(out) == ["CLASSIFICATION.csv"]
    csv, _, _ = reference_query_outputs(oracle, os.path.join(GOLD, "db"), recs, 1.0, 100, False, False)
    assert open(os.path.join(out, "CLASSIFICATION.csv")).read() == csv
    out2 = str(tmp_path / "out2")
    run("query", "-r", str(reads_dir), "-o", out2, "-d", os.path.join(GOLD, "db"), "--neg-filter", "-f", "0.8")
    assert sorted(os.listdir(out2)) == ["CLASSIFICATION.csv", "NEG_FILTERING.fa"]
    # files are popped from the end of the (sorted) listing: b.fa then a.fasta.gz
    order = [r for r in recs[120:]] + [r for r in recs[:120]]
    _, _, neg = reference_query_outputs(oracle, os.path.join(GOLD, "db"), [type(r)(r.sequence, None, r.id) for r in order],
                                        0.8, 100, False, True)
    got = parse_filter_file(os.path.join(out2, "NEG_FILTERING.fa"))
    assert Counter((i, s, q) for i, _, s, q in got) == Counter(neg)


def test_build_and_add_match_oracle(oracle, tmp_path):
    """`build` then `add` write the same tree.bin and filter payloads as the oracle's BloomTree::insert/save."""
    from phagefilter_b200.file_parser import read_records
    genomes = list(read_records(os.path.join(GOLD, "genomes.fa")))
    first, rest = tmp_path / "first.fa", tmp_path / "rest.fa"
    with open(first, "wb") as f:
        for g in genomes[:5]:
            f.write(b">%s\n%s\n" % (g.id.encode(), g.sequence))
    with open(rest, "wb") as f:
        for g in genomes[5:]:
            f.write(b">%s\n%s\n" % (g.id.encode(), g.sequence))
    db = str(tmp_path / "db")
    out = run("build", "-g", str(first), "-d", db, "-k", "20", "-f", "0.01", "-l", "3000", "--seed-one", str(0x5EED0001),
              "--seed-two", str(0x5EED0002), "--node-names", "counter")
    assert "Building the SBT..." in out and "Finished." in out
    t = oracle.Tree(20, 0.01, 3000, 0x5EED0001, 0x5EED0002)
    for g in genomes[:5]:
        t.insert(g.id, g.sequence)
    ref1 = str(tmp_path / "ref1")
    t.save(ref1)
    assert open(os.path.join(db, "tree.bin"), "rb").read() == open(os.path.join(ref1, "tree.bin"), "rb").read()
    out = run("add", "-g", str(rest), "-d", db)
    assert "Adding new genomes to the SBT..." in out
    # `add` names new interior nodes with fresh u16 values: compare topology, leaves and filter payloads
    got, want = oracle.Tree.load(db), oracle.Tree.load(os.path.join(GOLD, "db"))
    assert [(l, d) for _, l, d in got.preorder()] == [(l, d) for _, l, d in want.preorder()]
    assert got.leaf_ids() == want.leaf_ids()
    for (n1, _, _), (n2, _, _) in zip(got.preorder(), want.preorder()):
        a = oracle.Filter.load(os.path.join(db, n1 + ".bf"))
        b = oracle.Filter.load(os.path.join(GOLD, "db", n2 + ".bf"))
        assert (a.words() == b.words()).all()
