"""The `phage_filter query` driver end to end on a machine WITHOUT a GPU: the database/query entry points of the C ABI are
replaced by a preloaded test stub (tests/host/stub_pfgpu.cpp) that fabricates hits by a fixed rule, so everything the
driver itself does is checked byte for byte -- ingest thread and chunking (tiny parse buffers force many chunks and
carried blocks), GPU batches smaller than a chunk, the per-block ResultMap semantics with duplicated ids, POS/NEG files,
CLASSIFICATION.csv, the re-created output directory, stdout lines and the panic exit code.  The real query path is NOT
involved (that is what the -m gpu tests are for); the 2-bit packer is the real one."""
import gzip
import os
import subprocess

import numpy as np
import pytest

from tests.test_host_outputs_cpu import expected

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "phagefilter_b200", "bin", "phage_filter")
STUB_SRC = os.path.join(ROOT, "tests", "host", "stub_pfgpu.cpp")


@pytest.fixture(scope="module")
def stub(tmp_path_factory):
    if not os.path.exists(BIN):
        pytest.skip("phage_filter binary not built")
    so = str(tmp_path_factory.mktemp("stub") / "stub_pfgpu.so")
    subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-o", so, STUB_SRC], check=True)
    return so


def run_query(stub, args, env=None):
    e = dict(os.environ, LD_PRELOAD=stub, **(env or {}))
    return subprocess.run([BIN, "query", *args], capture_output=True, env=e, timeout=120)


def make_records(rng, n, fastq):
    alpha = np.frombuffer(b"ACGTacgtNn", dtype=np.uint8)
    recs = []
    for i in range(n):
        L = int(rng.choice([0, 1, 19, 20, 30, 100, 150]))
        rid = f"dup{i % 23}" if i % 5 == 0 else f"r{i}"
        q = bytes(rng.integers(46, 74, size=L, dtype=np.uint8)) if fastq else None
        recs.append((rid, alpha[rng.integers(0, 10, size=L)].tobytes(), q))
    return recs


def write_records(path, recs, fastq, gz=False):
    blob = b"".join((b"@" + r.encode() + b" d\n" + s + b"\n+\n" + q + b"\n") if fastq else (b">" + r.encode() + b" d\n" + s + b"\n")
                    for r, s, q in recs)
    with (gzip.open(path, "wb") if gz else open(path, "wb")) as f:
        f.write(blob)


def expected_csv(n_reads, n_leaves=9):
    counts = [0] * n_leaves
    for g in range(n_reads):
        for l in {(g * 7 + j * 5) % n_leaves for j in range(0 if g % 4 == 0 else g % 3)}:
            counts[l] += 1
    return "".join(f"genome_{l},{c}\n" for l, c in enumerate(counts) if c)


@pytest.mark.parametrize("fastq", [True, False])
def test_query_driver_outputs(stub, tmp_path, fastq):
    rng = np.random.default_rng(31 + fastq)
    recs = make_records(rng, 3000, fastq)
    reads = tmp_path / ("reads.fq" if fastq else "reads.fa")
    write_records(reads, recs, fastq)
    ext = "fq" if fastq else "fa"
    want = [(r, s, q if fastq else b"") for r, s, q in recs]
    for block, buf, batch, threads, pos, neg in ((100, None, None, None, 1, 1), (7, 3000, 50, 5, 1, 1), (64, 20000, 1000, 2, 1, 0),
                                                 (1, 900, 3, 8, 0, 1), (1000, 50000, None, 3, 1, 1), (5000, 4096, 100, 4, 1, 1)):
        out = tmp_path / f"out_{block}"
        out.mkdir()
        (out / "stale.txt").write_text("x")  # --out is deleted and re-created (main.rs:380-391)
        args = ["-r", str(reads), "-o", str(out), "-d", str(tmp_path), "-b", str(block), "-f", "0.8", "-t", "4", "-c", "10"]
        args += ["--pos-filter"] if pos else []
        args += ["--neg-filter"] if neg else []
        args += ["--gpu-batch-reads", str(batch)] if batch else []
        args += ["--host-threads", str(threads)] if threads else []
        p = run_query(stub, args, {"PF_PARSE_BUF_BYTES": str(buf)} if buf else None)
        assert p.returncode == 0, p.stderr.decode()
        so = p.stdout.decode()
        assert so.startswith("Querying reads...\n") and so.rstrip().endswith("Finished.")
        assert f"Filtering settings: positive={'true' if pos else 'false'}; negative={'true' if neg else 'false'}" in so
        files = ["CLASSIFICATION.csv"] + (["NEG_FILTERING." + ext] if neg else []) + (["POS_FILTERING." + ext] if pos else [])
        assert sorted(os.listdir(out)) == sorted(files)
        want_pos, want_neg = expected(want, block, 9, pos, neg, fastq)
        if pos:
            assert open(out / f"POS_FILTERING.{ext}", "rb").read() == want_pos, (block, buf, batch)
        if neg:
            assert open(out / f"NEG_FILTERING.{ext}", "rb").read() == want_neg, (block, buf, batch)
        assert open(out / "CLASSIFICATION.csv").read() == expected_csv(len(recs))


def test_query_driver_directory_no_flags_and_errors(stub, tmp_path):
    rng = np.random.default_rng(40)
    d = tmp_path / "reads"
    d.mkdir()
    parts = {"a.fasta.gz": make_records(rng, 120, False), "b.fa": make_records(rng, 333, False), "c.fna.gz": make_records(rng, 1, False)}
    for name, recs in parts.items():
        write_records(d / name, recs, False, gz=name.endswith(".gz"))
    (d / "ignored.txt").write_text("not a sequence file")
    order = [x for name in sorted(parts, reverse=True) for x in parts[name]]  # files are popped from the end of the listing
    out = tmp_path / "o"
    p = run_query(stub, ["--reads", str(d), "--out", str(out), "--db-path", str(tmp_path)], {"PF_PARSE_BUF_BYTES": "5000"})
    assert p.returncode == 0, p.stderr.decode()
    assert "positive=false; negative=false" in p.stdout.decode() and os.listdir(out) == ["CLASSIFICATION.csv"]
    assert open(out / "CLASSIFICATION.csv").read() == expected_csv(len(order))
    # --search-depth prints its two lines; blocks run across file boundaries
    p = run_query(stub, ["-r", str(d), "-o", str(out), "-d", str(tmp_path), "--neg-filter", "--search-depth", "2", "-b", "50"],
                  {"PF_PARSE_BUF_BYTES": "7000"})
    assert p.returncode == 0 and "Search depth settings: 2" in p.stdout.decode()
    _, want_neg = expected([(r, s, b"") for r, s, _ in order], 50, 9, 0, 1, False)
    assert open(out / "NEG_FILTERING.fa", "rb").read() == want_neg
    # empty input: outputs exist and are empty
    e = tmp_path / "empty.fq"
    e.write_bytes(b"")
    p = run_query(stub, ["-r", str(e), "-o", str(out), "-d", str(tmp_path), "--pos-filter", "--neg-filter"])
    assert p.returncode == 0 and sorted(os.listdir(out)) == ["CLASSIFICATION.csv", "NEG_FILTERING.fq", "POS_FILTERING.fq"]
    assert all(os.path.getsize(out / f) == 0 for f in os.listdir(out))
    # errors panic like the reference (exit code 101, message on stderr)
    p = run_query(stub, ["-r", str(e), "-o", str(out), "-d", str(tmp_path / "missing")])
    assert p.returncode == 101 and b"panicked" in p.stderr and b"BloomTree::load" in p.stderr
    p = run_query(stub, ["-r", str(tmp_path / "nope.fq"), "-o", str(out), "-d", str(tmp_path)])
    assert p.returncode == 101 and b"No such file" in p.stderr
    p = run_query(stub, ["-r", str(e), "-o", str(out)])
    assert p.returncode == 101 and b"required arguments" in p.stderr
    bad = tmp_path / "bad.fq"
    bad.write_bytes(b"@r1\nACGT\n+\nIIII\nnot a header\nAC\n+\nII\n")
    p = run_query(stub, ["-r", str(bad), "-o", str(out), "-d", str(tmp_path)])
    assert p.returncode == 101 and b"Expected @" in p.stderr


def test_build_and_add_driver(stub, tmp_path):
    """`build` / `add`: one leaf per record (main.rs:173-195), files popped from the end of the sorted listing, multi-line
    and gzip FASTA, the flags of main.rs:38-99 -- against a log the stub builder writes."""
    rng = np.random.default_rng(50)
    d = tmp_path / "genomes"
    d.mkdir()
    parts = {}
    for k, name in enumerate(["a.fna", "b.fasta.gz", "c.fa", "d.fas"]):
        recs = [(f"{name}_{i}", np.frombuffer(b"ACGTN", dtype=np.uint8)[rng.integers(0, 5, size=int(rng.integers(1, 4000)))].tobytes())
                for i in range(1 + k)]
        blob = b"".join(b">" + r.encode() + b" description\n" + b"".join(s[j:j + 70] + b"\n" for j in range(0, len(s), 70)) for r, s in recs)
        with (gzip.open(d / name, "wb") if name.endswith(".gz") else open(d / name, "wb")) as f:
            f.write(blob)
        parts[name] = recs
    (d / "README.txt").write_text("not a genome")
    order = [x for name in sorted(parts, reverse=True) for x in parts[name]]
    db = tmp_path / "db.log"
    e = dict(os.environ, LD_PRELOAD=stub)
    p = subprocess.run([BIN, "build", "-g", str(d), "-d", str(db), "-k", "21", "-f", "0.01", "-l", "5000", "--seed-one", "11", "--seed-two", "12",
                        "-t", "4", "-c", "7", "--host-threads", "3"], capture_output=True, env=e, timeout=120)
    assert p.returncode == 0, p.stderr.decode()
    assert p.stdout.decode().splitlines() == ["Building the SBT...", "Finished."]
    log = open(db, "rb").read().split(b"\n")
    assert log[0] == b"create k=21 fpr=0.01 largest=5000 seeds=11,12 names=1"
    assert log[1:-1] == [b"insert " + r.encode() + b" " + s for r, s in order]
    p = subprocess.run([BIN, "add", "-g", str(d / "c.fa"), "-d", str(db)], capture_output=True, env=e, timeout=120)
    assert p.returncode == 0 and p.stdout.decode().splitlines() == ["Adding new genomes to the SBT...", "Finished."]
    log = open(db, "rb").read().split(b"\n")
    assert log[0].startswith(b"open ") and log[1:-1] == [b"insert " + r.encode() + b" " + s for r, s in parts["c.fa"]]
    p = subprocess.run([BIN, "add", "-g", str(d), "-d", str(tmp_path / "missing")], capture_output=True, env=e, timeout=120)
    assert p.returncode == 101 and b"BloomTree::load" in p.stderr
    p = subprocess.run([BIN, "build", "-g", str(d)], capture_output=True, env=e, timeout=120)
    assert p.returncode == 101 and b"required arguments" in p.stderr


def test_global_verbosity_flags_and_duplicate_genome_ids(stub, tmp_path):
    """clap-verbosity-flag is global in the reference (`phage_filter -vv query ...`, main.rs:38-44), and ResultMap keeps a
    HashSet<String> of genome ids (result_map.rs:10), so two leaves that carry the same tax_id are listed once."""
    rng = np.random.default_rng(5)
    recs = [(f"r{i}", b"ACGT" * 10, None) for i in range(400)]
    reads = tmp_path / "reads.fa"
    write_records(reads, recs, False)
    out = tmp_path / "out"
    e = dict(os.environ, LD_PRELOAD=stub, PF_STUB_DUP_NAMES="1")
    p = subprocess.run([BIN, "-vv", "query", "-r", str(reads), "-o", str(out), "-d", str(tmp_path), "--pos-filter"],
                       capture_output=True, env=e, timeout=120)
    assert p.returncode == 0, p.stderr.decode()
    heads = [ln for ln in (out / "POS_FILTERING.fa").read_text().splitlines() if ln.startswith(">")]
    assert heads
    both = 0
    for h in heads:
        rid, ids = h[1:].split(" |")
        g = int(rid[1:])
        want = {("genome_5" if l == 1 else f"genome_{l}") for l in {(g * 7 + j * 5) % 9 for j in range(g % 3)}}
        got = ids.split(",")
        assert len(got) == len(set(got)) and set(got) == want, h
        both += {(g * 7 + j * 5) % 9 for j in range(g % 3)} >= {1, 5}
    assert both > 0  # the case really occurred
    q = subprocess.run([BIN, "-q", "--version"], capture_output=True, env=e, timeout=60)
    assert q.returncode == 0 and b"PhageFilter" in q.stdout
