"""CPU test of the host driver's ingest pipeline (`phage_filter parse` runs the same producer thread the `query`
command uses, without the GPU): FASTA/FASTQ, multi-line records, CRLF, blank lines, missing trailing newline, gzip,
directories, tiny buffers that force refills and growth, several parser threads with tiny segments (speculative
FASTQ cut points, including quality lines that look like headers), block carry-over and the packer's view of the
records -- against the simple Python reader."""
import gzip
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "phagefilter_b200", "bin", "phage_filter")


def parse_cli(path, buf=1 << 20, chunk=None, fmt=None, threads=None, min_segment=None, block=None, pack=False, env=None):
    """chunk (the old reader's record cap) is accepted and unused.  Unless told otherwise the run uses 1 or 5
    parser threads and small segments (derived from buf), so every call site also covers the threaded path."""
    if threads is None:
        threads = 1 if buf % 2 else 5
    if min_segment is None:
        min_segment = 64 if buf < (1 << 20) else 1000
    args = [BIN, "parse", "-r", str(path), "--buf-bytes", str(buf), "--host-threads", str(threads),
            "--min-segment", str(min_segment)]
    if block:
        args += ["-b", str(block)]
    if pack:
        args += ["--pack"]
    if fmt:
        args += ["-F", fmt]
    p = subprocess.run(args, capture_output=True, env=dict(os.environ, **(env or {})))
    assert p.returncode == 0, p.stderr.decode()
    out = []
    for line in p.stdout.split(b"\n"):
        if not line:
            continue
        rid, seq, qual = line.split(b"\t")
        out.append((rid.decode(), seq, None if qual == b"-" else qual))
    return out


def want(path, fmt="auto"):
    from phagefilter_b200.file_parser import read_records
    return [(r.id, r.sequence, r.quality) for r in read_records(str(path), fmt)]


@pytest.fixture(scope="module", autouse=True)
def _built():
    if not os.path.exists(BIN):
        pytest.skip("phage_filter binary not built")


def test_fastq_variants(tmp_path):
    rng = np.random.default_rng(0)
    acgt = np.frombuffer(b"ACGTN", dtype=np.uint8)
    recs = []
    for i in range(3000):
        L = int(rng.choice([0, 1, 20, 100, 150, 151, 5000])) if i % 50 == 0 else 100
        s = acgt[rng.integers(0, 5, size=L)].tobytes()
        q = bytes(rng.integers(33, 74, size=L, dtype=np.uint8))  # may contain '@' and '+'
        recs.append((f"read{i}", s, q))
    p = tmp_path / "a.fq"
    with open(p, "wb") as f:
        for i, (rid, s, q) in enumerate(recs):
            desc = b" extra words" if i % 3 == 0 else b""
            f.write(b"@" + rid.encode() + desc + b"\n" + s + b"\n+" + (rid.encode() if i % 7 == 0 else b"") + b"\n" + q + b"\n")
    got = parse_cli(p)
    assert got == [(r, s, q) for r, s, q in recs]
    for buf in (64, 300, 4096, 20000):  # refills, carry-over and buffer growth (a 5000-base record)
        assert parse_cli(p, buf=buf, chunk=7) == got
    # CRLF + blank lines + no trailing newline + gzip
    p2 = tmp_path / "b.fastq.gz"
    with gzip.open(p2, "wb") as f:
        f.write(b"\r\n@x1 d\r\nACGT\r\n+\r\nIIII\r\n\r\n@x2\r\nGG\r\n+\r\n##")
    assert parse_cli(p2) == [("x1", b"ACGT", b"IIII"), ("x2", b"GG", b"##")]
    # multi-line FASTQ
    p3 = tmp_path / "c.fq"
    p3.write_bytes(b"@m1\nACGT\nTTGA\nC\n+\nIIII\nJJJJ\nK\n@m2\nAC\n+\n@@\n")
    assert parse_cli(p3, buf=16) == [("m1", b"ACGTTTGAC", b"IIIIJJJJK"), ("m2", b"AC", b"@@")]


def test_fasta_variants_and_directory(tmp_path):
    rng = np.random.default_rng(1)
    acgt = np.frombuffer(b"ACGTacgtN", dtype=np.uint8)
    d = tmp_path / "reads"
    d.mkdir()
    all_recs = {}
    for name, wrap, crlf in (("a.fa", 60, False), ("b.fasta", 0, True), ("c.fna.gz", 13, False)):
        recs = []
        for i in range(200):
            L = int(rng.choice([0, 5, 59, 60, 61, 500, 3000]))
            recs.append((f"{name}_{i}", acgt[rng.integers(0, 9, size=L)].tobytes()))
        nl = b"\r\n" if crlf else b"\n"
        blob = b""
        for i, (rid, s) in enumerate(recs):
            blob += b">" + rid.encode() + (b"\tdesc > with gt" if i % 5 == 0 else b"") + nl
            if wrap:
                for j in range(0, len(s), wrap):
                    blob += s[j:j + wrap] + nl
            else:
                blob += s + nl
            if i % 17 == 0:
                blob += nl
        if name.endswith(".gz"):
            with gzip.open(d / name, "wb") as f:
                f.write(blob.rstrip(b"\r\n"))  # no trailing newline
        else:
            (d / name).write_bytes(blob)
        all_recs[name] = [(r, s, None) for r, s in recs]
    (d / "notes.txt").write_text("ignored")
    for name in all_recs:
        for buf in (1 << 20, 97, 5000):
            assert parse_cli(d / name, buf=buf, chunk=11) == all_recs[name], (name, buf)
            assert want(d / name) == all_recs[name]
    # directory: files are popped from the END of the sorted listing
    got = parse_cli(d)
    assert got == all_recs["c.fna.gz"] + all_recs["b.fasta"] + all_recs["a.fa"]


def test_golden_reads_match_python_reader():
    p = os.path.join(ROOT, "tests", "golden", "reads.fq")
    assert parse_cli(p, buf=1000) == want(p)
    g = os.path.join(ROOT, "tests", "golden", "genomes.fa")
    assert parse_cli(g, buf=500) == want(g)


def test_threaded_fastq_adversarial_cut_points(tmp_path):
    """Quality lines starting with '@' or '+', sequence-free records, '+' lines repeating the id, a multi-line
    record in the middle (the speculative segments must hand over to the full parser there), blank lines."""
    rng = np.random.default_rng(7)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    recs, blob = [], b""
    for i in range(4000):
        L = int(rng.choice([0, 1, 2, 3, 30, 100]))
        s = acgt[rng.integers(0, 4, size=L)].tobytes()
        q = bytearray(rng.integers(46, 74, size=L, dtype=np.uint8).tobytes())  # not "-": the dump marks "no quality" with it
        if L and i % 3 == 0:
            q[0] = ord("@")
        if L and i % 3 == 1:
            q[0] = ord("+")
        q = bytes(q)
        rid = f"@r{i}" if i % 11 == 0 else f"r{i}"  # ids may start with '@' too
        recs.append((rid, s, q))
        if i == 2000 and L >= 30:
            blob += b"@" + rid.encode() + b"\n" + s[:10] + b"\n" + s[10:] + b"\n+\n" + q[:7] + b"\n" + q[7:] + b"\n"
        else:
            blob += b"@" + rid.encode() + b"\n" + s + b"\n+" + (rid.encode() if i % 5 == 0 else b"") + b"\n" + q + b"\n"
        if i % 97 == 0:
            blob += b"\n"
    p = tmp_path / "adv.fq"
    p.write_bytes(blob)
    serial = parse_cli(p, threads=1)
    assert serial == recs
    for threads, seg, buf in ((2, 1, 1 << 20), (8, 50, 1 << 20), (3, 200, 4096), (16, 1, 700), (5, 10_000, 1 << 22)):
        assert parse_cli(p, threads=threads, min_segment=seg, buf=buf) == recs, (threads, seg, buf)
    # a fully multi-line file: every segment bails, the full parser does the work
    p2 = tmp_path / "ml.fq"
    with open(p2, "wb") as f:
        for rid, s, q in recs[:500]:
            h = len(s) // 2
            f.write(b"@" + rid.encode() + b"\n" + s[:h] + b"\n" + s[h:] + b"\n+\n" + q[:h] + b"\n" + q[h:] + b"\n")
    # (a quality piece starting with '@' is ambiguous in multi-line FASTQ for ANY parser; only compare the two modes)
    assert parse_cli(p2, threads=6, min_segment=1) == parse_cli(p2, threads=1)


def test_blocks_carry_and_packer_view(tmp_path):
    """Chunks handed to the query thread hold whole blocks (also across files and tiny buffers); the 2-bit batches
    built by the ingest thread describe exactly those records (checked inside `parse --pack`)."""
    rng = np.random.default_rng(9)
    acgt = np.frombuffer(b"ACGTN", dtype=np.uint8)
    d = tmp_path / "many"
    d.mkdir()
    expect = {}
    for k in range(12):
        recs = []
        for i in range(int(rng.integers(1, 90))):
            L = int(rng.choice([0, 19, 20, 21, 150]))
            recs.append((f"f{k}_{i}", acgt[rng.integers(0, 5 if i % 9 == 0 else 4, size=L)].tobytes()))
        name = f"s{k:02d}.fa" + (".gz" if k % 4 == 0 else "")
        blob = b"".join(b">" + r.encode() + b"\n" + s + b"\n" for r, s in recs)
        if name.endswith(".gz"):
            with gzip.open(d / name, "wb") as f:
                f.write(blob)
        else:
            (d / name).write_bytes(blob)
        expect[name] = [(r, s, None) for r, s in recs]
    order = [x for name in sorted(expect, reverse=True) for x in expect[name]]
    for block, buf, batch in ((1, 1 << 20, None), (7, 300, None), (64, 5000, None), (1000, 1 << 20, None), (100, 1 << 20, None)):
        got = parse_cli(d, block=block, buf=buf, pack=True)
        assert got == order, (block, buf)


def test_gzip_files_inflated_ahead(tmp_path):
    """A directory of gzip files (plus plain ones in between): the next files are inflated on their own threads while
    the current one is parsed; tiny inflate blocks force the bounded queue to fill and drain many times."""
    rng = np.random.default_rng(21)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    d = tmp_path / "lanes"
    d.mkdir()
    expect = {}
    for k in range(9):
        recs = [(f"l{k}_{i}", acgt[rng.integers(0, 4, size=int(rng.choice([1, 50, 150])))].tobytes(),
                 bytes(rng.integers(46, 74, size=1, dtype=np.uint8)) * 0) for i in range(int(rng.integers(1, 400)))]
        recs = [(r, s, bytes(rng.integers(46, 74, size=len(s), dtype=np.uint8))) for r, s, _ in recs]
        blob = b"".join(b"@" + r.encode() + b"\n" + s + b"\n+\n" + q + b"\n" for r, s, q in recs)
        name = f"lane{k}.fastq" + ("" if k in (3, 4) else ".gz")
        if name.endswith(".gz"):
            with gzip.open(d / name, "wb") as f:
                f.write(blob)
        else:
            (d / name).write_bytes(blob)
        expect[name] = recs
    (d / "empty.fq.gz").write_bytes(gzip.compress(b""))
    order = [x for name in sorted(expect, reverse=True) for x in expect[name]]
    for threads, blk in ((1, "100000"), (4, "512"), (12, "97"), (3, "16777216")):
        got = parse_cli(d, threads=threads, buf=3000 if threads != 3 else 1 << 20, block=50, pack=True, env={"PF_GZ_BLOCK": blk})
        assert got == order, (threads, blk)
