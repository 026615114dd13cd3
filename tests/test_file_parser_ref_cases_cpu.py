"""The reference's own file_parser unit tests (src/file_parser.rs:410-604) restated for the two host-side readers of
this repo: the C++ driver's threaded reader (`phage_filter parse`, host/seq_reader.h) and the Python mirror
(phagefilter_b200/file_parser.py).  Each test names the reference test it follows."""
import gzip
import os
import subprocess

import pytest

from phagefilter_b200.file_parser import (ReadQueue, detect_format, format_from_extension, has_supported_extension,
                                          read_records)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "phagefilter_b200", "bin", "phage_filter")


def cli(path, *extra):
    if not os.path.exists(BIN):
        pytest.skip("phage_filter binary not built")
    p = subprocess.run([BIN, "parse", "-r", str(path), "--host-threads", "3", "--min-segment", "8", *extra], capture_output=True)
    assert p.returncode == 0, p.stderr.decode()
    return [tuple(line.split(b"\t")) for line in p.stdout.split(b"\n") if line]


def test_detect_format_by_content(tmp_path):  # test_detect_format_fasta_content / _fastq_content (:410-431)
    fa, fq = tmp_path / "a" / "test.dat", tmp_path / "q" / "test.dat"
    fa.parent.mkdir(), fq.parent.mkdir()
    fa.write_bytes(b">seq1\nACGTACGT\n")
    fq.write_bytes(b"@read1\nACGTACGT\n+\nIIIIIIII\n")
    assert detect_format(str(fa)) == "fasta" and detect_format(str(fq)) == "fastq"
    assert cli(fa) == [(b"seq1", b"ACGTACGT", b"-")]
    assert cli(fq) == [(b"read1", b"ACGTACGT", b"IIIIIIII")]


def test_detect_format_override(tmp_path):  # :433-448
    p = tmp_path / "test.fasta"
    p.write_bytes(b">seq1\nACGTACGT\n")
    assert detect_format(str(p), "fastq") == "fastq" and detect_format(str(p), "fasta") == "fasta"
    assert cli(p, "-F", "fasta") == [(b"seq1", b"ACGTACGT", b"-")]
    # forcing FASTQ on FASTA content is a parse error in the reference (bio's reader) and here
    bad = subprocess.run([BIN, "parse", "-r", str(p), "-F", "fastq"], capture_output=True)
    assert bad.returncode == 101 and b"Expected @" in bad.stderr


def test_detect_format_gzipped(tmp_path):  # test_detect_format_gzipped_fasta / _fastq (:450-474)
    fa, fq = tmp_path / "test.fa.gz", tmp_path / "test.fq.gz"
    fa.write_bytes(gzip.compress(b">seq1\nACGTACGT\n"))
    fq.write_bytes(gzip.compress(b"@read1\nACGTACGT\n+\nIIIIIIII\n"))
    assert detect_format(str(fa)) == "fasta" and detect_format(str(fq)) == "fastq"
    assert cli(fa) == [(b"seq1", b"ACGTACGT", b"-")]
    assert cli(fq) == [(b"read1", b"ACGTACGT", b"IIIIIIII")]


def test_format_from_extension():  # :476-486
    for name in ("reads.fq", "reads.fastq", "reads.fq.gz"):
        assert format_from_extension(name) == "fastq"
    for name in ("genome.fa", "genome.fasta", "genome.fna", "genome.fa.gz", "genome.fasta.gz"):
        assert format_from_extension(name) == "fasta"


def test_has_supported_extension(tmp_path):  # :488-503
    yes = ["test.fa", "test.fasta", "test.fna", "test.fsa", "test.fas", "test.fq", "test.fastq", "test.fq.gz", "test.fasta.gz",
           "test.fna.gzip"]
    no = ["test.txt", "test.gz", "test"]
    for n in yes:
        assert has_supported_extension(n), n
    for n in no:
        assert not has_supported_extension(n), n
    # the C++ reader applies the same rule when it scans a directory: one record per accepted file
    for n in yes + no:
        blob = b">" + n.encode() + b"\nACGT\n" if "fq" not in n and "fastq" not in n else b"@" + n.encode() + b"\nACGT\n+\nIIII\n"
        (tmp_path / n).write_bytes(gzip.compress(blob) if n.endswith((".gz", ".gzip")) else blob)
    got = sorted(r[0].decode() for r in cli(tmp_path))
    assert got == sorted(yes)


def test_read_queue_gzipped(tmp_path):  # test_read_queue_gzipped_fasta / _fastq (:505-532)
    fa = tmp_path / "genome.fa.gz"
    fa.write_bytes(gzip.compress(b">seq1\nACGTACGTACGTACGTACGTACGTACGT\n"))
    q = ReadQueue(str(fa), 10, 5, False)
    block = q.next_block()
    assert len(block) == 1 and block[0].id == "seq1"
    seq = b"ACGTACGTACGTACGTACGTACGTACGT"
    fq = tmp_path / "reads.fq.gz"
    fq.write_bytes(gzip.compress(b"@read1\n" + seq + b"\n+\n" + b"I" * len(seq) + b"\n"))
    block = ReadQueue(str(fq), 10, 5, False).next_block()
    assert len(block) == 1 and block[0].id == "read1"
    assert cli(fa) == [(b"seq1", seq, b"-")] and cli(fq) == [(b"read1", seq, b"I" * len(seq))]


def test_directory_scan_includes_gz_files(tmp_path):  # :534-554
    (tmp_path / "genome.fa").write_bytes(b">s1\nACGT\n")
    (tmp_path / "reads.fq").write_bytes(b"@r1\nACGT\n+\nIIII\n")
    (tmp_path / "compressed.fasta.gz").write_bytes(gzip.compress(b">s2\nACGT\n"))
    (tmp_path / "notes.txt").write_bytes(b"")
    (tmp_path / "random.gz").write_bytes(b"")
    q = ReadQueue(str(tmp_path), 10, 3, False)
    names = sorted(os.path.basename(f) for f in q.filequeue)
    assert names == ["compressed.fasta.gz", "genome.fa", "reads.fq"]
    assert sorted(r[0] for r in cli(tmp_path)) == [b"r1", b"s1", b"s2"]


def test_fastq_quality_preserved(tmp_path):  # :556-577 -- note: the reference's fixture has 28 bases and 29 quality values
    seq, qual = b"ACGTACGTACGTACGTACGTACGTACGT", b"IIIIIIIIIIIIIIIIIIIIIIIIIIIII"
    assert (len(seq), len(qual)) == (28, 29)
    p = tmp_path / "reads.fq"
    p.write_bytes(b"@read1\n" + seq + b"\n+\n" + qual + b"\n")
    block = ReadQueue(str(p), 10, 5, True).next_block()
    assert len(block) == 1 and block[0].id == "read1" and block[0].quality == qual
    assert cli(p) == [(b"read1", seq, qual)]  # kept as read, one value longer than the sequence


def test_fasta_has_no_quality(tmp_path):  # :579-588
    p = tmp_path / "genome.fa"
    p.write_bytes(b">seq1\nACGTACGTACGTACGTACGTACGTACGT\n")
    block = ReadQueue(str(p), 10, 5, True).next_block()
    assert len(block) == 1 and block[0].quality is None
    assert cli(p)[0][2] == b"-"


def test_peek_format(tmp_path):  # :590-604
    (tmp_path / "genome.fa").write_bytes(b">s1\nACGT\n")
    (tmp_path / "reads.fq").write_bytes(b"@r1\nACGT\n+\nIIII\n")
    assert ReadQueue(str(tmp_path / "genome.fa"), 10, 5, False).peek_format() == "fasta"
    assert ReadQueue(str(tmp_path / "reads.fq"), 10, 5, False).peek_format() == "fastq"
    assert list(read_records(str(tmp_path / "reads.fq")))[0].id == "r1"
