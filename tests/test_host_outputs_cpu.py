"""CPU test of the host driver's POS/NEG writer (phagefilter_b200/host/outputs.h, several threads) against a plain
restatement of the reference's per-block ResultMap logic (src/main.rs:345-364, result_map.rs:20-41): a read is
mapped when any record of ITS BLOCK with the same id has a hit, the header lists the union of their genomes,
sequences are upper-cased, FASTQ keeps its quality.  Hits are fabricated by a fixed rule (no GPU)."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "host", "filter_writer_harness.cpp")


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("harness") / "filter_writer_harness")
    subprocess.run(["g++", "-O1", "-std=c++17", "-pthread", "-o", exe, SRC, "-lz"], check=True)
    return exe


def expected(records, block, n_leaves, pos, neg, fastq):
    hits = []
    for g in range(len(records)):
        cnt = 0 if g % 4 == 0 else g % 3
        hits.append(sorted({(g * 7 + j * 5) % n_leaves for j in range(cnt)}))
    pos_out, neg_out = [], []
    for lo in range(0, len(records), block):
        blk = range(lo, min(len(records), lo + block))
        result_map = {}
        for i in blk:
            for l in hits[i]:
                result_map.setdefault(records[i][0], set()).add(l)
        for i in blk:
            rid, seq, qual = records[i]
            body = seq.upper() + (b"\n+\n" + qual if fastq else b"") + b"\n"
            mark = b"@" if fastq else b">"
            if rid in result_map:
                if pos:
                    names = ",".join(f"genome_{l}" for l in sorted(result_map[rid]))
                    pos_out.append(mark + rid.encode() + b" |" + names.encode() + b"\n" + body)
            elif neg:
                neg_out.append(mark + rid.encode() + b"\n" + body)
    return b"".join(pos_out), b"".join(neg_out)


@pytest.mark.parametrize("fastq", [True, False])
def test_filter_writer_matches_reference_logic(harness, tmp_path, fastq):
    rng = np.random.default_rng(5 + fastq)
    alpha = np.frombuffer(b"ACGTacgtNn", dtype=np.uint8)
    records = []
    for i in range(2500):
        L = int(rng.choice([0, 1, 30, 100]))
        rid = f"dup{i % 40}" if i % 6 == 0 else f"r{i}"  # duplicated ids, inside and across blocks
        records.append((rid, alpha[rng.integers(0, 10, size=L)].tobytes(), bytes(rng.integers(46, 74, size=L, dtype=np.uint8))))
    path = tmp_path / ("r.fq" if fastq else "r.fa")
    with open(path, "wb") as f:
        for rid, s, q in records:
            if fastq:
                f.write(b"@" + rid.encode() + b" desc\n" + s + b"\n+\n" + q + b"\n")
            else:
                f.write(b">" + rid.encode() + b" desc\n" + s + b"\n")
    ext = "fq" if fastq else "fa"
    for block, threads, pos, neg, batch_blocks in ((100, 4, 1, 1, 3), (7, 8, 1, 1, 50), (1, 3, 1, 0, 1000), (1000, 1, 0, 1, 1),
                                                   (64, 16, 1, 1, 1), (5000, 2, 1, 1, 1)):
        out = tmp_path / f"out_{block}_{threads}_{pos}{neg}"
        if threads % 2 == 0:  # an existing directory is emptied (main.rs:380-391); a missing one is created
            (out / "sub").mkdir(parents=True)
            (out / "sub" / "stale.txt").write_text("x")
            (out / "POS_FILTERING.fa").write_text("stale")
        subprocess.run([harness, str(path), str(out), str(block), str(threads), str(pos), str(neg), "9", str(block * batch_blocks)], check=True)
        want_pos, want_neg = expected(records, block, 9, pos, neg, fastq)
        assert sorted(os.listdir(out)) == sorted((["NEG_FILTERING." + ext] if neg else []) + (["POS_FILTERING." + ext] if pos else []))
        if pos:
            assert open(out / f"POS_FILTERING.{ext}", "rb").read() == want_pos, (block, threads)
        if neg:
            assert open(out / f"NEG_FILTERING.{ext}", "rb").read() == want_neg, (block, threads)
