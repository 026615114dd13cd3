"""CPU tests of the product's host logic: the C-ABI library loads and exports every symbol
include/pfgpu.h declares, the packer produces the documented 2-bit layout, geometry matches the oracle,
ResultMap strings match the reference's tests.  No compute entry point is called without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from phagefilter_b200 import _lib
    L = _lib.lib()
    hdr = open(os.path.join(ROOT, "include", "pfgpu.h")).read()
    declared = set(re.findall(r"\b(pf_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations found"
    for name in sorted(declared):
        assert hasattr(L, name), f"libpfgpu.so does not export {name}"
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    assert b"sm_100a" in L.pf_version()


def test_no_cpu_fallback_without_gpu(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from phagefilter_b200 import BloomTree, _lib
    with pytest.raises(_lib.PfError) as e:
        BloomTree.load(str(tmp_path))
    assert e.value.status == 4  # PF_ERR_CUDA: fails loudly, never computes on the CPU


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "phagefilter_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "pf_oracle" not in src and "oracle/" not in src, f


def test_packer_layout():
    from phagefilter_b200.query import PackedReads
    reads = [b"ACGTACGTACGTACGTAC", b"ACGN", b"", b"T" * 33, b"acgt"]
    p = PackedReads(reads)
    b = p.batch.contents
    assert b.n_reads == 5 and b.n_exc == 2
    assert [b.lengths[i] for i in range(5)] == [18, 4, 0, 33, 4]
    offs = [b.word_off[i] for i in range(5)]
    assert all(o % 2 == 0 for o in offs) and offs == sorted(offs)
    for r, s in enumerate(reads):
        if any(c not in b"ACGT" for c in s):
            e = b.exc_index[r]
            assert e != 0xFFFFFFFF
            got = bytes(b.exc_bytes[b.exc_off[e] + j] for j in range(len(s)))
            assert got == s
        else:
            assert b.exc_index[r] == 0xFFFFFFFF
            for j, c in enumerate(s):
                w = b.packed[offs[r] + j // 16]
                assert (w >> (2 * (j % 16))) & 3 == b"ACGT".index(c)
    assert b.n_words >= offs[-1] + 2 + 4  # trailing pad words
    p.close()


def test_geometry_matches_oracle(oracle):
    from phagefilter_b200 import _lib
    L = _lib.lib()
    for fpr, n in [(0.001, 1_000_000), (0.001, 1000), (1e-5, 500_000), (0.01, 77), (0.001, 450_000), (0.5, 1)]:
        bits = L.pf_needed_bits(C.c_float(fpr), n)
        assert bits == oracle.needed_bits(fpr, n)
        assert L.pf_optimal_num_hashes(bits, n) == oracle.optimal_num_hashes(bits, n)


def test_result_map_reference_strings():
    """result_map.rs:53-123"""
    from phagefilter_b200 import ResultMap
    rm = ResultMap()
    assert not rm.read_mapped("any")
    assert rm.get_ext_id("unknown") == "unknown |"
    rm.add_read_map("read1", "genomeA")
    assert rm.read_mapped("read1") and rm.get_ext_id("read1") == "read1 |genomeA"
    rm.add_read_map("read1", "genomeB")
    rm.add_read_map("read1", "genomeA")
    head, genomes = rm.get_ext_id("read1").split(" |")
    assert head == "read1" and sorted(genomes.split(",")) == ["genomeA", "genomeB"]
    rm.empty_read_map()
    assert not rm.read_mapped("read1")


def test_read_queue_and_formats(tmp_path):
    """file_parser.rs:410-604: format sniffing, gzip, ids up to the first whitespace, block sizes."""
    import gzip
    from phagefilter_b200.file_parser import ReadQueue, detect_format, has_supported_extension
    fq = tmp_path / "a.fq"
    fq.write_bytes(b"@r1 desc\nACGT\n+\n####\n@r2\nGGCC\n+\n!!!!\n")
    fa = tmp_path / "b.fasta.gz"
    with gzip.open(fa, "wb") as f:
        f.write(b">s1 something\nAC\nGT\n>s2\nTTTT\n")
    (tmp_path / "notes.txt").write_text("x")
    assert detect_format(str(fq)) == "fastq" and detect_format(str(fa)) == "fasta"
    assert has_supported_extension(str(fa)) and not has_supported_extension(str(tmp_path / "notes.txt"))
    q = ReadQueue(str(tmp_path), 3, 3, True)
    blocks = []
    while True:
        b = q.next_block()
        if not b:
            break
        blocks.append(b)
    recs = [r for b in blocks for r in b]
    assert sorted(r.id for r in recs) == ["r1", "r2", "s1", "s2"]
    by = {r.id: r for r in recs}
    assert by["s1"].sequence == b"ACGT" and by["s1"].quality is None
    assert by["r2"].sequence == b"GGCC" and by["r2"].quality == b"!!!!"
    assert [len(b) for b in blocks] == [3, 1]
    assert by["r1"].num_kmers(3) == 2 and by["r1"].num_kmers(5) == 0 and by["r1"].num_kmers(0) == 0


def test_packer_avx2_matches_table_loop(monkeypatch):
    """The AVX2 body of the packer (pf_pack_simd.cpp) against the table-driven loop (PF_PACK_SCALAR=1): every length
    around the 16/32-base steps, invalid bytes (lower case, N, bytes that differ from A/C/G/T in one bit, >= 0x80) at
    every position class."""
    import ctypes as C
    from phagefilter_b200.query import PackedReads
    rng = np.random.default_rng(11)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    reads = []
    for L in list(range(0, 100)) + [127, 128, 129, 150, 151, 1000, 4097]:
        reads.append(acgt[rng.integers(0, 4, size=L)].tobytes())
    for bad in (b"N", b"a", b"@", b"E", b"U", b"\xc1", b"\x00", b"\n"):
        for L in (1, 31, 32, 33, 64, 150):
            for pos in {0, L // 2, L - 1}:
                r = bytearray(acgt[rng.integers(0, 4, size=L)].tobytes())
                r[pos:pos + 1] = bad
                reads.append(bytes(r))

    def snapshot():
        p = PackedReads(reads)
        b = p.batch.contents
        words = np.ctypeslib.as_array(b.packed, shape=(b.n_words,)).copy()
        offs = np.ctypeslib.as_array(b.word_off, shape=(b.n_reads,)).copy()
        exc = None if not b.n_exc else np.ctypeslib.as_array(b.exc_index, shape=(b.n_reads,)).copy()
        n_exc = b.n_exc
        p.close()
        return words, offs, exc, n_exc

    monkeypatch.delenv("PF_PACK_SCALAR", raising=False)
    fast = snapshot()
    monkeypatch.setenv("PF_PACK_SCALAR", "1")
    slow = snapshot()
    assert fast[3] == slow[3] == sum(1 for r in reads if any(c not in b"ACGT" for c in r))  # every poisoned read is an exception read
    assert (fast[0] == slow[0]).all() and (fast[1] == slow[1]).all() and (fast[2] == slow[2]).all()


def test_packer_threads_give_identical_batches(monkeypatch):
    """Ranges of reads packed by several threads (per-thread size sums + prefix) against the one-thread pass,
    with exception reads spread over the ranges and ragged lengths."""
    from phagefilter_b200.query import PackedReads
    rng = np.random.default_rng(12)
    alpha = np.frombuffer(b"ACGTN", dtype=np.uint8)
    lens = rng.choice([0, 1, 15, 16, 17, 100, 150, 151, 700], size=30_000)
    lens[0] = 150  # the thread heuristic looks at the first read
    reads = [alpha[rng.integers(0, 5 if i % 97 == 0 else 4, size=int(L))].tobytes() for i, L in enumerate(lens)]

    def snapshot(threads):
        monkeypatch.setenv("PF_PACK_THREADS", str(threads))
        p = PackedReads(reads)
        b = p.batch.contents
        out = (np.ctypeslib.as_array(b.packed, shape=(b.n_words,)).copy(),
               np.ctypeslib.as_array(b.word_off, shape=(b.n_reads,)).copy(),
               np.ctypeslib.as_array(b.lengths, shape=(b.n_reads,)).copy(),
               np.ctypeslib.as_array(b.exc_index, shape=(b.n_reads,)).copy(),
               np.ctypeslib.as_array(b.exc_off, shape=(b.n_exc + 1,)).copy(),
               bytes(bytearray(b.exc_bytes[i] for i in range(int(b.exc_off[b.n_exc])))),
               (b.n_exc, b.max_length, b.total_bases, b.n_words))
        p.close()
        return out

    one = snapshot(1)
    assert one[6][0] == sum(1 for r in reads if b"N" in r) and one[6][1] == 700 and one[6][2] == int(lens.sum())
    for threads in (2, 7, 32):
        got = snapshot(threads)
        assert got[6] == one[6] and got[5] == one[5]
        for a, b in zip(got[:5], one[:5]):
            assert (a == b).all()
