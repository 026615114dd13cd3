"""Read records and the block queue (reference: src/file_parser.rs).

The reference materialises every canonical k-mer as a byte vector (get_kmers, :135-148); here a
DNASequence only carries the raw bytes -- k-mers are re-created on the GPU from 2-bit codes.
Parsing stays on the host (BASELINE.json north_star); this module is the minimal reader the tests
and the query driver need: FASTA / FASTQ, optionally gzip, id = header up to the first whitespace.
"""
from __future__ import annotations

import gzip
import os
from dataclasses import dataclass
from typing import Iterator, List, Optional

SEQ_EXTENSIONS = ("fa", "fasta", "fna", "fsa", "fas", "fq", "fastq")  # file_parser.rs:303
COMPRESSED_EXTENSIONS = ("gz", "gzip")


@dataclass
class DNASequence:
    """file_parser.rs:151-157 (without the materialised `kmers`)."""
    sequence: Optional[bytes]
    quality: Optional[bytes]
    id: str

    def num_kmers(self, kmer_size: int) -> int:
        """len(get_kmers(sequence, k)) -- file_parser.rs:136-139"""
        n = len(self.sequence or b"")
        return 0 if kmer_size == 0 or kmer_size > n else n - kmer_size + 1


def _open(path: str):
    with open(path, "rb") as f:
        magic = f.read(2)
    return gzip.open(path, "rb") if magic == b"\x1f\x8b" else open(path, "rb")  # open_reader :89-101


def detect_format(path: str, override: str = "auto") -> str:
    """file_parser.rs:33-66"""
    if override in ("fasta", "fastq"):
        return override
    try:
        with _open(path) as f:
            b = f.read(1)
        if b == b">":
            return "fasta"
        if b == b"@":
            return "fastq"
    except OSError:
        pass
    return format_from_extension(path)


def format_from_extension(path: str) -> str:
    """file_parser.rs:69-86: the extension inside .gz/.gzip decides; anything that is not fq/fastq is FASTA."""
    parts = os.path.basename(path).split(".")
    ext = parts[-1] if len(parts) > 1 else ""
    if ext.lower() in COMPRESSED_EXTENSIONS and len(parts) > 2:
        ext = parts[-2]
    return "fastq" if ext in ("fq", "fastq") else "fasta"


def has_supported_extension(path: str) -> bool:
    """file_parser.rs:323-344"""
    parts = os.path.basename(path).split(".")
    if len(parts) < 2:
        return False
    if parts[-1] in SEQ_EXTENSIONS:
        return True
    return parts[-1] in COMPRESSED_EXTENSIONS and len(parts) > 2 and parts[-2] in SEQ_EXTENSIONS


def read_records(path: str, fmt: str = "auto") -> Iterator[DNASequence]:
    fmt = detect_format(path, fmt)
    with _open(path) as f:
        if fmt == "fastq":
            while True:
                h = f.readline()
                if not h:
                    return
                if not h.strip():
                    continue
                seq = f.readline().rstrip(b"\r\n")
                f.readline()
                qual = f.readline().rstrip(b"\r\n")
                yield DNASequence(seq, qual, h[1:].split()[0].decode() if h[1:].split() else "")
        else:
            rid, chunks = None, []
            for line in f:
                if line.startswith(b">"):
                    if rid is not None:
                        yield DNASequence(b"".join(chunks), None, rid)
                    toks = line[1:].split()
                    rid, chunks = (toks[0].decode() if toks else ""), []
                else:
                    chunks.append(line.strip())
            if rid is not None:
                yield DNASequence(b"".join(chunks), None, rid)


class ReadQueue:
    """file_parser.rs:227-301: files are popped from the END of the directory listing (:238)."""

    def __init__(self, file_path: str, block_size: int, kmer_size: int, filtering: bool, fmt: str = "auto"):
        if os.path.isfile(file_path):
            self.filequeue: List[str] = [file_path]
        else:
            self.filequeue = [os.path.join(file_path, n) for n in os.listdir(file_path)
                              if has_supported_extension(os.path.join(file_path, n))]
        self.block_size, self.kmer_size, self.filtering, self.fmt = block_size, kmer_size, filtering, fmt
        self._cur: Optional[Iterator[DNASequence]] = None

    def peek_format(self) -> str:
        return detect_format(self.filequeue[-1], self.fmt) if self.filequeue else "fasta"

    def next_block(self) -> List[DNASequence]:
        block: List[DNASequence] = []
        while len(block) < self.block_size:
            if self._cur is None:
                if not self.filequeue:
                    break
                self._cur = read_records(self.filequeue.pop(), self.fmt)
            rec = next(self._cur, None)
            if rec is None:
                self._cur = None
                continue
            block.append(rec)
        return block
