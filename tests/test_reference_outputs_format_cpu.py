"""Real outputs of the reference binary kept in its repository (benchmarking/results/RasPiData/*/POS_FILTERING.fa,
written by main.rs:345-361,394-404 with --search-depth, hence the Internal_Node_<u16> genome names) against the
formatting rule the writer tests of this repo use: `>{id} |{g1,g2,...}\\n{SEQUENCE}\\n`, sequence upper-case on one
line.  tests/test_host_outputs_cpu.py holds the product's writer byte-exactly to that rule, so this closes the chain
real reference output == rule == our writer.  The files cannot travel (they live in /root/reference): skipped where
the reference is absent."""
import glob
import os
import re

import pytest

from tests.util import parse_filter_file

FILES = sorted(glob.glob("/root/reference/benchmarking/results/RasPiData/*/POS_FILTERING.fa"))


@pytest.mark.skipif(not FILES, reason="reference results not present")
@pytest.mark.parametrize("path", FILES)
def test_real_pos_filtering_files_follow_the_writer_rule(path):
    raw = open(path, "rb").read()
    recs = parse_filter_file(path)
    assert len(recs) > 100
    head = re.compile(rb"^>(\S+) \|([^,\s]+(?:,[^,\s]+)*)$")
    lines = raw.split(b"\n")
    assert lines[-1] == b"" and len(lines) == 2 * len(recs) + 1  # header + one sequence line per record, final newline
    rebuilt = []
    for i, (rid, genomes, seq, qual) in enumerate(recs):
        m = head.match(lines[2 * i])
        assert m, lines[2 * i][:80]
        order = m.group(2).decode().split(",")  # the reference's order is that of a HashSet: keep it as found
        assert rid == m.group(1).decode() and set(order) == set(genomes) and len(order) == len(set(order))
        assert qual is None and seq == seq.upper() and re.fullmatch(rb"[A-Z]*", seq)
        rebuilt.append(b">" + rid.encode() + b" |" + ",".join(order).encode() + b"\n" + seq + b"\n")
    assert b"".join(rebuilt) == raw
