"""Regression for the round-1 oracle race: the complement table was filled lazily (identity first, pairs after)
and, in a process that only LOADS a database, its first use happened inside the OpenMP region of the query
(oracle/pf_oracle.c, query_impl) - a late thread re-running the identity loop made other threads canonicalise
with 'A'->'A' and drop hits.  Here: a database built in THIS process, then many fresh subprocesses that load it
and whose FIRST call is a 16-thread query; every one of them must equal the single-thread answer."""
import json
import os
import subprocess
import sys

import numpy as np

from tests.util import oracle_build_db, random_genomes, sample_reads

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_CHILD = r"""
import json, sys
sys.path.insert(0, sys.argv[1])
from oracle import pf_oracle
t = pf_oracle.Tree.load(sys.argv[2])
reads = [bytes.fromhex(x) for x in json.load(open(sys.argv[3]))]
res = t.query_batch(reads, 1.0, want_hits=True, threads=int(sys.argv[4]))
print(json.dumps(sorted((int(r), int(l)) for r, l in res.hits)))
"""


def test_first_query_in_fresh_process_is_thread_count_independent(oracle, tmp_path):
    rng = np.random.default_rng(77)
    genomes = random_genomes(rng, 12, 3000, 5000)
    db = str(tmp_path / "db")
    oracle_build_db(oracle, genomes, 20, db, largest=6000)
    reads = sample_reads(rng, genomes, 1500, 100, 0.0)
    rfile = str(tmp_path / "reads.json")
    with open(rfile, "w") as f:
        json.dump([r.hex() for r in reads], f)

    def run(threads):
        out = subprocess.run([sys.executable, "-c", _CHILD, ROOT, db, rfile, str(threads)], check=True,
                             capture_output=True, text=True, env={**os.environ, "OMP_DYNAMIC": "FALSE"})
        return json.loads(out.stdout.strip().splitlines()[-1])

    want = run(1)
    assert len(want) >= len(reads)  # error-free reads hit at least their own genome
    for _ in range(50):
        assert run(16) == want
