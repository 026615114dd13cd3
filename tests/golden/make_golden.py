"""Generates the golden fixture under tests/golden/ with the CPU oracle (run from the repo root:
`python tests/golden/make_golden.py`).  The reference binary cannot be built in this image (no Rust
toolchain), so these vectors are ORACLE-generated: they pin the oracle and the GPU path against
regressions, not against the Rust build ("parity unpinned" at the rustc-hash boundary)."""
import json
import os
import shutil
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pf_oracle  # noqa: E402
from phagefilter_b200.synth import make_genomes, simulate_reads  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
K, FPR, LARGEST = 20, 0.01, 3000
THETAS = (1.0, 0.8, 0.3)


def main():
    genomes = make_genomes(3, 3, seed=11, len_lo=1800, len_hi=2600, divergence=0.03)
    g0 = bytearray(genomes[0][1])
    g0[200:206] = b"NNNNNN"
    g0[400:430] = bytes(g0[400:430]).lower()
    genomes[0] = (genomes[0][0], bytes(g0))
    db = os.path.join(HERE, "db")
    shutil.rmtree(db, ignore_errors=True)
    tree = pf_oracle.Tree(K, FPR, LARGEST, 0x5EED0001, 0x5EED0002)
    for gid, seq in genomes:
        tree.insert(gid, seq)
    tree.save(db)
    reads, _ = simulate_reads(genomes, 240, 100, seed=12, error_rates=(0.0, 0.02, 0.0, 0.1), background_frac=0.2)
    read_list = [r.tobytes() for r in reads]
    read_list += [b"", b"ACGT", genomes[0][1][190:290], genomes[0][1][380:480], genomes[0][1][380:480].upper(),
                  genomes[1][1][:20], genomes[2][1][100:700], b"N" * 60]
    with open(os.path.join(HERE, "reads.fq"), "wb") as f:
        for i, r in enumerate(read_list):
            f.write(b"@read_%d some description\n%s\n+\n%s\n" % (i, r, b"#" * len(r)))
    with open(os.path.join(HERE, "genomes.fa"), "wb") as f:
        for gid, seq in genomes:
            f.write(b">%s synthetic\n" % gid.encode())
            for i in range(0, len(seq), 70):
                f.write(seq[i:i + 70] + b"\n")
    expected = {"k": K, "fpr": FPR, "largest": LARGEST, "num_bits": tree.num_bits, "num_hashes": tree.num_hashes,
                "leaf_ids": tree.leaf_ids(), "preorder": [[n, int(l), d] for n, l, d in tree.preorder()], "cases": {}}
    for th in THETAS:
        tree.reset_counts()
        res = tree.query_batch(read_list, th)
        expected["cases"][str(th)] = {
            "hits": [sorted(s) for s in res.hit_sets(len(read_list))],
            "csv": tree.classification_csv(), "pairs": res.pairs, "probes_ref": res.probes_ref,
        }
    for depth in (0, 2):
        t2 = pf_oracle.Tree.load(db)
        t2.prune_tree(depth)
        res = t2.query_batch(read_list, 0.8)
        expected["cases"][f"depth{depth}"] = {"leaf_ids": t2.leaf_ids(), "hits": [sorted(s) for s in res.hit_sets(len(read_list))],
                                              "csv": t2.classification_csv()}
    with open(os.path.join(HERE, "expected.json"), "w") as f:
        json.dump(expected, f, indent=0)
    print("golden fixture written:", sorted(os.listdir(HERE)), sum(os.path.getsize(os.path.join(db, x)) for x in os.listdir(db)), "bytes of DB")


if __name__ == "__main__":
    main()
