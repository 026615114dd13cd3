#!/usr/bin/env python
"""Generates tests/golden/ref_adapter_vectors.json by IMPORTING the reference's own benchmarking code (Python, runnable in
this container): the command lines its PhageFilter adapter issues (benchmarking/bench/tools/phage_filter.py:68-118) and
what its output parser (:27-66) and metric helpers (benchmarking/bench/utils.py:229-337) return for fixed small inputs.
The GPU box has no /root/reference; the committed JSON is what travels.

    python tests/golden/make_ref_adapter_vectors.py
"""
import json
import os
import random
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/benchmarking"
sys.path.insert(0, REF)
from bench import utils as U  # noqa: E402
from bench.tools.phage_filter import PhageFilter  # noqa: E402


def main():
    out = {"source": "Dreycey/PhageFilter benchmarking/bench/tools/phage_filter.py + bench/utils.py, imported", "commands": [], "parse": [], "metrics": []}
    for k, theta, threads in ((20, 0.3, 1), (25, 1.0, 4)):
        pf = PhageFilter(k, theta, threads=threads)
        build = pf.build("{DB}", "{GENOMES}")
        for filt, depth in ((False, None), (True, None), (True, 3)):
            run = pf.run("{READS}", "{OUT}", filter_reads=filt, depth=depth)
            out["commands"].append({"kmer": k, "theta": theta, "threads": threads, "filter_reads": filt, "depth": depth,
                                    "build": build, "run": run})
    rng = random.Random(7)
    genomes = [f"NC_{rng.randrange(1000, 999999):06d}.{rng.randrange(1, 4)}" for _ in range(12)]
    with tempfile.TemporaryDirectory() as d:
        for case in range(4):
            counts = {g: rng.choice([0, 1, 3, 40, 900, 5000]) for g in rng.sample(genomes, 8)}
            csv = "".join(f"{g},{c}\n" for g, c in counts.items() if c > 0)
            fa = "".join(f">{g}_{i} |{','.join(rng.sample(genomes, rng.randrange(1, 3)))}\nACGT\n" for g in counts for i in range(counts[g] % 7))
            os.makedirs(os.path.join(d, str(case)))
            open(os.path.join(d, str(case), "CLASSIFICATION.csv"), "w").write(csv)
            open(os.path.join(d, str(case), "POS_FILTERING.fa"), "w").write(fa)
            pf = PhageFilter(20, 0.3)
            out["parse"].append({"classification_csv": csv, "pos_filtering_fa": fa,
                                 "classification": pf.parse_output(os.path.join(d, str(case))),
                                 "filter": dict(pf.parse_output(os.path.join(d, str(case)), filter_reads=True))})
    for case in range(6):
        truth = {g: rng.randrange(1, 6000) for g in rng.sample(genomes, 6)}
        got = {g: max(0, c + rng.randrange(-600, 600)) for g, c in truth.items() if rng.random() < 0.8}
        got.update({f"other_{i}": rng.randrange(1, 4000) for i in range(rng.randrange(0, 3))})
        out["metrics"].append({"true": truth, "out": got,
                               "classification": list(U.get_classification_metrics(truth, got)),
                               "filter": list(U.get_filter_metrics(truth, got)),
                               "readcount": list(U.get_readcount_metrics(truth, got))})
    with open(os.path.join(HERE, "ref_adapter_vectors.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", len(out["commands"]), "command sets,", len(out["parse"]), "parser cases,", len(out["metrics"]), "metric cases")


if __name__ == "__main__":
    main()
