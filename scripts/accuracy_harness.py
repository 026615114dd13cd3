#!/usr/bin/env python
"""Accuracy table of the `phage_filter` host driver in the layout of the reference's benchmarking harness.

The reference's harness (benchmarking/bench.py, subcommand `parameterization` -> benchtest_parameter_sweep,
benchmarking/bench/benchmarking_tests.py:387-503) builds a DB, queries simulated reads with `--pos-filter`, and
writes one CSV row per (k-mer size, theta, error rate) with classification and filter recall/precision.  This
script drives the B200 driver the same way (the adapter's command lines, bench/tools/phage_filter.py:68-118) on the
reference's own example data (BASELINE config #1: 107 viral genomes; 9 read files named sim_reads_c<count>_n<genomes>_e<error>.fq)
and writes the same columns.  Needs a GPU (the driver has no CPU path).

    python scripts/accuracy_harness.py [--data tests/_cfg1_data] [--out gpurun_out/accuracy_cfg1.csv]
                                       [--thetas 0.3,0.5,0.8,1.0] [--kmers 20] [--check-oracle]
"""
import argparse
import os
import re
import resource
import shutil
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from phagefilter_b200 import accuracy as A  # noqa: E402

BIN = os.path.join(ROOT, "phagefilter_b200", "bin", "phage_filter")
# the command lines of the reference's own adapter (benchmarking/bench/tools/phage_filter.py:68-118), captured by importing
# it (tests/golden/make_ref_adapter_vectors.py); only the binary path and the four placeholders are substituted
import json  # noqa: E402
ADAPTER = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_adapter_vectors.json")))["commands"]


def adapter_cmd(which, k, theta, sub, filter_reads=True, depth=None):
    tpl = next(c for c in ADAPTER if c["filter_reads"] == filter_reads and c["depth"] == depth)[which][0]
    out, skip = [], False
    for i, a in enumerate(tpl):
        if skip:
            skip = False
            continue
        if a == "./target/release/phage_filter":
            out.append(BIN)
        elif a == "--kmer-size":
            out += [a, str(k)]
            skip = True
        elif a == "--filter-threshold":
            out += [a, str(theta)]
            skip = True
        else:
            out.append(sub.get(a, a))
    return out
COLUMNS = ["replicate", "kmer size", "theta", "error rate", "number of genomes", "read count", "time", "memory",
           "classification recall", "classification precision", "filter recall", "filter precision", "avg read count error"]


def run(cmd):
    """(elapsed seconds, max RSS of the child in bytes) like utils.run_command's BenchmarkResult."""
    before = resource.getrusage(resource.RUSAGE_CHILDREN).ru_maxrss
    t0 = time.perf_counter()
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL)
    dt = time.perf_counter() - t0
    return dt, max(resource.getrusage(resource.RUSAGE_CHILDREN).ru_maxrss, before) * 1024


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--data", default=None)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "accuracy_cfg1.csv"))
    ap.add_argument("--thetas", default="0.3,0.5,0.8,1.0")
    ap.add_argument("--kmers", default="20")
    ap.add_argument("--check-oracle", action="store_true", help="also compare one CLASSIFICATION.csv with the CPU oracle's")
    args = ap.parse_args()
    data = args.data
    if data is None:
        for cand in (os.path.join(ROOT, "tests", "_cfg1_data"), "/root/reference/examples"):
            if os.path.isdir(cand):
                data = cand
                break
    genomes = os.path.join(data, "viral_genome_dir")
    if not os.path.isdir(genomes):
        genomes = os.path.join(data, "genomes", "viral_genome_dir")
    reads_dir = os.path.join(data, "test_reads")
    read_files = sorted(f for f in os.listdir(reads_dir) if f.endswith(".fq"))
    n_genomes = len(os.listdir(genomes))
    work = tempfile.mkdtemp(prefix="pf_acc_")
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    rows = []
    try:
        for k in [int(x) for x in args.kmers.split(",")]:
            db = os.path.join(work, f"db_k{k}")
            # the adapter's build line (phage_filter.py:79-88) + fixed seeds so that the table is reproducible
            t_build, _ = run(adapter_cmd("build", k, 0, {"{GENOMES}": genomes, "{DB}": db}) +
                             ["--seed-one", str(0x5EED0001), "--seed-two", str(0x5EED0002)])
            print(f"build k={k}: {n_genomes} genome files in {t_build:.2f} s")
            for rf in read_files:
                m = re.match(r"sim_reads_c(\d+)_n(\d+)_e([0-9.]+)\.fq", rf)
                read_count, err = int(m.group(1)), float(m.group(3))
                reads = os.path.join(reads_dir, rf)
                truth = A.get_true_maps(reads)
                for theta in [float(x) for x in args.thetas.split(",")]:
                    out = os.path.join(work, "out")
                    # the adapter's query line (phage_filter.py:104-117)
                    dt, rss = run(adapter_cmd("run", k, theta, {"{READS}": reads, "{DB}": db, "{OUT}": out}))
                    cls = A.parse_classification(os.path.join(out, "CLASSIFICATION.csv"))
                    c_rec, c_prec = A.get_classification_metrics(truth, cls)
                    diffs = A.get_readcount_metrics(truth, cls)
                    flt = A.parse_pos_filtering(os.path.join(out, "POS_FILTERING.fq"))
                    f_rec, f_prec = A.get_filter_metrics(truth, flt)
                    rows.append([1, k, theta, err, n_genomes, read_count, f"{dt:.3f}", rss, c_rec, c_prec, f_rec, f_prec,
                                 sum(diffs) / len(diffs) if diffs else float("nan")])
                    print(f"{rf} theta={theta}: classification R={c_rec:.3f} P={c_prec:.3f}  filter R={f_rec:.3f} P={f_prec:.3f}  {dt:.2f} s")
                    if args.check_oracle and theta == 1.0 and rf == read_files[0]:
                        from oracle import pf_oracle  # checker only
                        from phagefilter_b200.file_parser import read_records
                        t = pf_oracle.Tree.load(db)
                        recs = list(read_records(reads))
                        for lo in range(0, len(recs), 1000):
                            t.query_batch([r.sequence for r in recs[lo:lo + 1000]], theta, want_hits=False)
                        same = t.classification_csv() == open(os.path.join(out, "CLASSIFICATION.csv")).read()
                        print("oracle CLASSIFICATION.csv identical:", same)
                        assert same
        with open(args.out, "w") as f:
            f.write(",".join(COLUMNS) + "\n")
            for r in rows:
                f.write(",".join(str(x) for x in r) + "\n")
        print("wrote", args.out)
    finally:
        shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main()
