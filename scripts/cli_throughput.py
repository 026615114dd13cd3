#!/usr/bin/env python
"""Times the host driver end to end on a synthetic FASTQ (parse + pack + GPU + outputs), config-2 shaped."""
import os, subprocess, sys, tempfile, time, shutil
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from phagefilter_b200.synth import make_genomes, simulate_reads

BIN = os.path.join(ROOT, "phagefilter_b200", "bin", "phage_filter")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
d = tempfile.mkdtemp(prefix="pf_cli_")
genomes = make_genomes(10, 10, 1001)
with open(os.path.join(d, "genomes.fa"), "wb") as f:
    for gid, seq in genomes:
        f.write(b">%s\n%s\n" % (gid.encode(), seq))
reads, _ = simulate_reads(genomes, n, 150, 2001, error_rates=(0.0, 0.01))
t0 = time.perf_counter()
with open(os.path.join(d, "reads.fq"), "wb") as f:
    q = b"#" * 150
    for lo in range(0, n, 100_000):
        f.write(b"".join(b"@r%d\n%s\n+\n%s\n" % (i, reads[i].tobytes(), q) for i in range(lo, min(n, lo + 100_000))))
print("wrote fastq in %.1f s, %.0f MB" % (time.perf_counter() - t0, os.path.getsize(os.path.join(d, "reads.fq")) / 1e6))
def run(*args):
    t0 = time.perf_counter()
    p = subprocess.run([BIN, *args], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)
    if p.stderr.strip():
        print("   ", p.stderr.strip())
    return time.perf_counter() - t0
t = run("build", "-g", os.path.join(d, "genomes.fa"), "-d", os.path.join(d, "db"), "--seed-one", "1", "--seed-two", "2")
print("build 100 genomes: %.2f s" % t)
for extra, label in (([], "counts only"), (["--pos-filter", "--neg-filter"], "pos+neg filter"), (["-b", "100000"], "counts, -b 100000")):
    t = run("query", "-r", os.path.join(d, "reads.fq"), "-o", os.path.join(d, "out"), "-d", os.path.join(d, "db"), "--stats", *extra)
    print("query %-20s %.2f s  -> %.2f M reads/s (process start, DB load, parse, pack, GPU, outputs)" % (label, t, n / t / 1e6))
shutil.rmtree(d)
