#!/usr/bin/env python
"""Diagnostic: per-call timing of the host driver's GPU calls (PF_TIMING=1) for several driver settings."""
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
with open(os.path.join(d, "reads.fq"), "wb") as f:
    q = b"#" * 150
    for lo in range(0, n, 100_000):
        f.write(b"".join(b"@r%d\n%s\n+\n%s\n" % (i, reads[i].tobytes(), q) for i in range(lo, min(n, lo + 100_000))))
env = dict(os.environ, PF_TIMING="1")
def run(label, *args):
    t0 = time.perf_counter()
    p = subprocess.run([BIN, *args], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True, env=env)
    print("== %s: %.2f s" % (label, time.perf_counter() - t0))
    print(p.stderr.strip())
run("build", "build", "-g", os.path.join(d, "genomes.fa"), "-d", os.path.join(d, "db"), "--seed-one", "1", "--seed-two", "2")
base = ["query", "-r", os.path.join(d, "reads.fq"), "-o", os.path.join(d, "out"), "-d", os.path.join(d, "db"), "--stats"]
for label, extra in (("b100", []), ("b100000", ["-b", "100000"]), ("b100 again", []), ("b100 host-threads 4", ["--host-threads", "4"]),
                     ("b100 pos+neg", ["--pos-filter", "--neg-filter"]), ("b100000 again", ["-b", "100000"])):
    run(label, *base, *extra)
shutil.rmtree(d)
