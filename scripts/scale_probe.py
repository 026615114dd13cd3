#!/usr/bin/env python
"""Exploratory scale runs of the query path on other BASELINE.json shapes (not the driver's bench):
   python scripts/scale_probe.py --families 200 --family-size 10 --reads 10000000 --read-len 150 --theta 0.8 \
          --background 0.9 --error 0.01
Builds the DB on the GPU (reference on-disk format), opens it, runs the query twice and prints one JSON line."""
import argparse
import ctypes as C
import json
import os
import shutil
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--families", type=int, default=100)
    ap.add_argument("--family-size", type=int, default=10)
    ap.add_argument("--reads", type=int, default=2_000_000)
    ap.add_argument("--read-len", type=int, default=150)
    ap.add_argument("--theta", type=float, default=0.8)
    ap.add_argument("--background", type=float, default=0.9)
    ap.add_argument("--error", type=float, default=0.01)
    ap.add_argument("--largest", type=int, default=1_000_000)
    ap.add_argument("--genome-len", type=int, default=50_000)
    ap.add_argument("--modes", default="1,2", help="comma list of pf_db_set_mode values to run (0 auto, 1 node-at-a-time, 2 sliced)")
    ap.add_argument("--reuse", action="store_true", help="keep the database under the temp dir and reuse it when the parameters match")
    ap.add_argument("--exact", action="store_true", help="evaluate every node exactly (no step-limited pre-test)")
    ap.add_argument("--oracle-sample", type=int, default=0, help="check this many reads against the CPU oracle")
    a = ap.parse_args()
    from phagefilter_b200 import BloomTree, _lib
    from phagefilter_b200.bloom_tree import BloomTreeBuilder
    from phagefilter_b200.query import PackedReads, query_packed
    from phagefilter_b200.synth import make_genomes, reads_to_concat, simulate_reads
    L = _lib.lib()
    t0 = time.perf_counter()
    lo, hi = int(a.genome_len * 0.8), int(a.genome_len * 1.2)
    genomes = make_genomes(a.families, a.family_size, 1003, len_lo=lo, len_hi=hi)
    d = os.path.join(tempfile.gettempdir(), "pf_scale_db")
    key = f"{a.families}x{a.family_size} len{a.genome_len} largest{a.largest}"
    marker = os.path.join(d, "scale_probe.key")
    t_build = t_save = 0.0
    if not (a.reuse and os.path.exists(marker) and open(marker).read() == key):
        shutil.rmtree(d, ignore_errors=True)
        b = BloomTreeBuilder(20, 0.001, a.largest)
        for gid, seq in genomes:
            b.insert(gid, seq)
        t_build = time.perf_counter() - t0
        t0 = time.perf_counter()
        b.save(d)
        b.close()
        t_save = time.perf_counter() - t0
        with open(marker, "w") as f:
            f.write(key)
    t0 = time.perf_counter()
    tree = BloomTree.load(d)
    t_open = time.perf_counter() - t0
    if a.exact:
        tree.set_lazy(False)
    info = tree.info
    reads, src = simulate_reads(genomes, a.reads, a.read_len, 2003, error_rates=(a.error,), background_frac=a.background)
    blob, offs = reads_to_concat(reads)
    packed = PackedReads.from_concat(blob, offs)
    out = {}
    results = {}
    for mode in [int(x) for x in a.modes.split(",")]:
      tree.set_mode(mode)
      for rep in range(2):  # the second run is the warm one
          tree.reset_stats()
          tree.reset_counts()
          t0 = time.perf_counter()
          off, leaf = query_packed(tree, packed, a.theta, want_hits=True)
          wall = time.perf_counter() - t0
          st = tree.stats()
          out = {"reads": a.reads, "read_len": a.read_len, "theta": a.theta, "leaves": int(info.n_leaves), "nodes": int(info.n_nodes),
                 "levels": int(info.n_levels), "filter_gb": round(info.filter_bytes / 1e9, 2), "monotone": f"{info.n_monotone}/{info.n_internal}",
                 "build_s": round(t_build, 1), "save_s": round(t_save, 1), "open_s": round(t_open, 1),
                 "wall_ms": round(wall * 1e3, 1), "device_ms": round(st.device_ms, 1), "probe_ms": round(st.probe_kernel_ms, 1),
                 "reads_per_s_wall": round(a.reads / wall), "pairs": int(st.pairs), "probes": int(st.probes_issued),
                 "probes_per_s": round(st.probes_issued / max(st.probe_kernel_ms * 1e-3, 1e-9)), "hits": int(len(leaf)),
                 "sectors": int(st.sector_loads), "sliced_ms": round(st.sliced_kernel_ms, 1), "tile_pairs": int(st.sliced_pairs),
                 "sectors_per_s": round(st.sector_loads / max(st.sliced_kernel_ms * 1e-3, 1e-9)),
                 "group_rounds": int(st.group_rounds), "lazy": not a.exact, "mode": mode,
                 "sliced_blocks": int(st.sliced_blocks), "tiles": int(st.sliced_tiles), "table_gb": round(st.sliced_table_bytes / 1e9, 2),
                 "levels_run": int(st.levels)}
      results[mode] = (off.copy(), leaf.copy())
      print(json.dumps(out), flush=True)
    ms = sorted(results)
    for m in ms[1:]:
        same = bool((results[m][0] == results[ms[0]][0]).all() and (results[m][1] == results[ms[0]][1]).all())
        print(json.dumps({"modes_identical": [ms[0], m], "same": same}), flush=True)
    if a.oracle_sample:
        import numpy as np
        from oracle import pf_oracle
        sel = np.linspace(0, a.reads - 1, a.oracle_sample).astype(np.int64)
        ot = pf_oracle.Tree.load(d)
        res = ot.query_batch([reads[i].tobytes() for i in sel], a.theta)
        want = res.hit_sets(len(sel))
        bad = sum(1 for j, r in enumerate(sel) if frozenset(int(x) for x in leaf[int(off[r]):int(off[r + 1])]) != want[j])
        print(json.dumps({"oracle_sample": int(a.oracle_sample), "mismatches": bad}))
    packed.close()
    tree.close()
    if not a.reuse:
        shutil.rmtree(d, ignore_errors=True)


if __name__ == "__main__":
    main()
