#!/usr/bin/env python
"""bench.py -- reads/s of PhageFilter's `query` hot path on B200 (BASELINE.json metric).

A step = one pass of the hot path (gSBT descent, per-read leaf lists, per-genome counts) over one batch of
synthetic reads.  `--config` picks the BASELINE.json configuration; the default, cfg3, is the
largest configuration one GPU holds:

  cfg1  examples/genomes/viral_genome_dir (107 genomes) vs examples/test_reads (90,000 x 100 bp), -f 1.0
  cfg2  100 synthetic phage genomes (~50 kb) vs 1 M x 150 bp reads per step, -f 1.0
  cfg3  10,000 synthetic genomes (35.9 GB of filters) vs 5 M x 150 bp reads per step (20 steps = the 100 M reads of
        the configuration), 10 % phage spike-in, 90 % background, -f 0.8
  cfg4  the cfg3 database vs 200,000 x 10 kb reads per step (5 steps = 1 M reads), -f 0.9
  cfg5  100,000 genomes, --largest-genome 450000 (162 GB replica): explicit only -- the database has to pass through the
        file system between build and query, and this pool's boxes have 80 GB of disk and 196 GB of RAM
  cfg5s a fifth of cfg5 (20,000 genomes, 32 GB), same geometry, reads and threshold (-f 1.0)

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config cfgN] [--reads R]

`value`  : whole-job reads/s, inputs resident in HBM, timed with CUDA events on the library's stream.
`e2e`    : reads/s through the C ABI with pinned HOST buffers (H2D of the 2-bit batch and D2H of the hit lists
           inside the timed region).
`--impl reference`: the reference's algorithm on the host CPU cores (the C oracle port; the Rust binary cannot be
           built in this image), all threads, each step a bounded sample of the same workload.
The database is built on the GPU once per box and kept under $PF_BENCH_CACHE (default <tmp>/pf_bench_cache).
"""
from __future__ import annotations

import argparse
import ctypes as C
import datetime
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "query_reads_per_s"
UNIT = "reads/s"
K_MER, FPR = 20, 0.001

CONFIGS = {
    "cfg1": dict(kind="examples", largest=1_000_000, theta=1.0, read_len=100, reads=90_000, steps=5,
                 cpu_sample=20_000, lru_sample=500,
                 name="cfg1: examples/genomes/viral_genome_dir (107 genomes) vs examples/test_reads (9 files, 90000 x 100 bp), -f 1.0"),
    "cfg2": dict(kind="synth", families=10, family_size=10, largest=1_000_000, theta=1.0, read_len=150, reads=1_000_000,
                 errors=(0.0, 0.01), background=0.0, seed_g=1001, seed_r=2001, steps=5, cpu_sample=200_000, lru_sample=2_000,
                 name="cfg2: 100 synthetic phage genomes (~50 kb) vs {reads} simulated 150 bp reads, -f 1.0"),
    "cfg3": dict(kind="synth", families=1000, family_size=10, largest=1_000_000, theta=0.8, read_len=150, reads=5_000_000,
                 errors=(0.01,), background=0.9, seed_g=1003, seed_r=2003, steps=20, cpu_sample=20_000, lru_sample=500,
                 name="cfg3: 10000 synthetic phage genomes (35.9 GB gSBT) vs {reads} x 150 bp reads per step "
                      "(20 steps = 100 M reads), 10 % phage spike-in, -f 0.8"),
    "cfg3s": dict(kind="synth", families=100, family_size=10, largest=1_000_000, theta=0.8, read_len=150, reads=5_000_000,
                  errors=(0.01,), background=0.9, seed_g=1003, seed_r=2003, steps=5, cpu_sample=20_000, lru_sample=500,
                  name="cfg3s (profiling stand-in for cfg3: same reads, filter geometry and tile-table size, a tenth of the "
                       "genomes so that ncu's kernel replay can save and restore device memory): 1000 synthetic genomes "
                       "(3.6 GB gSBT) vs {reads} x 150 bp reads per step, 10 % spike-in, -f 0.8"),
    "cfg4": dict(kind="synth", families=1000, family_size=10, largest=1_000_000, theta=0.9, read_len=10_000, reads=200_000,
                 errors=(0.001,), background=0.9, seed_g=1003, seed_r=2004, steps=5, cpu_sample=400, lru_sample=100,
                 name="cfg4: 10000-genome gSBT (35.9 GB) vs {reads} x 10 kb reads per step (5 steps = 1 M reads), "
                      "10 % spike-in, -f 0.9"),
    "cfg5s": dict(kind="synth", families=2_000, family_size=10, largest=450_000, theta=1.0, read_len=150, reads=5_000_000,
                  errors=(0.0, 0.01), background=0.9, seed_g=1005, seed_r=2005, steps=10, cpu_sample=5_000, lru_sample=200,
                  name="cfg5s (a fifth of cfg5: the same geometry, reads and threshold, 20000 genomes -- the greedy build of all "
                       "100000 takes hours): 20000 synthetic genomes (--largest-genome 450000, 32 GB gSBT) vs {reads} x 150 bp "
                       "reads per GPU per step, 10 % spike-in, -f 1.0"),
    "cfg5": dict(kind="synth", families=10_000, family_size=10, largest=450_000, theta=1.0, read_len=150, reads=5_000_000,
                 errors=(0.0, 0.01), background=0.9, seed_g=1005, seed_r=2005, steps=20, cpu_sample=5_000, lru_sample=200,
                 name="cfg5: 100000 synthetic genomes (--largest-genome 450000, 162 GB gSBT) vs {reads} x 150 bp reads "
                      "per GPU per step, 10 % spike-in, -f 1.0"),
}


def cache_root() -> str:
    return os.environ.get("PF_BENCH_CACHE", os.path.join(tempfile.gettempdir(), "pf_bench_cache"))


def db_key(cfg) -> str:
    if cfg["kind"] == "examples":
        return "examples_viral107"
    return f"synth_{cfg['families']}x{cfg['family_size']}_g{cfg['seed_g']}_l{cfg['largest']}"


def cfg1_data():
    d = os.path.join(ROOT, "tests", "_cfg1_data")
    if not os.path.isdir(os.path.join(d, "test_reads")):
        raise SystemExit("bench.py --config cfg1 needs tests/_cfg1_data (python tests/golden/fetch_cfg1.py where the "
                         "reference is present; the copy travels with the repo snapshot)")
    return d


def load_genomes(cfg):
    if cfg["kind"] == "examples":
        from phagefilter_b200.file_parser import read_records
        gdir = os.path.join(cfg1_data(), "viral_genome_dir")
        out = []
        for fn in sorted(os.listdir(gdir)):
            for rec in read_records(os.path.join(gdir, fn)):
                out.append((rec.id, rec.sequence))
        return out
    from phagefilter_b200.synth import make_genomes
    return make_genomes(cfg["families"], cfg["family_size"], cfg["seed_g"])


def make_reads(cfg, genomes, n_reads: int, seed: int):
    """(blob, offs) of one batch.  cfg1: the example reads, cycled to n_reads."""
    from phagefilter_b200.synth import reads_to_concat, simulate_reads_fast
    if cfg["kind"] == "examples":
        from phagefilter_b200.file_parser import read_records
        rdir = os.path.join(cfg1_data(), "test_reads")
        seqs = [rec.sequence for fn in sorted(os.listdir(rdir)) for rec in read_records(os.path.join(rdir, fn))]
        seqs = [seqs[i % len(seqs)] for i in range(n_reads)]
        offs = np.zeros(n_reads + 1, dtype=np.uint64)
        offs[1:] = np.cumsum([len(s) for s in seqs], dtype=np.uint64)
        return b"".join(seqs), offs
    reads, _ = simulate_reads_fast(genomes, n_reads, cfg["read_len"], seed, error_rates=cfg["errors"],
                                   background_frac=cfg["background"])
    return reads_to_concat(reads)


def ensure_db(cfg, genomes, device: int) -> tuple:
    """Builds the database on the GPU (reference on-disk format) unless this box already has it.  Returns (dir, seconds)."""
    from phagefilter_b200.bloom_tree import BloomTreeBuilder
    d = os.path.join(cache_root(), db_key(cfg))
    marker = os.path.join(d, "bench.complete")
    if os.path.exists(marker):
        return d, 0.0
    shutil.rmtree(d, ignore_errors=True)
    os.makedirs(cache_root(), exist_ok=True)
    # the GPU boxes have ~80 GB of scratch disk: make room by dropping databases of other configurations
    from phagefilter_b200 import _lib
    need = 2 * len(genomes) * (int(_lib.lib().pf_needed_bits(C.c_float(FPR), cfg["largest"])) // 8 + 256) * 1.02
    if shutil.disk_usage(cache_root()).free < need:
        for other in os.listdir(cache_root()):
            if other != db_key(cfg):
                shutil.rmtree(os.path.join(cache_root(), other), ignore_errors=True)
    t0 = time.perf_counter()
    b = BloomTreeBuilder(K_MER, FPR, cfg["largest"], device=device)
    for gid, seq in genomes:
        b.insert(gid, seq)
    b.save(d)
    b.close()
    with open(marker, "w") as f:
        f.write("ok\n")
    return d, time.perf_counter() - t0


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---- the reference's algorithm on the host cores (oracle port) --------------------------------------------------------
def cpu_resident(db_dir, cfg, blob, offs, n_sample, steps, warmup, threads=0, target_s=None):
    """Every filter in RAM (the reference with --cache-size >= nodes), one block per step: its fastest configuration.
    target_s: size the sample from a pilot of 200 reads so that a step takes about that long (bounded by n_sample).
    Returns (reads per step, seconds per timed step, reference-semantics probes and pairs per read)."""
    from oracle import pf_oracle
    tree = pf_oracle.Tree.load(db_dir)
    n_sample = min(n_sample, len(offs) - 1)
    if target_s and n_sample > 200:
        po = np.ascontiguousarray(offs[:201])
        t0 = time.perf_counter()
        tree.query_batch(None, cfg["theta"], threads=threads, want_hits=True, concat=(blob[: int(po[-1])], po))
        per_read = (time.perf_counter() - t0) / 200
        n_sample = int(max(200, min(n_sample, target_s / max(per_read, 1e-9))))
        tree.reset_counts()
    sub_offs = np.ascontiguousarray(offs[: n_sample + 1])
    sub_blob = blob[: int(sub_offs[-1])]
    times, res = [], None
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        res = tree.query_batch(None, cfg["theta"], threads=threads, want_hits=True, concat=(sub_blob, sub_offs))
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return n_sample, times, res.probes_ref / max(n_sample, 1), res.pairs / max(n_sample, 1), res.hits


def cpu_faithful(db_dir, cfg, blob, offs, n_sample, threads, cache_size=10, block_size=100):
    """The reference with its default flags (main.rs:111-122): filters through an LRU of --cache-size 10 entries,
    re-read from disk on every miss (cache.rs:56-77), query_batch per block of --block-size-reads 100."""
    from oracle import pf_oracle
    tree = pf_oracle.Tree.load_lazy(db_dir, cache_size=cache_size)
    n_sample = min(n_sample, len(offs) - 1)
    sub_offs = np.ascontiguousarray(offs[: n_sample + 1])
    sub_blob = blob[: int(sub_offs[-1])]
    t0 = time.perf_counter()
    tree.query_blocks(None, cfg["theta"], block_size=block_size, threads=threads, want_hits=True, concat=(sub_blob, sub_offs))
    dt = time.perf_counter() - t0
    loads, hits, nbytes = tree.cache_stats()
    return {"value": n_sample / dt, "reads": n_sample, "threads": threads, "cache_size": cache_size, "block_size": block_size,
            "filter_files_loaded": loads, "filter_bytes_decoded": nbytes, "seconds": dt}


def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    genomes = load_genomes(cfg)
    db_dir = os.path.join(cache_root(), db_key(cfg))
    built_with = "found on this box"
    if not os.path.exists(os.path.join(db_dir, "bench.complete")):
        import torch
        if torch.cuda.is_available():
            # database files are setup, not the path that is timed; the GPU builder's files are byte-identical to the
            # port's own (tests/test_gpu_parity.py, tests/test_cli_gpu.py) and 100x faster to produce
            db_dir, _ = ensure_db(cfg, genomes, 0)
            built_with = "pf_builder (GPU), files byte-identical to the port's"
        else:
            from oracle import pf_oracle
            t = pf_oracle.Tree(K_MER, FPR, cfg["largest"])
            for gid, seq in genomes:
                t.insert(gid, seq)
            os.makedirs(cache_root(), exist_ok=True)
            t.save(db_dir)
            open(os.path.join(db_dir, "bench.complete"), "w").write("ok\n")
            built_with = "oracle port (CPU)"
    cores = os.cpu_count() or 1
    n_sample = cfg["cpu_sample"]
    blob, offs = make_reads(cfg, genomes, n_sample, cfg.get("seed_r", 0))
    warm = max(args.warmup, 1)
    n_sample, times, p_ref, pairs_ref, _ = cpu_resident(db_dir, cfg, blob, offs, n_sample, args.steps, warm, target_s=4.0)
    total = sum(times)
    v = n_sample * len(times) / total
    sample = (f"first {n_sample} reads of the workload per step, all {cores} host threads (OpenMP), every filter resident in "
              f"RAM and one block per step (the reference's fastest setting: --cache-size >= nodes, --block-size-reads >= reads)")
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": warm, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": cfg["name"].format(reads=args.reads), "sample_reads_per_step": n_sample,
                   "database": built_with,
                   "note": "C port of the reference's query path (Rust toolchain absent in this image); ASCII k-mers "
                           "re-hashed at every node as in src/query.rs"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "reference_probes_per_read": p_ref, "reference_pairs_per_read": pairs_ref},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def run_ours(args, cfg):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the GPU path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), timeout=datetime.timedelta(hours=2))

    from phagefilter_b200 import BloomTree, _lib
    from phagefilter_b200.query import PackedReads, query_sharded
    L = _lib.lib()
    theta = cfg["theta"]

    genomes = load_genomes(cfg)
    build_s = 0.0
    if rank == 0:
        db_dir, build_s = ensure_db(cfg, genomes, local_rank)
    if world > 1:
        dist.barrier()
    db_dir = os.path.join(cache_root(), db_key(cfg))
    # one NCCL communicator owned by the library: id made on rank 0, handed over by the host
    from phagefilter_b200.shard import exchange_nccl_id
    nccl_id = exchange_nccl_id(rank) if (world > 1 or args.shard_tree) else None
    t0 = time.perf_counter()
    if args.shard_tree:
        # subtree shards (trees larger than HBM): top replicated, subtrees owned by ranks, frontier all-to-all
        tree = BloomTree.open_sharded(db_dir, local_rank, rank, world, nccl_id,
                                      cut_level=None if args.cut_level < 0 else args.cut_level)
    else:
        tree = BloomTree(db_dir, local_rank)
        if world > 1:
            _lib.check(L.pf_comm_init(tree._h, world, rank, nccl_id))
    open_s = time.perf_counter() - t0
    if args.mode is not None:
        tree.set_mode(args.mode)
    info = tree.info
    query_device = L.pf_query_sharded_device if args.shard_tree else L.pf_query_device

    # two distinct batches per rank, alternated over the steps
    n_batches = 1 if cfg["kind"] == "examples" else 2
    host_batches, packed = [], []
    for b in range(n_batches):
        blob, offs = make_reads(cfg, genomes, args.reads, cfg.get("seed_r", 0) + 1000 * b + rank)
        host_batches.append((blob, offs))
        packed.append(PackedReads.from_concat(blob, offs))
    dev_batches = []
    for p in packed:
        h = C.c_void_p()
        _lib.check(L.pf_batch_upload(tree._h, p.batch, C.byref(h)))
        dev_batches.append(h)
    stream = torch.cuda.ExternalStream(L.pf_db_stream(tree._h), device=torch.device("cuda", local_rank))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def flush_l2():
        flush.add_(1)
        torch.cuda.synchronize()

    hits = _lib.Hits()

    def step_device(i):
        _lib.check(query_device(tree._h, dev_batches[i % n_batches], C.c_float(theta), 1, C.byref(hits)))

    def combine():
        # per-genome counts are combined with ONE NCCL reduce at the end of the run (north_star)
        if world > 1:
            _lib.check(L.pf_allreduce_counts(tree._h))

    pipe = [C.c_void_p(), C.c_void_p()]  # two device batches: upload of block i+1 overlaps the query of block i

    def run_e2e(n_steps):
        _lib.check(L.pf_batch_upload_async(tree._h, packed[0].batch, C.byref(pipe[0])))
        for i in range(n_steps):
            if i + 1 < n_steps:
                _lib.check(L.pf_batch_upload_async(tree._h, packed[(i + 1) % n_batches].batch, C.byref(pipe[(i + 1) % 2])))
            _lib.check(query_device(tree._h, pipe[i % 2], C.c_float(theta), 1, C.byref(hits)))
        combine()

    sampler = ClockSampler(local_rank)
    t0 = time.perf_counter()
    for i in range(args.warmup):
        step_device(i)
    combine()
    warm_s = time.perf_counter() - t0
    # ---- timed region: device-resident inputs, CUDA events on the launching stream -------------
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    tree.reset_stats()
    sampler.start()
    ms = 0.0
    for i in range(args.steps + 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if i < args.steps:
            flush_l2()
            e0.record(stream)
            step_device(i)
        else:
            e0.record(stream)
            combine()
        e1.record(stream)
        e1.synchronize()
        ms += e0.elapsed_time(e1)
    torch.cuda.synchronize()
    clocks = sampler.stop()
    if world > 1:
        dist.barrier()
    st = tree.stats()
    n_hits = int(hits.n_hits)
    last_off = np.ctypeslib.as_array(hits.read_off, shape=(args.reads + 1,)).copy()
    last_leaf = np.ctypeslib.as_array(hits.leaf, shape=(n_hits,)).copy() if n_hits else np.zeros(0, dtype=np.uint32)
    last_batch = (args.steps - 1) % n_batches
    if last_batch == 0:
        first_off, first_leaf = last_off, last_leaf
    else:  # the CPU sample below is the head of batch 0
        _lib.check(query_device(tree._h, dev_batches[0], C.c_float(theta), 1, C.byref(hits)))
        first_off = np.ctypeslib.as_array(hits.read_off, shape=(args.reads + 1,)).copy()
        first_leaf = (np.ctypeslib.as_array(hits.leaf, shape=(int(hits.n_hits),)).copy() if hits.n_hits
                      else np.zeros(0, dtype=np.uint32))
    # ---- end-to-end through the C-ABI call with host buffers --------------------------------------
    run_e2e(2)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    tree.reset_stats()
    t0 = time.perf_counter()
    run_e2e(args.steps)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    st2 = tree.stats()
    for b in pipe:
        L.pf_batch_free(tree._h, b)

    tmax = torch.tensor([ms, e2e_s * 1e3], dtype=torch.float64, device="cuda")
    per_rank = [tmax.clone() for _ in range(world)]
    if world > 1:
        dist.all_gather(per_rank, tmax)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_all, e2e_ms_all = float(tmax[0]), float(tmax[1])
    total_reads = args.reads * world * args.steps
    value = total_reads / (ms_all * 1e-3)
    e2e_value = total_reads / (e2e_ms_all * 1e-3)

    counts = tree.leaf_counts()
    assert int(counts.sum()) > 0
    shard = None
    if args.shard_tree:
        si, ss = tree.shard_info(), tree.shard_stats()
        shard = {"cut_level": int(si.cut_level), "top_nodes": int(si.top_nodes), "owned_nodes_rank0": int(si.owned_nodes),
                 "resident_filter_bytes_rank0": int(si.resident_bytes),
                 "pairs_sent_per_query_rank0": int(ss.pairs_sent) // max(int(ss.queries), 1),
                 "bytes_sent_per_query_rank0": int(ss.bytes_sent) // max(int(ss.queries), 1),
                 "bytes_received_per_query_rank0": int(ss.bytes_received) // max(int(ss.queries), 1)}

    # ---- N > 1, replicated tree: the subtree-shard path on the same database and reads (the path trees larger than HBM
    # take): every rank sends a slice of its last batch through pf_query_sharded and must get exactly the hit lists the
    # replicated tree gave it for those reads ------------------------------------------------------------------------------
    sharded_check = None
    if world > 1 and not args.shard_tree and not args.no_shard_check:
        n_sl = min(args.reads, max(1000, int(200_000 * 150 / max(cfg["read_len"], 1))))
        blob, offs = host_batches[last_batch]
        sub_offs = np.ascontiguousarray(offs[: n_sl + 1])
        sp = PackedReads.from_concat(blob[: int(sub_offs[-1])], sub_offs)
        t0 = time.perf_counter()
        stree = BloomTree.open_sharded(db_dir, local_rank, rank, world, exchange_nccl_id(rank))
        sopen = time.perf_counter() - t0
        query_sharded(stree, sp, theta)  # warm-up (allocations)
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        soff, sleaf = query_sharded(stree, sp, theta)
        torch.cuda.synchronize()
        sq = time.perf_counter() - t0
        same = bool((soff == last_off[: n_sl + 1]).all() and (sleaf == last_leaf[: int(last_off[n_sl])]).all())
        # and at -f 1.0, where the top of the tree is not skipped by the step plan, so that (read, node) pairs really cross
        # between ranks at the cut: against the replicated tree's answer for the same reads at that threshold
        from phagefilter_b200.query import query_packed
        tree.set_mode(1)  # no new tiling for one cross-check block
        roff, rleaf = query_packed(tree, sp, 1.0)
        tree.set_mode(0 if args.mode is None else args.mode)
        s1off, s1leaf = query_sharded(stree, sp, 1.0)
        same = same and bool((s1off == roff).all() and (s1leaf == rleaf).all())
        si, ss = stree.shard_info(), stree.shard_stats()
        flag = torch.tensor([1 if same else 0, int(ss.pairs_sent), int(ss.pairs_received), int(ss.hits_sent)], dtype=torch.int64, device="cuda")
        mn = flag.clone()
        dist.all_reduce(mn, op=dist.ReduceOp.MIN)
        sm = flag.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        tq = torch.tensor([sq], dtype=torch.float64, device="cuda")
        dist.all_reduce(tq, op=dist.ReduceOp.MAX)
        sharded_check = {"reads_per_rank": n_sl, "identical_to_replicated_on_every_rank": bool(int(mn[0]) == 1),
                         "cut_level": int(si.cut_level), "top_nodes": int(si.top_nodes),
                         "resident_filter_bytes_rank0": int(si.resident_bytes),
                         "thresholds": [theta, 1.0],
                         "pairs_sent_all_ranks": int(sm[1]), "pairs_received_all_ranks": int(sm[2]), "hits_sent_all_ranks": int(sm[3]),
                         "reads_per_s_all_ranks": n_sl * world / float(tq[0]), "open_s_rank0": round(sopen, 2),
                         "note": "subtree-shard path (pf_db_open_sharded / pf_query_sharded): replicated top, subtrees owned "
                                 "by ranks, frontier and hits exchanged with NCCL; node-at-a-time descent"}
        sp.close()
        stree.close()

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        steps = args.steps
        probes = int(st.probes_issued)
        pairs = int(st.pairs)
        sliced = int(st.sliced_blocks) > 0
        read_bytes = (cfg["read_len"] + 3) // 4
        memo_lookups, memo_hits = int(st.memo_lookups), int(st.memo_hits)
        probe_ms = float(st.probe_kernel_ms)
        n_launch = max(int(st.probe_launches), 1)
        sectors, sliced_ms, sliced_pairs = int(st.sector_loads), float(st.sliced_kernel_ms), int(st.sliced_pairs)
        lines = int(st.line_loads)  # 128-byte line loads of the entry line kernel (a row of up to four entry tiles each)
        entry_ms = float(st.entry_kernel_ms)
        n_sliced_launch = max(int(st.sliced_launches), 1)  # one timed launch per tile-tree depth
        # DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture of this same command
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            tj = tj.get(args.config) or tj.get({"cfg3": "cfg3s", "cfg4": "cfg3s"}.get(args.config, ""), {})
            if sliced:
                # mean over the launches of both sliced kernels (entry line kernel + the depths below), as `achieved` is
                tot = sum(tj.get(k + "_dram_bytes_per_launch", 0.0) * tj.get(k + "_launches", 0)
                          for k in ("sliced_entry_quad_kernel", "sliced_probe_kernel"))
                cnt = sum(tj.get(k + "_launches", 0) for k in ("sliced_entry_quad_kernel", "sliced_probe_kernel")
                          if k + "_dram_bytes_per_launch" in tj)
                traffic = tot / cnt if cnt else None
            else:
                traffic = tj.get("probe_kernel_dram_bytes_per_launch")
            traffic_src = tj.get("source")
        rate = C.c_double(0)
        if sliced:
            # one 32 B sector per row load (each answers one probe step for every node of the tile) + the
            # cached 8 B hash value per k-mer and pair
            alg_bytes = 32 * sectors + 128 * lines
            achieved = alg_bytes / (sliced_ms * 1e-3) / 1e9 if sliced_ms > 0 else 0.0
            peaks = {}
            for name, nbytes in (("one_256_column_table", int(info.words_per_filter) * 8 * 256),
                                 ("all_tables", min(max(int(st.sliced_table_bytes), 1 << 20), 48 << 30))):
                _lib.check(L.pf_microbench_sectors(local_rank, max(nbytes, 1 << 20), 60, C.byref(rate)))
                peaks[name] = rate.value
            rs_peak = peaks["all_tables"]
            sect_rate = (sectors + lines) / (sliced_ms * 1e-3) if sliced_ms > 0 else 0.0  # random accesses (sector or line)
            roofline = {
                "bound": "hbm",
                "kernel": "sliced_entry_quad_kernel (entry depth: one 128 B line per k-mer and probe step answers up to 1024 "
                          "nodes) + sliced_probe_kernel (depths below: one 32 B sector per k-mer and step, up to 256 nodes)",
                "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "note": "random gathers over tables far larger than L2 and than the 256 MB reach of the first-level TLB: "
                        "what binds is the rate of random ACCESSES (about 37 G/s on this part whether an access brings 4, 32 "
                        "or -- four adjacent lanes on one line -- 128 bytes: scripts/mb/mb_coop.cu, profiles/r2x_mb_*), not "
                        "bytes; `random_access_peak_per_s` is that rate measured in this process over the same footprint, "
                        "`frac_of_random_access_peak` the kernels' accesses (sector or line loads) against it, `frac` the "
                        "useful bytes they bring against the streaming-copy peak.  `traffic`: dram bytes per launch from the "
                        "ncu capture named in traffic_source (stand-in config for cfg3/cfg4: ncu's kernel replay has to save "
                        "and restore 72 GB of device memory there)",
                "algorithmic_bytes_per_launch": alg_bytes / n_sliced_launch, "launches_per_step": n_sliced_launch // steps,
                "avg_launch_ms": sliced_ms / n_sliced_launch,
                "kernel_share_of_step": sliced_ms / float(st.device_ms) if st.device_ms else None,
                "random_accesses_per_s": sect_rate,
                "per_kernel": {
                    "entry_depth (sliced_entry_quad_kernel when the entry tiles share lines)": {
                        "ms_per_step": entry_ms / steps, "line_loads_per_step": lines // steps,
                        "lines_per_s": lines / (entry_ms * 1e-3) if entry_ms > 0 and lines else None,
                        "gbytes_per_s": 128 * lines / (entry_ms * 1e-3) / 1e9 if entry_ms > 0 and lines else None},
                    "depths_below (sliced_probe_kernel)": {
                        "ms_per_step": (sliced_ms - entry_ms) / steps,
                        "sector_loads_per_step": (sectors // steps) if lines else None,
                        "sectors_per_s": sectors / ((sliced_ms - entry_ms) * 1e-3) if lines and sliced_ms > entry_ms else None}},
                "random_access_peak_per_s": peaks,
                "frac_of_random_access_peak": sect_rate / rs_peak if rs_peak else None,
                "tiles": int(st.sliced_tiles), "table_bytes": int(st.sliced_table_bytes),
                "handed_over_to_probe_kernel": {
                    "note": "reads the tiles could not rule out continue node by node (probe_kernel, L2-resident filters)",
                    "pairs_per_step": (pairs - sliced_pairs) // steps, "probes_per_step": probes // steps,
                    "memo_lookups_per_step": memo_lookups // steps, "kernel_ms_per_step": probe_ms / steps,
                    "share_of_step": probe_ms / float(st.device_ms) if st.device_ms else None,
                    "probes_per_s": (probes + memo_lookups) / (probe_ms * 1e-3) if probe_ms > 0 else 0.0},
            }
        else:
            # one 32 B sector per probe issued (and per k-mer memo look-up) + the 2-bit read per pair
            alg_bytes = 32 * (probes + memo_lookups) + read_bytes * pairs
            achieved = alg_bytes / (probe_ms * 1e-3) / 1e9 if probe_ms > 0 else 0.0
            l2_rate, hbm_rate = C.c_double(0), C.c_double(0)
            _lib.check(L.pf_microbench_sectors(local_rank, int(info.words_per_filter) * 8, 200, C.byref(l2_rate)))
            _lib.check(L.pf_microbench_sectors(local_rank, 8 << 30, 100, C.byref(hbm_rate)))
            probes_per_s = (probes + memo_lookups) / (probe_ms * 1e-3) if probe_ms > 0 else 0.0
            l2_peak_gbs = l2_rate.value * 32 / 1e9
            roofline = {
                "bound": "l2_sector", "kernel": f"probe_kernel<G={int(st.group_rounds)},small_m>", "achieved": achieved,
                "peak": l2_peak_gbs, "unit": "GB/s", "frac": achieved / l2_peak_gbs if l2_peak_gbs else None,
                "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": "measured in this run: random 32 B sector loads over one L2-resident filter (pf_microbench_sectors)",
                "note": "the frontier is node-major, so the probes are L2 hits by design: the roof is the L2 random-sector "
                        "rate, not HBM; hbm_equivalent_frac = the same bytes against MEASURED_PEAKS.json hbm_gbs",
                "hbm_copy_peak": peak, "hbm_copy_peak_source": peak_src, "hbm_equivalent_frac": achieved / peak,
                "algorithmic_bytes_per_launch": alg_bytes / n_launch, "launches_per_step": int(st.probe_launches) // steps,
                "avg_launch_ms": probe_ms / n_launch,
                "kernel_share_of_step": probe_ms / float(st.device_ms) if st.device_ms else None,
                "probes_per_s": probes_per_s,
                "l2_random_sector_peak_per_s": l2_rate.value, "hbm_random_sector_peak_per_s": hbm_rate.value,
            }
        # CPU baseline beside it: the oracle port on this box's host cores, bounded samples
        cores = os.cpu_count() or 1
        blob, offs = host_batches[0]
        n_cpu = 500 if args.profile else cfg["cpu_sample"]
        n_sample, times, p_ref, pairs_ref, cpu_hits = cpu_resident(db_dir, cfg, blob, offs, n_cpu, 1, 0, target_s=10.0)
        cpu_v = n_sample * len(times) / sum(times)
        # parity guard at the full database size: the oracle's (read, leaf) hits for its sample are exactly the GPU's
        order = np.lexsort((cpu_hits[:, 1], cpu_hits[:, 0])) if len(cpu_hits) else np.zeros(0, dtype=np.int64)
        want_off = np.zeros(n_sample + 1, dtype=np.uint64)
        if len(cpu_hits):
            np.add.at(want_off, cpu_hits[:, 0].astype(np.int64) + 1, 1)
        want_off = np.cumsum(want_off, dtype=np.uint64)
        want_leaf = cpu_hits[order, 1] if len(cpu_hits) else np.zeros(0, dtype=np.uint32)
        parity_ok = bool((first_off[: n_sample + 1] == want_off).all() and
                         (first_leaf[: int(want_off[-1])] == want_leaf).all())
        assert parity_ok, "GPU hit lists differ from the CPU oracle's on the sampled reads"
        modes = {"resident_all_cores": {"value": cpu_v, "reads": n_sample, "threads": cores}}
        if not args.profile:
            n_lru = cfg["lru_sample"]
            modes["faithful_lru10_block100_t4"] = cpu_faithful(db_dir, cfg, blob, offs, n_lru, 4)
            modes["faithful_lru10_block100_all_cores"] = cpu_faithful(db_dir, cfg, blob, offs, n_lru, cores)
        # correctness guard on the timed work: the CPU sample's hit lists are the GPU's for the same reads
        faithful = modes.get("faithful_lru10_block100_all_cores")
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": args.warmup,
            "ms_per_step": ms_all / steps, "ms_per_step_by_rank": [float(t[0]) / steps for t in per_rank],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic",
            "config": {"workload": cfg["name"].format(reads=args.reads), "config": args.config,
                       "reads_per_gpu_per_step": args.reads, "theta": theta,
                       "nodes": int(info.n_nodes), "leaves": int(info.n_leaves), "levels": int(info.n_levels),
                       "filter_bytes": int(info.filter_bytes), "want_hits": True,
                       "evaluation": "bit-sliced tiles (pf_sliced.cu)" if sliced else "node-at-a-time descent (pf_query.cu)",
                       "l2": "filters / tile tables and the read batches exceed the 126 MB L2; a 256 MB buffer is also "
                             "written between timed steps; two distinct batches alternate",
                       "parallelism": (f"reads sharded x{world}, tree cut into a replicated top and subtrees owned by ranks; "
                                       "frontier + hit all-to-all over NCCL") if args.shard_tree else
                       f"reads sharded x{world}, tree replicated"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(st2.h2d_bytes) // steps,
                    "d2h_bytes_per_step": int(st2.d2h_bytes) // steps, "ms_per_step": e2e_ms_all / steps,
                    "how": "wall clock over K steps of pf_batch_upload_async (pinned host 2-bit batch -> HBM) + "
                           "pf_query_device (descent, hit lists copied back to pinned host memory); the upload of "
                           "step i+1 overlaps the query of step i"},
            "gpu_launches": int(st.probe_launches + st.other_launches),
            "roofline": roofline,
            "cpu_baseline": {"value": (faithful or modes["resident_all_cores"])["value"], "unit": UNIT, "cores": cores,
                             "kind": "port",
                             "sample": (f"first {faithful['reads']} reads of the same workload, the reference's default flags: LRU of "
                                        f"--cache-size 10 filters re-read from disk, --block-size-reads 100, all {cores} host threads")
                             if faithful else f"first {n_sample} reads, filters resident, all {cores} host threads",
                             "modes": modes},
            "work": {"pairs_per_step": pairs // steps,
                     "sector_loads_per_step": sectors // steps, "line_loads_per_step": lines // steps,
                     "tile_pairs_per_step": sliced_pairs // steps,
                     "probes_issued_per_step": probes // steps,
                     "memo_lookups_per_step": memo_lookups // steps, "memo_hits_per_step": memo_hits // steps,
                     "hits_last_step": n_hits,
                     "reference_semantics_probes_per_step": int(p_ref * args.reads),
                     "reference_semantics_pairs_per_step": int(pairs_ref * args.reads),
                     # bloom probes per second under the REFERENCE's semantics -- the probes its descent would
                     # have issued for these reads, divided by the time this path takes to give the same answers (all ranks)
                     "reference_semantics_probes_per_s": p_ref * args.reads * world / (ms_all / steps * 1e-3) if ms_all > 0 else None,
                     "reference_semantics_sample": f"oracle on the first {n_sample} reads, scaled to the step",
                     "parity_check": {"reads": n_sample, "hits": int(len(cpu_hits)), "identical_to_oracle": parity_ok},
                     "db_build_s": round(build_s, 2), "db_open_s": round(open_s, 2), "warmup_s": round(warm_s, 2)},
        }
        if shard:
            line["shard"] = shard
        if sharded_check:
            line["sharded_check"] = sharded_check
        print(json.dumps(line))
    for h in dev_batches:
        L.pf_batch_free(tree._h, h)
    for p in packed:
        p.close()
    tree.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg3", choices=sorted(CONFIGS))
    ap.add_argument("--reads", type=int, default=None, help="reads per GPU per step (default: the configuration's)")
    ap.add_argument("--mode", type=int, default=None, choices=[0, 1, 2],
                    help="pf_db_set_mode: 0 cost model (default), 1 node-at-a-time, 2 bit-sliced tiles")
    ap.add_argument("--shard-tree", action="store_true",
                    help="subtree-sharded tree (pf_db_open_sharded) instead of one replica per GPU")
    ap.add_argument("--cut-level", type=int, default=-1, help="cut level of the sharded tree (-1: automatic)")
    ap.add_argument("--no-shard-check", action="store_true", help="N > 1: skip the subtree-shard cross-check")
    ap.add_argument("--profile", action="store_true",
                    help="for runs under ncu: shrink the CPU-baseline samples (numbers printed under a profiler are not bench values)")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    if args.steps is None:
        args.steps = cfg["steps"] if args.impl == "ours" else 3
    if args.reads is None:
        args.reads = cfg["reads"]
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    return run_reference(args, cfg) if args.impl == "reference" else run_ours(args, cfg)


if __name__ == "__main__":
    sys.exit(main())
