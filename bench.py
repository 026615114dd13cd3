#!/usr/bin/env python
"""bench.py -- reads/s of PhageFilter's `query` hot path on B200 (BASELINE.json metric).

A step = one pass of the hot path (level-synchronous gSBT descent, per-read leaf lists, per-genome
counts) over one batch of synthetic reads.  Workload at N=1 = BASELINE.json configs[1]:
100 synthetic phage genomes (~50 kb, 10 families x 10) vs 1,000,000 simulated 150 bp reads, -f 1.0,
default geometry (k=20, fpr 0.001, largest genome 1e6 => m=14,377,587 bits, K=10).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--reads R]

`value`  : whole-job reads/s, inputs resident in HBM, timed with CUDA events on the library's stream.
`e2e`    : reads/s through the C-ABI call pf_query_block with pinned HOST buffers (H2D of the 2-bit batch
           and D2H of the hit lists inside the timed region).
`--impl reference`: the reference's algorithm on the host CPU cores (the C oracle port; the Rust binary
           cannot be built in this image), all threads, each step a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import ctypes as C
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
K_MER, FPR, LARGEST, THETA = 20, 0.001, 1_000_000, 1.0
N_FAMILIES, FAMILY_SIZE, READ_LEN = 10, 10, 150
SEED_GENOMES, SEED_READS = 1001, 2001
CPU_SAMPLE_READS = 1_000_000


def workload_name(n_reads: int) -> str:
    return (f"cfg2: {N_FAMILIES * FAMILY_SIZE} synthetic phage genomes (~50 kb) vs {n_reads} simulated "
            f"{READ_LEN} bp reads, -f {THETA}, k={K_MER}, fpr={FPR}, largest-genome={LARGEST}")


def make_inputs(n_reads: int, rank: int):
    from phagefilter_b200.synth import make_genomes, reads_to_concat, simulate_reads
    genomes = make_genomes(N_FAMILIES, FAMILY_SIZE, SEED_GENOMES)
    reads, src = simulate_reads(genomes, n_reads, READ_LEN, SEED_READS + rank, error_rates=(0.0, 0.01))
    blob, offs = reads_to_concat(reads)
    return genomes, blob, offs, src


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


def cpu_reference_run(genomes, blob, offs, n_sample, steps, warmup, db_dir=None):
    """The reference's algorithm on host cores: oracle port, all threads, counting only its own work."""
    from oracle import pf_oracle
    if db_dir and os.path.exists(os.path.join(db_dir, "tree.bin")):
        tree = pf_oracle.Tree.load(db_dir)
    else:
        tree = pf_oracle.Tree(K_MER, FPR, LARGEST)
        for gid, seq in genomes:
            tree.insert(gid, seq)
    n_sample = min(n_sample, len(offs) - 1)
    sub_offs = np.ascontiguousarray(offs[: n_sample + 1])
    sub_blob = blob[: int(sub_offs[-1])]
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        tree.query_batch(None, THETA, threads=0, want_hits=True, concat=(sub_blob, sub_offs))
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return n_sample, times


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    genomes, blob, offs, _ = make_inputs(max(CPU_SAMPLE_READS, 1), 0)
    n_sample, times = cpu_reference_run(genomes, blob, offs, CPU_SAMPLE_READS, args.steps, max(args.warmup, 1))
    total = sum(times)
    v = n_sample * len(times) / total
    cores = os.cpu_count() or 1
    sample = f"first {n_sample} reads of the workload per step, all {cores} host threads (OpenMP)"
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": max(args.warmup, 1), "ms_per_step": 1e3 * total / len(times), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": workload_name(args.reads), "sample_reads_per_step": n_sample,
                   "note": "C port of the reference's query path (Rust toolchain absent); filters resident in RAM, "
                           "one block, ASCII k-mers re-hashed at every node as in src/query.rs"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the GPU path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from phagefilter_b200 import BloomTree, _lib
    from phagefilter_b200.bloom_tree import BloomTreeBuilder
    from phagefilter_b200.query import PackedReads
    L = _lib.lib()

    genomes, blob, offs, src = make_inputs(args.reads, rank)
    db_dir = os.path.join(tempfile.gettempdir(), f"pf_bench_db_{os.getuid()}")
    if rank == 0:
        shutil.rmtree(db_dir, ignore_errors=True)
        t0 = time.perf_counter()
        b = BloomTreeBuilder(K_MER, FPR, LARGEST, device=local_rank)  # GPU build, reference on-disk format
        for gid, seq in genomes:
            b.insert(gid, seq)
        b.save(db_dir)
        b.close()
        build_s = time.perf_counter() - t0
    if world > 1:
        dist.barrier()
    # one NCCL communicator owned by the library: id made on rank 0, handed over by the host
    from phagefilter_b200.shard import exchange_nccl_id
    nccl_id = exchange_nccl_id(rank) if (world > 1 or args.shard_tree) else None
    if args.shard_tree:
        # subtree shards (trees larger than HBM): top replicated, subtrees owned by ranks, frontier all-to-all
        tree = BloomTree.open_sharded(db_dir, local_rank, rank, world, nccl_id,
                                      cut_level=None if args.cut_level < 0 else args.cut_level)
    else:
        tree = BloomTree(db_dir, local_rank)
        if world > 1:
            _lib.check(L.pf_comm_init(tree._h, world, rank, nccl_id))
    info = tree.info
    query_device = L.pf_query_sharded_device if args.shard_tree else L.pf_query_device

    packed = PackedReads.from_concat(blob, offs)
    dev_batch = C.c_void_p()
    _lib.check(L.pf_batch_upload(tree._h, packed.batch, C.byref(dev_batch)))
    stream = torch.cuda.ExternalStream(L.pf_db_stream(tree._h), device=torch.device("cuda", local_rank))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def flush_l2():
        flush.add_(1)
        torch.cuda.synchronize()

    hits = _lib.Hits()

    def step_device():
        _lib.check(query_device(tree._h, dev_batch, C.c_float(THETA), 1, C.byref(hits)))

    def combine():
        # per-genome counts are combined with ONE NCCL reduce at the end of the run (north_star)
        if world > 1:
            _lib.check(L.pf_allreduce_counts(tree._h))

    pipe = [C.c_void_p(), C.c_void_p()]  # two device batches: upload of block i+1 overlaps the query of block i

    def run_e2e(n_steps):
        _lib.check(L.pf_batch_upload_async(tree._h, packed.batch, C.byref(pipe[0])))
        for i in range(n_steps):
            if i + 1 < n_steps:
                _lib.check(L.pf_batch_upload_async(tree._h, packed.batch, C.byref(pipe[(i + 1) % 2])))
            _lib.check(query_device(tree._h, pipe[i % 2], C.c_float(THETA), 1, C.byref(hits)))
        combine()

    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        step_device()
    combine()
    # ---- timed region: device-resident inputs, CUDA events on the launching stream -------------
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    tree.reset_stats()
    ms = 0.0
    for i in range(args.steps + 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if i < args.steps:
            flush_l2()
            e0.record(stream)
            step_device()
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
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_all, e2e_ms_all = float(tmax[0]), float(tmax[1])
    total_reads = args.reads * world * args.steps
    value = total_reads / (ms_all * 1e-3)
    e2e_value = total_reads / (e2e_ms_all * 1e-3)

    # correctness guard on the timed work: every error-free read must hit its source genome's leaf
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

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        steps = args.steps
        probes = int(st.probes_issued)
        pairs = int(st.pairs)
        read_bytes = (READ_LEN + 3) // 4
        memo_lookups, memo_hits = int(st.memo_lookups), int(st.memo_hits)
        # one 32 B sector per probe issued (and per k-mer memo look-up) + the 2-bit read per pair
        alg_bytes = 32 * (probes + memo_lookups) + read_bytes * pairs
        probe_ms = float(st.probe_kernel_ms)
        achieved = alg_bytes / (probe_ms * 1e-3) / 1e9 if probe_ms > 0 else 0.0
        l2_rate, hbm_rate = C.c_double(0), C.c_double(0)
        _lib.check(L.pf_microbench_sectors(local_rank, int(info.words_per_filter) * 8, 200, C.byref(l2_rate)))
        _lib.check(L.pf_microbench_sectors(local_rank, 8 << 30, 100, C.byref(hbm_rate)))
        probes_per_s = (probes + memo_lookups) / (probe_ms * 1e-3) if probe_ms > 0 else 0.0
        # DRAM bytes per probe launch from the committed `ncu --set full` capture of this same command
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            traffic, traffic_src = tj.get("probe_kernel_dram_bytes_per_launch"), tj.get("source")
        roofline = {
            "bound": "hbm", "kernel": f"probe_kernel<G={int(st.group_rounds)},small_m>", "achieved": achieved,
            "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
            "peak_source": peak_src,
            "note": "achieved counts one 32 B sector per bloom probe and per k-mer memo look-up; they are L2 hits by design (node-major "
                    "frontier), so achieved may exceed the HBM copy peak; the binding roof is the L2 random-sector "
                    "peak measured below",
            "algorithmic_bytes_per_launch": alg_bytes / max(int(st.probe_launches), 1),
            "launches_per_step": int(st.probe_launches) // steps,
            "avg_launch_ms": probe_ms / max(int(st.probe_launches), 1),
            "probe_share_of_step": probe_ms / float(st.device_ms) if st.device_ms else None,
            "probes_per_s": probes_per_s,
            "l2_random_sector_peak_per_s": l2_rate.value, "hbm_random_sector_peak_per_s": hbm_rate.value,
            "frac_of_l2_random_sector_peak": probes_per_s / l2_rate.value if l2_rate.value else None,
        }
        # CPU baseline beside it: the oracle port on this box's host cores, bounded sample
        n_sample, times = cpu_reference_run(genomes, blob, offs, 2000 if args.profile else CPU_SAMPLE_READS, 1, 1, db_dir)
        cores = os.cpu_count() or 1
        cpu_v = n_sample * len(times) / sum(times)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": args.warmup,
            "ms_per_step": ms_all / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic",
            "config": {"workload": workload_name(args.reads), "reads_per_gpu_per_step": args.reads,
                       "nodes": int(info.n_nodes), "leaves": int(info.n_leaves), "levels": int(info.n_levels),
                       "filter_bytes": int(info.filter_bytes), "want_hits": True,
                       "l2": "filters (358 MB) + reads exceed the 126 MB L2; a 256 MB buffer is also written "
                             "between timed steps", "parallelism": (f"reads sharded x{world}, tree cut into a replicated top and subtrees owned by ranks; "
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
            "cpu_baseline": {"value": cpu_v, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"first {n_sample} reads of the same workload, one pass, all {cores} host threads"},
            "work": {"pairs_per_step": pairs // steps, "probes_issued_per_step": probes // steps,
                     "memo_lookups_per_step": memo_lookups // steps, "memo_hits_per_step": memo_hits // steps,
                     "hits_per_step": n_hits, "db_build_s": round(build_s, 2)},
        }
        if shard:
            line["shard"] = shard
        print(json.dumps(line))
    L.pf_batch_free(tree._h, dev_batch)
    packed.close()
    tree.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads", type=int, default=1_000_000, help="reads per GPU per step")
    ap.add_argument("--shard-tree", action="store_true",
                    help="subtree-sharded tree (pf_db_open_sharded) instead of one replica per GPU")
    ap.add_argument("--cut-level", type=int, default=-1, help="cut level of the sharded tree (-1: automatic)")
    ap.add_argument("--profile", action="store_true",
                    help="for runs under ncu: shrink the CPU-baseline sample (numbers printed under a profiler are not bench values)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    return run_reference(args) if args.impl == "reference" else run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
