#!/usr/bin/env python
"""One-call profiling of the bench step (cfg 2) and the summaries that go under profiles/.

On the GPU box (one gpurun call of about 4 minutes -- the full-set pass alone replays ten launches ~40 times each and
every ncu pass is preceded by the pool's own plain run; the recipe is /opt/skills/guides/B200_PROFILING.md):
    python scripts/profile_round.py run --tag r2a
  1. plain `python bench.py --steps 5 --warmup 3`                      -> gpurun_out/<tag>/bench.json
  2. plain `python bench.py --steps 1 --warmup 3 --profile` (must exit 0 before anything runs under ncu)
  3. the same command under `ncu --metrics gpu__time_duration.sum`     -> gpurun_out/<tag>/launches_raw.csv
  4. the same command under `ncu --set full` for the hash + probe launches of the TIMED step only
     (launch-skip/-count derived from pass 3), raw page exported on the box -> gpurun_out/<tag>/full_raw.csv
Here (no GPU):
    python scripts/profile_round.py summarise --tag r2a
  writes profiles/<tag>_launches_bench_steps1.csv, profiles/<tag>_ncu_full_summary.csv, profiles/bench_<tag>.json and
  profiles/traffic.json (mean DRAM bytes per probe launch, read by bench.py for roofline.traffic), and prints the
  kernel shares of the step.

With `--steps 1 --warmup 3` the process runs 4 device steps (3 warm-up + the timed one) and then 3 end-to-end steps;
every step starts with one hash_kernel launch, so the timed step is the 4th hash_kernel launch up to the 5th.
"""
import argparse
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CMD = [sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "3", "--profile"]
CONFIG = "cfg2"
HOT = ("hash_kernel", "probe_kernel", "sliced_entry_quad_kernel")  # the kernels the full-set pass captures
TIMED_STEP = 4  # 1-based index of the timed step among the hash_kernel launches
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "smsp__inst_executed.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__cycles_elapsed.max", "smsp__sass_average_branch_targets_threads_uniform.pct"]


def read_launch_list(path):
    """[(id, kernel name, ns)] from an `ncu --csv --log-file` launch list (lines before the header are ncu chatter)."""
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    rd = csv.reader(lines)
    header = next(rd)
    i_id, i_name, i_metric, i_val = header.index("ID"), header.index("Kernel Name"), header.index("Metric Name"), header.index("Metric Value")
    for r in rd:
        if r[i_metric] == "gpu__time_duration.sum":
            rows.append((int(r[i_id]), r[i_name], float(r[i_val].replace(",", ""))))
    return rows


def step_marker(rows):
    """The kernel a step starts with: the entry line kernel when it hashes on the fly (hash_kernel then runs in the middle
    of the step, for the surviving reads only), else hash_kernel."""
    return "sliced_entry_quad_kernel" if any("sliced_entry_quad_kernel" in r[1] for r in rows) else "hash_kernel"


def timed_step(rows):
    marker = step_marker(rows)
    hashes = [i for i, r in enumerate(rows) if marker in r[1]]
    if len(hashes) < TIMED_STEP:
        raise SystemExit(f"expected at least {TIMED_STEP} {marker} launches, found {len(hashes)}")
    lo = hashes[TIMED_STEP - 1]
    hi = hashes[TIMED_STEP] if len(hashes) > TIMED_STEP else len(rows)
    return [r for r in rows[lo:hi] if "pf::" in r[1] or any(k in r[1] for k in HOT)]


def cmd_run(tag, no_bench=False, launches_only=False):
    out = os.path.join(ROOT, "gpurun_out", tag)
    os.makedirs(out, exist_ok=True)
    if not no_bench:
        with open(os.path.join(out, "bench.json"), "w") as f:
            subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "5", "--warmup", "3", *CMD[CMD.index("--profile") + 1:]],
                           check=True, stdout=f)
    subprocess.run(CMD, check=True, stdout=subprocess.DEVNULL)  # must exit 0 without ncu first
    raw = os.path.join(out, "launches_raw.csv")
    subprocess.run(["ncu", "--metrics", "gpu__time_duration.sum", "--clock-control", "none", "--csv", "--log-file", raw, *CMD],
                   check=True, stdout=subprocess.DEVNULL)
    rows = read_launch_list(raw)
    if launches_only:
        print("launch list only:", len(rows), "launches")
        return
    sel = [r for r in rows if any(k in r[1] for k in HOT)]
    step = [r for r in timed_step(rows) if any(k in r[1] for k in HOT)]
    skip = sel.index(step[0])
    rep = os.path.join(out, "full")
    subprocess.run(["ncu", "--set", "full", "--clock-control", "none", "--import-source", "on", "-k", "regex:" + "|".join(HOT),
                    "-s", str(skip), "-c", str(len(step)), "-f", "-o", rep, *CMD], check=True, stdout=subprocess.DEVNULL)
    with open(os.path.join(out, "full_raw.csv"), "w") as f:
        subprocess.run(["ncu", "-i", rep + ".ncu-rep", "--page", "raw", "--csv"], check=True, stdout=f)
    if os.path.getsize(rep + ".ncu-rep") > (40 << 20):
        os.remove(rep + ".ncu-rep")  # gpurun_out/ returns at most 64 MiB; the raw page is what the summary needs
    print("profiled", len(step), "launches of the timed step; skip =", skip)


def cmd_summarise(tag):
    src = os.path.join(ROOT, "gpurun_out", tag)
    prof = os.path.join(ROOT, "profiles")
    rows = timed_step(read_launch_list(os.path.join(src, "launches_raw.csv")))
    with open(os.path.join(prof, f"{tag}_launches_bench_steps1.csv"), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["ID", "Kernel Name", "gpu__time_duration.sum [ns]"])
        for r in rows:
            w.writerow([r[0], r[1], int(r[2])])
    total = sum(r[2] for r in rows)
    share = {}
    for _, name, ns in rows:
        key = name.split("(")[0].replace("void ", "").replace("pf::", "").split("<")[0]
        share[key] = share.get(key, 0.0) + ns
    print(f"{len(rows)} launches, {total / 1e6:.2f} ms of kernels in the timed step")
    for k, v in sorted(share.items(), key=lambda kv: -kv[1]):
        print(f"  {k:28s} {v / 1e6:7.3f} ms  {100 * v / total:5.1f} %")
    probe = [r[2] / 1e6 for r in rows if "probe_kernel" in r[1]]  # also matches sliced_probe_kernel
    print("  probe levels (ms):", ", ".join(f"{x:.2f}" for x in probe))
    # full-set summary: metrics as rows, launches as columns
    if not os.path.exists(os.path.join(src, "full_raw.csv")):
        print("no full_raw.csv (the full-set pass did not finish): launch list only")
        return
    with open(os.path.join(src, "full_raw.csv"), newline="") as f:
        rd = list(csv.reader(ln for ln in f if ln.startswith('"')))
    header, units, launches = rd[0], rd[1], rd[2:]
    col = {name: i for i, name in enumerate(header)}
    keep = [m for m in KEEP if m in col] + sorted(m for m in col if "issue_stalled" in m and m.endswith("per_issue_active.ratio"))
    names, lvl = [], 0
    for r in launches:
        kn = r[col["Kernel Name"]]
        if "hash_kernel" in kn:
            names.append(kn.split("(")[0].replace("void ", ""))
        else:
            names.append(kn.split("(")[0].replace("void ", "") + f" level {lvl}")
            lvl += 1
    with open(os.path.join(prof, f"{tag}_ncu_full_summary.csv"), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit", *names])
        w.writerow(["Kernel Name", "", *[r[col["Kernel Name"]] for r in launches]])
        for m in keep:
            w.writerow([m, units[col[m]], *[r[col[m]] for r in launches]])

    def as_bytes(r, m):
        v, u = float(r[col[m]].replace(",", "")), units[col[m]].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
    tpath = os.path.join(prof, "traffic.json")
    tj = json.load(open(tpath)) if os.path.exists(tpath) else {}
    if "probe_kernel_dram_bytes_per_launch" in tj:  # round-1 layout: one flat entry for cfg2
        tj = {"cfg2": tj}
    entry = {}
    for key in ("sliced_entry_quad_kernel", "sliced_probe_kernel", "probe_kernel"):
        pl = [r for r in launches if key in r[col["Kernel Name"]] and (key != "probe_kernel" or "sliced" not in r[col["Kernel Name"]])]
        if pl and "dram__bytes_read.sum" in col:
            mean = sum(as_bytes(r, "dram__bytes_read.sum") + as_bytes(r, "dram__bytes_write.sum") for r in pl) / len(pl)
            entry[key + "_dram_bytes_per_launch"] = mean
            entry[key + "_launches"] = len(pl)
            print(f"  DRAM traffic per {key} launch: {mean / 1e9:.2f} GB over {len(pl)} launches")
    if entry:
        forced = os.environ.get("PF_SLICED_FORCE_G")
        entry["source"] = (f"profiles/{tag}_ncu_full_summary.csv (ncu --set full on bench.py --config {CONFIG}"
                           + (f" with PF_SLICED_FORCE_G={forced}" if forced else "") + ", dram__bytes_read.sum + "
                           "dram__bytes_write.sum, mean over that kernel's launches of one step)")
        tj[CONFIG] = entry
        json.dump(tj, open(tpath, "w"), indent=1)
    if os.path.exists(os.path.join(src, "bench.json")):
        shutil.copy(os.path.join(src, "bench.json"), os.path.join(prof, f"bench_{tag}.json"))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("mode", choices=["run", "summarise"])
    ap.add_argument("--tag", required=True)
    ap.add_argument("--no-bench", action="store_true", help="run: skip the plain 5-step bench line")
    ap.add_argument("--config", default="cfg2", help="bench.py --config to profile")
    ap.add_argument("--launches-only", action="store_true", help="run: stop after the launch list (no full-set pass)")
    ap.add_argument("--mode", dest="mode_arg", type=int, default=None, help="bench.py --mode (pf_db_set_mode) for every run of this call")
    a = ap.parse_args()
    CONFIG = a.config
    CMD += ["--config", CONFIG]
    if a.mode_arg is not None:
        CMD += ["--mode", str(a.mode_arg)]
    if a.mode == "run":
        cmd_run(a.tag, a.no_bench, a.launches_only)
    else:
        cmd_summarise(a.tag)
