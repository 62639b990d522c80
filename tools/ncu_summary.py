#!/usr/bin/env python
"""Summarise ncu outputs into profiles/: per-kernel launch shares from a launch list CSV and the key
raw metrics of a --set full capture (.ncu-rep, read with `ncu -i … --page raw --csv`)."""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "l1tex__t_set_accesses_pipe_lsu_mem_global_op_red.sum", "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum"]


def launches(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        us = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
        a = agg[row["Kernel Name"].split("(")[0]]
        a[0] += 1
        a[1] += us
        a[2] = max(a[2], us)
    tot = sum(a[1] for a in agg.values())
    out = [f"# launch list {path}: {sum(a[0] for a in agg.values())} launches, {tot / 1e3:.3f} ms total (cold-cache, serialised: compare shares)",
           f"{'kernel':46s} {'n':>5s} {'total_ms':>10s} {'max_ms':>9s} {'share':>7s}"]
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        out.append(f"{k[:46]:46s} {a[0]:5d} {a[1] / 1e3:10.3f} {a[2] / 1e3:9.3f} {a[1] / tot * 100:6.1f}%")
    return "\n".join(out)


def full(path):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(txt.splitlines()))
    hdr, units, vals = r[0], r[1], r[2]
    out = [f"# ncu --set full {path}: {vals[hdr.index('Kernel Name')]}"]
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            out.append(f"{k:90s} {vals[i]:>16s} {units[i]}")
    return "\n".join(out)


if __name__ == "__main__":
    for p in sys.argv[1:]:
        print(launches(p) if p.endswith(".csv") else full(p))
        print()
