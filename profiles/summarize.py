#!/usr/bin/env python
"""Summaries of ncu output for profiles/ (run here, on files brought back in gpurun_out/).

  python profiles/summarize.py launches <launches.csv>            # per-kernel share of a step
  python profiles/summarize.py raw <report.ncu-rep> [kernel-substr]  # key counters of one capture
"""
import collections
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
    "l1tex__t_bytes_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_bytes_pipe_lsu_mem_local_op_st.sum",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
]


def launches(path):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    agg = collections.OrderedDict()
    for r in rows[1:]:
        d = dict(zip(hdr, r))
        k = d["Kernel Name"].split("(")[0]
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += float(d["Metric Value"])
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: {sum(v[0] for v in agg.values())} launches, {tot / 1e6:.3f} ms total (ncu: cold-cache, serialised)")
    print(f"{'kernel':44s} {'n':>5s} {'total_ms':>10s} {'avg_us':>9s} {'share':>7s}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:44s} {v[0]:5d} {v[1] / 1e6:10.3f} {v[1] / v[0] / 1e3:9.1f} {v[1] / tot * 100:6.1f}%")


def raw(path, substr=""):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(l for l in out.splitlines() if l.startswith('"')))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        if substr and substr not in d["Kernel Name"]:
            continue
        print(f"## {d['Kernel Name'][:90]}  grid={d.get('Grid Size')} block={d.get('Block Size')}")
        for k in KEYS:
            if k in d:
                print(f"  {k:62s} {d[k]:>18s} {units[hdr.index(k)]}")
        stalls = {h[len('smsp__pcsamp_warps_issue_stalled_'):]: float(d[h]) for h in hdr
                  if h.startswith("smsp__pcsamp_warps_issue_stalled_") and "not_issued" not in h and d[h]}
        tot = sum(stalls.values()) or 1.0
        top = sorted(stalls.items(), key=lambda kv: -kv[1])[:6]
        print("  stall samples: " + ", ".join(f"{k} {v / tot * 100:.0f}%" for k, v in top))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        raw(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "")
