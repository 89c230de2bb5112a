#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) or a launch-list CSV into the few numbers DESIGN.md / profiles/ quote."""
import collections
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
]


def rep(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    name_i = hdr.index("Kernel Name")
    for d in data:
        print(f"== {d[name_i][:70]}")
        for k in KEYS:
            if k in hdr:
                print(f"  {k:75s} {d[hdr.index(k)]:>18s} {units[hdr.index(k)]}")
        for i, h in enumerate(hdr):
            if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and float(d[i] or 0) >= 0.15:
                print(f"  stall {h.split('issue_stalled_')[1].split('_per_issue')[0]:68s} {float(d[i]):18.2f} warps/issue")


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(list)
    for row in csv.DictReader(lines):
        if row.get("Metric Name") == "gpu__time_duration.sum":
            v = float(row["Metric Value"].replace(",", ""))
            u = row.get("Metric Unit", "ns")
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1e-3)
            agg[row["Kernel Name"][:80]].append(v)
    tot = sum(sum(v) for v in agg.values())
    print(f"{'kernel':82s} {'n':>4s} {'avg us':>10s} {'share':>7s}")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"{k:82s} {len(v):4d} {sum(v)/len(v):10.2f} {100*sum(v)/tot:6.1f}%")


if __name__ == "__main__":
    for p in sys.argv[1:]:
        print(f"# {p}")
        (launches if p.endswith(".csv") else rep)(p)
