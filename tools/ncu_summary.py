#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU): tools/ncu_summary.py gpurun_out/prof.ncu-rep [kernel-regex]"""
import csv, io, re, subprocess, sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "launch__waves_per_multiprocessor",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "l1tex__t_bytes.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.max",
]


def main():
    rep = sys.argv[1]
    pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
    out = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], stderr=subprocess.DEVNULL).decode()
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for row in rows[2:]:
        name = row[idx["Kernel Name"]]
        if pat and not pat.search(name):
            continue
        print("=== %s  (id %s)" % (name[:100], row[idx["ID"]]))
        for k in KEYS:
            if k in idx:
                print("  %-88s %-14s %s" % (k, units[idx[k]], row[idx[k]]))
        stalls = [(float(row[i].replace(",", "") or 0), h) for h, i in idx.items()
                  if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio") and row[i]]
        if not stalls:
            stalls = [(float(row[i].replace(",", "") or 0), h) for h, i in idx.items()
                      if "issue_stalled" in h and h.endswith(".ratio") and row[i]]
        for v, h in sorted(stalls, reverse=True)[:10]:
            print("  stall %-82s %.3f" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("smsp__average_warp_latency_issue_stalled_", ""), v))


if __name__ == "__main__":
    main()
