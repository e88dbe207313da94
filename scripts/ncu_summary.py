#!/usr/bin/env python
"""Prints the handful of ncu metrics we track from a .ncu-rep (run where ncu is installed)."""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "lts__t_sector_hit_rate.pct", "launch__waves_per_multiprocessor"]


def main(path):
  out = subprocess.check_output(["ncu", "-i", path, "--page", "raw", "--csv"]).decode()
  rows = list(csv.reader(io.StringIO(out)))
  hdr, units, data = rows[0], rows[1], rows[2:]
  for r in data:
    d = dict(zip(hdr, r))
    print("== %s  grid %s block %s" % (d.get("Kernel Name", "?")[:80], d.get("Grid Size"), d.get("Block Size")))
    for k in KEYS:
      if k in d:
        print("   %-75s %s %s" % (k, d[k], units[hdr.index(k)]))


if __name__ == "__main__":
  for p in sys.argv[1:]:
    main(p)
