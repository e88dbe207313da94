#!/usr/bin/env python
"""Prints the handful of ncu metrics we track from a .ncu-rep (run where ncu is installed)."""
import csv
import io
import os
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


UNIT_BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}

# bench.py stage -> substring of the kernel that dominates it
STAGE_KERNELS = {"segment_sum+apply": "apply_plan_kernel", "gather": "expand_plan_kernel",
                 "unique": "unique_insert_kernel"}


def main(path, traffic):
  out = subprocess.check_output(["ncu", "-i", path, "--page", "raw", "--csv"]).decode()
  rows = list(csv.reader(io.StringIO(out)))
  hdr, units, data = rows[0], rows[1], rows[2:]
  for r in data:
    d = dict(zip(hdr, r))
    name = d.get("Kernel Name", "?")
    print("== %s  grid %s block %s" % (name[:80], d.get("Grid Size"), d.get("Block Size")))
    for k in KEYS:
      if k in d:
        print("   %-75s %s %s" % (k, d[k], units[hdr.index(k)]))
    try:
      tot = 0.0
      for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        tot += float(d[k].replace(",", "")) * UNIT_BYTES[units[hdr.index(k)]]
      for stage, sub in STAGE_KERNELS.items():
        if sub in name:
          traffic[stage] = {"name": name.split("(")[0][-70:], "dram_bytes": tot,
                            "duration_us": float(d["gpu__time_duration.sum"].replace(",", "")) *
                            {"us": 1.0, "ns": 1e-3, "ms": 1e3}.get(units[hdr.index("gpu__time_duration.sum")], 1.0),
                            "report": os.path.basename(path)}
    except (KeyError, ValueError):
      pass


if __name__ == "__main__":
  import json
  import os
  args = [a for a in sys.argv[1:] if not a.startswith("--")]
  traffic = {}
  for p in args:
    main(p, traffic)
  if "--json" in sys.argv:
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    try:
      git = subprocess.check_output(["git", "-C", root, "rev-parse", "--short", "HEAD"]).decode().strip()
    except Exception:
      git = None
    dst = os.path.join(root, "profiles", "ncu_traffic.json")
    json.dump({"git": git, "how": "ncu --set full --clock-control none, one launch per kernel, "
               "dram__bytes_read.sum + dram__bytes_write.sum", "kernels": traffic},
              open(dst, "w"), indent=1)
    print("wrote", dst)
