#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into the handful of numbers DESIGN.md / bench.py quote.
Usage: tools/ncu_summary.py <report.ncu-rep> [> profiles/xxx.txt]"""
import csv, io, subprocess, sys

KEYS = [
  "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
  "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
  "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
  "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active",
  "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
  "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
  "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
  "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
  "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
  "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
  "smsp__average_warp_latency_issue_stalled_barrier.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
  "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
  "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
  "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
  "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
  "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
  "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
  "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
  "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
  "smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio",
]

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
print(f"# ncu --set full summary of {rep} (values per launch; cold caches, serialised replays)")
for r in rows[2:]:
  d = dict(zip(hdr, r))
  print(f"\n## {d['Kernel Name']}   grid {d.get('Grid Size')} block {d.get('Block Size')}")
  for k in KEYS:
    if k in d and d[k] != "":
      print(f"{k:85s} {d[k]:>18s} {units[hdr.index(k)]}")
  try:
    t = float(d["gpu__time_duration.sum"]); tu = units[hdr.index("gpu__time_duration.sum")]
    t_s = t * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}[tu]
    def b(k):
      v = float(d[k]); u = units[hdr.index(k)]
      return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    tr = b("dram__bytes_read.sum") + b("dram__bytes_write.sum")
    print(f"{'derived: dram traffic (read+write) per launch':85s} {tr/1e6:18.3f} MB")
    print(f"{'derived: dram GB/s over the launch':85s} {tr/t_s/1e9:18.1f} GB/s")
  except Exception as e:
    print("derived: n/a", e)
