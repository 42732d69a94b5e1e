#!/usr/bin/env python
"""Condense an ncu report (--set full) into the small CSV kept under profiles/: one row per captured launch of the
kernels matching a name prefix, the metrics the bench lines and DESIGN.md quote.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep k_tile "comment line" > profiles/r2_x_ncu_k_tile.csv
"""
import csv
import subprocess
import sys

METRICS = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "gpu__time_duration.sum",
           "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
           "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.sum",
           "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
           "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "lts__t_sector_hit_rate.pct"]


def main():
	rep, prefix = sys.argv[1], sys.argv[2]
	comment = sys.argv[3] if len(sys.argv) > 3 else ""
	raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
	rows = list(csv.reader(raw.splitlines()))
	hdr, units = rows[0], rows[1]
	keep = [i for i, h in enumerate(hdr) if h in ("ID", "Kernel Name", "Block Size", "Grid Size") or h in METRICS]
	w = csv.writer(sys.stdout, quoting=csv.QUOTE_ALL)
	if comment:
		print("# " + comment)
	w.writerow([hdr[i] for i in keep])
	w.writerow([units[i] for i in keep])
	for r in rows[2:]:
		if r[hdr.index("Kernel Name")].startswith(prefix):
			w.writerow([r[i] for i in keep])


if __name__ == '__main__':
	main()
