"""Selected metrics of every kernel in an .ncu-rep as a small CSV (what profiles/*_ncu_full.csv hold):
    python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/r2_top_kernels_ncu_full.csv
"""
import csv, io, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__shared_mem_per_block_static", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "launch__waves_per_multiprocessor", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"]

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ki = hdr.index("Kernel Name")
with open(out, "w", newline="") as fh:
    w = csv.writer(fh)
    w.writerow(["metric", "unit"] + [r[ki].split("(")[0] for r in rows[2:]])
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            w.writerow([k, units[i]] + [r[i] for r in rows[2:]])
print(open(out).read()[:600])
