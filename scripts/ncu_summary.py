"""Key metrics of every kernel launch in an ncu report (ncu -i rep --page raw --csv), in the format of profiles/*/…_ncu_summary.txt.
Usage: python scripts/ncu_summary.py report.ncu-rep [> profiles/rNN/name_ncu_summary.txt]"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tensor.sum",
        "sm__cycles_elapsed.avg", "smsp__cycles_active.avg", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warp_latency_per_inst_issued.ratio", "sm__warps_active.avg.per_cycle_active", "lts__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
    print(f"== {sys.argv[1]} {r[col['Kernel Name']][:90]}")
    for k in KEYS + sorted(h for h in hdr if "issue_stalled" in h and h.endswith("per_issue_active.ratio")):
        if k in col and r[col[k]] not in ("", "0"):
            print(f"  {k:100s} {r[col[k]]:>18s} {units[col[k]]}")
