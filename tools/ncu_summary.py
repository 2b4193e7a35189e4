"""Condense an `ncu -i x.ncu-rep --page raw --csv` export into the handful of counters the design
notes cite (one row per metric): python tools/ncu_summary.py raw.csv out.csv"""
import csv
import re
import sys

KEEP = [r"^Kernel Name$", r"^gpu__time_duration\.sum$", r"^launch__registers_per_thread$", r"^launch__grid_size$",
        r"^launch__block_size$", r"^launch__shared_mem_per_block_dynamic$", r"^launch__occupancy_limit",
        r"^sm__warps_active\.avg\.pct_of_peak_sustained_active$", r"^dram__bytes_(read|write)\.sum$",
        r"^smsp__inst_executed\.sum$", r"^smsp__issue_active\.avg\.per_cycle_active$",
        r"^sm__inst_executed_pipe_(fp64|lsu|alu|fma|xu|tensor_subpipe_dmma)\.avg\.pct_of_peak_sustained_active$",
        r"^sm__pipe_(fp64|tensor)_cycles_active\.avg\.pct_of_peak_sustained_active$",
        r"^sm__ops_path_tensor_src_fp64\.avg\.pct_of_peak_sustained_elapsed$",
        r"^sass__inst_executed_(shared|local|global)_(loads|stores)$", r"^l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum$",
        r"^smsp__average_warps_issue_stalled_.*_per_issue_active\.ratio$", r"^memory_l2_theoretical_sectors_local$",
        r"^smsp__inst_executed\.(max|min)$", r"^lts__t_sector_hit_rate\.pct$"]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
with open(sys.argv[2], "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["metric", "unit"] + [f"launch{i}" for i in range(len(rows) - 2)])
    for j, h in enumerate(hdr):
        if any(re.search(p, h) for p in KEEP):
            vals = [r[j] for r in rows[2:]]
            if all(v in ("0", "0.000000", "") for v in vals) and "stalled" in h:
                continue
            w.writerow([h, units[j]] + vals)
