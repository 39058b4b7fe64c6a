#!/usr/bin/env python3
"""Text summary of an `ncu --page raw --csv` export (one block per profiled launch): the metrics the roofline
discussion uses. usage: ncu_summary.py raw.csv "header line" > profiles/<name>.txt"""
import csv
import re
import sys

KEEP = [r"^dram__bytes_(read|write)\.sum$", r"^dram__cycles_active\.avg\.pct_of_peak_sustained_elapsed$",
        r"^gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed$", r"^gpu__time_duration\.sum$", r"^launch__block_size$",
        r"^launch__grid_size$", r"^launch__registers_per_thread$", r"^launch__shared_mem_per_block_(dynamic|static)$",
        r"^launch__occupancy_limit_", r"^lts__t_sector_hit_rate\.pct$", r"^lts__t_bytes\.sum$", r"^l1tex__t_sector_hit_rate\.pct$",
        r"^sm__issue_active\.avg\.pct_of_peak_sustained_elapsed$", r"^sm__warps_active\.avg\.pct_of_peak_sustained_active$",
        r"^sm__throughput\.avg\.pct_of_peak_sustained_elapsed$", r"^smsp__average_warps_issue_stalled_.*_per_issue_active\.ratio$",
        r"^smsp__inst_executed\.sum$", r"^sm__inst_executed_pipe_fp64\.", r"^l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum$",
        r"^smsp__cycles_active\.avg$", r"^derived__smsp__sass_thread_inst_executed_op_d(fma|add|mul)_pred_on"]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
if len(sys.argv) > 2:
    print(sys.argv[2])
for n, r in enumerate(data):
    if len(r) != len(hdr):
        continue
    print("\n--- launch %d: %s  block %s grid %s" % (n, r[hdr.index("Kernel Name")], r[hdr.index("Block Size")], r[hdr.index("Grid Size")]))
    for i, h in enumerate(hdr):
        name = h.split(".", 2)[-1] if h.startswith(("FBSP.", "Breakdown.")) else h
        if any(re.search(k, h) for k in KEEP) or any(re.search(k, name) for k in KEEP):
            print("%s [%s] = %s" % (h, units[i], r[i]))
