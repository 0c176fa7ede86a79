#!/usr/bin/env python
"""Print the key metrics of an ncu --set full report as a markdown table: tools/ncu_metrics.py <file.ncu-rep>"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed_pipe_xu.sum", "smsp__cycles_active.avg",
        "sm__cycles_elapsed.max", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor"]
print("| metric | " + " | ".join(f"launch {i+1}" for i in range(len(rows) - 2)) + " |")
print("|---|" + "---|" * (len(rows) - 2))
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print(f"| {w} ({units[i]}) | " + " | ".join(r[i][:60].replace("|", "/") for r in rows[2:]) + " |")
