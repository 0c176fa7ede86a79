#!/usr/bin/env python
"""Summarise `ncu --set full` captures (.ncu-rep) into a markdown table: one row per captured launch with the metrics the
roofline discussion uses.  Runs on the CPU box (ncu -i ... --page raw --csv).

    python tools/ncu_full_summary.py gpurun_out/prof_a.ncu-rep gpurun_out/prof_b.ncu-rep > profiles/r02_full.md
"""
import csv
import io
import re
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "time"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe % active"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("lts__t_sectors_srcunit_tex_op_read.sum", "L2->SM read sectors"),
    ("launch__waves_per_multiprocessor", "waves/SM"),
    ("launch__registers_per_thread", "regs"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
]


def short(name):
    name = re.sub(r"^void\s+", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name.replace("tg::", "")


def main():
    print("# `ncu --set full --clock-control none` captures of the launches that set the headline (tools/profile_r02*.sh)\n")
    print("Launches inside bench.py's timed region (`--nvtx-include tg_timed/`); per-launch times under ncu are cold-cache and "
          "serialised.  One row per captured launch.\n")
    hdr = ["kernel", "grid"] + [m[1] for m in METRICS]
    print("| " + " | ".join(hdr) + " |")
    print("|" + "---|" * len(hdr))
    for path in sys.argv[1:]:
        raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        if len(rows) < 3:
            continue
        h, units = rows[0], rows[1]
        for r in rows[2:]:
            cells = [f"`{short(r[h.index('Kernel Name')])}`", r[h.index("Grid Size")]]
            for m, _ in METRICS:
                if m in h:
                    i = h.index(m)
                    v = r[i]
                    try:
                        v = f"{float(v.replace(',', '')):.4g}"
                    except ValueError:
                        pass
                    cells.append(f"{v} {units[i]}".strip())
                else:
                    cells.append("")
            print("| " + " | ".join(cells) + " |")


if __name__ == "__main__":
    main()
