#!/usr/bin/env python
"""gpurun_out/gemm_dram_<tag>.csv (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum over the
GEMM launches of bench.py's timed region) -> profiles/gemm_traffic.json (mean DRAM bytes per launch; bench.py puts it
into roofline.traffic) and a markdown summary on stdout."""
import csv
import json
import sys
from collections import defaultdict

path, tag = sys.argv[1], sys.argv[2]
unit_scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6}
per = defaultdict(dict)
with open(path, newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if "gemm_tc_kernel" not in r["Kernel Name"]:  # the launch list may hold every kernel of the timed region
        continue
    v = float(r["Metric Value"].replace(",", "")) * unit_scale.get(r.get("Metric Unit", ""), 1)
    per[r["ID"]][r["Metric Name"]] = v
    per[r["ID"]]["name"] = r["Kernel Name"].split("(")[0].replace("void ", "")
n = len(per)
rd = sum(p.get("dram__bytes_read.sum", 0) for p in per.values())
wr = sum(p.get("dram__bytes_write.sum", 0) for p in per.values())
ns = sum(p.get("gpu__time_duration.sum", 0) for p in per.values())
out = {"mean_dram_bytes_per_launch": round((rd + wr) / n), "launches": n, "dram_read_bytes": rd, "dram_write_bytes": wr,
       "source": f"ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum over the {n} gemm_tc_kernel launches of bench.py's timed region ({tag}; cold-cache, serialised replays)"}
json.dump(out, open("profiles/gemm_traffic.json", "w"), indent=1)
print(f"# {tag} -- DRAM traffic of the GEMM family over the timed region of bench.py (2 steps)\n")
print(f"{n} launches, {rd / 1e9:.3f} GB read + {wr / 1e9:.3f} GB written = {(rd + wr) / n / 1e6:.2f} MB per launch; {ns / 1e6:.2f} ms serialised.\n")
by = defaultdict(lambda: [0, 0.0, 0.0])
for p in per.values():
    a = by[p["name"]]
    a[0] += 1; a[1] += p.get("dram__bytes_read.sum", 0) + p.get("dram__bytes_write.sum", 0); a[2] += p.get("gpu__time_duration.sum", 0)
print("| kernel | launches | DRAM MB / launch | avg us | DRAM GB/s |\n|---|---|---|---|---|")
for k, (c, b, t) in sorted(by.items(), key=lambda kv: -kv[1][2]):
    print(f"| {k} | {c} | {b / c / 1e6:.2f} | {t / c / 1e3:.1f} | {b / t:.0f} |")
