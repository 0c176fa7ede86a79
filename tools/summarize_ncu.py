#!/usr/bin/env python
"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) into a per-kernel share table.

    python tools/summarize_ncu.py gpurun_out/launches_<tag>.csv [--by-grid] > profiles/<name>.md
Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.
"""
import csv
import re
import sys
from collections import defaultdict


def short(name):
    name = re.sub(r"^void\s+", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name.replace("tg::", "")


def main():
    path = sys.argv[1]
    by_grid = "--by-grid" in sys.argv
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        val = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ns = val * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
        rows.append((short(r["Kernel Name"]), r.get("Grid Size", ""), r.get("Block Size", ""), ns))
    tot = sum(r[3] for r in rows)
    agg = defaultdict(lambda: [0, 0.0])
    for k, grid, blk, ns in rows:
        key = (k, grid) if by_grid else (k,)
        agg[key][0] += 1
        agg[key][1] += ns
    print(f"launches: {len(rows)}   total device time (serialised, cold cache): {tot / 1e6:.3f} ms\n")
    hdr = "| kernel |" + (" grid |" if by_grid else "") + " launches | total us | avg us | share |"
    print(hdr)
    print("|" + "---|" * (hdr.count("|") - 1))
    for key, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        cols = [key[0]] + ([key[1]] if by_grid else []) + [str(n), f"{ns / 1e3:.1f}", f"{ns / 1e3 / n:.2f}",
                                                            f"{100 * ns / tot:.1f}%"]
        print("| " + " | ".join(cols) + " |")


if __name__ == "__main__":
    main()
