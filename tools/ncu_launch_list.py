#!/usr/bin/env python
"""Compact an ncu launch list (--csv, long format: one row per launch and metric) into one row per launch of a libssb
kernel: id, kernel, time_us, dram_read_MB, dram_write_MB [, warp_instructions].

    python tools/ncu_launch_list.py gpurun_out/r2_ncu_launches_c2.csv > profiles/r2_ncu_launches_c2.csv
"""
import csv
import re
import sys

SC = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3,
      "usecond": 1.0, "msecond": 1e3, "inst": 1.0}
rows = {}
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    e = rows.setdefault(int(r["ID"]), {"name": r["Kernel Name"], "grid": r["Grid Size"], "block": r["Block Size"]})
    e[r["Metric Name"]] = float(r["Metric Value"].replace(",", "")) * SC.get(r["Metric Unit"], 1.0)
w = csv.writer(sys.stdout)
w.writerow(["id", "kernel", "grid", "block", "time_us", "dram_read_MB", "dram_write_MB", "warp_instructions"])
for i in sorted(rows):
    e = rows[i]
    n = e["name"]
    if "at::" in n or "cutlass" in n or "elementwise" in n or "gemm" in n.lower():
        continue  # torch's own kernels of the synthetic-data setup
    short = re.sub(r"\(.*", "", n).replace("void ", "").replace("<unnamed>::", "")
    w.writerow([i, short, e["grid"], e["block"], round(e.get("gpu__time_duration.sum", 0.0), 2),
                round(e.get("dram__bytes_read.sum", 0.0) / 1e6, 2), round(e.get("dram__bytes_write.sum", 0.0) / 1e6, 2),
                int(e["smsp__inst_executed.sum"]) if "smsp__inst_executed.sum" in e else ""])
