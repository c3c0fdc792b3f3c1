#!/usr/bin/env python
"""Summarise ncu launch lists (gpu__time_duration.sum + dram__bytes_read/write.sum, --csv) into
profiles/r2_ncu_traffic.json: DRAM bytes per launch of each libssb kernel (single-plan run) and DRAM bytes of one
update_once step of the chunked run (four chunk plans on four streams, the engine path bench.py times).

    python tools/ncu_traffic_summary.py gpurun_out/r2_ncu_launches_c2.csv gpurun_out/r2_ncu_launches_c2_chunked.csv 3
"""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# ncu kernel name -> the label libssb's own profiler (ssb_profile_begin/end) reports in bench.py
LABELS = [("kf_basis_coop", "coop_basis"), ("kf_activation_coop", "coop_activation"), ("kf_phi_cov", "fused_phi_cov"),
          ("kf_cov_coop", "coop_phi_cov"), ("kc_cov_mma8", "mma_phi_cov"), ("kf_normalize", "fused_normalize"),
          ("kf_ip1_n2", "fused_ip1_n2"), ("kq_ip1", "update_by_ip1"), ("kq_ip2", "update_by_ip2"),
          ("kf_vsplit", "coop_vsplit"), ("kt_tile", "tma_tile")]


def read(path):
    rows = {}
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        e = rows.setdefault(int(r["ID"]), {"name": r["Kernel Name"]})
        e[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
        e["unit:" + r["Metric Name"]] = r["Metric Unit"]
    out = []
    for i in sorted(rows):
        e = rows[i]
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        b = sum(e.get(m, 0.0) * scale.get(e.get("unit:" + m, "byte"), 1.0)
                for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        tscale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "usecond": 1e-3, "nsecond": 1e-6, "msecond": 1.0}
        t = e.get("gpu__time_duration.sum", 0.0) * tscale.get(e.get("unit:gpu__time_duration.sum", "ns"), 1e-6)
        out.append((e["name"], b, t))
    return out


def label(name):
    for pat, lab in LABELS:
        if re.search(r"\b%s\b" % pat, name) or pat in name:
            return lab
    return None


def is_ssb(name):
    return "<unnamed>::k" in name or name.startswith("k") or "ssb" in name


def main():
    single, chunked, steps = sys.argv[1], sys.argv[2], int(sys.argv[3])
    per = {}
    for name, b, t in read(single):
        lab = label(name)
        if lab:
            per.setdefault(lab, []).append((b, t))
    per_launch = {k: sum(x[0] for x in v) / len(v) for k, v in per.items()}
    per_ms = {k: sum(x[1] for x in v) / len(v) for k, v in per.items()}
    # chunked run: every kernel of the steady-state iterations (the preparation of the plan runs once, before them)
    it_kernels = ("kf_basis_coop", "kf_activation_coop", "kf_phi_cov", "kf_cov_coop", "kc_cov_mma8", "kf_normalize",
                  "kf_ip1_n2", "kq_ip1", "kq_ip2", "kt_tile")
    tot = sum(b for name, b, t in read(chunked) if any(k in name for k in it_kernels))
    tsum = sum(t for name, b, t in read(chunked) if any(k in name for k in it_kernels))
    out = {"workload": ["GaussILRMA", "IP", 2, 1025, 512, 16, 64],
           "source": [os.path.basename(single), os.path.basename(chunked)],
           "note": "ncu --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum "
                   "python tools/ncu_target.py --steps %d [--chunked]; per-launch values are means over the launches of "
                   "the single-plan run (one launch covers the whole 64-mixture batch); the per-step figure sums the "
                   "iteration kernels of the chunked run (4 chunk plans) and divides by the steps" % steps,
           "dram_bytes_per_launch": per_launch, "ncu_ms_per_launch": per_ms,
           "dram_bytes_per_step_chunked": tot / steps, "ncu_serialised_ms_per_step_chunked": tsum / steps}
    with open(os.path.join(ROOT, "profiles", "r2_ncu_traffic.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
