#!/usr/bin/env python
"""Summarise `ncu -i X.ncu-rep --page raw --csv` output: the metrics DESIGN.md quotes, one line per kernel launch.
    python tools/ncu_summary.py raw.csv [--traffic-json out.json --kernel-note "..." --source "..."]
"""
import csv
import json
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("sm__inst_executed_pipe_tensor", "tensor pipe"),
    ("sm__pipe_tensor", "tensor pipe"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm throughput %"),
    ("sm__memory_throughput", "sm memory throughput"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared", "smem wavefronts"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__cycles_elapsed.avg", "sm cycles"),
    ("launch__registers_per_thread", "registers/thread"),
    ("smsp__pcsamp", None),
]


def main():
    path = sys.argv[1]
    rows = list(csv.reader(open(path, newline="")))
    hdr_i = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[hdr_i]
    units = rows[hdr_i + 1] if len(rows) > hdr_i + 1 else [""] * len(hdr)
    out = {}
    for r in rows[hdr_i + 2:]:
        if len(r) != len(hdr):
            continue
        d = dict(zip(hdr, r))
        print("kernel:", d["Kernel Name"][:90], "grid", d.get("Grid Size"), "block", d.get("Block Size"))
        for col, u in zip(hdr, units):
            if any(k in col for k, _ in KEYS if _):
                print(f"   {col:110s} {d[col]:>18s} {u}")
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        for col, u in zip(hdr, units):
            if col.endswith("dram__bytes_read.sum"):
                out["dram_bytes_read"] = float(d[col].replace(",", "")) * scale.get(u, 1.0)
            if col.endswith("dram__bytes_write.sum"):
                out["dram_bytes_write"] = float(d[col].replace(",", "")) * scale.get(u, 1.0)
        if "--traffic-json" in sys.argv and "dram_bytes_read" in out:
            out["dram_bytes_per_launch"] = out["dram_bytes_read"] + out["dram_bytes_write"]
            if "--kernel-note" in sys.argv:
                out["kernel"] = sys.argv[sys.argv.index("--kernel-note") + 1]
            if "--source" in sys.argv:
                out["source"] = sys.argv[sys.argv.index("--source") + 1]
            json.dump(out, open(sys.argv[sys.argv.index("--traffic-json") + 1], "w"), indent=1)
            break


if __name__ == "__main__":
    main()
