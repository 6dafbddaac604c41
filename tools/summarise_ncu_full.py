#!/usr/bin/env python
"""Per-launch table of the metrics that matter from an `ncu --set full` report.
usage: python tools/summarise_ncu_full.py gpurun_out/x.ncu-rep "<command>" [--traffic-json profiles/ncu_traffic.json]
       > profiles/rNN_ncu_full.md
--traffic-json also writes, per kernel name, the DRAM bytes (read + write) of its LONGEST launch: bench.py's
roofline.traffic reads that file (it never carries a constant of its own)."""
import csv
import json
import os
import subprocess
import sys

COLS = [("gpu__time_duration.sum", "time us", 1.0), ("dram__bytes_read.sum", "DRAM rd MB", 1e3),
        ("dram__bytes_write.sum", "DRAM wr MB", 1e3), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %", 1.0),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe %", 1.0),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %", 1.0),
        ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2->SM MB", 1e3), ("lts__t_sector_hit_rate.pct", "L2 hit %", 1.0),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %", 1.0), ("launch__registers_per_thread", "regs", 1.0)]


def main():
    rep, cmd = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else "")
    tj = sys.argv[sys.argv.index("--traffic-json") + 1] if "--traffic-json" in sys.argv else None
    traffic = {}
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader([l for l in raw.splitlines() if not l.startswith("==")]))
    head, units, data = rows[0], rows[1], rows[2:]
    ix = {n: i for i, n in enumerate(head)}
    print(f"# ncu --set full, per launch\n\ncommand: `{cmd}`\n")
    print("DRAM bytes are `dram__bytes_read.sum` / `dram__bytes_write.sum` per launch (the `traffic` of bench.py's roofline);"
          " times are cold-cache and serialised under the profiler.\n")
    print("| kernel | grid x block | " + " | ".join(c[1] for c in COLS) + " |\n|---|---|" + "---:|" * len(COLS))
    for r in data:
        name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "")
        vals = []
        for m, _, sc in COLS:
            try:
                v = float(r[ix[m]].replace(",", ""))
                u = units[ix[m]]
                if u == "Gbyte":
                    v *= 1e3
                elif u == "Kbyte":
                    v *= 1e-3
                elif u == "byte":
                    v *= 1e-6
                elif u in ("ms", "msecond"):
                    v *= 1e3
                elif u in ("ns", "nsecond"):
                    v *= 1e-3
                vals.append(f"{v:.1f}")
            except (KeyError, ValueError):
                vals.append("-")
        if tj:
            def num(m):
                v = float(r[ix[m]].replace(",", ""))
                return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e3, "msecond": 1e3, "us": 1.0,
                            "usecond": 1.0, "ns": 1e-3, "nsecond": 1e-3}.get(units[ix[m]], 1.0)
            try:
                t_us, by = num("gpu__time_duration.sum"), num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
                if name not in traffic or t_us > traffic[name]["time_us"]:
                    traffic[name] = {"dram_bytes": by, "time_us": t_us, "dram_read_bytes": num("dram__bytes_read.sum"),
                                     "dram_write_bytes": num("dram__bytes_write.sum")}
            except (KeyError, ValueError):
                pass
        g = r[ix["Grid Size"]].strip("()").split(",")[0]
        b = r[ix["Block Size"]].strip("()").split(",")[0]
        print(f"| `{name}` | {g} x {b} | " + " | ".join(vals) + " |")
    if tj:
        json.dump({"file": os.path.basename(rep), "command": cmd, "note": "per kernel: the longest launch of the capture",
                   "kernels": traffic}, open(tj, "w"), indent=1)


if __name__ == "__main__":
    main()
