#!/usr/bin/env python
"""Turn an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel markdown table.
usage: python tools/summarise_launches.py gpurun_out/launches.csv "<command that was profiled>" > profiles/rNN_launches.md"""
import collections
import csv
import sys


def main():
    path, cmd = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    tot = collections.defaultdict(lambda: [0, 0.0, 0.0])
    scale = {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}
    for x in rows:
        name = x["Kernel Name"].split("(")[0].replace("void ", "")[:70]
        v = float(x["Metric Value"].replace(",", "")) * scale.get(x["Metric Unit"], 1.0)
        t = tot[name]
        t[0] += 1
        t[1] += v
        t[2] = max(t[2], v)
    T = sum(v[1] for v in tot.values())
    ours = sum(v[1] for k, v in tot.items() if k.startswith("k_"))
    print(f"# ncu launch list summary\n\ncommand: `{cmd}`\n")
    print(f"{len(rows)} launches, {T / 1e3:.2f} ms total device time (cold-cache, serialised: compare SHARES, not absolutes); "
          f"hand-written `k_*` kernels: {100 * ours / T:.1f} % of device time, torch glue kernels: {100 * (1 - ours / T):.1f} %\n")
    print("| kernel | launches | total us | share | max us / launch |\n|---|---:|---:|---:|---:|")
    for n, (c, t, m) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:40]:
        print(f"| `{n}` | {c} | {t:.1f} | {100 * t / T:.1f} % | {m:.1f} |")


if __name__ == "__main__":
    main()
