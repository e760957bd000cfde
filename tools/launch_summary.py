"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (markdown table to stdout).

  python tools/launch_summary.py gpurun_out/r1_launches.csv STEPS > profiles/....md
STEPS = number of optimiser steps the profiled command ran (to print per-step figures).
"""
import csv
import re
import sys
from collections import defaultdict

path, steps = sys.argv[1], float(sys.argv[2])
rows = []
with open(path, newline="") as fh:
    lines = [l for l in fh if not l.startswith("==")]
rd = csv.DictReader(lines)
tot = defaultdict(lambda: [0, 0.0])
n = 0
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").strip()
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    tot[name][0] += 1
    tot[name][1] += us
    n += 1
total_us = sum(t for _, t in tot.values())
print(f"{n} launches, {steps:g} optimiser steps; sum of kernel durations {total_us / 1e3:.1f} ms = "
      f"{total_us / 1e3 / steps:.2f} ms per step (serialised, cold-cache ncu replay: compare SHARES, not absolutes).\n")
print("| kernel | launches/step | us/launch | ms/step | share |")
print("|---|---:|---:|---:|---:|")
for name, (c, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{name}` | {c / steps:.1f} | {t / c:.1f} | {t / 1e3 / steps:.3f} | {100 * t / total_us:.1f}% |")
