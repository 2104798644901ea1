"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel time and share of one step.

    python tools/launch_list_summary.py launches.csv [steps_in_capture] > summary.txt

The capture holds warm-up + timed + profile-pass steps; kernels are grouped by name and divided by the step count."""
import collections
import csv
import re
import sys

path = sys.argv[1]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rows = []
with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        rows.append((r["Kernel Name"], float(r["Metric Value"].replace(",", ""))))
agg = collections.OrderedDict()
for name, ns in rows:
    key = re.sub(r"\(.*", "", name)[:70]
    a = agg.setdefault(key, [0, 0.0])
    a[0] += 1
    a[1] += ns
tot = sum(v[1] for v in agg.values())
print(f"{len(rows)} launches captured, {steps} step(s); times in ms per step (cold-cache, serialised: compare SHARES)")
for k, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{ns / 1e6 / steps:8.3f} ms  x{c / steps:<5.1f} {100 * ns / tot:5.1f}%  {k}")
print(f"total {tot / 1e6 / steps:.3f} ms per step")
