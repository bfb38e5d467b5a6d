"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel (name + grid) count, total and
average duration, share of the summed kernel time.  usage: python tools/launch_list.py launches.csv [title]"""
import collections
import csv
import re
import sys

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"^void ", "", r["Kernel Name"])
    name = re.sub(r"\(.*", "", name)
    rows.append((name, r["Grid Size"], float(r["Metric Value"]) / 1e3))
agg = collections.OrderedDict()
for name, grid, us in rows:
    k = (name, grid)
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += us
total = sum(a[1] for a in agg.values())
if len(sys.argv) > 2:
    print(sys.argv[2])
print(f"launches {len(rows)}, sum of kernel durations {total:.1f} us (cold-cache, serialised: compare shares)")
for (name, grid), (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{name:44s} grid {grid:>16s} n={n:4d} total_us={us:9.1f} avg_us={us / n:8.2f} share={100 * us / total:6.2f}%")
