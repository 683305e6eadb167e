"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name.
usage: python tests/runs/summarize_launches.py launches.csv "header line" > profiles/xxx_launches_summary.txt"""
import csv
import re
import sys
from collections import defaultdict

path, header = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else ""
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = v / 1e3 if unit in ("ns", "nsecond") else v * 1e3 if unit in ("ms", "msecond") else v
    name = re.sub(r"^void ", "", r["Kernel Name"])
    name = re.sub(r"\(.*$", "", name).replace("<unnamed>::", "")
    rows.append((name, us))
agg = defaultdict(lambda: [0, 0.0])
for n, us in rows:
    agg[n][0] += 1
    agg[n][1] += us
total = sum(v[1] for v in agg.values())
print(header)
for n, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    if us / total < 0.0005:
        continue
    print(f"{n[:72]:72s} launches={c:5d} total_ms={us / 1e3:9.3f} share={100 * us / total:5.1f}% avg_us={us / c:9.1f}")
print(f"total_ms={total / 1e3:.3f} launches={len(rows)}")
