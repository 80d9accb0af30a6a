"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import csv
import collections
import re
import sys

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = v / 1e3 if unit in ("nsecond", "ns") else (v if unit in ("usecond", "us") else v * 1e3)
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        rows.append((name, us, r.get("Grid Size", ""), r.get("Block Size", "")))
agg = collections.defaultdict(lambda: [0, 0.0])
for name, us, *_ in rows:
    agg[name][0] += 1
    agg[name][1] += us
total = sum(v[1] for v in agg.values())
print(f"{len(rows)} launches, {total:.1f} us total")
for name, (cnt, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{us:10.1f} us {100 * us / total:5.1f}%  x{cnt:4d}  avg {us / cnt:8.1f} us  {name}")
if len(sys.argv) > 2:
    print("--- top individual launches")
    for name, us, grid, block in sorted(rows, key=lambda r: -r[1])[: int(sys.argv[2])]:
        print(f"{us:10.1f} us  grid {grid} block {block}  {name}")
