"""Aggregate an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv` launch list by
kernel name: launches, total / average duration and, when the DRAM metrics are present, bytes moved and GB/s."""
import collections
import csv
import re
import sys

with open(sys.argv[1]) as f:
    lines = [ln for ln in f if not ln.startswith("==")]
launch = collections.OrderedDict()
for r in csv.DictReader(lines):
    key = r["ID"]
    d = launch.setdefault(key, dict(name=re.sub(r"\(.*", "", r["Kernel Name"]).replace("void fb::", ""),
                                    grid=r.get("Grid Size", ""), us=0.0, rd=0.0, wr=0.0))
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    m = r["Metric Name"]
    if m == "gpu__time_duration.sum":
        d["us"] = v / 1e3 if unit in ("nsecond", "ns") else (v if unit in ("usecond", "us") else v * 1e3)
    elif m.startswith("dram__bytes"):
        scale = dict(byte=1.0, Kbyte=1e3, Mbyte=1e6, Gbyte=1e9).get(unit, 1.0)
        d["rd" if "read" in m else "wr"] = v * scale
rows = list(launch.values())
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for d in rows:
    a = agg[d["name"]]
    a[0] += 1
    a[1] += d["us"]
    a[2] += d["rd"]
    a[3] += d["wr"]
total = sum(v[1] for v in agg.values())
print(f"{len(rows)} launches, {total:.1f} us total (ncu: serialised, cold caches)")
for name, (cnt, us, rd, wr) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    bw = f"  dram {rd / 1e6:9.1f} MB rd {wr / 1e6:9.1f} MB wr  {(rd + wr) / us / 1e3:7.1f} GB/s" if rd + wr > 0 else ""
    print(f"{us:10.1f} us {100 * us / total:5.1f}%  x{cnt:4d}  avg {us / cnt:8.1f} us  {name}{bw}")
if len(sys.argv) > 2:
    print("--- top individual launches")
    for d in sorted(rows, key=lambda r: -r["us"])[: int(sys.argv[2])]:
        bw = f"  dram {(d['rd'] + d['wr']) / 1e6:8.1f} MB {(d['rd'] + d['wr']) / d['us'] / 1e3:7.1f} GB/s" if d["rd"] + d["wr"] else ""
        print(f"{d['us']:10.1f} us  grid {d['grid']:>16s}  {d['name']}{bw}")
