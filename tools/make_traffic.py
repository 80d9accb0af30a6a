"""DRAM traffic per kernel family from an ncu launch list (`--metrics gpu__time_duration.sum,dram__bytes_read.sum,
dram__bytes_write.sum --csv` of tools/profile_step.py = ONE group launch):
    python tools/make_traffic.py <launches.csv> <workload:precision> [out.json]
Writes / updates profiles/r2_traffic.json: per family the measured DRAM bytes per C-ABI call (bench.py's `launches`
count calls: fb_bn_bwd is two kernels, fb_flat_sqnorm two, ...), next to the duration shares under ncu."""
import collections
import csv
import json
import os
import re
import sys

FAMILY = [("conv_gemm_kernel", "conv_gemm", 1), ("wgrad_kernel", "conv_wgrad", 1), ("bn_apply_kernel", "bn_fwd", 1),
          ("bn_bwd_reduce_kernel", "bn_bwd", 1), ("bn_bwd_apply_kernel", "bn_bwd", 0), ("weight_prep", "weight_prep", 1),
          ("avgpool2", "pool", 1), ("reduce_multi_kernel", "wgrad_reduce", 1), ("stem_im2col", "stem_im2col", 1),
          ("head_fwd_kernel", "head", 1), ("head_bwd", "head", 0), ("fd_combine", "flat", 1), ("sqnorm_partial", "flat", 1),
          ("sqnorm_final", "flat", 0), ("mean_accumulate", "flat", 1), ("perturb_ranges", "flat", 1),
          ("flat_relayout", "flat", 1), ("flat_scale", "flat", 1), ("bn_ema", "misc", 1), ("group_finish", "misc", 1)]

src, key = sys.argv[1], sys.argv[2]
out = sys.argv[3] if len(sys.argv) > 3 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                         "profiles", "r2_traffic.json")
with open(src) as f:
    lines = [ln for ln in f if not ln.startswith("==")]
launch = collections.OrderedDict()
for r in csv.DictReader(lines):
    d = launch.setdefault(r["ID"], dict(name=r["Kernel Name"], us=0.0, bytes=0.0))
    v = float(r["Metric Value"].replace(",", ""))
    if r["Metric Name"] == "gpu__time_duration.sum":
        d["us"] = v / 1e3 if r["Metric Unit"] in ("nsecond", "ns") else v
    elif r["Metric Name"].startswith("dram__bytes"):
        d["bytes"] += v * dict(byte=1.0, Kbyte=1e3, Mbyte=1e6, Gbyte=1e9).get(r["Metric Unit"], 1.0)
fam = collections.defaultdict(lambda: dict(calls=0, kernels=0, dram_bytes=0.0, us=0.0))
for d in launch.values():
    for pat, name, is_call in FAMILY:
        if re.search(pat, d["name"]):
            f = fam[name]
            f["calls"] += is_call
            f["kernels"] += 1
            f["dram_bytes"] += d["bytes"]
            f["us"] += d["us"]
            break
total_us = sum(f["us"] for f in fam.values())
res = json.load(open(out)) if os.path.exists(out) else {}
res["source"] = "ncu dram__bytes_read.sum + dram__bytes_write.sum per kernel, one group launch (see profiles/README.md)"
res[key] = {name: dict(dram_bytes_per_launch=round(f["dram_bytes"] / max(f["calls"], 1)), calls=f["calls"],
                       kernels=f["kernels"], dram_gb_per_s_under_ncu=round(f["dram_bytes"] / f["us"] / 1e3, 1),
                       share_of_step_under_ncu=round(f["us"] / total_us, 4), capture=os.path.basename(src))
            for name, f in sorted(fam.items(), key=lambda kv: -kv[1]["us"])}
with open(out, "w") as f:
    json.dump(res, f, indent=1)
print(json.dumps(res[key], indent=1))
