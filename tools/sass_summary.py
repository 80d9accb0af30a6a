"""Per-kernel counts of the Blackwell-specific SASS instructions in the shipped library (proof that the tensor-core
kernels are tcgen05 / TMEM / TMA code): `python tools/sass_summary.py > profiles/r2_sass_summary.txt`.
UTCHMMA = tcgen05.mma, UTMALDG = TMA tile load, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, SYNCS = mbarrier ops."""
import collections
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                          "fullbatchtraining_b200", "libfullbatch_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
demangle = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True,
                          text=True).stdout.split("\n")
names = iter(demangle)
OPS = ["UTCHMMA", "UTMALDG", "LDTM", "UTCBAR", "SYNCS", "HMMA", "LDG", "STG", "LDS", "STS", "LD.E", "ST.E"]
cur, counts, order = None, {}, []
for line in sass.split("\n"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = re.sub(r"\(.*", "", re.sub(r"\((int|bool)\)", "", next(names))).replace("void ", "")
        counts[cur] = collections.Counter()
        order.append(cur)
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        counts[cur]["total"] += 1
        for o in OPS:
            if op == o or op.startswith(o + "."):
                counts[cur][o] += 1
print(f"{os.path.basename(lib)}: {len(order)} kernels, sm_100a SASS (cuobjdump -sass)")
print(f"{'kernel':60s} " + " ".join(f"{o:>8s}" for o in ["total"] + OPS))
for k in order:
    print(f"{k[:60]:60s} " + " ".join(f"{counts[k][o]:8d}" for o in ["total"] + OPS))
tot = collections.Counter()
for k in order:
    tot.update(counts[k])
print(f"{'ALL':60s} " + " ".join(f"{tot[o]:8d}" for o in ["total"] + OPS))
