#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that prove which hardware paths the library uses (profiles/rNN_sass_summary.txt):
UTCIMMA / UTCBAR / LDTM (tcgen05 int8 MMA, its commit barrier, TMEM loads), DMMA (FP64 tensor), UBLKCP (1-D bulk copies:
cp.async.bulk, the TMA unit WITHOUT a tensor map), UTMALDG (tensor-map TMA: none here), SYNCS (mbarrier), MUFU, I2F / F2F
(conversion unit), DFMA / DADD / DMUL, STL / LDL (local memory).

    python tools/sass_summary.py [covid19uk_b200/lib/libseir_b200.so] > profiles/r02_sass_summary.txt
"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "covid19uk_b200/lib/libseir_b200.so"
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
keys = ["UTCIMMA", "UTCBAR", "LDTM", "DMMA", "UBLKCP", "UTMALDG", "SYNCS", "MUFU", "I2F", "F2F", "DFMA", "DADD", "DMUL", "STL", "LDL"]
counts, cur, total = collections.OrderedDict(), None, collections.Counter()
it = iter(names)
for line in sass.split("\n"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = next(it).split("(")[0].replace("void ", "")
        counts[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and cur:
        counts[cur]["_all"] += 1
        for k in keys:
            if m.group(1) == k or m.group(1).startswith(k):
                counts[cur][k] += 1
print(f"# cuobjdump -sass {lib}: instruction counts per kernel (sm_100a)")
print("kernel".ljust(64) + "".join(k.rjust(9) for k in ["instrs"] + keys))
for name, c in counts.items():
    if not name.startswith("seir_"):
        continue
    print(name[:63].ljust(64) + "".join(str(c[k]).rjust(9) for k in ["_all"] + keys))
