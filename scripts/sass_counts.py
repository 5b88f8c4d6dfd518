#!/usr/bin/env python
"""Per-kernel counts of the memory / sync instructions in the shipped SASS (cuobjdump -sass libbulletb200.so).
usage: scripts/sass_counts.py > profiles/r2_sass_counts.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "bullet_js_b200", "csrc", "libbulletb200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
cols = ["LDGSTS", "UBLKCP", "SYNCS", "ATOM", "LDG", "STG", "LDS", "STS", "BAR", "SHFL", "VOTE"]
print("# SASS instruction counts per kernel (cuobjdump -sass libbulletb200.so, sm_100a): static evidence of the memory paths used")
print("# LDGSTS = cp.async global->shared, UBLKCP = cp.async.bulk (the TMA / bulk-copy engine), SYNCS = mbarrier ops, ATOM = atomics (ATOM*/RED*)")
print(f"{'kernel':78s}" + "".join(f"{c:>7s}" for c in cols))
cur, counts, order = None, collections.defaultdict(collections.Counter), []
it = iter(names)
for line in sass.split("\n"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = next(it).split("(")[0].replace("void ", "")
        order.append(cur)
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        for c in cols:
            if op.startswith(c) or (c == "ATOM" and (op.startswith("ATOM") or op.startswith("RED"))):
                counts[cur][c] += 1
                break
for k in order:
    print(f"{k[:78]:78s}" + "".join(f"{counts[k][c]:7d}" for c in cols))
