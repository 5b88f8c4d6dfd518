"""Attribute executed SASS instructions (ncu --page source --csv) to CUDA source lines
(nvdisasm -g of the cubin).  usage: sass_lines.py <all.sass> <ncu_source.csv> <mangled kernel substr> [top]"""
import collections
import csv
import re
import sys

sass, ncu_csv, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
lines = open(sass).read().split("\n")
start = [i for i, l in enumerate(lines) if l.startswith("//--------------------- .text.") and kern in l][0]
cur, off2line = None, {}
for l in lines[start + 1:]:
    if l.startswith("//--------------------- .text."):
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*?);", l)
    if m:
        off2line[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(ncu_csv)))
agg, samp, tot, inst = collections.Counter(), collections.Counter(), 0, 0
hdr = None
for r in rows:
    if r and r[0] == "Address":
        hdr, base = r, None
        inst += 1
        if inst > 1:
            break
        ia, ie, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
        continue
    if hdr is None or len(r) <= ie:
        continue
    a = int(r[ia], 16)
    if base is None:
        base = a
    loc = off2line.get(a - base)
    agg[loc] += int(r[ie])
    samp[loc] += int(r[isamp])
    tot += int(r[ie])
print("total warp-instructions", tot, "samples", sum(samp.values()))
src = {}
for loc, n in agg.most_common(top):
    f, ln = loc if loc else ("?", 0)
    if f not in src:
        try:
            src[f] = open("/root/repo/bullet_js_b200/csrc/" + f).read().split("\n")
        except OSError:
            src[f] = None
    text = src[f][ln - 1].strip()[:86] if src[f] else ""
    print(f"{n:>9} {100 * n / tot:5.1f}%  smp {samp[loc]:>6}  {f}:{ln}  {text}")
