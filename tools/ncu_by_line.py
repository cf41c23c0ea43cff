#!/usr/bin/env python
"""Join an ncu SASS source page (ncu -i rep --page source --csv) with nvdisasm -g line info of the same cubin and
aggregate executed instructions / stall samples per CUDA source line.
usage: ncu_by_line.py <src.csv> <kernel.sass from `nvdisasm -g -c cubin`> [top] [section-substring] [sort: inst|samples]"""
import csv
import re
import sys
from collections import defaultdict

src_csv, sass, top = sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 50
section = sys.argv[4] if len(sys.argv) > 4 else "chain_kernel"
sort_by = sys.argv[5] if len(sys.argv) > 5 else "inst"
# nvdisasm: find the chain kernel section, map offset -> (file, line)
off2line = {}
cur = None
in_kernel = False
for line in open(sass):
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if line.startswith(".text."):
        in_kernel = section in line
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", line)
    if m and in_kernel:
        off2line[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src_csv)))
hdr = next(i for i, r in enumerate(rows) if "Address" in r and "Source" in r)
h = rows[hdr]
iA, iS, iI, iW = h.index("Address"), h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
stall_cols = [(i, c) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
base = int(rows[hdr + 1][iA], 16)
agg = defaultdict(lambda: [0, 0, defaultdict(int), defaultdict(int)])
tot_i = tot_w = 0
for r in rows[hdr + 1:]:
    if len(r) < len(h):
        continue
    off = int(r[iA], 16) - base
    key = off2line.get(off, ("?", 0))
    ni, nw = int(r[iI]), int(r[iW])
    a = agg[key]
    a[0] += ni
    a[1] += nw
    op = r[iS].split()[0] if r[iS].split() else "?"
    if op.startswith("@"):
        op = r[iS].split()[1]
    a[2][op.split(".")[0]] += ni
    for i, c in stall_cols:
        try:
            a[3][c[6:]] += int(r[i])
        except ValueError:
            pass
    tot_i += ni
    tot_w += nw
print(f"total warp-instructions {tot_i}, samples {tot_w}, sass lines {len(rows) - hdr - 1}")
items = sorted(agg.items(), key=lambda kv: -(kv[1][0] if sort_by == "inst" else kv[1][1]))
src_cache = {}
def src_line(f, l):
    import glob
    if f not in src_cache:
        cand = glob.glob(f"/root/repo/phyloacc_b200/csrc/{f}")
        src_cache[f] = open(cand[0]).read().split("\n") if cand else []
    s = src_cache[f]
    return s[l - 1].strip()[:90] if 0 < l <= len(s) else ""
for (f, l), (ni, nw, ops, st) in items[:top]:
    topops = ",".join(f"{k}:{v * 100 // max(ni, 1)}" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:4])
    tops = ",".join(f"{k}:{v * 100 // max(nw, 1)}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{ni / tot_i * 100:5.1f}%i {nw / max(tot_w, 1) * 100:5.1f}%s {f}:{l:<4d} {src_line(f, l)}   [{topops}] <{tops}>")
