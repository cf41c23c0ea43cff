#!/usr/bin/env python
"""Summarise one `ncu --set full --import-source on` capture of a chain kernel for profiles/:
   headline metrics, stall-reason totals, instructions per phase of pa_chain.cu (by source-line range), hottest lines.
usage: ncu_summary.py <rep.ncu-rep> <object-name-in-lib e.g. pa_chain.sm> <kernel-substring e.g. Li96ELi4> <chain-iterations in the capture> [title]"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, obj, ksub, n_iter = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
title = sys.argv[5] if len(sys.argv) > 5 else os.path.basename(rep)
SRC = os.path.join(ROOT, "phyloacc_b200", "csrc", os.environ.get("PA_NCU_SRC", "pa_chain.cu"))  # PA_NCU_SRC=pa_gbgc_chain.cu for the GBGC chain


def ncu_csv(page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


raw = ncu_csv("raw")
h, u, v = raw[0], raw[1], raw[2]
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
print(f"# {title}")
vals = {}
for i, k in enumerate(h):
    if k in want:
        vals[k] = v[i]
        print(f"{k} [{u[i]}] = {v[i]}")
inst = float(vals.get("smsp__inst_executed.sum", 0))
print(f"warp-instructions per chain-iteration = {inst / n_iter:.0f}   (chain-iterations in this capture: {n_iter:.0f})")
stalls = [(k.replace("smsp__pcsamp_warps_issue_stalled_", ""), float(v[i])) for i, k in enumerate(h)
          if k.startswith("smsp__pcsamp_warps_issue_stalled_") and "not_issued" not in k]
tot = sum(x for _, x in stalls) or 1
print("stall reasons (share of warp samples): " + ", ".join(f"{k} {100 * x / tot:.1f}%" for k, x in sorted(stalls, key=lambda t: -t[1])[:9]))

# ---- per phase / per line
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", obj, os.path.join(ROOT, "phyloacc_b200", "libphyloacc_b200.so")], cwd=tmp, check=True, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
sass = subprocess.run(["nvdisasm", "-g", "-c", cubin], cwd=tmp, capture_output=True, text=True).stdout
off2, cur, ins = {}, None, False
for l in sass.splitlines():
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if l.startswith(".text."):
        ins = ksub in l
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", l)
    if m and ins:
        off2[int(m.group(1), 16)] = cur
src_lines = open(SRC).read().split("\n")
marks = []  # (line, label) from the phase comments of the source
for i, l in enumerate(src_lines, 1):
    m = re.match(r"\s*// -{4,} (.*)", l)
    if m:
        marks.append((i, m.group(1)[:60]))
    m = re.match(r"(?:template.*\n)?__device__ .*? (\w+)\(", l)
    if m and "forceinline" in l or (m and "noinline" in l):
        marks.append((i, "fn " + m.group(1)))
marks.sort()


def phase(f, ln):
    if f != os.path.basename(SRC):
        return None
    lab = "prologue"
    for a, b in marks:
        if a <= ln:
            lab = b
    return lab


rows = ncu_csv("source")
hdr = next(i for i, r in enumerate(rows) if "Address" in r and "Source" in r)
hh = rows[hdr]
iA, iI, iW = hh.index("Address"), hh.index("Instructions Executed"), hh.index("# Samples")
base = int(rows[hdr + 1][iA], 16)
agg, smp, stat, line_agg = collections.Counter(), collections.Counter(), collections.Counter(), collections.Counter()
last = "prologue"
for r in rows[hdr + 1:]:
    if len(r) < len(hh):
        continue
    f, ln = off2.get(int(r[iA], 16) - base, ("?", 0))
    p = phase(f, ln)
    if p:
        last = p
    agg[last] += int(r[iI])
    smp[last] += int(r[iW])
    stat[last] += 1
    line_agg[(f, ln)] += int(r[iI])
T, TS = sum(agg.values()) or 1, sum(smp.values()) or 1
print(f"static SASS instructions of the kernel: {len(off2)}")
print("instructions / stall samples by phase (library code attributed to the phase that precedes it in address order):")
for k, x in agg.most_common(16):
    print(f"  {100 * x / T:5.1f}%i {x / n_iter:9.0f} instr/iter {100 * smp[k] / TS:5.1f}%s  static {stat[k]:6d}  {k}")
print("hottest source lines (share of executed warp-instructions):")
for (f, ln), x in line_agg.most_common(int(os.environ.get("PA_NCU_TOP", "14"))):
    text = src_lines[ln - 1].strip()[:110] if f == os.path.basename(SRC) and 0 < ln <= len(src_lines) else ""
    print(f"  {100 * x / T:5.1f}%  {f}:{ln}  {text}")
