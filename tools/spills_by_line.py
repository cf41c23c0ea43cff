#!/usr/bin/env python
"""List the source lines that carry local-memory (spill / stack) instructions in a kernel of the built library.
usage: spills_by_line.py <object-name-in-lib, e.g. pa_chain.sm> <kernel-substring, e.g. Li96ELi4> [top]"""
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
obj, kern_sub, top = sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 25
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", obj, os.path.join(ROOT, "phyloacc_b200", "libphyloacc_b200.so")], cwd=tmp, check=True, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
sass = subprocess.run(["nvdisasm", "-g", "-c", cubin], cwd=tmp, capture_output=True, text=True).stdout
cur, kern = None, None
cnt, tot, ops = collections.Counter(), 0, collections.Counter()
for line in sass.splitlines():
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if line.startswith(".text."):
        kern = line.strip()
        continue
    if kern and kern_sub in kern:
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(.*)", line)
        if m:
            ins = m.group(1)
            tot += 1
            op = ins.split()[1] if ins.startswith("@") else ins.split()[0]
            ops[op.split(".")[0]] += 1
            if "STL" in ins or "LDL" in ins:
                cnt[cur] += 1
print("instructions:", tot, "local-memory instructions:", sum(cnt.values()))
print("top opcodes:", ops.most_common(14))
for k, v in sorted(cnt.items(), key=lambda kv: -kv[1])[:top]:
    print(k, v)
