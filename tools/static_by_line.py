#!/usr/bin/env python
"""Static SASS instruction counts of one object file by source region (code-size / instruction-cache work).
usage: static_by_line.py <file.o> [source-file-name] [bucket-lines]"""
import collections
import os
import re
import subprocess
import sys
import tempfile

obj = os.path.abspath(sys.argv[1])
src = sys.argv[2] if len(sys.argv) > 2 else None
bucket = int(sys.argv[3]) if len(sys.argv) > 3 else 25
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", obj], cwd=tmp, check=True, capture_output=True)
cnt = collections.Counter()
for cubin in os.listdir(tmp):
    sass = subprocess.run(["nvdisasm", "-g", "-c", cubin], cwd=tmp, capture_output=True, text=True).stdout
    cur = None
    for l in sass.splitlines():
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        if re.match(r"\s*/\*[0-9a-f]{4,}\*/", l):
            cnt[cur] += 1
print("static instructions:", sum(cnt.values()))
b = collections.Counter()
for (f, ln), c in ((k, v) for k, v in cnt.items() if k):
    b[(f, ln // bucket * bucket if f == src else 0)] += c
for k, c in b.most_common(40):
    print(f"{c:7d}  {k[0]}:{k[1]}")
