#!/usr/bin/env python
"""Join an ncu SASS source page (instruction counts per address) with nvdisasm line info:
per CUDA source line, the warp instructions executed, average active threads and stall samples.

usage: tools/sass_lines.py report.ncu-rep <mangled-kernel-substring> [lib.so]
"""
import collections
import csv
import glob
import os
import re
import subprocess
import sys
import tempfile

rep, kern = sys.argv[1], sys.argv[2]
so = sys.argv[3] if len(sys.argv) > 3 else os.path.join(os.path.dirname(__file__), "..", "rbepwt_b200", "_lib", "librbepwt_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
cubin = glob.glob(os.path.join(tmp, "*.cubin"))[0]
sass = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout.splitlines()

# offset -> (file, line) for the kernel's text section
line_of, cur, infn = {}, None, False
for ln in sass:
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
    if m:
        infn = kern in m.group(1)
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*?);", ln)
    if m and cur:
        line_of[int(m.group(1), 16)] = cur

out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
tables, i = [], 0
while i < len(rows):
    if rows[i] and rows[i][0] == "Kernel Name":
        name = rows[i][1]
        hdr = rows[i + 1]
        j = i + 2
        body = []
        while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
            body.append(rows[j]); j += 1
        tables.append((name, hdr, body)); i = j
    else:
        i += 1
for name, hdr, body in tables:
    ia, ii, it, ism = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
    base = int(body[0][ia], 16)
    agg = collections.defaultdict(lambda: [0, 0, 0])
    tot = 0
    for r in body:
        off = int(r[ia], 16) - base
        key = line_of.get(off, ("?", 0))
        n = int(r[ii]); agg[key][0] += n; agg[key][1] += int(r[it]); agg[key][2] += int(r[ism] or 0); tot += n
    print("==", name, " total warp instructions", tot)
    srcs = {}
    for (f, l), (n, th, sm) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(os.environ.get("TOP", "40"))]:
        if f not in srcs:
            p = os.path.join(os.path.dirname(__file__), "..", "rbepwt_b200", "csrc", f)
            srcs[f] = open(p).read().splitlines() if os.path.isfile(p) else []
        text = srcs[f][l - 1].strip()[:90] if 0 < l <= len(srcs[f]) else ""
        print("%5.1f%%  thr/inst %4.1f  samples %6d  %s:%d  %s" % (100.0 * n / tot, th / max(n, 1), sm, f, l, text))
