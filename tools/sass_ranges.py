#!/usr/bin/env python
"""Like sass_lines.py, but aggregated over source line ranges given as name:file:lo-hi ... (unmatched -> other)."""
import collections, csv, glob, os, re, subprocess, sys, tempfile
rep, kern, so = sys.argv[1], sys.argv[2], sys.argv[3]
ranges = []
for a in sys.argv[4:]:
    name, f, r = a.split(":")
    lo, hi = r.split("-")
    ranges.append((name, f, int(lo), int(hi)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
cubin = glob.glob(os.path.join(tmp, "*.cubin"))[0]
sass = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout.splitlines()
line_of, cur, infn = {}, None, False
for ln in sass:
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
    if m:
        infn = kern in m.group(1); continue
    if not infn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*?);", ln)
    if m and cur: line_of[int(m.group(1), 16)] = cur
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
i = 0
while i < len(rows):
    if rows[i] and rows[i][0] == "Kernel Name":
        name, hdr = rows[i][1], rows[i + 1]
        j = i + 2; body = []
        while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
            body.append(rows[j]); j += 1
        ia, ii, it, ism = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
        base = int(body[0][ia], 16)
        agg = collections.defaultdict(lambda: [0, 0, 0]); tot = 0
        for r in body:
            f, l = line_of.get(int(r[ia], 16) - base, ("?", 0))
            key = "other"
            for nm, rf, lo, hi in ranges:
                if f == rf and lo <= l <= hi: key = nm; break
            else:
                if f not in ("walk.cuh", "paths.cuh"): key = "hdr:" + f
            n = int(r[ii]); agg[key][0] += n; agg[key][1] += int(r[it]); agg[key][2] += int(r[ism] or 0); tot += n
        print("==", name, "total warp instructions", tot)
        for k, (n, th, sm) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
            print("%6.1f%%  thr/inst %4.1f  samples %7d  %s" % (100.0 * n / tot, th / max(n, 1), sm, k))
        i = j
    else:
        i += 1
