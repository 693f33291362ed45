#!/usr/bin/env python
"""Per CUDA source line: stall samples by reason (ncu source page joined with nvdisasm line info).
usage: tools/sass_stalls.py report.ncu-rep <mangled-kernel-substring> [lib.so]   (TOP=n lines, default 30)"""
import collections, csv, glob, os, re, subprocess, sys, tempfile
rep, kern = sys.argv[1], sys.argv[2]
so = sys.argv[3] if len(sys.argv) > 3 else os.path.join(os.path.dirname(__file__), "..", "rbepwt_b200", "_lib", "librbepwt_b200.so")
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
        i = j
        if kern.split("ILi")[0].replace("_ZN6rbepwt", "") not in name and "k1" not in name: continue
        ia = hdr.index("Address")
        reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        idx = {h: hdr.index(h) for h in reasons}
        base = int(body[0][ia], 16)
        agg = collections.defaultdict(collections.Counter); tot = collections.Counter()
        for r in body:
            key = line_of.get(int(r[ia], 16) - base, ("?", 0))
            for h in reasons:
                v = int(r[idx[h]] or 0)
                agg[key][h] += v; tot[h] += v
        allsum = sum(tot.values())
        print("==", name, "samples", allsum, {h[6:]: round(100 * v / allsum, 1) for h, v in tot.most_common(8)})
        for key, c in sorted(agg.items(), key=lambda kv: -sum(kv[1].values()))[:int(os.environ.get("TOP", "30"))]:
            sm = sum(c.values())
            print("%5.1f%% %s:%d  %s" % (100 * sm / allsum, key[0], key[1], " ".join("%s=%.0f%%" % (h[6:], 100 * v / sm) for h, v in c.most_common(4))))
    else:
        i += 1
