#!/usr/bin/env python
"""gpurun_out/*_<tag>.{csv,ncu-rep} -> profiles/<tag>_*.txt (the committed, judged summaries).
usage: tools/summarize_profiles.py <tag>"""
import collections
import csv
import glob
import os
import subprocess
import sys

tag = sys.argv[1]
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
G = os.environ.get("PROF_SRC", os.path.join(ROOT, "gpurun_out"))      # the .ncu-rep files
GL = os.environ.get("PROF_LAUNCHES", os.path.join(ROOT, "gpurun_out"))  # the launch list
P = os.environ.get("PROF_DST", os.path.join(ROOT, "profiles"))
os.makedirs(P, exist_ok=True)

# ---- launch list
rows = list(csv.reader(open(os.path.join(GL, "launches_%s.csv" % tag))))
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        h, start = r, i
        break
kn, mv, idc, gs, bs = h.index("Kernel Name"), h.index("Metric Value"), h.index("ID"), h.index("Grid Size"), h.index("Block Size")
data = [(int(r[idc]), r[kn].split("(")[0], float(r[mv].replace(",", "")), r[gs], r[bs]) for r in rows[start + 2:] if len(r) > mv]
with open(os.path.join(P, "%s_launches.txt" % tag), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k[0-9q_]  python bench.py --batch 512 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-extras --parity-images 0\n")
    f.write("# per-launch device time (ns), cold-cache and serialised by ncu: compare SHARES, not absolutes.\n")
    f.write("# The list covers the warm-up steps, the timed step and the serial kernel-timing steps of bench.py.\n")
    agg = collections.OrderedDict()
    for d in data:
        agg.setdefault(d[1], [0, 0.0])
        agg[d[1]][0] += 1
        agg[d[1]][1] += d[2]
    tot = sum(v[1] for v in agg.values())
    f.write("\n%-28s %8s %12s %10s %7s\n" % ("kernel", "launches", "total_ms", "avg_us", "share"))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write("%-28s %8d %12.3f %10.1f %6.1f%%\n" % (k, v[0], v[1] / 1e6, v[1] / v[0] / 1e3, 100 * v[1] / tot))
    f.write("\n# every launch: id kernel grid block ns\n")
    for d in data:
        f.write("%d %s %s %s %.0f\n" % (d[0], d[1], d[3].replace(" ", ""), d[4].replace(" ", ""), d[2]))

# ---- full captures
WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit",
        "launch__cluster", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sectors.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__pcsamp_warps_issue_stalled"]
for rep in sorted(glob.glob(os.path.join(G, "prof_*_%s.ncu-rep" % tag))):
    name = os.path.basename(rep)[len("prof_"):-len("_%s.ncu-rep" % tag)]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(out.splitlines()))
    hdr, units = rr[0], rr[1]
    with open(os.path.join(P, "%s_ncu_%s.txt" % (tag, name)), "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on -k regex:%s -s <3 warm-up steps> -c 1|2  python bench.py --batch 512 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-extras --parity-images 0   (tools/make_profiles.sh)\n" % name)
        for r in rr[2:]:
            f.write("== %s  grid %s block %s\n" % (r[hdr.index("Kernel Name")], r[hdr.index("Grid Size")] if "Grid Size" in hdr else "?", r[hdr.index("Block Size")] if "Block Size" in hdr else "?"))
            for hh, u, v in zip(hdr, units, r):
                if any(hh.startswith(w) for w in WANT) and "not_issued" not in hh and v not in ("0", ""):
                    if hh.endswith((".sum", ".ratio", "active", "elapsed")) or "pcsamp" in hh or "launch" in hh or "per_second" in hh:
                        f.write("  %-78s %-14s %s\n" % (hh, u, v))
    print("wrote", name)

# ---- DRAM traffic of the hot kernels (bench.py scales roofline.traffic from this file)
import json
traffic = {"source": "profiles/%s_ncu_*.txt (ncu --set full, one launch each: k1_walk = both instantiations of one 256-image path group; k3/k5 = the level-1 launch of one sub-batch; tools/make_profiles.sh)" % tag}
for rep in sorted(glob.glob(os.path.join(G, "prof_*_%s.ncu-rep" % tag))):
    name = os.path.basename(rep)[len("prof_"):-len("_%s.ncu-rep" % tag)]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(out.splitlines()))
    hdr, units = rr[0], rr[1]
    def col(r, n):
        i = hdr.index(n)
        v = float(r[i].replace(",", ""))
        u = units[i]
        return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e3, "us": 1.0, "ns": 1e-3, "s": 1e6}.get(u, 1.0)
    e = {"dram_bytes_read": 0.0, "dram_bytes_write": 0.0, "duration_us": 0.0, "launches_summed": 0}
    for r in rr[2:]:
        e["dram_bytes_read"] += col(r, "dram__bytes_read.sum")
        e["dram_bytes_write"] += col(r, "dram__bytes_write.sum")
        e["duration_us"] += col(r, "gpu__time_duration.sum")
        e["launches_summed"] += 1
        gy = r[hdr.index("Grid Size")].strip("()").replace(" ", "").split(",") if "Grid Size" in hdr else []
    e["images_per_launch"] = int(gy[1]) if name in ("k3_dwt_level", "k5_idwt_level") and len(gy) > 1 else (int(gy[0]) if name in ("k4_select", "k3_dwt_tail") else int(os.environ.get("PROF_GROUP", "256")))
    traffic[name] = e
with open(os.path.join(P, "%s_traffic.json" % tag), "w") as f:
    json.dump(traffic, f, indent=1)
print("wrote traffic")
