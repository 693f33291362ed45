#!/bin/bash
# usage (GPU box): tools/ht_ncu.sh   -- the path kernels of the heavy-tailed workload one by one under ncu (durations in isolation)
ncu --metrics gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__warps_active.avg.pct_of_peak_sustained_active \
    --clock-control none -k regex:"k1_|k2_" -s 24 -c 12 --csv --log-file /tmp/ht_ncu.csv python tools/heavytail_timing.py 256 > /tmp/ht_ncu.log 2>&1
python - <<'PY'
import csv
rows = list(csv.reader(open("/tmp/ht_ncu.csv")))
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        h = r; start = i; break
kn, mn, mv, idc = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
agg = {}
for r in rows[start + 1:]:
    if len(r) <= mv: continue
    agg.setdefault((int(r[idc]), r[kn].split("(")[0]), {})[r[mn]] = r[mv]
for (i, k), m in sorted(agg.items()):
    print("%3d %-28s %10s ns  issue %5s%%  inst %12s  lanes/inst %5s  warps_active %5s%%" % (i, k, m.get("gpu__time_duration.sum"), m.get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
          m.get("smsp__inst_executed.sum"), m.get("smsp__thread_inst_executed_per_inst_executed.ratio"), m.get("sm__warps_active.avg.pct_of_peak_sustained_active")))
PY
