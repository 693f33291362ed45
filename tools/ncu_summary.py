#!/usr/bin/env python
"""Print the handful of ncu raw-page metrics the roofline discussion needs.
usage: tools/ncu_summary.py report.ncu-rep [kernel-regex]"""
import csv
import re
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput", "lts__t_bytes.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__pcsamp_warps_issue_stalled"]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    kcol = hdr.index("Kernel Name")
    for r in rows[2:]:
        if len(sys.argv) > 2 and not re.search(sys.argv[2], r[kcol]):
            continue
        print("==", r[kcol][:100])
        for h, u, v in zip(hdr, units, r):
            if any(h.startswith(w) for w in WANT) and "not_issued" not in h and v not in ("0", ""):
                if h.endswith((".sum", ".ratio", "_pct", "active", "elapsed")) or "pcsamp" in h or "launch" in h or "per_second" in h:
                    print("  %-75s %-12s %s" % (h, u, v))


main()
