#!/usr/bin/env python
"""Debug: build the library with -DTPR_STATS into a scratch .so and print k1_paths_tpr's trips per level
and trip kind for one batch of the bench workload (run on the GPU box)."""
import ctypes, os, subprocess, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from rbepwt_b200 import build as b
so = os.path.join(ROOT, "gpurun_out", "librbepwt_stats.so")
cmd = ["nvcc"] + b.NVCC_FLAGS + ["-DTPR_STATS", "-o", so, os.path.join(b.CSRC, "rbepwt_b200.cu")]
subprocess.check_call(cmd)
b.LIB_PATH = so
b.needs_build = lambda: False
import numpy as np, torch
import rbepwt_b200 as rb
from rbepwt_b200 import synth, _capi
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
imgs, labs = synth.torch_batch(B, 512, 512, 1024, 1000, device="cuda")
c = rb.BatchCodec()
L = _capi.lib()
out = (ctypes.c_ulonglong * 160)()
c.encode(imgs, labs, 16, "bior4.4"); c.sync()
L.rbepwt_debug_tpr_stats(c._ctx, out, 1)
c.encode(imgs, labs, 16, "bior4.4"); c.sync()
L.rbepwt_debug_tpr_stats(c._ctx, out, 1)
v = np.array(list(out), dtype=np.float64)
names = ["lut3x3", "rows3", "wide4", "list"]
tot = 0
print("level  warp-trips/img  lanes/trip   lane-trips/img by kind: " + "  ".join(names))
for lv in range(16):
    trips = v[128 + lv]
    if not trips:
        continue
    lanes = [v[(lv * 4 + k) * 2 + 1] for k in range(4)]
    tot += trips
    print("%5d  %12.0f  %10.1f   %s" % (lv + 1, trips / B, sum(lanes) / trips, "  ".join("%9.0f" % (l / B) for l in lanes)))
print("warp trips per image: %.0f" % (tot / B))
