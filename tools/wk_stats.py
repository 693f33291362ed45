#!/usr/bin/env python
"""Debug: build the library with -DWK_STATS into a scratch .so and print k1_walk's warp trips and lane units by
kind for one batch of the bench workload (run on the GPU box)."""
import ctypes, os, subprocess, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from rbepwt_b200 import build as b
so = "/tmp/librbepwt_stats.so"
extra = [a for a in sys.argv[2:] if a.startswith("-D")]
fam = [a for a in sys.argv[2:] if not a.startswith("-D")]
subprocess.check_call(["nvcc"] + b.NVCC_FLAGS + ["-DWK_STATS"] + extra + ["-o", so, os.path.join(b.CSRC, "rbepwt_b200.cu")])
b.LIB_PATH = so
b.needs_build = lambda: False
import numpy as np, torch
import rbepwt_b200 as rb
from rbepwt_b200 import synth, _capi
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
imgs, labs = synth.torch_batch(B, 512, 512, 1024, 1000, device="cuda")
if fam:  # e.g. "heavytail 600" or "voronoi 64"
    gen = (lambda s: synth.heavytail_labels(512, int(fam[1]), s)) if fam[0] == "heavytail" else (lambda s: synth.voronoi_labels(512, 512, int(fam[1]), seed=s))
    l8 = np.stack([gen(100 + i) for i in range(8)])
    labs = torch.from_numpy(np.concatenate([l8] * (B // 8))).cuda()
c = rb.BatchCodec()
L = _capi.lib()
out = (ctypes.c_ulonglong * 16)()
c.encode(imgs, labs, 8, "bior4.4"); c.sync()
L.rbepwt_debug_wk_stats(c._ctx, out, 1)
c.encode(imgs, labs, 8, "bior4.4"); c.sync()
L.rbepwt_debug_wk_stats(c._ctx, out, 1)
v = np.array(list(out), dtype=np.float64) / B
names = ["done", "near", "far", "list", "level", "commit", "error"]
print("per image: warp trips %.0f, regions per chunk %.1f | lanes per trip at its start: %s"
      % (v[0], v[10] / v[0], "  ".join("%s %.1f" % (n, v[1 + i] / v[0]) for i, n in enumerate(names))))
print("per image: searches beyond the window: %.0f (narrow planes) + %.0f (wide planes); bitmap-mode steps %.0f; trips of the windowed instantiation %.0f"
      % (v[11], v[12], v[13], v[14]))
