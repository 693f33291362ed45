"""EPWT (one 512^2 image = one region): time as a function of the number of levels.  GPU box."""
import sys, time, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
import rbepwt_b200 as rb
from rbepwt_b200 import synth
img, _ = synth.config_inputs("epwt512")
imgs = torch.from_numpy(img[None].copy()).cuda(); out = torch.empty_like(imgs)
c = rb.BatchCodec()
prev = 0.0
for L in (1, 2, 3, 4, 5, 6, 8, 12, 16):
    c.transcode(imgs, None, L, "haar", 2048, "epwt-easypath", True, out); c.sync()
    t = time.perf_counter()
    for _ in range(2): c.transcode(imgs, None, L, "haar", 2048, "epwt-easypath", True, out)
    c.sync(); ms = (time.perf_counter() - t) / 2 * 1e3
    print("levels %2d: %.1f ms (+%.1f)" % (L, ms, ms - prev), flush=True); prev = ms
