"""EPWT (BASELINE config 3: one 512^2 image = one region) latency with the library named by RBEPWT_B200_LIB.  GPU box."""
import sys, time, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
import rbepwt_b200 as rb
from rbepwt_b200 import synth
img, _ = synth.config_inputs("epwt512")
for B in (1, 16):
    imgs = torch.from_numpy(np.stack([img] * B)).cuda(); out = torch.empty_like(imgs)
    c = rb.BatchCodec()
    c.transcode(imgs, None, 16, "haar", 2048, "epwt-easypath", True, out); c.sync()
    t = time.perf_counter()
    for _ in range(3): c.transcode(imgs, None, 16, "haar", 2048, "epwt-easypath", True, out)
    c.sync()
    print("EPWT 512^2 B=%d: %.1f ms/batch" % (B, (time.perf_counter() - t) / 3 * 1e3), flush=True)
    c.close()
