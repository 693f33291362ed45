import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
import rbepwt_b200 as rb
from rbepwt_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
img = synth.smooth_field_image(n, n, seed=2)
c = rb.BatchCodec()
t = torch.from_numpy(img[None].copy()).cuda(); out = torch.empty_like(t)
for _ in range(2):
    c.transcode(t, None, 2, "haar", 100, "epwt-easypath", True, out); c.sync()
print("ok")
