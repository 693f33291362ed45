"""Debug: paths of images of BOTH path groups of a 512-image device-resident batch against the CPU port (fresh context,
one call: nothing stale to hide behind).  GPU box."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
import rbepwt_b200 as rb
from rbepwt_b200 import synth
from oracle import c_oracle
B = 512
imgs, labs = synth.torch_batch(B, 512, 512, 1024, 1000, device="cuda")
c = rb.BatchCodec()
out = torch.empty_like(imgs)
c.transcode(imgs, labs, 16, "bior4.4", 2048, "easypath", True, out); c.sync()
bad = 0
for b in (0, 100, 255, 256, 300, 511):
    o = c_oracle.run(imgs[b].cpu().numpy(), labs[b].cpu().numpy(), 16, rb.filter_bank("bior4.4"), "easypath", True, ncoefs=2048)
    for lev in range(1, 17):
        pix = c.paths(b, lev)
        if not np.array_equal(np.stack([pix // 512, pix % 512], axis=1), o["points"][lev]):
            print("image %d level %d: paths differ" % (b, lev)); bad += 1; break
    else:
        print("image %d ok, decoded maxdiff %.3g" % (b, np.max(np.abs(out[b].cpu().numpy() - o["decoded"]))))
print("BAD" if bad else "ALL OK")
