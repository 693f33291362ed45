import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
import rbepwt_b200 as rb
from rbepwt_b200 import synth
for (n, seeds, B) in ((2048, 16, 1), (512, 4, 1), (512, 4, 64), (512, 40, 64)):
    lab = synth.voronoi_labels(n, n, seeds, seed=5, warp=6.0)
    img = synth.piecewise_smooth_image(lab, seed=5)
    imgs = torch.from_numpy(np.stack([img] * B)).cuda(); labs = torch.from_numpy(np.stack([lab] * B)).cuda(); out = torch.empty_like(imgs)
    c = rb.BatchCodec()
    c.transcode(imgs, labs, 16, "haar", 4096, out=out); c.sync()
    c.enable_timing(True); c.timings()
    t = time.perf_counter(); c.transcode(imgs, labs, 16, "haar", 4096, out=out); c.sync(); dt = time.perf_counter() - t
    st = c.timings()
    sizes = np.bincount(lab.ravel() - lab.min())
    print("%d^2 %3d regions (largest %d px) B=%-3d: %.1f ms  %s" % (n, seeds, sizes.max(), B, dt * 1e3, {k: round(v, 2) for k, v in st.items() if v > 0.005}))
    c.close()
