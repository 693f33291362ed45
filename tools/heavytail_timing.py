"""Throughput on heavy-tailed label maps (a few 10^4-pixel background regions next to hundreds of small ones --
what Felzenszwalb produces on natural images), next to the uniform-Voronoi workload of bench.py (GPU box)."""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
import rbepwt_b200 as rb
from rbepwt_b200 import synth

heavytail_labels = synth.heavytail_labels

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
for name, gen in (("uniform voronoi 1024", lambda s: synth.voronoi_labels(512, 512, 1024, seed=s)),
                  ("heavy-tailed 600", lambda s: heavytail_labels(512, 600, s))):
    labs = np.stack([gen(100 + i) for i in range(8)])
    sizes = np.concatenate([np.bincount(l.ravel()) for l in labs]); sizes = sizes[sizes > 0]
    imgs = np.stack([synth.piecewise_smooth_image(l, seed=3) for l in labs])
    reps = B // 8
    timg = torch.from_numpy(np.concatenate([imgs] * reps)).cuda(); tlab = torch.from_numpy(np.concatenate([labs] * reps)).cuda()
    out = torch.empty_like(timg)
    c = rb.BatchCodec()
    c.transcode(timg, tlab, 16, "bior4.4", 2048, out=out); c.sync()
    c.enable_timing(True); c.timings()
    t = time.perf_counter()
    for _ in range(3): c.transcode(timg, tlab, 16, "bior4.4", 2048, out=out)
    c.sync(); dt = (time.perf_counter() - t) / 3
    st = c.timings()
    print("%-22s B=%d: %.1f ms/batch = %.0f img/s | region sizes: median %d, p99 %d, max %d | stages/batch %s" % (
        name, B, dt * 1e3, B / dt, np.median(sizes), np.percentile(sizes, 99), sizes.max(), {k: round(v / 3, 2) for k, v in st.items() if v > 0.01}))
    c.close()
