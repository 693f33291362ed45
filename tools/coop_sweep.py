"""Tuning: ms per batch of the path stage for several label-map families as a function of RBEPWT_OPT_COOP_LIMIT (how many
of a group's largest regions get a whole warp).  GPU box.  usage: tools/coop_sweep.py [B]"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
import rbepwt_b200 as rb
from rbepwt_b200 import synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
LIMITS = [0, -1] if len(sys.argv) > 2 else [0, 74, 148, 296, 592, 1184, 2368]
fams = [("voronoi 1024", lambda s: synth.voronoi_labels(512, 512, 1024, seed=s)),
        ("voronoi 256", lambda s: synth.voronoi_labels(512, 512, 256, seed=s)),
        ("voronoi 64", lambda s: synth.voronoi_labels(512, 512, 64, seed=s)),
        ("voronoi 16", lambda s: synth.voronoi_labels(512, 512, 16, seed=s)),
        ("heavy-tailed 600", lambda s: synth.heavytail_labels(512, 600, s)),
        ("heavy-tailed 150", lambda s: synth.heavytail_labels(512, 150, s))]
for name, gen in fams:
    labs = np.stack([gen(100 + i) for i in range(8)])
    imgs = np.stack([synth.piecewise_smooth_image(l, seed=3) for l in labs])
    for b in ((16, 8, 1) if len(sys.argv) > 2 else (B, 16)):
        reps = max(b // 8, 1)
        timg = torch.from_numpy(np.concatenate([imgs] * reps)[:b]).cuda(); tlab = torch.from_numpy(np.concatenate([labs] * reps)[:b]).cuda()
        out = torch.empty_like(timg)
        res = []
        for lim in LIMITS:
            c = rb.BatchCodec()
            c.set_option(coop_limit=lim)
            c.transcode(timg, tlab, 8, "bior4.4", 2048, out=out); c.sync()
            t = time.perf_counter()
            for _ in range(3): c.transcode(timg, tlab, 8, "bior4.4", 2048, out=out)
            c.sync(); res.append((time.perf_counter() - t) / 3 * 1e3)
            c.close()
        print("%-18s B=%-4d ms/batch by coop limit %s: %s" % (name, b, LIMITS, "  ".join("%.2f" % r for r in res)), flush=True)
