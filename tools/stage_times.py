"""Stage times (serial, one stream) of one transcode for a label-map family.  GPU box.
usage: tools/stage_times.py <family> <seeds> <B> [levels] [coop_limit]"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
import rbepwt_b200 as rb
from rbepwt_b200 import synth

fam, seeds, B = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
levels = int(sys.argv[4]) if len(sys.argv) > 4 else 8
gen = (lambda s: synth.heavytail_labels(512, seeds, s)) if fam == "heavytail" else (lambda s: synth.voronoi_labels(512, 512, seeds, seed=s))
labs = np.stack([gen(100 + i) for i in range(min(B, 8))])
imgs = np.stack([synth.piecewise_smooth_image(l, seed=3) for l in labs])
reps = max(B // 8, 1)
timg = torch.from_numpy(np.concatenate([imgs] * reps)[:B]).cuda(); tlab = torch.from_numpy(np.concatenate([labs] * reps)[:B]).cuda()
out = torch.empty_like(timg)
c = rb.BatchCodec()
if len(sys.argv) > 5: c.set_option(coop_limit=int(sys.argv[5]))
c.set_option(streams=1)
c.transcode(timg, tlab, levels, "bior4.4", 2048, out=out); c.sync()
c.enable_timing(True); c.timings()
t = time.perf_counter()
for _ in range(3): c.transcode(timg, tlab, levels, "bior4.4", 2048, out=out)
c.sync(); dt = (time.perf_counter() - t) / 3
st = c.timings()
sizes = np.concatenate([np.bincount(l.ravel()) for l in labs]); sizes = sizes[sizes > 0]
print("%s %d B=%d L=%d: %.2f ms | sizes median %d max %d | %s" % (fam, seeds, B, levels, dt * 1e3, np.median(sizes), sizes.max(), {k: round(v / 3, 2) for k, v in st.items() if v > 0.01}))
