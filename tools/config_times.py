#!/usr/bin/env python
"""Latency of the non-headline BASELINE.json configurations on one GPU (device-resident inputs), and the C port
beside them: config 1 (256^2), config 3 (EPWT, 512^2 as ONE region), config 4 (2048^2, 65k regions)."""
import sys, time, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
import rbepwt_b200 as rb
from rbepwt_b200 import synth
from oracle import c_oracle

def gpu_ms(fn, reps=3):
    fn(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / reps * 1e3

rows = [("config1 256^2 easypath bior4.4 k=512", "cameraman256", "easypath", "bior4.4", 512, 1),
        ("config1 x64 batch", "cameraman256", "easypath", "bior4.4", 512, 64),
        ("config3 EPWT 512^2 haar k=2048", "epwt512", "epwt-easypath", "haar", 2048, 1),
        ("config3 EPWT 512^2 x16 batch", "epwt512", "epwt-easypath", "haar", 2048, 16),
        ("config4 2048^2 65k regions bior4.4 k=8192", "small2048", "easypath", "bior4.4", 8192, 1)]
for name, cfg, ptype, wav, k, B in rows:
    img, lab = synth.config_inputs(cfg)
    imgs = torch.from_numpy(np.stack([img] * B)).cuda()
    labs = None if lab is None else torch.from_numpy(np.stack([lab] * B)).cuda()
    out = torch.empty_like(imgs)
    c = rb.BatchCodec()
    t_all = gpu_ms(lambda: (c.transcode(imgs, labs, 16, wav, k, ptype, True, out), c.sync()))
    c.enable_timing(True); c.timings()
    c.transcode(imgs, labs, 16, wav, k, ptype, True, out); c.sync()
    st = c.timings()
    t0 = time.perf_counter()
    c_oracle.run(img, lab, 16, rb.filter_bank(wav), ptype, True, ncoefs=k)
    cpu = (time.perf_counter() - t0) * 1e3
    print("%-44s B=%-3d GPU %9.3f ms/batch (%8.3f ms/img)  stages %s | C port %8.1f ms/img" % (
        name, B, t_all, t_all / B, {k_: round(v, 3) for k_, v in st.items() if v > 0}, cpu))
    c.close()
