"""Stage durations in PIPELINED mode (streams = 2): a stage's CUDA-event time then includes what it waited for, so the sum
says where the step's time goes when everything competes.  GPU box.  usage: RBEPWT_B200_LIB=... tools/pipe_stages.py"""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
import rbepwt_b200 as rb
from rbepwt_b200 import synth
imgs, labs = synth.torch_batch(512, 512, 512, 1024, 1000, device="cuda")
out = torch.empty_like(imgs)
c = rb.BatchCodec()
for _ in range(3): c.transcode(imgs, labs, 16, "bior4.4", 2048, "easypath", True, out)
c.sync()
c.enable_timing(True); c.timings()
t = time.perf_counter()
for _ in range(5): c.transcode(imgs, labs, 16, "bior4.4", 2048, "easypath", True, out)
c.sync(); dt = (time.perf_counter() - t) / 5
st = c.timings()
print("%.2f ms/step | per step, pipelined: %s" % (dt * 1e3, {k: round(v / 5, 2) for k, v in st.items() if v > 0.01}))
