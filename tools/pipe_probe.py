#!/usr/bin/env python
"""(GPU) Stage times of the default pipelined mode next to the one-stream mode, same batch: which stage pays for
the overlap.  usage: tools/pipe_probe.py [batch]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import rbepwt_b200 as rb
from rbepwt_b200 import synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
imgs, labs = synth.torch_batch(B, 512, 512, 1024, 1000, device="cuda")
out = torch.empty_like(imgs)
stream = torch.cuda.Stream()  # the context runs on it and the timing events are recorded on it (a context's private
c = rb.BatchCodec(stream=stream.cuda_stream)  # stream is not ordered with torch's default stream: events there see nothing)
def step():
    c.transcode(imgs, labs, 16, "bior4.4", 2048, "easypath", True, out)
for streams in (2, 1, 2, 1):
    c.set_option(streams=streams)
    for timing in (False, True):
        c.enable_timing(timing)
        for _ in range(3):
            step()
        c.sync(); c.timings()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(stream)
        for _ in range(10):
            step()
        e1.record(stream)
        torch.cuda.synchronize()
        t = c.timings() if timing else {}
        print("streams %d timing %d: %.2f ms/step  %s" % (streams, timing, e0.elapsed_time(e1) / 10,
              " ".join("%s %.2f" % (k, v / 10) for k, v in t.items() if v > 0)))
