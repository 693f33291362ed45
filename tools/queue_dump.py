"""Debug (-DWK_STATS build): the queue tables of both path slots after one 512-image call.  GPU box."""
import ctypes, os, subprocess, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from rbepwt_b200 import build as b
so = "/tmp/librbepwt_stats.so"
subprocess.check_call(["nvcc"] + b.NVCC_FLAGS + ["-DWK_STATS", "-o", so, os.path.join(b.CSRC, "rbepwt_b200.cu")])
b.LIB_PATH = so
b.needs_build = lambda: False
import numpy as np, torch
import rbepwt_b200 as rb
from rbepwt_b200 import synth, _capi
imgs, labs = synth.torch_batch(512, 512, 512, 1024, 1000, device="cuda")
c = rb.BatchCodec()
L = _capi.lib()
out = torch.empty_like(imgs)
for it in range(2):
    c.transcode(imgs, labs, 16, "bior4.4", 2048, "easypath", True, out); c.sync()
    for slot in (0, 1):
        buf = (ctypes.c_int * 64)()
        L.rbepwt_debug_queue(c._ctx, slot, buf)
        v = list(buf)
        names = ["NBIG", "NREG", "CUR_BIG", "CUR_SMALL", "ERR", "NCHUNKS", "CHUNK_SPLIT", "CUR_WIDE", "COOP_SIZE", "CLS1_CHUNKS", "CUR_COOP"]
        print("call %d slot %d:" % (it, slot), dict(zip(names, v[:11])))
        print("   class counts:", v[16 + 13:16 + 26])
