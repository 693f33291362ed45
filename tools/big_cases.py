"""One-off large-input parity checks against the C oracle (GPU box): huge regions that need the global-memory
bitmap scratch of k1_paths_big, and a 4096^2 image."""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import rbepwt_b200 as rb
from rbepwt_b200 import synth
from oracle import c_oracle

def check(name, img, lab, levels, wav, k):
    t = time.perf_counter()
    c = rb.BatchCodec()
    c.encode(img[None], lab[None], levels, wav)
    coefs = c.coefs(0)
    c.threshold(k)
    dec = c.decode()[0]
    tg = time.perf_counter() - t
    t = time.perf_counter()
    orc = c_oracle.run(img, lab, levels, rb.filter_bank(wav), "easypath", True, ncoefs=k)
    to = time.perf_counter() - t
    H, W = img.shape
    for lev in (1, 2, 5, levels):
        pix = c.paths(0, lev)
        want = orc["points"][lev]
        assert np.array_equal(pix, want[:, 0].astype(np.int64) * W + want[:, 1]), "%s: paths differ at level %d" % (name, lev)
    assert np.max(np.abs(coefs - orc["coefs"])) <= 1e-9 * np.abs(orc["coefs"]).max()
    assert np.array_equal(np.flatnonzero(c.coefs(0)), orc["kept"])
    assert np.max(np.abs(dec - orc["decoded"])) <= 1e-9 * 255
    print("%-40s ok   R=%d  gpu %.2fs (first call, incl. allocation)  oracle %.2fs" % (name, c.region_count(0), tg, to))

lab = synth.voronoi_labels(2048, 2048, 16, seed=5, warp=6.0)
check("2048^2, 16 huge regions (gscratch)", synth.piecewise_smooth_image(lab, seed=5), lab, 16, "haar", 4096)
lab = synth.voronoi_labels(1024, 1024, 6, seed=6, warp=10.0)
lab[::7, ::5] = 99  # plus one scattered label class spanning the whole image
check("1024^2, 6 huge + 1 scattered region", synth.piecewise_smooth_image(lab, seed=6), lab, 14, "bior4.4", 1000)
lab = synth.voronoi_labels(4096, 4096, 50000, seed=7, warp=2.0)
check("4096^2, 50k regions", synth.piecewise_smooth_image(lab, seed=7), lab, 16, "bior4.4", 20000)
