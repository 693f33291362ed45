"""Helpers for the -m gpu parity tests: run the CUDA path through the public Python API (which calls
the C ABI) and lay the results out like oracle.c_oracle.run / ref_harness.run_reference."""
import numpy as np


def cuda_run(img, labels, levels, wavelet, path_type="easypath", euclidean_distance=True, ncoefs=None,
             with_perm=True, paths_first_level=False, copies=1, which=0):
    """One image through the batch API; `copies` > 1 submits that many copies as one batch (more than 4096 regions
    in a path group make the thread-per-region kernels walk them instead of a warp per region) and returns the
    results of copy `which`."""
    import rbepwt_b200 as rb

    img = np.asarray(img)
    H, W = img.shape
    c = rb.BatchCodec()
    imgs = np.ascontiguousarray(np.broadcast_to(img[None], (copies, H, W)))
    labs = None if labels is None else np.ascontiguousarray(np.broadcast_to(np.asarray(labels)[None], (copies, H, W)))
    c.encode(imgs, labs, levels, wavelet, path_type, euclidean_distance, paths_first_level=paths_first_level)
    return collect(c, which, imgs[which], levels, ncoefs, with_perm)


def collect(c, b, img, levels, ncoefs=None, with_perm=True):
    """Everything the codec holds for image `b` of its encoded batch, in the oracle's layout.  With `ncoefs` the
    whole batch is thresholded and decoded (state changes: call it for the images of interest AFTER reading
    whatever else is needed from the unthresholded encoding)."""
    W = c.shape[2]
    out = {"perm": {}, "roff": {}, "points": {}, "codec": c}
    for lev in range(1, levels + 2):
        out["roff"][lev] = c.region_offsets(b, lev)
        pix = c.paths(b, lev)
        out["points"][lev] = np.stack([pix // W, pix % W], axis=1).astype(np.int32)
        if lev <= levels and with_perm:
            out["perm"][lev] = c.perm(b, lev)
    out["coefs"] = c.coefs(b)
    if ncoefs is not None:
        c.threshold(ncoefs)
        th = c.coefs(b)
        out["thresholded"] = th
        out["kept"] = np.flatnonzero(th != 0).astype(np.int64)
        dec = c.decode()[b]
        out["decoded"] = dec
        out["psnr"] = float(c.psnr(np.asarray(img, dtype=np.float64)[None], dec[None])[0])
        out["nonzero_coefs"] = int(c.nonzero_coefs()[b])
    return out


def collect_batch(c, imgs, which, levels, ncoefs, with_perm=True):
    """collect() for several images of ONE encoded batch: everything unthresholded first, then one threshold +
    decode of the batch."""
    outs = {b: collect(c, b, imgs[b], levels, None, with_perm) for b in which}
    c.threshold(ncoefs)
    dec = c.decode()
    nz = c.nonzero_coefs()
    for b, out in outs.items():
        th = c.coefs(b)
        out["thresholded"] = th
        out["kept"] = np.flatnonzero(th != 0).astype(np.int64)
        out["decoded"] = dec[b]
        out["psnr"] = float(c.psnr(np.asarray(imgs[b], dtype=np.float64)[None], dec[b][None])[0])
        out["nonzero_coefs"] = int(nz[b])
    return outs


def assert_same_as_oracle(out, orc, levels, coef_rtol=1e-12):
    """CUDA vs C oracle on identical inputs and identical filter banks: integer work bit-exact; the
    fp64 arithmetic follows the same operation order, so coefficients agree to rounding."""
    for lev in range(1, levels + 2):
        np.testing.assert_array_equal(out["roff"][lev], orc["roff"][lev], err_msg="roff level %d" % lev)
        np.testing.assert_array_equal(out["points"][lev], orc["points"][lev], err_msg="points level %d" % lev)
    for lev in out["perm"]:
        np.testing.assert_array_equal(out["perm"][lev], orc["perm"][lev], err_msg="perm level %d" % lev)
    scale = np.max(np.abs(orc["coefs"]))
    assert np.max(np.abs(out["coefs"] - orc["coefs"])) <= coef_rtol * scale
    if "kept" in orc:
        np.testing.assert_array_equal(out["kept"], orc["kept"])
        assert np.max(np.abs(out["decoded"] - orc["decoded"])) <= 1e-9 * 255
        assert abs(out["psnr"] - orc["psnr"]) < 5e-7
