"""TEST INFRASTRUCTURE ONLY -- ctypes front end of oracle/rbepwt_oracle.c.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this.  The product (rbepwt_b200/) never does.
"""
import ctypes
import os
import subprocess

import numpy as np

from . import pywt_port

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None

MODE_EUCLID, MODE_CHEB, MODE_EPWT, MODE_GRAD_EUCLID, MODE_GRAD_CHEB = 0, 1, 2, 3, 4

_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")


def build(force=False):
    src = os.path.join(_HERE, "rbepwt_oracle.c")
    if force or not os.path.isfile(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "all"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        L.rbo_count_regions.restype = ctypes.c_int
        L.rbo_count_regions.argtypes = [_i32p, ctypes.c_int, ctypes.c_int]
        L.rbo_level_offset.restype = ctypes.c_long
        L.rbo_level_offset.argtypes = [ctypes.c_long, ctypes.c_int]
        L.rbo_encode.restype = ctypes.c_int
        L.rbo_encode.argtypes = [_f64p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                 ctypes.c_int, _f64p, _f64p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                 _i32p, _i32p, _i32p, _i32p, _f64p]
        L.rbo_threshold.restype = ctypes.c_int
        L.rbo_threshold.argtypes = [_f64p, ctypes.c_int64, ctypes.c_int64]
        L.rbo_decode.restype = ctypes.c_int
        L.rbo_decode.argtypes = [_f64p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, _f64p,
                                 _f64p, ctypes.c_int, _i32p, _i32p, _i32p, _f64p]
        L.rbo_threshold_percentage.restype = ctypes.c_int
        L.rbo_threshold_percentage.argtypes = [_f64p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, _i32p, ctypes.c_double]
        L.rbo_grad_tie_events.restype = ctypes.c_long
        L.rbo_grad_tie_events.argtypes = []
        L.rbo_psnr.restype = ctypes.c_double
        L.rbo_psnr.argtypes = [_f64p, _f64p, ctypes.c_int64]
        _lib = L
    return _lib


def path_mode(path_type, euclidean_distance):
    if path_type == "epwt-easypath":
        return MODE_EPWT
    if path_type == "easypath":
        return MODE_EUCLID if euclidean_distance else MODE_CHEB
    if path_type == "gradpath":
        return MODE_GRAD_EUCLID if euclidean_distance else MODE_GRAD_CHEB
    raise ValueError("unknown path_type %r" % (path_type,))


def level_offset(n, lev):
    return sum(n >> (l - 1) for l in range(1, lev))


def encode(img, labels, levels, wavelet, mode, u8wrap=None, paths_first_level=False):
    """Flat-array encode.  Returns dict(R, roff[(L+1),(R+1)], inc_pix, path_pix, perm, coefs)."""
    L = lib()
    if u8wrap is None:
        u8wrap = np.asarray(img).dtype == np.uint8
    img64 = np.ascontiguousarray(img, dtype=np.float64)
    H, W = img64.shape
    N = H * W
    dec_lo, dec_hi, _, _ = pywt_port.filter_bank(wavelet)
    if mode == MODE_EPWT:
        lab, labp, R = None, None, 1
    else:
        lab = np.ascontiguousarray(labels, dtype=np.int32)
        labp = lab.ctypes.data_as(ctypes.c_void_p)
        R = L.rbo_count_regions(lab, H, W)
    roff = np.zeros((levels + 1, R + 1), dtype=np.int32)
    inc_pix = np.zeros(level_offset(N, levels + 2), dtype=np.int32)
    path_pix = np.zeros(level_offset(N, levels + 1), dtype=np.int32)
    perm = np.zeros(level_offset(N, levels + 1), dtype=np.int32)
    coefs = np.zeros(N, dtype=np.float64)
    rc = L.rbo_encode(img64, labp, H, W, levels, len(dec_lo), np.ascontiguousarray(dec_lo),
                      np.ascontiguousarray(dec_hi), mode, int(bool(u8wrap)), int(bool(paths_first_level)), R, roff, inc_pix,
                      path_pix, perm, coefs)
    if rc == -1:
        raise Exception("Image size must be a power of 2")
    if rc == -2:
        raise Exception("2^levels must be smaller or equal to the number of pixels in the image")
    if rc == -4:
        raise ValueError("Shape of array too small to calculate a numerical gradient, at least 2 elements are required.")
    if rc:
        raise RuntimeError("rbo_encode failed: %d" % rc)
    out = dict(R=R, H=H, W=W, levels=levels, roff=roff, inc_pix=inc_pix, path_pix=path_pix, perm=perm, coefs=coefs)
    if mode in (MODE_GRAD_EUCLID, MODE_GRAD_CHEB):
        out["grad_tie_events"] = int(L.rbo_grad_tie_events())  # complete ties: the reference's answer is unpinned there
    return out


def threshold(coefs, k):
    out = np.array(coefs, dtype=np.float64, copy=True)
    lib().rbo_threshold(out, out.size, int(k))
    return out


def threshold_percentage(enc, coefs, perc):
    """Rbepwt.threshold_by_percentage (rbepwt.py:2120-2192) on a flat coefficient vector of the encoding `enc`."""
    out = np.array(coefs, dtype=np.float64, copy=True)
    lib().rbo_threshold_percentage(out, enc["H"], enc["W"], enc["levels"], enc["R"], np.ascontiguousarray(enc["roff"]), float(perc))
    return out


def decode(enc, coefs, wavelet):
    _, _, rec_lo, rec_hi = pywt_port.filter_bank(wavelet)
    H, W, N = enc["H"], enc["W"], enc["H"] * enc["W"]
    out = np.zeros((H, W), dtype=np.float64)
    lib().rbo_decode(np.ascontiguousarray(coefs, dtype=np.float64), H, W, enc["levels"], len(rec_lo),
                     np.ascontiguousarray(rec_lo), np.ascontiguousarray(rec_hi), enc["R"],
                     enc["roff"], np.ascontiguousarray(enc["inc_pix"][:N]), enc["perm"], out)
    return out


def psnr(a, b):
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    return float(lib().rbo_psnr(a.ravel(), b.ravel(), a.size))


def run(img, labels, levels, wavelet, path_type="easypath", euclidean_distance=True, ncoefs=None,
        paths_first_level=False):
    """Same output dict as ref_harness.run_reference (perm/roff/points per level, coefs, kept,
    decoded, psnr) so the two can be compared key by key."""
    mode = path_mode(path_type, euclidean_distance)
    enc = encode(img, labels, levels, wavelet, mode, paths_first_level=paths_first_level)
    H, W = enc["H"], enc["W"]
    N = H * W
    out = {"perm": {}, "roff": {}, "points": {}, "enc": enc}
    for lev in range(1, levels + 2):
        lo, nl = level_offset(N, lev), N >> (lev - 1)
        out["roff"][lev] = enc["roff"][lev - 1].copy()
        src = enc["path_pix"] if lev <= levels else enc["inc_pix"]
        pix = src[lo:lo + nl]
        out["points"][lev] = np.stack([pix // W, pix % W], axis=1).astype(np.int32)
        if lev <= levels:
            out["perm"][lev] = enc["perm"][lo:lo + nl].copy()
    out["coefs"] = enc["coefs"].copy()
    if ncoefs is not None:
        th = threshold(enc["coefs"], ncoefs)
        out["thresholded"] = th
        out["kept"] = np.flatnonzero(th != 0).astype(np.int64)
        dec = decode(enc, th, wavelet)
        out["decoded"] = dec
        out["psnr"] = psnr(np.asarray(img, dtype=np.float64), dec)
        out["nonzero_coefs"] = int(np.count_nonzero(th))
    return out
