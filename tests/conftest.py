import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load_golden(name):
    """Fixture produced by the unmodified reference (tests/golden/make_golden.py)."""
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    g = {k: z[k] for k in z.files}
    for k in ("levels", "ncoefs", "nonzero_coefs"):
        g[k] = int(g[k])
    for k in ("wavelet", "path_type"):
        g[k] = str(g[k])
    g["euclidean_distance"] = bool(g["euclidean_distance"])
    g["paths_first_level"] = bool(g["paths_first_level"]) if "paths_first_level" in g else False
    g["psnr"] = float(g["psnr"])
    if g["labels"].size == 0:
        g["labels"] = None
    n = g["img"].size
    L = g["levels"]
    lens = [n >> (l - 1) for l in range(1, L + 2)]
    offs = np.concatenate([[0], np.cumsum(lens)])
    g["perm_by_level"] = {l: g["perm"][offs[l - 1]:offs[l]] for l in range(1, L + 1)}
    g["points_by_level"] = {l: g["points"][offs[l - 1]:offs[l]].astype(np.int32) for l in range(1, L + 2)}
    g["roff_by_level"] = {l: g["roff"][l - 1] for l in range(1, L + 2)}
    return g


def assert_matches_golden(out, g, rtol=1e-9):
    """`out` in the layout of oracle.c_oracle.run / ref_harness.run_reference.
    Paths, permutations, region offsets, kept indices: bit-exact.  Coefficients and decoded
    pixels: 1e-9 relative (to the largest magnitude).  PSNR: 6 decimals."""
    L = g["levels"]
    for l in range(1, L + 2):
        np.testing.assert_array_equal(out["roff"][l], g["roff_by_level"][l], err_msg="roff level %d" % l)
        np.testing.assert_array_equal(out["points"][l], g["points_by_level"][l], err_msg="points level %d" % l)
    for l in range(1, L + 1):
        np.testing.assert_array_equal(out["perm"][l], g["perm_by_level"][l], err_msg="perm level %d" % l)
    scale = np.max(np.abs(g["coefs"]))
    assert np.max(np.abs(out["coefs"] - g["coefs"])) <= rtol * scale
    np.testing.assert_array_equal(out["kept"], g["kept"])
    assert np.max(np.abs(out["thresholded"] - g["thresholded"])) <= rtol * scale
    assert np.max(np.abs(out["decoded"] - g["decoded"])) <= rtol * 255.0
    assert abs(out["psnr"] - g["psnr"]) < 5e-7
    assert out["nonzero_coefs"] == g["nonzero_coefs"]


def perc_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "perc", "*.npz")))


def load_perc(name):
    """threshold_by_percentage fixture produced by the unmodified reference (tests/golden/make_golden.py --perc)."""
    z = np.load(os.path.join(GOLDEN_DIR, "perc", name + ".npz"))
    g = {k: z[k] for k in z.files}
    g["levels"] = int(g["levels"]); g["perc"] = float(g["perc"]); g["psnr"] = float(g["psnr"])
    g["wavelet"] = str(g["wavelet"]); g["path_type"] = str(g["path_type"]); g["euclidean_distance"] = bool(g["euclidean_distance"])
    if g["labels"].size == 0:
        g["labels"] = None
    return g
