"""The C oracle (oracle/rbepwt_oracle.c) must reproduce what the unmodified reference produced
(tests/golden/*.npz): paths, permutations, kept indices bit-exact; coefficients, pixels 1e-9."""
import numpy as np
import pytest

from conftest import assert_matches_golden, golden_names, load_golden
from oracle import c_oracle


@pytest.mark.parametrize("name", golden_names())
def test_c_oracle_reproduces_reference(name):
    g = load_golden(name)
    out = c_oracle.run(g["img"], g["labels"], g["levels"], g["wavelet"], g["path_type"],
                       g["euclidean_distance"], ncoefs=g["ncoefs"], paths_first_level=g["paths_first_level"])
    assert_matches_golden(out, g)


def test_golden_set_is_present():
    names = golden_names()
    assert "survey11_4x4_haar" in names and len(names) >= 25


def test_oracle_guards():
    img = np.zeros((6, 8))
    lab = np.zeros((6, 8), np.int32)
    with pytest.raises(Exception, match="power of 2"):
        c_oracle.encode(img, lab, 2, "haar", c_oracle.MODE_EUCLID)
    with pytest.raises(Exception, match="levels"):
        c_oracle.encode(np.zeros((4, 4)), np.zeros((4, 4), np.int32), 5, "haar", c_oracle.MODE_EUCLID)


def test_threshold_quirks():
    x = np.array([3.0, -5.0, 1.0, 5.0, 0.5])
    np.testing.assert_array_equal(c_oracle.threshold(x, 0), x)       # k=0 keeps everything
    np.testing.assert_array_equal(c_oracle.threshold(x, 5), x)
    np.testing.assert_array_equal(c_oracle.threshold(x, 99), x)
    np.testing.assert_array_equal(c_oracle.threshold(x, 1), [0, 0, 0, 5.0, 0])  # tie: highest index
    np.testing.assert_array_equal(c_oracle.threshold(x, 3), [3.0, -5.0, 0, 5.0, 0])


def test_oracle_threshold_by_percentage_reproduces_reference():
    from conftest import load_perc, perc_names
    from oracle import c_oracle, pywt_port

    assert perc_names()
    for name in perc_names():
        g = load_perc(name)
        fb = pywt_port.filter_bank(g["wavelet"])
        enc = c_oracle.encode(g["img"], g["labels"], g["levels"], fb, c_oracle.path_mode(g["path_type"], g["euclidean_distance"]))
        th = c_oracle.threshold_percentage(enc, enc["coefs"], g["perc"])
        np.testing.assert_array_equal(np.flatnonzero(th), np.flatnonzero(g["thresholded"]), err_msg=name)
        assert np.max(np.abs(th - g["thresholded"])) <= 1e-9 * np.abs(g["thresholded"]).max()
        dec = c_oracle.decode(enc, th, fb)
        assert np.max(np.abs(dec - g["decoded"])) <= 1e-9 * 255
        assert abs(c_oracle.psnr(g["img"], dec) - g["psnr"]) < 5e-7


def test_oracle_dwt2_baseline():
    """The 2-D baseline's restatement (oracle/pywt_port.py wavedec2 / waverec2; reference class Dwt, rbepwt.py:2249-2298).
    PyWavelets is not in this image, so this is pinned through what pywt documents: dwt2 = the 1-D periodized dwt (which
    the golden fixtures pin) along axis 0 then axis 1, quadrant names (cH = high along rows, cV = high along columns),
    the Haar closed form, orthogonality and perfect reconstruction."""
    from oracle import pywt_port

    rng = np.random.default_rng(3)
    x = rng.uniform(0, 255, (32, 32))
    co = pywt_port.wavedec2(x, "haar", 1)
    a, b, c, d = x[0::2, 0::2], x[0::2, 1::2], x[1::2, 0::2], x[1::2, 1::2]
    np.testing.assert_allclose(co[0], (a + b + c + d) / 2, atol=1e-12 * 255)
    cH, cV, cD = co[1]
    np.testing.assert_allclose(np.abs(cH), np.abs(a + b - c - d) / 2, atol=1e-12 * 255)  # detail along the rows axis
    np.testing.assert_allclose(np.abs(cV), np.abs(a - b + c - d) / 2, atol=1e-12 * 255)
    np.testing.assert_allclose(np.abs(cD), np.abs(a - b - c + d) / 2, atol=1e-12 * 255)
    for wav, lev in (("db3", 3), ("bior4.4", 5), ("haar", 5)):
        co = pywt_port.wavedec2(x, wav, lev)
        assert co[0].shape == (32 >> lev, 32 >> lev) and len(co) == lev + 1
        np.testing.assert_allclose(pywt_port.waverec2(co, wav), x, atol=1e-9 * 255)
        if wav != "bior4.4":  # orthogonal banks keep the energy
            e = np.sum(co[0] ** 2) + sum(np.sum(q ** 2) for t in co[1:] for q in t)
            assert abs(e - np.sum(x ** 2)) <= 1e-9 * np.sum(x ** 2)
    # separable: one level = 1-D dwt of every column, then of every row
    lo = np.stack([pywt_port.dwt(x[:, j], "db3")[0] for j in range(32)], axis=1)
    aa = np.stack([pywt_port.dwt(lo[i], "db3")[0] for i in range(16)], axis=0)
    np.testing.assert_allclose(pywt_port.wavedec2(x, "db3", 1)[0], aa, atol=1e-12 * 255)
    dec, nz, mags = pywt_port.dwt2_baseline(x, 3, "db3", 40)
    assert nz == 40 and len(mags) == 40 and dec.shape == x.shape


def _roi_fixtures():
    import glob
    import os

    from conftest import GOLDEN_DIR
    return sorted(glob.glob(os.path.join(GOLDEN_DIR, "roi", "*.npz")))


def test_roi_select_reproduces_reference():
    """Region-of-interest thresholding (Roi.compute_dual_roi_coeffs, rbepwt.py:1718-1789): the host-side selection
    (rbepwt_b200/roi.py, numpy) on the oracle's paths and coefficients against what the unmodified reference kept
    (tests/golden/roi/*.npz, tests/golden/make_golden.py --roi): counts and the set of surviving coefficients."""
    import rbepwt_b200 as rb
    from rbepwt_b200.roi import global_perm, roi_select

    files = _roi_fixtures()
    assert len(files) >= 5
    for f in files:
        z = np.load(f)
        img, lab, L = z["img"], z["labels"], int(z["levels"])
        out = c_oracle.run(img, lab, L, rb.filter_bank(str(z["wavelet"])), "easypath", bool(z["euclidean_distance"]))
        N = img.size
        # `regions` are region indices (rank of first appearance), whatever the label values are
        off = out["roff"][1]
        in_mask = np.zeros(N, dtype=bool)
        for r in set(int(x) for x in z["regions"]):
            if r < off.size - 1:
                in_mask[off[r]:off[r + 1]] = True
        flat = out["coefs"].copy()
        det_off = {lev: N - (N >> (lev - 1)) for lev in range(1, L + 2)}
        details = {lev: flat[det_off[lev]:det_off[lev + 1]] for lev in range(1, L + 1)}
        perms = {lev: global_perm(out["perm"][lev], out["roff"][lev]) for lev in range(2, L + 1)}
        keep, nin, nout = roi_select(details, in_mask, perms, float(z["perc_in"]), float(z["perc_out"]))
        assert (nin, nout) == (int(z["nin"]), int(z["nout"])), f
        for lev in range(1, L + 1):
            details[lev][~keep[lev]] = 0.0  # views into flat
        want = z["thresholded"]
        np.testing.assert_array_equal(flat != 0, want != 0, err_msg=f)
        assert np.max(np.abs(flat - want)) <= 1e-9 * np.max(np.abs(want)), f
    with pytest.raises(Exception, match="between 0 and 1"):
        roi_select({1: np.zeros(2)}, np.zeros(4, dtype=bool), {}, 1.5, 0.0)
