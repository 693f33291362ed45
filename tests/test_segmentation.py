"""Felzenszwalb segmentation (Image.segment, rbepwt.py:220-245, 779-785): the library's host-side restatement of
scikit-image's felzenszwalb (csrc/segment.hpp, C ABI rbepwt_felzenszwalb) against an independently written numpy
restatement (oracle/fh_port.py) and the algorithm's invariants.  Parity with scikit-image itself is UNPINNED: it is a
third-party dependency that is not available here.  No GPU needed: the function is host code."""
import numpy as np
import pytest

from oracle import fh_port


def _cases():
    from rbepwt_b200 import synth

    rng = np.random.default_rng(5)
    lab = synth.voronoi_labels(48, 48, 14, seed=4)
    smooth = synth.piecewise_smooth_image(lab, seed=4)
    yield "piecewise u8", np.round(smooth).astype(np.uint8), 200, 2.0, 10
    yield "piecewise u8 fine", np.round(smooth).astype(np.uint8), 20, 0.8, 2
    yield "smooth field u8", np.round(synth.smooth_field_image(64, 32, seed=6, sigma=3.0)).astype(np.uint8), 100, 1.0, 5
    yield "float in [0,1]", synth.smooth_field_image(32, 32, seed=7, sigma=2.0) / 255.0, 50, 1.5, 8
    yield "noise u8", rng.integers(0, 256, size=(24, 40)).astype(np.uint8), 300, 0.5, 4
    yield "constant", np.full((16, 16), 9, dtype=np.uint8), 200, 2.0, 10
    yield "no smoothing", np.round(smooth).astype(np.uint8), 150, 0.0, 6


def _components_8(lab):
    """number of 8-connected components of equal labels (union-find over right / down / diagonal neighbours)"""
    H, W = lab.shape
    parent = np.arange(H * W)

    def find(i):
        while parent[i] != i:
            parent[i] = parent[parent[i]]
            i = parent[i]
        return i

    for i in range(H):
        for j in range(W):
            for di, dj in ((0, 1), (1, 0), (1, 1), (1, -1)):
                a, b = i + di, j + dj
                if a < H and 0 <= b < W and lab[i, j] == lab[a, b]:
                    ra, rb = find(i * W + j), find(a * W + b)
                    if ra != rb:
                        parent[max(ra, rb)] = min(ra, rb)
    return len({find(i) for i in range(H * W)})


@pytest.mark.parametrize("name,img,scale,sigma,min_size", list(_cases()), ids=[c[0] for c in _cases()])
def test_felzenszwalb_matches_restatement_and_invariants(name, img, scale, sigma, min_size):
    import rbepwt_b200 as rb

    got = rb.felzenszwalb_labels(img, scale, sigma, min_size)
    img01 = img.astype(np.float64) / 255.0 if img.dtype == np.uint8 else img.astype(np.float64)
    want = fh_port.felzenszwalb(img01, scale, sigma, min_size)
    assert got.dtype == np.int32 and got.shape == img.shape
    np.testing.assert_array_equal(got, want)
    n = int(got.max()) + 1
    # numbered in order of first appearance, row-major: what compute_label_dict (rbepwt.py:840-848) ranks regions by
    _, first = np.unique(got.ravel(), return_index=True)
    assert np.all(np.diff(first) > 0) and np.array_equal(np.unique(got), np.arange(n))
    # every segment is 8-connected, and none is smaller than min_size unless it is the only one
    assert _components_8(got) == n
    if n > 1:
        assert np.bincount(got.ravel()).min() >= min_size
    if name == "constant":
        assert n == 1


def test_felzenszwalb_facade_and_guards():
    import rbepwt_b200 as rb
    from rbepwt_b200 import synth

    img = np.round(synth.piecewise_smooth_image(synth.voronoi_labels(32, 32, 8, seed=2), seed=2)).astype(np.uint8)
    im = rb.Image()
    im.read_array(img)
    im.segment(scale=200, sigma=2, min_size=10)  # the reference's defaults (rbepwt.py:224)
    assert im.has_segmentation and im.segmentation_method == "felzenszwalb" and im.label_img.shape == img.shape
    assert (im.felz_scale, im.felz_sigma, im.felz_min_size) == (200, 2, 10)
    np.testing.assert_array_equal(im.label_img, rb.felzenszwalb_labels(img, 200, 2, 10))
    # a larger scale merges more
    assert rb.felzenszwalb_labels(img, 2000, 2, 10).max() <= rb.felzenszwalb_labels(img, 20, 2, 10).max()
    with pytest.raises(NotImplementedError):
        im.segment(method="kmeans")
    with pytest.raises(TypeError):
        rb.felzenszwalb_labels(img.astype(np.int64))
    with pytest.raises(ValueError):
        rb.felzenszwalb_labels(np.zeros((2, 2, 3)))
    with pytest.raises(Exception):
        rb.felzenszwalb_labels(img, scale=-1.0)
