"""TEST INFRASTRUCTURE: numpy restatement of scikit-image's felzenszwalb (what the reference's Image.segment calls,
/root/reference/rbepwt.py:779-785) for checking rbepwt_b200/csrc/segment.hpp.  PARITY UNPINNED: scikit-image is a
third-party dependency, absent from /root/reference and from this image; this follows the published algorithm
(Felzenszwalb & Huttenlocher 2004) in the form scikit-image's _felzenszwalb_cy.pyx gives it, written independently of
the C++ (vectorised edge construction, scipy's own gaussian_filter when scipy is present).  Only tests/ import this."""
import numpy as np


def gaussian(img, sigma):
    try:
        from scipy import ndimage as ndi
        return ndi.gaussian_filter(np.asarray(img, dtype=np.float64), sigma=sigma)  # mode='reflect', truncate=4.0
    except ImportError:  # the same filter by hand
        radius = int(4.0 * sigma + 0.5)
        x = np.arange(-radius, radius + 1)
        w = np.exp(-0.5 / sigma ** 2 * x ** 2)
        w /= w.sum()
        out = np.asarray(img, dtype=np.float64)
        for axis in (0, 1):
            pad = [(radius, radius) if a == axis else (0, 0) for a in (0, 1)]
            p = np.pad(out, pad, mode="symmetric")
            out = sum(w[k] * np.take(p, np.arange(out.shape[axis]) + k, axis=axis) for k in range(2 * radius + 1))
        return out


def felzenszwalb(img01, scale=1.0, sigma=0.8, min_size=20):
    """img01: 2-D float64 as img_as_float64 would deliver it.  Returns int32 labels numbered by ascending root."""
    img = np.asarray(img01, dtype=np.float64)
    H, W = img.shape
    scale = float(scale) / 255.0
    a = gaussian(img, sigma) if sigma > 0 else img
    seg = np.arange(H * W).reshape(H, W)
    costs = np.hstack([np.abs(a[:, 1:] - a[:, :-1]).ravel(), np.abs(a[1:, :] - a[:-1, :]).ravel(),
                       np.abs(a[1:, 1:] - a[:-1, :-1]).ravel(), np.abs(a[1:, :-1] - a[:-1, 1:]).ravel()])
    edges = np.vstack([np.c_[seg[:, 1:].ravel(), seg[:, :-1].ravel()], np.c_[seg[1:, :].ravel(), seg[:-1, :].ravel()],
                       np.c_[seg[1:, 1:].ravel(), seg[:-1, :-1].ravel()], np.c_[seg[:-1, 1:].ravel(), seg[1:, :-1].ravel()]])
    order = np.argsort(costs, kind="stable")
    edges, costs = edges[order], costs[order]
    forest = np.arange(H * W)
    size = np.ones(H * W, dtype=np.int64)
    cint = np.zeros(H * W)

    def find(i):
        while forest[i] != i:
            forest[i] = forest[forest[i]]
            i = forest[i]
        return i

    for (p, q), c in zip(edges, costs):
        r0, r1 = find(p), find(q)
        if r0 == r1:
            continue
        if c < min(cint[r0] + scale / size[r0], cint[r1] + scale / size[r1]):
            root, child = min(r0, r1), max(r0, r1)
            forest[child] = root
            size[root] = size[r0] + size[r1]
            cint[root] = c
    for p, q in edges:
        r0, r1 = find(p), find(q)
        if r0 != r1 and (size[r0] < min_size or size[r1] < min_size):
            root, child = min(r0, r1), max(r0, r1)
            forest[child] = root
            size[root] = size[r0] + size[r1]
    roots = np.array([find(i) for i in range(H * W)])
    return np.unique(roots, return_inverse=True)[1].reshape(H, W).astype(np.int32)
