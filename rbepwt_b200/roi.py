"""Region-of-interest thresholding (reference: class Roi, /root/reference/rbepwt.py:1684-1789).

`compute_dual_roi_coeffs(regionsidx, perc_in, perc_out)` keeps the proportion perc_in of the detail coefficients whose
support (in the tree the transform builds: coefficient o of a level comes from the path points 2o and 2o+1 of that level)
reaches a pixel of the regions `regionsidx`, and perc_out of the others; the approximation is never touched
(rbepwt.py:1777 "TODO").  Like the reference it follows the supports as if the wavelet were Haar (rbepwt.py:1759).

This is host-side bookkeeping over what the transform produced -- region offsets, the per-level permutations and the
coefficients, all read from the GPU state through the codec's accessors; the selection itself is a few numpy passes per
level (the reference does it with Python sets and list.index, O(N^2)).  `roi_select` is the pure function, tested on
the CPU against fixtures produced by the reference.
"""
import numpy as np


def roi_select(details, level1_in_mask, perms, perc_in, perc_out):
    """details: {level: |coefficients| source array of length N_level / 2}, level = 1..L.
    level1_in_mask: bool [N] over the level-1 signal in PATH order -- True for the points of the ROI regions.
    perms: {level: int [N_level]} for level = 2..L, the position in the level's incoming order of its t-th path point
    (the reference's global_perm, rbepwt.py:1742-1751).
    Returns (keep, nin, nout): keep = {level: bool mask of the coefficients that survive}."""
    if perc_in < 0 or perc_in > 1 or perc_out < 0 or perc_out > 1:
        raise Exception("perc_in and perc_out must be floats between 0 and 1")  # rbepwt.py:1720-1721
    L = len(details)
    in_mask = np.asarray(level1_in_mask, dtype=bool)
    ent_in, ent_out = [], []  # (magnitude, level, index) of the coefficients of each side
    for level in range(1, L + 1):
        mag = np.abs(np.asarray(details[level], dtype=np.float64))
        pair = in_mask.reshape(-1, 2)
        cin = pair[:, 0] | pair[:, 1]      # a child inside the ROI  -> the coefficient is an "in" coefficient
        cout = ~pair[:, 0] | ~pair[:, 1]   # a child outside         -> (also) an "out" coefficient: rbepwt.py:1755-1760
        for side, m in ((ent_in, cin), (ent_out, cout)):
            idx = np.flatnonzero(m)
            side.append(np.stack([mag[idx], np.full(idx.size, level, dtype=np.float64), idx.astype(np.float64)], axis=1))
        if level < L:  # the selected children's parents, as points of the next level's path (rbepwt.py:1758)
            in_mask = cin[np.asarray(perms[level + 1])]
    keep = {level: np.zeros(np.asarray(details[level]).shape[0], dtype=bool) for level in range(1, L + 1)}
    counts = []
    for ent, perc in ((ent_in, perc_in), (ent_out, perc_out)):
        e = np.concatenate(ent) if ent else np.zeros((0, 3))
        n = int(perc * e.shape[0])
        counts.append(n)
        if n:
            # largest magnitudes first; equal magnitudes: unpinned in the reference (a sort of a list built from a set)
            order = np.argsort(-e[:, 0], kind="stable")[:n]
            for level in range(1, L + 1):
                sel = order[e[order, 1] == level]
                keep[level][e[sel, 2].astype(np.int64)] = True
    return keep, counts[0], counts[1]


def global_perm(local_perm, region_offsets):
    """Region.permutation of every region, concatenated (what rbepwt_get_perm returns), + the regions' offsets in the
    level's signal = the reference's global_perm (rbepwt.py:1742-1751)."""
    off = np.asarray(region_offsets, dtype=np.int64)
    return np.asarray(local_perm, dtype=np.int64) + np.repeat(off[:-1], np.diff(off))


class Roi:
    """Region of Interest class. Takes as initialization argument an Image instance (rbepwt.py:1684-1688)."""

    def __init__(self, img):
        self.img = img

    def find_intersecting_regions(self, rect):
        """Label values met by the rectangle (row0, col0, row1, col1), corners included (rbepwt.py:1690-1698) -- region
        indices for label maps numbered in order of first appearance, see compute_dual_roi_coeffs."""
        sub = np.asarray(self.img.label_img)[rect[0]:rect[2] + 1, rect[1]:rect[3] + 1]
        print("%d points in the regions intersecting the rectangle" % sub.size)
        return set(int(v) for v in np.unique(sub))

    def compute_roi_coeffs(self, regionsidx, perc=1, threshold=True):
        nin, _ = self.compute_dual_roi_coeffs(regionsidx, perc, 0, threshold)
        return nin

    def compute_dual_roi_coeffs(self, regionsidx, perc_in, perc_out, threshold=True):
        """Keeps perc_in percentage of coefficients in regions in regionsidx and perc_out percentage for other regions"""
        rb = self.img.rbepwt
        codec, L = rb._codec, rb.levels
        rb._upload_if_mirrored()
        # `regionsidx` holds region INDICES -- the keys of the reference's region dict, i.e. the rank of a label's first
        # appearance in the row-major scan (rbepwt.py:840-848), which equals the label value only for label maps numbered
        # in that order (Felzenszwalb's output is)
        off = codec.region_offsets(0, 1)
        in_mask = np.zeros(int(off[-1]), dtype=bool)
        for r in set(int(r) for r in regionsidx):
            if 0 <= r < off.size - 1:
                in_mask[off[r]:off[r + 1]] = True
        details = {lev: np.asarray(rb.wavelet_details[lev]) for lev in range(1, L + 1)}
        perms = {lev: global_perm(codec.perm(0, lev), codec.region_offsets(0, lev)) for lev in range(2, L + 1)}
        keep, nin, nout = roi_select(details, in_mask, perms, perc_in, perc_out)
        print("coeffs in = %5d, coeffs out = %5d" % (nin, nout))
        if threshold:
            for lev in range(1, L + 1):
                rb.wavelet_details[lev][~keep[lev]] = 0  # live host mirrors: uploaded before the next decode
            print("self.img.nonzero_coefs() = %d" % self.img.nonzero_coefs())
        return nin, nout
