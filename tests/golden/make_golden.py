"""Generate the golden fixtures by executing the UNMODIFIED reference (/root/reference/rbepwt.py).

Run in the build container only (the reference does not exist on the GPU box):

    python tests/golden/make_golden.py [--only NAME ...] [--big]

Each fixture tests/golden/<name>.npz holds the inputs (img, labels) and what the reference
produced for encode_rbepwt -> threshold_coefs -> decode_rbepwt -> psnr, in the flat layout of
oracle/ref_harness.run_reference.  The only non-reference arithmetic involved is the
pywt.dwt/idwt restatement (oracle/pywt_port.py) -- PyWavelets is not installed here.
`--big` adds the 256x256 BASELINE.json config-1 case (about 4 minutes of reference time).
"""
import argparse
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))

from oracle import ref_harness  # noqa: E402
from rbepwt_b200 import synth  # noqa: E402


def noise_labels(h, w, nlab, seed):
    """Salt-and-pepper labels: every region is a sparse, disconnected point set, so almost
    every step is a jump (r >= 2) with an arbitrary preferred direction -- the tie-break stress."""
    return np.random.default_rng(seed).integers(0, nlab, size=(h, w)).astype(np.int32) * 7 - 3


def cases(big):
    c = []
    img = np.array([[10, 12, 50, 52], [11, 49, 51, 53], [13, 14, 15, 54], [90, 91, 16, 55]], dtype=np.float64)
    lab = np.array([[5, 5, 2, 2], [5, 2, 2, 2], [5, 5, 5, 2], [7, 7, 5, 2]], dtype=np.int32)
    c.append(("survey11_4x4_haar", img, lab, 4, "haar", "easypath", True, 4))
    for s in range(6):  # tie-break stress, euclid + chebyshev
        lab = noise_labels(32, 32, 5 + 3 * s, 100 + s)
        img = synth.noise_image(32, 32, seed=200 + s)
        c.append(("noise32_euclid_s%d" % s, img, lab, 10, "bior4.4" if s % 2 else "haar", "easypath", True, 64))
        c.append(("noise32_cheb_s%d" % s, img, lab, 10, "db2", "easypath", False, 100))
    lab = synth.voronoi_labels(32, 32, 23, seed=1)
    img = synth.piecewise_smooth_image(lab, seed=1)
    c.append(("vor32_euclid_bior44", img, lab, 10, "bior4.4", "easypath", True, 37))
    c.append(("vor32_cheb_haar", img, lab, 10, "haar", "easypath", False, 37))
    c.append(("vor32_euclid_db3_L3", img, lab, 3, "db3", "easypath", True, 200))
    c.append(("vor32_u8img_db4", np.round(img).astype(np.uint8), lab, 10, "db4", "easypath", True, 50))
    c.append(("one_label32_euclid", img, np.zeros((32, 32), np.int32), 10, "bior4.4", "easypath", True, 64))
    lab = synth.voronoi_labels(32, 64, 30, seed=2)  # H*W power of two, H != W (rbepwt.py:301)
    c.append(("vor32x64_euclid_bior44", synth.piecewise_smooth_image(lab, seed=2), lab, 11, "bior4.4", "easypath", True, 128))
    lab = (np.arange(64)[:, None] // 3 + 0 * np.arange(64)[None, :]).astype(np.int32)  # stripes
    c.append(("stripes64_euclid_haar", synth.noise_image(64, 64, seed=3), lab, 12, "haar", "easypath", True, 256))
    lab = synth.voronoi_labels(64, 64, 40, seed=4)
    img = synth.piecewise_smooth_image(lab, seed=4)
    c.append(("vor64_euclid_bior44", img, lab, 12, "bior4.4", "easypath", True, 256))
    c.append(("vor64_k0_keeps_all", img, lab, 12, "bior4.4", "easypath", True, 0))
    c.append(("epwt16_bior44", synth.noise_image(16, 16, seed=5), None, 8, "bior4.4", "epwt-easypath", True, 32))
    c.append(("epwt32_smooth_haar", synth.smooth_field_image(32, 32, seed=6, sigma=2.0), None, 10, "haar", "epwt-easypath", True, 64))
    c.append(("epwt64_noise_bior44", synth.noise_image(64, 64, seed=7), None, 12, "bior4.4", "epwt-easypath", True, 256))
    c.append(("epwt64_smooth_db2", synth.smooth_field_image(64, 64, seed=8, sigma=3.0), None, 12, "db2", "epwt-easypath", True, 256))
    lab = synth.voronoi_labels(128, 128, 96, seed=9)
    c.append(("vor128_euclid_bior44", synth.piecewise_smooth_image(lab, seed=9), lab, 14, "bior4.4", "easypath", True, 512))
    c.append(("epwt128_smooth_bior44", synth.smooth_field_image(128, 128, seed=10, sigma=4.0), None, 14, "bior4.4", "epwt-easypath", True, 512))
    # paths_first_level=True: Region.same_path (identity permutation) at every level >= 2 (rbepwt.py:1183-1188, 2024-2025)
    lab = synth.voronoi_labels(32, 32, 17, seed=11)
    img = synth.piecewise_smooth_image(lab, seed=11)
    c.append(("pfl32_euclid_bior44", img, lab, 10, "bior4.4", "easypath", True, 40, True))
    c.append(("pfl32_cheb_haar", img, noise_labels(32, 32, 6, 12), 10, "haar", "easypath", False, 40, True))
    c.append(("pfl32_epwt_db2", synth.smooth_field_image(32, 32, seed=13, sigma=2.0), None, 10, "db2", "epwt-easypath", True, 40, True))
    # uint8 EPWT: |value[cur] - value[cand]| wraps modulo 256 at level 1 (numpy uint8 scalars, rbepwt.py:1302, 844-846)
    c.append(("epwt16_u8_noise_haar", np.random.default_rng(14).integers(0, 256, size=(16, 16)).astype(np.uint8), None, 8,
              "haar", "epwt-easypath", True, 32))
    c.append(("epwt32_u8_smooth_bior44", np.round(synth.smooth_field_image(32, 32, seed=15, sigma=2.0)).astype(np.uint8), None,
              10, "bior4.4", "epwt-easypath", True, 64))
    # gradpath (Region.grad_path, rbepwt.py:1190-1271) with the reference's set iteration pinned to row-major order
    # (oracle/ref_harness.RowMajorSet): complete ties of the gradient preference are otherwise interpreter-dependent
    lab = synth.voronoi_labels(16, 16, 6, seed=21)
    img = synth.piecewise_smooth_image(lab, seed=21)
    c.append(("grad16_euclid_haar", img, lab, 6, "haar", "gradpath", True, 30))
    c.append(("grad16_cheb_db2", img, lab, 6, "db2", "gradpath", False, 30))
    lab = synth.voronoi_labels(32, 32, 14, seed=22)
    img = synth.piecewise_smooth_image(lab, seed=22)
    c.append(("grad32_euclid_bior44", img, lab, 10, "bior4.4", "gradpath", True, 64))
    c.append(("grad32_cheb_haar", img, noise_labels(32, 32, 4, 23), 10, "haar", "gradpath", False, 64))
    c.append(("grad32_u8_pfl_euclid", np.round(img).astype(np.uint8), lab, 10, "haar", "gradpath", True, 64, True))
    c.append(("grad32_flat_regions", np.full((32, 32), 7.0), lab, 10, "haar", "gradpath", True, 16))  # zero gradients: NaN directions
    if big:
        img, lab = synth.config_inputs("cameraman256")  # BASELINE.json configs[0]
        c.append(("config1_cameraman256", img, lab, 16, "bior4.4", "easypath", True, 512))
    return c


def perc_cases():
    """Rbepwt.threshold_by_percentage (rbepwt.py:2120-2192): fixtures in tests/golden/perc/."""
    c = []
    lab = synth.voronoi_labels(32, 32, 9, seed=3)
    img = synth.piecewise_smooth_image(lab, seed=3)
    c.append(("perc32_haar_25", img, lab, 6, "haar", "easypath", True, 0.25))
    c.append(("perc32_bior44_10", img, lab, 10, "bior4.4", "easypath", True, 0.1))
    c.append(("perc32_cheb_db2_50", img, noise_labels(32, 32, 5, 31), 8, "db2", "easypath", False, 0.5))
    c.append(("perc32_haar_3", img, lab, 7, "haar", "easypath", True, 0.03))
    c.append(("perc32_haar_150", img, lab, 6, "haar", "easypath", True, 1.5))
    c.append(("perc16_epwt_haar_20", synth.smooth_field_image(16, 16, seed=32, sigma=2.0), None, 6, "haar", "epwt-easypath", True, 0.2))
    return c


def make_perc(only, force):
    import contextlib
    import io

    os.makedirs(os.path.join(HERE, "perc"), exist_ok=True)
    for name, img, lab, levels, wav, ptype, euclid, perc in perc_cases():
        if only and name not in only:
            continue
        path = os.path.join(HERE, "perc", name + ".npz")
        if os.path.exists(path) and not force:
            continue
        im = ref_harness.make_image(img, lab if ptype != "epwt-easypath" else None)
        with contextlib.redirect_stdout(io.StringIO()):
            im.encode_rbepwt(levels, wav, path_type=ptype, euclidean_distance=euclid)
            im.rbepwt.threshold_by_percentage(perc)
            flat = np.asarray(im.rbepwt.flat_wavelet(), dtype=np.float64).copy()
            im.decode_rbepwt()
        np.savez_compressed(path, img=img, labels=(lab if lab is not None else np.zeros((0, 0), np.int32)), levels=levels,
                            wavelet=wav, path_type=ptype, euclidean_distance=euclid, perc=perc, thresholded=flat,
                            decoded=np.asarray(im.decoded_img, dtype=np.float64), psnr=float(im.psnr()))
        print("%-28s kept %d" % (name, np.count_nonzero(flat)), flush=True)


def roi_cases():
    """Roi.compute_dual_roi_coeffs (rbepwt.py:1718-1789): fixtures in tests/golden/roi/.  Label values are not 0..R-1 in
    one case: the reference matches `regionsidx` against the keys of its region dict, which are region indices (ranks of
    first appearance), not label values -- indices beyond R select nothing."""
    c = []
    lab = synth.voronoi_labels(16, 16, 6, seed=1)
    c.append(("roi16_haar_l3", synth.piecewise_smooth_image(lab, seed=1), lab, 3, "haar", True, [0, 1], 0.5, 0.1))
    lab = synth.voronoi_labels(32, 32, 9, seed=3)
    img = synth.piecewise_smooth_image(lab, seed=3)
    c.append(("roi32_haar_l5", img, lab, 5, "haar", True, [2, 3, 5], 1.0, 0.0))
    c.append(("roi32_haar_l10_all", img, lab, 10, "haar", True, [4], 0.3, 0.05))
    c.append(("roi32_bior44_l6", img, lab, 6, "bior4.4", True, [0, 7, 8], 0.25, 0.02))
    c.append(("roi32_cheb_sparse_labels", img, lab * 7 + 3, 6, "haar", False, [3, 24, 59], 0.6, 0.0))
    return c


def make_roi(only, force):
    import contextlib
    import io

    ref = ref_harness.load_reference()
    os.makedirs(os.path.join(HERE, "roi"), exist_ok=True)
    for name, img, lab, levels, wav, euclid, regs, pin, pout in roi_cases():
        if only and name not in only:
            continue
        path = os.path.join(HERE, "roi", name + ".npz")
        if os.path.exists(path) and not force:
            continue
        im = ref_harness.make_image(img, lab)
        with contextlib.redirect_stdout(io.StringIO()):
            im.encode_rbepwt(levels, wav, euclidean_distance=euclid)
            nin, nout = ref.Roi(im).compute_dual_roi_coeffs(regs, pin, pout)
            flat = np.asarray(im.rbepwt.flat_wavelet(), dtype=np.float64).copy()
            im.decode_rbepwt()
        np.savez_compressed(path, img=img, labels=lab, levels=levels, wavelet=wav, euclidean_distance=euclid,
                            regions=np.asarray(regs, dtype=np.int32), perc_in=pin, perc_out=pout, nin=nin, nout=nout,
                            thresholded=flat, decoded=np.asarray(im.decoded_img, dtype=np.float64), psnr=float(im.psnr()))
        print("%-28s nin %d nout %d kept %d" % (name, nin, nout, np.count_nonzero(flat)), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", nargs="*")
    ap.add_argument("--big", action="store_true")
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--perc", action="store_true", help="the threshold_by_percentage fixtures (tests/golden/perc/)")
    ap.add_argument("--roi", action="store_true", help="the region-of-interest thresholding fixtures (tests/golden/roi/)")
    args = ap.parse_args()
    if args.perc:
        make_perc(args.only, args.force)
        return
    if args.roi:
        make_roi(args.only, args.force)
        return
    for case in cases(args.big):
        name, img, lab, levels, wav, ptype, euclid, k = case[:8]
        pfl = bool(case[8]) if len(case) > 8 else False
        if args.only and name not in args.only:
            continue
        path = os.path.join(HERE, name + ".npz")
        if os.path.exists(path) and not args.force:
            continue
        t0 = time.perf_counter()
        out = ref_harness.run_reference(img, lab, levels, wav, ptype, euclid, ncoefs=k, paths_first_level=pfl)
        dt = time.perf_counter() - t0
        np.savez_compressed(
            path,
            img=img,
            labels=(lab if lab is not None else np.zeros((0, 0), np.int32)),
            levels=levels, wavelet=wav, path_type=ptype, euclidean_distance=euclid, ncoefs=k, paths_first_level=pfl,
            perm=np.concatenate([out["perm"][l] for l in range(1, levels + 1)]).astype(np.int32),
            roff=np.stack([out["roff"][l] for l in range(1, levels + 2)]).astype(np.int32),
            points=np.concatenate([out["points"][l] for l in range(1, levels + 2)]).astype(np.int16),
            coefs=out["coefs"], thresholded=out["thresholded"], kept=out["kept"],
            decoded=out["decoded"], psnr=out["psnr"], nonzero_coefs=out["nonzero_coefs"],
            reference_seconds=dt,
        )
        print("%-28s %6.1fs  %d KB" % (name, dt, os.path.getsize(path) // 1024), flush=True)


if __name__ == "__main__":
    main()
