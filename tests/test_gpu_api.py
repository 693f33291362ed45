"""-m gpu: the reference-shaped facade (rbepwt_b200.Image) used the way the reference's scripts use it."""
import contextlib
import io

import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu


def _image(g):
    import rbepwt_b200 as rbepwt

    im = rbepwt.Image()
    im.read_array(g["img"])
    if g["labels"] is not None:
        im.set_labels(g["labels"])
    return im


def test_image_facade_matches_golden():
    g = load_golden("vor64_euclid_bior44")
    im = _image(g)
    with contextlib.redirect_stdout(io.StringIO()):
        im.encode_rbepwt(g["levels"], g["wavelet"], path_type=g["path_type"], euclidean_distance=True)
        assert im.method == "rbepwt" and im.rbepwt.levels == g["levels"]
        flat = im.rbepwt.flat_wavelet()
        assert np.max(np.abs(flat - g["coefs"])) <= 1e-9 * np.abs(g["coefs"]).max()
        # region views: permutation, path-ordered base points
        rc1 = im.rbepwt.region_collection_at_level[1]
        off = g["roff_by_level"][1]
        for r in (0, 1, len(off) - 2):
            reg = rc1[r]
            a, b = off[r], off[r + 1]
            assert reg.permutation == list(g["perm_by_level"][1][a:b])
            assert reg.base_points == tuple(map(tuple, g["points_by_level"][1][a:b]))
        im.threshold_coefs(g["ncoefs"])
        assert im.nonzero_coefs() == g["nonzero_coefs"]
        im.decode_rbepwt()
    assert im.has_decoded_img
    assert np.max(np.abs(im.decoded_img - g["decoded"])) <= 1e-9 * 255
    assert abs(im.psnr() - g["psnr"]) < 5e-7
    # decoded_region_collection carries the DECODED values (rbepwt.py:308-310), level-1 regions in incoming order
    drc = im.decoded_region_collection
    pts = drc.points
    assert len(pts) == g["img"].size
    rebuilt = np.zeros_like(g["decoded"])
    for coord, value in pts.items():
        rebuilt[coord] = value
    assert np.max(np.abs(np.clip(rebuilt, 0, 255) - g["decoded"])) <= 1e-9 * 255
    assert drc[0].base_points == tuple(sorted(drc[0].base_points))  # row-major inside the region
    rcl = im.rbepwt.region_collection_at_level
    assert len(rcl) == g["levels"] + 1 and list(rcl) == list(range(1, g["levels"] + 2)) and (g["levels"] + 1) in rcl
    assert [k for k, _ in rcl.items()] == list(rcl.keys())


def test_epwt_facade_and_method_quirk():
    g = load_golden("epwt32_smooth_haar")
    im = _image(g)
    with contextlib.redirect_stdout(io.StringIO()):
        im.encode_epwt(g["levels"], g["wavelet"])
        assert im.method == "rbepwt"  # encode_rbepwt overwrites 'epwt' (rbepwt.py:335-337)
        im.threshold_coefs(g["ncoefs"])
        im.decode_epwt()
    assert np.max(np.abs(im.decoded_img - g["decoded"])) <= 1e-9 * 255
    rc = im.rbepwt.region_collection_at_level
    assert len(rc[1]) == 1 and len(rc[1][0]) == g["img"].size
    assert rc[2][0].base_points == tuple(map(tuple, g["points_by_level"][2]))


def test_gradpath_facade_matches_reference_fixture():
    g = load_golden("grad32_euclid_bior44")
    im = _image(g)
    with contextlib.redirect_stdout(io.StringIO()):
        im.encode_rbepwt(g["levels"], g["wavelet"], path_type="gradpath", euclidean_distance=True)
        assert im.rbepwt_path_type == "gradpath"
        im.threshold_coefs(g["ncoefs"])
        im.decode_rbepwt()
    assert np.max(np.abs(im.decoded_img - g["decoded"])) <= 1e-9 * 255
    assert abs(im.psnr() - g["psnr"]) < 5e-7


def test_edit_coefficients_then_decode_like_compute_basis_elements():
    """scripts/compute_basis_elements.py:58-81: zero every array, set one coefficient, decode."""
    g = load_golden("vor32_euclid_bior44")
    im = _image(g)
    with contextlib.redirect_stdout(io.StringIO()):
        im.encode_rbepwt(g["levels"], g["wavelet"])
        L = im.rbepwt.levels
        approx = im.rbepwt.region_collection_at_level[L + 1].values
        for lev in range(1, L + 1):
            im.rbepwt.wavelet_details[lev] = np.zeros_like(im.rbepwt.wavelet_details[lev])
        im.rbepwt.region_collection_at_level[L + 1].values = np.zeros_like(approx)
        im.rbepwt.region_collection_at_level[L + 1].values[0] = 1
        im.decode_rbepwt()
    from oracle import c_oracle
    import rbepwt_b200 as rbepwt

    enc = c_oracle.encode(g["img"], g["labels"], L, rbepwt.filter_bank(g["wavelet"]), c_oracle.MODE_EUCLID)
    flat = np.zeros(g["img"].size)
    flat[-1] = 1.0
    want = c_oracle.decode(enc, flat, rbepwt.filter_bank(g["wavelet"]))
    assert np.max(np.abs(im.decoded_img - want)) < 1e-12
    assert im.nonzero_coefs() == 1


def test_full_decode_equals_fast_decode():
    """scripts/check_decode.py:47-74: decode with stored paths == decode with paths recomputed from labels."""
    import rbepwt_b200 as rbepwt

    g = load_golden("vor64_euclid_bior44")
    im = _image(g)
    with contextlib.redirect_stdout(io.StringIO()):
        im.encode_rbepwt(g["levels"], g["wavelet"])
        im.rbepwt.threshold_coefs(51)
        im.decode_rbepwt()
        L = g["levels"]
        fdi = rbepwt.full_decode(im.rbepwt.wavelet_details, im.rbepwt.region_collection_at_level[L + 1].values,
                                 im.label_img, g["wavelet"], "easypath")
    np.testing.assert_array_equal(fdi, im.decoded_img)


def test_full_decode_vs_oracle_decode():
    """rbepwt_full_decode (paths regenerated from the labels alone, rbepwt.py:106-130) against the C oracle's decode of
    the same coefficient vectors: thresholded coefficients of the image, and arbitrary ones (the decoder must
    not depend on having seen the image)."""
    import rbepwt_b200 as rbepwt
    from oracle import c_oracle
    from rbepwt_b200 import synth

    rng = np.random.default_rng(8)
    for name, euclid in (("vor64_euclid_bior44", True), ("vor32_cheb_haar", False)):
        g = load_golden(name)
        fb = rbepwt.filter_bank(g["wavelet"])
        L = g["levels"]
        enc = c_oracle.encode(g["img"], g["labels"], L, fb, c_oracle.MODE_EUCLID if euclid else c_oracle.MODE_CHEB)
        vecs = np.stack([g["thresholded"], rng.normal(0, 30, g["img"].size)])
        labs = np.stack([g["labels"], g["labels"]])
        got = rbepwt.BatchCodec().full_decode(vecs, labs, L, g["wavelet"], "easypath", euclid)
        for i in range(2):
            want = c_oracle.decode(enc, vecs[i], fb)
            assert np.max(np.abs(got[i] - want)) <= 1e-9 * 255
        assert np.max(np.abs(got[0] - g["decoded"])) <= 1e-9 * 255  # and what the reference itself decoded
    # 512^2, the benchmark's label maps
    lab = synth.voronoi_labels(512, 512, 1024, seed=5)
    fb = rbepwt.filter_bank("bior4.4")
    enc = c_oracle.encode(np.zeros((512, 512)), lab, 16, fb, c_oracle.MODE_EUCLID)
    vec = rng.normal(0, 20, 512 * 512) * (rng.uniform(size=512 * 512) < 0.01)
    got = rbepwt.BatchCodec().full_decode(vec[None], lab[None], 16, "bior4.4")
    assert np.max(np.abs(got[0] - c_oracle.decode(enc, vec, fb))) <= 1e-9 * 255


def test_guards_raise_the_reference_messages():
    import rbepwt_b200 as rbepwt

    im = rbepwt.Image()
    im.read_array(np.zeros((6, 8)))
    im.set_labels(np.zeros((6, 8), np.int32))
    with pytest.raises(Exception, match="Image size must be a power of 2"):
        im.encode_rbepwt(2, "haar")
    im = rbepwt.Image()
    im.read_array(np.zeros((4, 4)))
    im.set_labels(np.zeros((4, 4), np.int32))
    with pytest.raises(Exception, match="2\\^levels must be smaller or equal"):
        im.encode_rbepwt(5, "haar")
    c = rbepwt.BatchCodec()
    with pytest.raises(Exception, match="There is no saved encoding to decode"):
        c.threshold(3)


def test_level_values_views():
    g = load_golden("vor32_euclid_bior44")
    im = _image(g)
    with contextlib.redirect_stdout(io.StringIO()):
        im.encode_rbepwt(g["levels"], g["wavelet"])
    from oracle import c_oracle
    import rbepwt_b200 as rbepwt

    rc = im.rbepwt.region_collection_at_level
    # level-1 collection values = pixel values in region order / row-major
    pix0 = np.array([r * 32 + c for r, c in rc[1].base_points])
    np.testing.assert_array_equal(rc[1].values, g["img"].ravel()[pix0])
    # level-2 values = level-1 low-pass
    enc = c_oracle.encode(g["img"], g["labels"], 1, rbepwt.filter_bank(g["wavelet"]), c_oracle.MODE_EUCLID)
    np.testing.assert_allclose(rc[2].values, enc["coefs"][g["img"].size // 2:], rtol=0, atol=1e-10)


def test_transcode_ex_narrow_types_and_codec_outputs():
    """rbepwt_transcode_ex: uint8 / float32 pixels and uint16 labels give bit-identical results to the float64 / int32
    path on the same values (widening is exact); PSNR and the kept (index, value) pairs equal what the getters of the
    plain path report; narrow outputs are the float64 output converted; the encode-only form decodes nothing."""
    torch = pytest.importorskip("torch")
    import rbepwt_b200 as rb
    from rbepwt_b200 import synth

    B, n, k = 6, 64, 300
    labs = np.stack([synth.voronoi_labels(n, n, 20 + 3 * s, seed=50 + s, shuffle_ids=False) for s in range(B)])
    img8 = np.stack([np.round(synth.piecewise_smooth_image(l, seed=50 + i)).astype(np.uint8) for i, l in enumerate(labs)])
    ref = rb.BatchCodec()
    want = ref.transcode(img8.astype(np.float64), labs, 12, "bior4.4", k)
    want_coefs = np.stack([ref.coefs(b) for b in range(B)])
    want_psnr = ref.psnr(img8.astype(np.float64), want)
    c = rb.BatchCodec()
    for pix, lab in ((img8, labs.astype(np.uint16)), (img8.astype(np.float32), labs.astype(np.int32)),
                     (img8.astype(np.float64), labs.astype(np.uint16))):
        r = c.transcode_ex(pix, lab, 12, "bior4.4", k, want_psnr=True, want_kept=True)
        np.testing.assert_array_equal(r["image"], want)
        np.testing.assert_allclose(r["psnr"], want_psnr, rtol=0, atol=1e-9)
        for b in range(B):
            nzi = np.flatnonzero(want_coefs[b])
            np.testing.assert_array_equal(r["kept_idx"][b], nzi)
            np.testing.assert_array_equal(r["kept_val"][b], want_coefs[b][nzi])
            np.testing.assert_array_equal(c.coefs(b), want_coefs[b])  # same state as the plain call
    # narrow outputs
    r32 = c.transcode_ex(img8, labs.astype(np.uint16), 12, "bior4.4", k, out_dtype="float32")
    np.testing.assert_array_equal(r32["image"], want.astype(np.float32))
    r8 = c.transcode_ex(img8, labs.astype(np.uint16), 12, "bior4.4", k, out_dtype="uint8")
    np.testing.assert_array_equal(r8["image"], np.rint(want).astype(np.uint8))
    # encode-only: the compact representation, nothing decoded
    enc = c.transcode_ex(img8, labs.astype(np.uint16), 12, "bior4.4", k, want_image=False, want_kept=True)
    assert set(enc) == {"kept_idx", "kept_val"}
    np.testing.assert_array_equal(enc["kept_idx"], r["kept_idx"])
    # ... from which the decoder side rebuilds the images (full_decode: labels + coefficients)
    flat = np.zeros((B, n * n))
    for b in range(B):
        flat[b, enc["kept_idx"][b]] = enc["kept_val"][b]
    np.testing.assert_array_equal(rb.BatchCodec().full_decode(flat, labs, 12, "bior4.4"), want)
    # device pointers, narrow types in and out
    s = torch.cuda.Stream()
    d = rb.BatchCodec(stream=s.cuda_stream)
    t8, t16 = torch.from_numpy(img8).cuda(), torch.from_numpy(labs.astype(np.int32)).cuda().to(torch.uint16)
    torch.cuda.synchronize()
    rd = d.transcode_ex(t8, t16, 12, "bior4.4", k, out_dtype="uint8", want_psnr=True)
    d.sync()
    np.testing.assert_array_equal(rd["image"].cpu().numpy(), np.rint(want).astype(np.uint8))
    np.testing.assert_allclose(rd["psnr"], want_psnr, rtol=0, atol=1e-9)
    # uint8 EPWT wraps like the reference's uint8 arrays do (rbepwt.py:1302): same as the plain path with a uint8 array
    g = load_golden("epwt16_u8_noise_haar")
    re = rb.BatchCodec().transcode_ex(g["img"][None], None, g["levels"], g["wavelet"], g["ncoefs"], path_type="epwt-easypath")
    assert np.max(np.abs(re["image"][0] - g["decoded"])) <= 1e-9 * 255


def test_box_codec_shards_by_image():
    """BoxCodec over every visible GPU (one here is fine: two codecs on the same device exercise the same code):
    results in batch order, identical to one BatchCodec."""
    torch = pytest.importorskip("torch")
    import rbepwt_b200 as rb
    from rbepwt_b200 import synth

    B, n = 7, 64
    labs = np.stack([synth.voronoi_labels(n, n, 25, seed=80 + s) for s in range(B)])
    imgs = np.stack([synth.piecewise_smooth_image(l, seed=80 + i) for i, l in enumerate(labs)])
    want = rb.BatchCodec().transcode(imgs, labs, 12, "bior4.4", 200)
    ndev = torch.cuda.device_count()
    devices = list(range(ndev)) if ndev > 1 else [0, 0, 0]
    box = rb.BoxCodec(devices)
    np.testing.assert_array_equal(box.transcode(imgs, labs, 12, "bior4.4", 200), want)
    r = box.transcode_ex(imgs.astype(np.float32).astype(np.float64), labs.astype(np.uint16), 12, "bior4.4", 200, want_psnr=True,
                         want_kept=True)
    assert r["image"].shape == imgs.shape and r["psnr"].shape == (B,) and r["kept_idx"].shape == (B, 200)
    box.close()


def test_threshold_by_percentage_matches_reference_fixtures():
    """Rbepwt.threshold_by_percentage (rbepwt.py:2120-2192) through the facade and the batch API, against what the
    unmodified reference produced (tests/golden/perc/), plus a 512x512 run against the C port."""
    import rbepwt_b200 as rbepwt
    from conftest import load_perc, perc_names
    from oracle import c_oracle
    from rbepwt_b200 import synth

    for name in perc_names():
        g = load_perc(name)
        im = rbepwt.Image()
        im.read_array(g["img"])
        if g["labels"] is not None:
            im.set_labels(g["labels"])
        with contextlib.redirect_stdout(io.StringIO()):
            im.encode_rbepwt(g["levels"], g["wavelet"], path_type=g["path_type"], euclidean_distance=g["euclidean_distance"])
            im.rbepwt.threshold_by_percentage(g["perc"])
            flat = im.rbepwt.flat_wavelet()
            im.decode_rbepwt()
        np.testing.assert_array_equal(np.flatnonzero(flat), np.flatnonzero(g["thresholded"]), err_msg=name)
        assert np.max(np.abs(flat - g["thresholded"])) <= 1e-9 * np.abs(g["thresholded"]).max()
        assert np.max(np.abs(im.decoded_img - g["decoded"])) <= 1e-9 * 255
        assert abs(im.psnr() - g["psnr"]) < 5e-7
    img, lab = synth.config_inputs("synthetic512", seed=12)
    fb = rbepwt.filter_bank("bior4.4")
    enc = c_oracle.encode(img, lab, 16, fb, c_oracle.MODE_EUCLID)
    c = rbepwt.BatchCodec()
    c.encode(np.stack([img, img]), np.stack([lab, lab]), 16, "bior4.4")
    c.threshold_by_percentage(0.02)
    want = c_oracle.threshold_percentage(enc, enc["coefs"], 0.02)
    for b in (0, 1):
        got = c.coefs(b)
        np.testing.assert_array_equal(np.flatnonzero(got), np.flatnonzero(want))
    # ties at the cut: the later entries of the region's list survive (the C port's rule)
    c.set_coefs(np.where(enc["coefs"] != 0, np.sign(enc["coefs"]) * 3.0, 0.0), 0)
    c.threshold_by_percentage(0.3)
    np.testing.assert_array_equal(c.coefs(0), c_oracle.threshold_percentage(enc, np.where(enc["coefs"] != 0, np.sign(enc["coefs"]) * 3.0, 0.0), 0.3))


def test_save_and_load_pickle(tmp_path):
    """Image.save_pickle / load_pickle (rbepwt.py:447-472): a thresholded encoding survives the round trip -- same
    coefficients, same decoded image -- and the restored object keeps working (threshold again, decode, psnr)."""
    import rbepwt_b200 as rbepwt

    g = load_golden("vor32_cheb_haar")
    im = _image(g)
    with contextlib.redirect_stdout(io.StringIO()):
        im.encode_rbepwt(g["levels"], g["wavelet"], euclidean_distance=False)
        im.threshold_coefs(g["ncoefs"])
        im.decode_rbepwt()
    path = str(tmp_path / "im.pickle")
    im.save_pickle(path)
    flat, dec = im.rbepwt.flat_wavelet().copy(), im.decoded_img.copy()
    del im
    im2 = rbepwt.Image()
    with contextlib.redirect_stdout(io.StringIO()):
        im2.load_pickle(path)
        assert im2.has_segmentation and im2.has_decoded_img and im2.method == "rbepwt"
        np.testing.assert_array_equal(im2.rbepwt.flat_wavelet(), flat)
        np.testing.assert_array_equal(im2.decoded_img, dec)
        assert im2.nonzero_coefs() == g["nonzero_coefs"]
        im2.decode_rbepwt()
        np.testing.assert_array_equal(im2.decoded_img, dec)
        assert abs(im2.psnr() - g["psnr"]) < 5e-7
        reg = im2.rbepwt.region_collection_at_level[1][0]
        assert reg.permutation == list(g["perm_by_level"][1][:len(reg)])
        im2.threshold_coefs(5)
        assert im2.nonzero_coefs() == 5


def test_dwt2_baseline_vs_oracle(tmp_path):
    """The tensor-product baseline (class Dwt, rbepwt.py:2249-2298) through the facade: wavedec2 coefficients,
    top-k thresholding, waverec2 + clip against the numpy restatement (oracle/pywt_port.py); batch API; pickle."""
    import rbepwt_b200 as rbepwt
    from oracle import pywt_port
    from rbepwt_b200 import synth

    for n, levels, wav, k in ((64, 4, "bior4.4", 300), (32, 5, "haar", 50), (128, 3, "db3", 2000), (16, 2, "db4", 0)):
        lab = synth.voronoi_labels(n, n, 12, seed=n)
        img = synth.piecewise_smooth_image(lab, seed=n)
        im = rbepwt.Image()
        im.read_array(img)
        im.encode_dwt(levels, wav)
        assert im.method == "dwt" and im.dwt_levels == levels
        want_co = pywt_port.wavedec2(img, rbepwt.filter_bank(wav), levels)
        got_co = im.dwt.wavelet_coefs
        assert len(got_co) == levels + 1
        assert np.max(np.abs(got_co[0] - want_co[0])) <= 1e-9 * np.abs(want_co[0]).max()
        for g3, w3 in zip(got_co[1:], want_co[1:]):
            for gq, wq in zip(g3, w3):
                assert gq.shape == wq.shape and np.max(np.abs(gq - wq)) <= 1e-9 * 255
        im.threshold_coefs(k)
        want_dec, want_nz, want_mags = pywt_port.dwt2_baseline(img, levels, rbepwt.filter_bank(wav), k)
        assert im.nonzero_coefs() == want_nz
        flat = np.concatenate([np.ravel(q) for c_ in im.dwt.wavelet_coefs for q in (c_ if isinstance(c_, tuple) else (c_,))])
        np.testing.assert_allclose(np.sort(np.abs(flat[flat != 0])), want_mags, rtol=0, atol=1e-9 * 255)
        im.decode_dwt()
        assert np.max(np.abs(im.decoded_img - want_dec)) <= 1e-9 * 255
        assert abs(im.psnr() - float(20 * np.log10(255 / np.sqrt(np.mean((img - want_dec) ** 2))))) < 5e-7
    path = str(tmp_path / "dwt.pickle")
    im.save_pickle(path)
    im2 = rbepwt.Image()
    im2.load_pickle(path)
    im2.decode_dwt()
    np.testing.assert_array_equal(im2.decoded_img, im.decoded_img)
    # batch API + guards
    imgs = np.stack([synth.noise_image(64, 64, seed=s) for s in range(3)])
    c = rbepwt.BatchCodec()
    c.dwt2_encode(imgs, 3, "bior4.4")
    np.testing.assert_allclose(c.decode(), np.clip(imgs, 0, 255), rtol=0, atol=1e-9 * 255)  # perfect reconstruction
    c.threshold(100)
    assert list(c.nonzero_coefs()) == [100] * 3
    with pytest.raises(Exception, match="square"):
        c.dwt2_encode(np.zeros((1, 32, 64)), 2, "haar")
    with pytest.raises(Exception):
        c.threshold_by_percentage(0.5)  # regions do not exist for this encoding


def test_roi_thresholding_vs_reference():
    """Roi.compute_dual_roi_coeffs / compute_roi_coeffs through the facade (rbepwt.py:1712-1789) against the fixtures the
    unmodified reference produced (tests/golden/roi): counts, surviving coefficients, decoded image, PSNR."""
    import glob
    import os

    import rbepwt_b200 as rbepwt
    from conftest import GOLDEN_DIR

    files = sorted(glob.glob(os.path.join(GOLDEN_DIR, "roi", "*.npz")))
    assert len(files) >= 5
    for f in files:
        z = np.load(f)
        im = rbepwt.Image()
        im.read_array(z["img"])
        im.set_labels(z["labels"])
        im.encode_rbepwt(int(z["levels"]), str(z["wavelet"]), euclidean_distance=bool(z["euclidean_distance"]))
        roi = rbepwt.Roi(im)
        nin, nout = roi.compute_dual_roi_coeffs([int(r) for r in z["regions"]], float(z["perc_in"]), float(z["perc_out"]))
        assert (nin, nout) == (int(z["nin"]), int(z["nout"])), f
        got, want = im.rbepwt.flat_wavelet(), z["thresholded"]
        np.testing.assert_array_equal(got != 0, want != 0, err_msg=f)
        assert np.max(np.abs(got - want)) <= 1e-9 * np.max(np.abs(want))
        assert im.nonzero_coefs() == int(np.count_nonzero(want))
        im.decode_rbepwt()
        assert np.max(np.abs(im.decoded_img - z["decoded"])) <= 1e-9 * 255
        assert abs(im.psnr() - float(z["psnr"])) < 5e-7
    # threshold=False only counts; compute_roi_coeffs = perc_out 0
    im = rbepwt.Image()
    im.read_array(z["img"])
    im.set_labels(z["labels"])
    im.encode_rbepwt(int(z["levels"]), str(z["wavelet"]), euclidean_distance=bool(z["euclidean_distance"]))
    before = im.nonzero_coefs()
    n = rbepwt.Roi(im).compute_roi_coeffs([0], 1, threshold=False)
    assert n > 0 and im.nonzero_coefs() == before
    assert rbepwt.Roi(im).find_intersecting_regions((0, 0, 3, 3)) == set(int(v) for v in np.unique(z["labels"][:4, :4]))
    with pytest.raises(Exception, match="between 0 and 1"):
        rbepwt.Roi(im).compute_dual_roi_coeffs([0], 2.0, 0.0)
