"""-m gpu: the CUDA path (through the public API -> C ABI) against
  (1) the golden fixtures produced by the unmodified reference,
  (2) the C oracle on seeded inputs at sizes it finishes in seconds,
  (3) size-independent properties at BASELINE.json's full sizes.
Bar: paths / permutations / kept indices bit-exact; coefficients and pixels within 1e-9 relative (fp64);
PSNR equal to 6 decimals."""
import numpy as np
import pytest

from conftest import assert_matches_golden, golden_names, load_golden
from gpu_util import assert_same_as_oracle, collect_batch, cuda_run

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", golden_names())
def test_cuda_reproduces_reference_golden(name):
    g = load_golden(name)
    out = cuda_run(g["img"], g["labels"], g["levels"], g["wavelet"], g["path_type"], g["euclidean_distance"],
                   ncoefs=g["ncoefs"], paths_first_level=g["paths_first_level"])
    assert_matches_golden(out, g)


@pytest.mark.parametrize("name", [n for n in golden_names() if not n.startswith("epwt")])
def test_golden_through_the_thread_per_region_kernels(name):
    """A single image's regions are each walked by a whole warp (latency); a batch with more than 4096 regions is
    walked thread per region -- the kernel the benchmark runs.  Enough copies of the fixture to get there, then the
    first and the last copy against what the reference produced."""
    g = load_golden(name)
    if g["path_type"] != "easypath":
        pytest.skip("EPWT has its own kernel")
    nreg = len(np.unique(g["labels"]))
    copies = 4096 // nreg + 2
    for which in (0, copies - 1):
        out = cuda_run(g["img"], g["labels"], g["levels"], g["wavelet"], g["path_type"], g["euclidean_distance"],
                       ncoefs=g["ncoefs"], paths_first_level=g["paths_first_level"], copies=copies, which=which)
        assert_matches_golden(out, g)


def _oracle(img, lab, levels, wavelet, ptype, euclid, k):
    from oracle import c_oracle
    import rbepwt_b200 as rb

    return c_oracle.run(img, lab, levels, rb.filter_bank(wavelet), ptype, euclid, ncoefs=k)


@pytest.mark.parametrize("wavelet,k", [("haar", 512), ("bior4.4", 2048), ("bior4.4", 8192), ("db6", 512)])
def test_config2_512_vs_oracle(wavelet, k):
    from rbepwt_b200 import synth

    img, lab = synth.config_inputs("synthetic512")
    out = cuda_run(img, lab, 16, wavelet, ncoefs=k)
    assert_same_as_oracle(out, _oracle(img, lab, 16, wavelet, "easypath", True, k), 16)


def test_bench_workload_batch_of_distinct_512_vs_oracle():
    """The kernel and the inputs bench.py times: a batch of DISTINCT 512x512 images from the bench generator (seeds
    1000, 1001, ...; 1024 Voronoi regions each), more than 4096 regions in the path group, so the regions are walked
    thread per region -- every image against the C oracle, paths and permutations bit-exact at all 16 levels."""
    torch = pytest.importorskip("torch")
    import rbepwt_b200 as rb
    from rbepwt_b200 import synth

    B = 8
    timg, tlab = synth.torch_batch(B, 512, 512, 1024, 1000, device="cuda")
    imgs, labs = timg.cpu().numpy(), tlab.cpu().numpy()
    c = rb.BatchCodec()
    c.encode(imgs, labs, 16, "bior4.4")
    assert sum(c.region_count(b) for b in range(B)) > 4096
    outs = collect_batch(c, imgs, range(B), 16, 2048)
    for b in range(B):
        assert_same_as_oracle(outs[b], _oracle(imgs[b], labs[b], 16, "bior4.4", "easypath", True, 2048), 16)
    # the one-call pipeline on device pointers (what the bench's timed region runs) gives the same pixels
    out = torch.empty_like(timg)
    c2 = rb.BatchCodec()
    c2.transcode(timg, tlab, 16, "bior4.4", 2048, out=out)
    c2.sync()
    got = out.cpu().numpy()
    for b in range(B):
        np.testing.assert_array_equal(got[b], outs[b]["decoded"])


@pytest.mark.parametrize("euclid", [True, False])
def test_heavy_tailed_label_maps_vs_oracle(euclid):
    """Felzenszwalb-like region sizes (median ~85 pixels, several regions of 5-10 thousand): the long chains take
    the windowed / whole-warp walkers next to the thread-per-region bulk.  Ten copies of two distinct maps make
    the group large enough (> 4096 regions) for the throughput kernels; one image alone takes the latency path."""
    import rbepwt_b200 as rb
    from rbepwt_b200 import synth

    labs = np.stack([synth.heavytail_labels(512, 600, 100 + i) for i in range(2)])
    imgs = np.stack([synth.piecewise_smooth_image(l, seed=3 + i) for i, l in enumerate(labs)])
    orc = [_oracle(imgs[i], labs[i], 16, "bior4.4", "easypath", euclid, 2048) for i in range(2)]
    bl, bi = np.concatenate([labs] * 5), np.concatenate([imgs] * 5)
    c = rb.BatchCodec()
    c.encode(bi, bl, 16, "bior4.4", euclidean_distance=euclid)
    assert sum(c.region_count(b) for b in range(10)) > 4096
    outs = collect_batch(c, bi, (0, 1, 9), 16, 2048)
    for b in (0, 1, 9):
        assert_same_as_oracle(outs[b], orc[b % 2], 16)
    one = cuda_run(imgs[1], labs[1], 16, "bior4.4", euclidean_distance=euclid, ncoefs=2048)
    assert_same_as_oracle(one, orc[1], 16)
    # who is walked by a whole warp is a tuning decision (RBEPWT_OPT_COOP_LIMIT: the largest n regions of the group;
    # default 4 per SM): the results must not depend on it -- none, a handful (the threshold falls inside the size
    # distribution), everything of >= 2048 pixels
    for lim in (0, 7, 1 << 20):
        c2 = rb.BatchCodec()
        c2.set_option(coop_limit=lim)
        c2.encode(bi, bl, 16, "bior4.4", euclidean_distance=euclid)
        outs2 = collect_batch(c2, bi, (1, 8), 16, 2048)
        for b in (1, 8):
            assert_same_as_oracle(outs2[b], orc[b % 2], 16)
        c2.close()


@pytest.mark.parametrize("euclid", [True, False])
def test_gradpath_vs_oracle(euclid):
    """path_type='gradpath' (Region.grad_path, rbepwt.py:1190-1271) at 256^2 and in a batch of distinct images, against
    the C port (itself pinned to the reference with its set iteration order fixed, tests/golden/grad*.npz): same
    tie rule on both sides, paths bit-exact."""
    import rbepwt_b200 as rb
    from rbepwt_b200 import synth

    img, lab = synth.config_inputs("cameraman256", seed=33)
    out = cuda_run(img, lab, 16, "bior4.4", "gradpath", euclid, ncoefs=512)
    assert_same_as_oracle(out, _oracle(img, lab, 16, "bior4.4", "gradpath", euclid, 512), 16)
    rng = np.random.default_rng(9)
    labs = np.stack([synth.voronoi_labels(64, 64, 12 + 5 * s, seed=60 + s) for s in range(4)])
    labs[3] = rng.integers(0, 6, size=(64, 64))  # scattered classes: every step a jump
    imgs = np.stack([synth.piecewise_smooth_image(l, seed=60 + i) for i, l in enumerate(labs)])
    imgs[2] = np.round(imgs[2])  # quantised values: equal gradients, more ties
    c = rb.BatchCodec()
    c.encode(imgs, labs, 12, "haar", "gradpath", euclid)
    outs = collect_batch(c, imgs, range(4), 12, 200)
    for b in range(4):
        assert_same_as_oracle(outs[b], _oracle(imgs[b], labs[b], 12, "haar", "gradpath", euclid, 200), 12)
    with pytest.raises(Exception, match="too small to calculate a numerical gradient"):
        rb.BatchCodec().encode(np.zeros((1, 1, 16)), np.zeros((1, 1, 16), np.int32), 2, "haar", "gradpath")


def test_config2_512_chebyshev_vs_oracle():
    from rbepwt_b200 import synth

    img, lab = synth.config_inputs("synthetic512", seed=11)
    out = cuda_run(img, lab, 16, "bior4.4", euclidean_distance=False, ncoefs=2048)
    assert_same_as_oracle(out, _oracle(img, lab, 16, "bior4.4", "easypath", False, 2048), 16)


@pytest.mark.parametrize("wavelet", ["haar", "bior4.4"])
def test_config3_epwt_512_vs_oracle(wavelet):
    from rbepwt_b200 import synth

    img, _ = synth.config_inputs("epwt512")
    out = cuda_run(img, None, 16, wavelet, "epwt-easypath", ncoefs=2048, with_perm=False)
    assert_same_as_oracle(out, _oracle(img, None, 16, wavelet, "epwt-easypath", True, 2048), 16)


def test_config1_256_vs_oracle_other_seeds():
    from rbepwt_b200 import synth

    for seed in (21, 22):
        img, lab = synth.config_inputs("cameraman256", seed=seed)
        out = cuda_run(img, lab, 16, "bior4.4", ncoefs=512)
        assert_same_as_oracle(out, _oracle(img, lab, 16, "bior4.4", "easypath", True, 512), 16)


def test_config4_2048_many_small_regions():
    """2048^2, ~65k regions: paths vs oracle bit-exact, perfect reconstruction, top-k count."""
    from rbepwt_b200 import synth
    import rbepwt_b200 as rb

    img, lab = synth.config_inputs("small2048")
    out = cuda_run(img, lab, 16, "bior4.4", ncoefs=None, with_perm=False)
    orc = _oracle(img, lab, 16, "bior4.4", "easypath", True, None)
    assert_same_as_oracle(out, orc, 16)
    c = out["codec"]
    dec = c.decode()[0]
    assert np.max(np.abs(dec - img)) < 1e-9 * 255
    c.threshold(8192)
    assert int(c.nonzero_coefs()[0]) == 8192


def test_hash_labels_and_disconnected_regions():
    """Arbitrary int32 label values (negative, huge -> open-addressing table) and label classes
    made of scattered pixels (every step is a jump)."""
    rng = np.random.default_rng(5)
    vals = np.array([-2147483648, -7, -1, 0, 3, 65536, 2147483647, 123456789], dtype=np.int64)
    lab = vals[rng.integers(0, len(vals), size=(64, 64))].astype(np.int32)
    img = rng.uniform(0, 255, size=(64, 64))
    for euclid in (True, False):
        out = cuda_run(img, lab, 12, "bior4.4", euclidean_distance=euclid, ncoefs=300)
        assert_same_as_oracle(out, _oracle(img, lab, 12, "bior4.4", "easypath", euclid, 300), 12)


def test_every_pixel_its_own_region_and_tiny_images():
    rng = np.random.default_rng(6)
    img = rng.uniform(0, 255, size=(16, 16))
    lab = rng.permutation(256).reshape(16, 16).astype(np.int32)  # R == N
    out = cuda_run(img, lab, 8, "db2", ncoefs=10)
    assert_same_as_oracle(out, _oracle(img, lab, 8, "db2", "easypath", True, 10), 8)
    for shape, levels in (((2, 2), 2), ((1, 4), 2), ((4, 1), 1), ((2, 8), 4)):
        img = rng.uniform(0, 255, size=shape)
        lab = rng.integers(0, 2, size=shape).astype(np.int32)
        for wav in ("haar", "bior4.4"):
            out = cuda_run(img, lab, levels, wav, ncoefs=2)
            assert_same_as_oracle(out, _oracle(img, lab, levels, wav, "easypath", True, 2), levels)


def test_large_regions_use_the_big_bitmap_path():
    """Two interleaved label classes spanning the whole 256x256 image (bitmap > shared-memory slot)."""
    rng = np.random.default_rng(7)
    ii, jj = np.meshgrid(np.arange(256), np.arange(256), indexing="ij")
    lab = (((ii // 5) + (jj // 3)) % 2).astype(np.int32)
    img = rng.uniform(0, 255, size=(256, 256))
    out = cuda_run(img, lab, 16, "haar", ncoefs=1000)
    assert_same_as_oracle(out, _oracle(img, lab, 16, "haar", "easypath", True, 1000), 16)


def test_many_mid_sized_bitmaps_and_wide_regions():
    """1024^2 tiled into 128^2 tiles of one-pixel anti-diagonal stripes: ~16k regions whose bounding-box bitmaps
    are 100..700 words (queue classes 1-4, the windowed path kernel, more chunks than the prebuilt-bitmap buffer
    holds, so the in-kernel bitmap build runs too), plus a band of long flat regions wider than one bitmap word."""
    rng = np.random.default_rng(11)
    ii, jj = np.meshgrid(np.arange(1024), np.arange(1024), indexing="ij")
    lab = ((ii // 128) * 8 + (jj // 128)) * 256 + (ii % 128) + (jj % 128)
    lab[:64] = 100000 + (ii[:64] // 2) * 16 + jj[:64] // 64  # 2 x 64 boxes: two words per bitmap row
    lab = lab.astype(np.int32)
    img = rng.uniform(0, 255, size=(1024, 1024))
    for euclid in (True, False):
        out = cuda_run(img, lab, 12, "haar", euclidean_distance=euclid, ncoefs=5000, with_perm=False)
        assert_same_as_oracle(out, _oracle(img, lab, 12, "haar", "easypath", euclid, 5000), 12)


def test_batch_equals_singles_and_threshold_properties():
    from rbepwt_b200 import synth
    import rbepwt_b200 as rb

    imgs, labs = [], []
    for s in range(5):
        lab = synth.voronoi_labels(128, 128, 60 + 10 * s, seed=40 + s)
        labs.append(lab)
        imgs.append(synth.piecewise_smooth_image(lab, seed=40 + s))
    imgs, labs = np.stack(imgs), np.stack(labs)
    c = rb.BatchCodec()
    c.encode(imgs, labs, 14, "bior4.4")
    full = np.stack([c.coefs(b) for b in range(5)])
    dec_all = c.decode()
    assert np.max(np.abs(dec_all - imgs)) < 1e-9 * 255  # perfect reconstruction without thresholding
    c.threshold(777)
    th = np.stack([c.coefs(b) for b in range(5)])
    np.testing.assert_array_equal(c.nonzero_coefs(), [777] * 5)
    for b in range(5):
        kept = th[b] != 0
        np.testing.assert_array_equal(th[b][kept], full[b][kept])          # survivors untouched
        assert np.abs(full[b][kept]).min() >= np.abs(full[b][~kept]).max()  # and they are the largest
    c.threshold(777)                                                       # idempotent
    np.testing.assert_array_equal(np.stack([c.coefs(b) for b in range(5)]), th)
    dec = c.decode()
    for b in (0, 3):
        one = cuda_run(imgs[b], labs[b], 14, "bior4.4", ncoefs=777, with_perm=False)
        np.testing.assert_array_equal(one["coefs"], full[b])
        np.testing.assert_array_equal(one["decoded"], dec[b])
    # k quirks of the reference: 0 and >= N keep everything
    c2 = rb.BatchCodec()
    c2.encode(imgs[:1], labs[:1], 14, "bior4.4")
    before = c2.coefs(0)
    for k in (0, -3, 128 * 128, 10 ** 9):
        c2.threshold(k)
        np.testing.assert_array_equal(c2.coefs(0), before)


def test_threshold_ties_keep_highest_index():
    import rbepwt_b200 as rb

    img = np.full((8, 8), 7.0)
    lab = np.zeros((8, 8), np.int32)
    c = rb.BatchCodec()
    c.encode(img[None], lab[None], 2, "haar")
    flat = np.zeros(64)
    flat[[3, 10, 20, 33, 50]] = [5.0, -5.0, 5.0, 9.0, -5.0]
    c.set_coefs(flat, 0)
    c.threshold(3)
    got = c.coefs(0)
    np.testing.assert_array_equal(np.flatnonzero(got), [20, 33, 50])
    from oracle import c_oracle
    np.testing.assert_array_equal(got, c_oracle.threshold(flat, 3))


@pytest.mark.parametrize("kind", ["one_binade", "few_values", "mixed"])
def test_threshold_512_crowded_magnitudes_vs_oracle(kind):
    """K4 keeps the candidates of the first digit in shared memory; these distributions overflow that buffer
    (every CTA slice holds more same-binade keys than fit) or tie massively, so the global-memory passes and
    the highest-index tie rule run at full size."""
    import rbepwt_b200 as rb
    from oracle import c_oracle

    rng = np.random.default_rng({"one_binade": 1, "few_values": 2, "mixed": 3}[kind])
    n = 512 * 512
    if kind == "one_binade":
        flat = rng.uniform(1.0, 1.24, n) * rng.choice([-1.0, 1.0], n)  # one quarter-binade: a single first digit
    elif kind == "few_values":
        flat = rng.choice([0.0, 0.5, -0.5, 3.0, -3.0, 7.25], n, p=[0.3, 0.25, 0.25, 0.1, 0.05, 0.05])
    else:
        flat = rng.normal(0, 2.0, n)
        flat[rng.integers(0, n, n // 3)] = 1.5  # a third of the slice in one bin, the rest spread
    img = np.zeros((512, 512))
    lab = np.zeros((512, 512), np.int32)
    c = rb.BatchCodec()
    c.encode(img[None], lab[None], 1, "haar")
    for k in (5, 2048, 100000, n - 7):
        c.set_coefs(flat, 0)
        c.threshold(k)
        got = c.coefs(0)
        assert np.count_nonzero(got) == min(k, np.count_nonzero(flat))
        np.testing.assert_array_equal(got, c_oracle.threshold(flat, k))


def test_device_pointer_path_matches_host_path():
    torch = pytest.importorskip("torch")
    from rbepwt_b200 import synth
    import rbepwt_b200 as rb

    lab = synth.voronoi_labels(64, 64, 30, seed=3)
    img = synth.piecewise_smooth_image(lab, seed=3)
    host = cuda_run(img, lab, 12, "bior4.4", ncoefs=200, with_perm=False)
    timg = torch.from_numpy(img[None].copy()).cuda()
    tlab = torch.from_numpy(lab[None].copy()).cuda()
    c = rb.BatchCodec(stream=torch.cuda.current_stream().cuda_stream)
    c.encode(timg, tlab, 12, "bior4.4")
    c.threshold(200)
    out = torch.empty_like(timg)
    c.decode(out)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(out.cpu().numpy()[0], host["decoded"])


def _small_batch(n=5, size=64, seed0=70):
    from rbepwt_b200 import synth

    imgs, labs = [], []
    for s in range(n):
        lab = synth.voronoi_labels(size, size, 20 + 3 * s, seed=seed0 + s)
        labs.append(lab)
        imgs.append(synth.piecewise_smooth_image(lab, seed=seed0 + s))
    return np.stack(imgs), np.stack(labs)


@pytest.mark.parametrize("streams,sub,grp", [(2, 0, 0), (1, 0, 0), (2, 2, 2), (2, 1, 3), (1, 3, 3), (2, 1, 1)])
def test_transcode_equals_three_calls_for_any_pipelining(streams, sub, grp):
    """rbepwt_transcode (one pipelined call) == encode, threshold, decode; the sub-batch size and the
    number of sub-batches in flight never change a bit of the result."""
    import rbepwt_b200 as rb

    imgs, labs = _small_batch()
    ref = rb.BatchCodec()
    ref.set_option(streams=1, sub_batch=0)
    ref.encode(imgs, labs, 12, "bior4.4")
    full = np.stack([ref.coefs(b) for b in range(len(imgs))])
    ref.threshold(300)
    want = ref.decode()
    want_coefs = np.stack([ref.coefs(b) for b in range(len(imgs))])
    want_paths = [ref.paths(b, 3) for b in range(len(imgs))]

    c = rb.BatchCodec()
    c.set_option(streams=streams, sub_batch=sub, path_group=grp)
    got = c.transcode(imgs, labs, 12, "bior4.4", 300)
    np.testing.assert_array_equal(got, want)
    for b in range(len(imgs)):
        np.testing.assert_array_equal(c.coefs(b), want_coefs[b])
        np.testing.assert_array_equal(c.paths(b, 3), want_paths[b])
        np.testing.assert_array_equal(c.region_offsets(b), ref.region_offsets(b))
    # and the three separate calls under the same options
    c.encode(imgs, labs, 12, "bior4.4")
    np.testing.assert_array_equal(np.stack([c.coefs(b) for b in range(len(imgs))]), full)
    c.threshold(300)
    np.testing.assert_array_equal(c.decode(), want)


def test_transcode_device_pointers_and_epwt():
    torch = pytest.importorskip("torch")
    import rbepwt_b200 as rb

    imgs, labs = _small_batch(4, 32, 90)
    want = rb.BatchCodec().transcode(imgs, labs, 10, "db2", 100)
    s = torch.cuda.Stream()
    c = rb.BatchCodec(stream=s.cuda_stream)
    c.set_option(sub_batch=1)
    timg, tlab = torch.from_numpy(imgs).cuda(), torch.from_numpy(labs).cuda()
    torch.cuda.synchronize()
    out = c.transcode(timg, tlab, 10, "db2", 100)
    c.sync()
    np.testing.assert_array_equal(out.cpu().numpy(), want)
    # EPWT through the pipeline: batch == singles
    e = rb.BatchCodec()
    e.set_option(sub_batch=3)
    got = e.transcode(imgs, None, 10, "haar", 50, path_type="epwt-easypath")
    for b in (0, 3):
        one = cuda_run(imgs[b], None, 10, "haar", "epwt-easypath", ncoefs=50, with_perm=False)
        np.testing.assert_array_equal(got[b], one["decoded"])


def test_region_arrays_grow_across_sub_batches():
    """More regions than the first capacity guess (B * 2048 + 4096): the arrays are regrown mid-pipeline."""
    import rbepwt_b200 as rb

    rng = np.random.default_rng(12)
    labs = np.stack([rng.permutation(128 * 128).reshape(128, 128).astype(np.int32) for _ in range(3)])  # R == N each
    imgs = rng.uniform(0, 255, size=labs.shape)
    c = rb.BatchCodec()
    c.set_option(sub_batch=1)
    c.encode(imgs, labs, 6, "haar")
    assert [c.region_count(b) for b in range(3)] == [128 * 128] * 3
    dec = c.decode()
    assert np.max(np.abs(dec - imgs)) < 1e-9 * 255
    for b in range(3):  # every region is one pixel: level-1 path order == order of first appearance == row-major
        np.testing.assert_array_equal(c.paths(b, 1), np.arange(128 * 128))


def test_paths_first_level_512_vs_oracle_and_facade():
    """paths_first_level=True (Region.same_path at levels >= 2) at full size, batch and facade."""
    from rbepwt_b200 import synth
    from oracle import c_oracle
    import rbepwt_b200 as rb

    img, lab = synth.config_inputs("synthetic512", seed=31)
    out = cuda_run(img, lab, 16, "bior4.4", ncoefs=2048, paths_first_level=True)
    orc = c_oracle.run(img, lab, 16, rb.filter_bank("bior4.4"), "easypath", True, ncoefs=2048, paths_first_level=True)
    assert_same_as_oracle(out, orc, 16)
    for lev in (2, 5, 16):  # identity permutation inside every region
        off = out["roff"][lev]
        want = np.concatenate([np.arange(off[r + 1] - off[r]) for r in range(len(off) - 1)]) if len(off) > 1 else []
        np.testing.assert_array_equal(out["perm"][lev], want)
    im = rb.Image()
    im.read_array(img)
    im.set_labels(lab)
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        im.encode_rbepwt(16, "bior4.4", paths_first_level=True)
        im.threshold_coefs(2048)
        im.decode_rbepwt()
    np.testing.assert_array_equal(im.decoded_img, out["decoded"])
    got = rb.BatchCodec().transcode(img[None], lab[None], 16, "bior4.4", 2048, paths_first_level=True)
    np.testing.assert_array_equal(got[0], out["decoded"])


def test_fuzz_random_small_cases_vs_oracle():
    """80 random small cases over every switch of the path: shape (H != W too), label structure (blobs, noise,
    stripes, few/many labels, arbitrary label values), path mode, paths_first_level, wavelet, levels, k."""
    from rbepwt_b200 import synth
    from oracle import c_oracle
    import rbepwt_b200 as rb

    rng = np.random.default_rng(2024)
    wavelets = ["haar", "db2", "db3", "db4", "db7", "bior4.4", "bior2.2", "rbio3.1"]
    for case in range(80):
        lh, lw = int(rng.integers(0, 7)), int(rng.integers(0, 7))
        if lh + lw < 2:
            lw = 2 - lh
        H, W = 1 << lh, 1 << lw
        N = H * W
        levels = int(rng.integers(1, lh + lw + 1))
        kind = int(rng.integers(0, 4))
        if kind == 0 and min(H, W) >= 8:
            lab = synth.voronoi_labels(H, W, int(rng.integers(1, max(2, N // 24))), seed=case)
        elif kind == 1:
            lab = rng.integers(0, int(rng.integers(1, 9)), size=(H, W)).astype(np.int32)
        elif kind == 2:
            lab = ((np.arange(H)[:, None] // int(rng.integers(1, 4))) * 3 + (np.arange(W)[None, :] // int(rng.integers(1, 5)))).astype(np.int32)
        else:
            lab = rng.integers(-5, N, size=(H, W)).astype(np.int32)
        lab = (lab.astype(np.int64) * int(rng.choice([1, 7, -3, 100003])) + int(rng.integers(-50, 50))).astype(np.int32)
        img = rng.uniform(0, 255, size=(H, W))
        mode = int(rng.integers(0, 3))
        ptype, euclid = ("easypath", True) if mode == 0 else ("easypath", False) if mode == 1 else ("epwt-easypath", True)
        pfl = bool(rng.integers(0, 2))
        wav = wavelets[int(rng.integers(0, len(wavelets)))]
        k = int(rng.integers(0, N + 2))
        labs = None if mode == 2 else lab
        try:
            out = cuda_run(img, labs, levels, wav, ptype, euclid, ncoefs=k, paths_first_level=pfl)
            orc = c_oracle.run(img, labs, levels, rb.filter_bank(wav), ptype, euclid, ncoefs=k, paths_first_level=pfl)
            assert_same_as_oracle(out, orc, levels)
        except AssertionError as e:
            raise AssertionError("fuzz case %d: %dx%d L=%d kind=%d mode=%d pfl=%s wav=%s k=%d: %s"
                                 % (case, H, W, levels, kind, mode, pfl, wav, k, e))


def test_batch_larger_than_one_chunk():
    """More images than one internal chunk (1024): chunks are processed one after the other with the label
    table reused; every image must equal its single-image result."""
    import rbepwt_b200 as rb

    rng = np.random.default_rng(77)
    B = 1100
    labs = rng.integers(0, 5, size=(B, 8, 8)).astype(np.int32)
    imgs = rng.uniform(0, 255, size=(B, 8, 8))
    c = rb.BatchCodec()
    got = c.transcode(imgs, labs, 6, "db2", 20)
    for b in (0, 1, 1023, 1024, 1025, 1099):
        one = cuda_run(imgs[b], labs[b], 6, "db2", ncoefs=20, with_perm=False)
        np.testing.assert_array_equal(got[b], one["decoded"])
        np.testing.assert_array_equal(c.paths(b, 1), one["points"][1][:, 0] * 8 + one["points"][1][:, 1])
        assert c.region_count(b) == len(one["roff"][1]) - 1
