"""TEST INFRASTRUCTURE ONLY -- runs the UNMODIFIED reference as the oracle.

Imports /root/reference/rbepwt.py as-is.  The packages it imports at module level
that are not installed in this image (ipdb, skimage.*, matplotlib.*, pywt) are
replaced by stub modules registered in sys.modules first; `pywt.dwt/idwt` are the
restatement in oracle/pywt_port.py (the only arithmetic on the path that is not
the reference's own code).  Everything else -- paths, permutations, reduce/expand
bookkeeping, top-k selection, clipping, PSNR -- is the reference's code executing.

/root/reference exists only in the build container, never on the GPU box: this
module is used by tests/golden/make_golden.py to produce committed fixtures, and by
optional local checks; nothing under `-m gpu`, smoke() or bench.py imports it.
"""
import contextlib
import io
import os
import sys
import types
import warnings

import numpy as np

REFERENCE_DIR = os.environ.get("RBEPWT_REFERENCE_DIR", "/root/reference")


class _Anything:
    """Attribute sink: plt.cm.gray etc. are evaluated in default args at import."""

    def __getattr__(self, name):
        return _Anything()

    def __call__(self, *a, **k):
        return _Anything()


def _stub(name, **attrs):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    return mod


_ref = None


def available():
    return os.path.isfile(os.path.join(REFERENCE_DIR, "rbepwt.py"))


def load_reference():
    """Import the reference module once, with stubs for absent third-party packages."""
    global _ref
    if _ref is not None:
        return _ref
    if not available():
        raise RuntimeError("reference not present at %s" % REFERENCE_DIR)
    from . import pywt_port

    for name in ("ipdb",):
        if name not in sys.modules:
            _stub(name, set_trace=lambda *a, **k: None)
    try:
        import skimage  # noqa: F401
    except ImportError:
        sk = _stub("skimage")
        sk.io = _stub("skimage.io", imread=_Anything(), imsave=_Anything())
        sk.filters = _stub("skimage.filters", gaussian=_Anything())
        sk.restoration = _stub("skimage.restoration")
        sk.segmentation = _stub("skimage.segmentation", felzenszwalb=_Anything())
        sk.measure = _stub("skimage.measure", compare_ssim=_Anything())
    try:
        import matplotlib  # noqa: F401
    except ImportError:
        mpl = _stub("matplotlib")
        mpl.pyplot = _stub("matplotlib.pyplot", cm=_Anything())
        mpl.patches = _stub("matplotlib.patches")
    if "pywt" not in sys.modules:
        _stub("pywt", dwt=pywt_port.dwt, idwt=pywt_port.idwt)
    sys.path.insert(0, REFERENCE_DIR)
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            import rbepwt as ref  # the reference, unmodified
    finally:
        sys.path.remove(REFERENCE_DIR)
    _ref = ref
    return ref


def make_image(img, labels=None):
    """Reference Image with externally supplied labels, injected exactly the way
    Image.load_mat_segmentation does (rbepwt.py:250-258)."""
    ref = load_reference()
    im = ref.Image()
    with contextlib.redirect_stdout(io.StringIO()):
        im.read_array(img)
        if labels is not None:
            im.segmentation = ref.Segmentation(im.img)
            im.segmentation_method = "external"
            im.label_img = labels
            im.segmentation.label_img = labels
            im.segmentation.nlabels = int(labels.max()) + 1
            im.segmentation.compute_label_dict()
            im.has_segmentation = True
    return im


class RowMajorSet(set):
    """A set that iterates in sorted order.  Region.grad_path resolves complete ties of its gradient preference by
    the iteration order of a Python set of (row, col) tuples (rbepwt.py:1224), which depends on the interpreter's
    hash-table internals: unpinned.  Binding the name `set` in the reference module's namespace to this class (the
    source file is not touched) makes the reference's own code iterate its candidates in row-major order, the tie
    rule the CUDA path and the C port follow; every other step of grad_path is order-independent."""

    def intersection(self, *others):
        return RowMajorSet(set.intersection(self, *others))

    def __iter__(self):
        return iter(sorted(set.__iter__(self)))


def run_reference(img, labels, levels, wavelet, path_type="easypath",
                  euclidean_distance=True, ncoefs=None, paths_first_level=False):
    """Reference run; path_type='gradpath' runs with the set iteration order pinned (RowMajorSet)."""
    ref = load_reference()
    if path_type != "gradpath":
        return _run_reference(img, labels, levels, wavelet, path_type, euclidean_distance, ncoefs, paths_first_level)
    ref.set = RowMajorSet
    try:
        return _run_reference(img, labels, levels, wavelet, path_type, euclidean_distance, ncoefs, paths_first_level)
    finally:
        del ref.set


def _run_reference(img, labels, levels, wavelet, path_type="easypath",
                   euclidean_distance=True, ncoefs=None, paths_first_level=False):
    """Encode (+ threshold + decode) with the reference; returns a dict of plain
    numpy arrays in the flat layout of SURVEY.md section 8a:

      perm[l]    int32 [N_l]  local permutation of every region, concatenated
      roff[l]    int32 [R+1]  region offsets at level l (l = 1..L+1)
      points[l]  int32 [N_l,2] (row,col) in path order, regions concatenated
      coefs      float64 [N]  details[1] | ... | details[L] | approx
      kept       int64 [k]    flat indices surviving threshold_coefs(ncoefs)
      decoded    float64 [H,W]; psnr float
    """
    im = make_image(img, labels if path_type != "epwt-easypath" else None)
    out = {}
    with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        im.encode_rbepwt(levels, wavelet, path_type=path_type,
                         euclidean_distance=euclidean_distance, paths_first_level=paths_first_level)
        rb = im.rbepwt
        perm, roff, points = {}, {}, {}
        for lev in range(1, levels + 2):
            rc = rb.region_collection_at_level[lev]
            offs, pts, pm = [0], [], []
            for _, reg in rc:
                n = len(reg.base_points)
                offs.append(offs[-1] + n)
                pts.extend(reg.base_points)
                if lev <= levels and n:
                    pm.extend(list(reg.permutation))
            roff[lev] = np.asarray(offs, dtype=np.int32)
            points[lev] = np.asarray(pts, dtype=np.int32).reshape(-1, 2)
            if lev <= levels:
                perm[lev] = np.asarray(pm, dtype=np.int32)
        out.update(perm=perm, roff=roff, points=points)
        out["coefs"] = np.asarray(rb.flat_wavelet(), dtype=np.float64).copy()
        if ncoefs is not None:
            im.threshold_coefs(ncoefs)
            flat = np.asarray(rb.flat_wavelet(), dtype=np.float64)
            out["thresholded"] = flat.copy()
            out["kept"] = np.flatnonzero(flat != 0).astype(np.int64)
            im.decode_rbepwt()
            out["decoded"] = np.asarray(im.decoded_img, dtype=np.float64).copy()
            out["psnr"] = float(im.psnr())
            out["nonzero_coefs"] = int(im.nonzero_coefs())
    return out
