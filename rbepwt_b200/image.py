"""Drop-in surface of the reference's `rbepwt.Image` for the encode -> threshold -> decode path.

Mirrors /root/reference/rbepwt.py: class Image (190-577, hot-path methods only), class Rbepwt
(1974-2246: encode, decode, threshold_coefs, flat_wavelet, wavelet_coefs_dict), and read-only-ish
views standing in for RegionCollection (1464-1682) / Region (995-1461).  Same method names, argument
meaning, attribute names, exception messages.  All arithmetic runs in the CUDA library through
BatchCodec; nothing here computes paths or wavelets on the CPU, and there is no fallback.

State lives on the GPU.  The attributes scripts read AND write --
`rbepwt.wavelet_details[l]`, `rbepwt.region_collection_at_level[L+1].values`
(scripts/compute_basis_elements.py:58-81, scripts/check_decode.py:49-64) -- are host mirrors created
on first access; once created they are the source of truth and are uploaded before the next
threshold / decode.
"""
import numpy as np

from . import _capi
from .codec import BatchCodec, path_mode


def ispowerof2(n):  # rbepwt.py:132-141
    n = int(n)
    return n >= 1 and (n & (n - 1)) == 0


def psnr(img1, img2):
    """rbepwt.py:156-162, evaluated on the GPU (K6)."""
    a = np.ascontiguousarray(img1, dtype=np.float64)
    b = np.ascontiguousarray(img2, dtype=np.float64)
    v = float(_shared_codec().psnr(a.reshape(1, -1, 1), b.reshape(1, -1, 1))[0])
    return -1 if v == -1.0 else v


_codec = None


def _shared_codec():
    global _codec
    if _codec is None:
        _codec = BatchCodec()
    return _codec


# GPU contexts of Rbepwt objects that were garbage-collected: a loop over many images re-uses their streams and
# device buffers instead of paying context creation and cudaMalloc for every image.
_idle_codecs = []
_MAX_IDLE_CODECS = 4


def _acquire_codec():
    while _idle_codecs:
        codec = _idle_codecs.pop()
        if getattr(codec, "_ctx", None):
            return codec
    return BatchCodec()


def _release_codec(codec):
    # (the cyclic garbage collector may already have finalised -- closed -- the codec of an unreachable Image)
    if codec is None or not getattr(codec, "_ctx", None):
        return
    if len(_idle_codecs) < _MAX_IDLE_CODECS:
        _idle_codecs.append(codec)
    else:
        codec.close()


def felzenszwalb_labels(img, scale=1.0, sigma=0.8, min_size=20):
    """skimage.segmentation.felzenszwalb for a 2-D image through the library (rbepwt_felzenszwalb, csrc/segment.hpp):
    int32 label map numbered in order of first appearance.  A uint8 image is scaled to [0, 1] first, as scikit-image's
    img_as_float64 does; other integer types are not accepted (their scaling rules differ), floats pass as they are."""
    import ctypes

    from . import _capi

    a = np.asarray(img)
    if a.ndim != 2:
        raise ValueError("a 2-D (grayscale) image is required")
    if a.dtype == np.uint8:
        a = a.astype(np.float64) / 255.0
    elif a.dtype.kind == "f":
        a = a.astype(np.float64)
    elif a.dtype == np.bool_:
        a = a.astype(np.float64)
    else:
        raise TypeError("felzenszwalb_labels takes uint8 or floating-point images")
    a = np.ascontiguousarray(a)
    lab = np.empty(a.shape, dtype=np.int32)
    n = ctypes.c_int32()
    _capi.check(_capi.lib().rbepwt_felzenszwalb(a.ctypes.data_as(ctypes.c_void_p), a.shape[0], a.shape[1], float(scale),
                                                float(sigma), int(min_size), lab.ctypes.data_as(ctypes.c_void_p),
                                                ctypes.byref(n)))
    return lab


class Segmentation:
    """Holder of an externally produced label map (rbepwt.py:770-848).  Region order = first
    appearance of the label in a row-major scan; computed on the GPU at encode time (K0)."""

    def __init__(self, image):
        self.img = image
        self.has_label_dict = False
        self.nlabels = -1
        self.label_img = None

    def compute_label_dict(self):
        self.has_label_dict = True  # the region records are built by K0 inside encode


class RegionView:
    """One region at one level (reference: Region).  base_points/values are in PATH order for
    levels 1..L (easy_path(inplace=True) reorders them, rbepwt.py:1338-1342), incoming order at L+1."""

    def __init__(self, base_points, values, permutation):
        self.base_points = tuple(map(tuple, base_points))
        self.values = values
        self.permutation = permutation
        self.trivial = len(self.base_points) == 0
        self.no_values = self.trivial

    @property
    def points(self):
        return {p: v for p, v in zip(self.base_points, self.values)}

    @property
    def start_point(self):
        return min(self.base_points) if self.base_points else None

    @property
    def top_left(self):
        if not self.base_points:
            return None
        return (min(p[0] for p in self.base_points), min(p[1] for p in self.base_points))

    @property
    def bottom_right(self):
        if not self.base_points:
            return None
        return (max(p[0] for p in self.base_points), max(p[1] for p in self.base_points))

    def __len__(self):
        return len(self.base_points)

    def __iter__(self):
        return iter(zip(self.base_points, self.values))

    def __getitem__(self, key):
        if isinstance(key, tuple):
            return self.points[key]
        return (self.base_points[key], self.values[key])


class RegionCollectionView:
    """All regions of one level (reference: RegionCollection), materialised from the GPU on demand."""

    def __init__(self, rb, level):
        self._rb = rb
        self.level = level
        self._values = None
        self._sub = None
        self._inc = None

    # incoming-order pixel list (what the reference's collection-level base_points hold: they are
    # not refreshed by the in-place path reorder)
    def _incoming(self):
        if self._inc is None:
            c, L = self._rb._codec, self._rb.levels
            lev = self.level
            if lev == 1:
                self._inc = c.paths(0, 0)
            elif lev == L + 1:
                self._inc = c.paths(0, L + 1)
            else:
                self._inc = c.paths(0, lev - 1)[0::2].copy()
        return self._inc

    def _rc(self, pix):
        W = self._rb.img.shape[1]
        return [(int(p) // W, int(p) % W) for p in pix]

    @property
    def offsets(self):
        return self._rb._codec.region_offsets(0, self.level)

    @property
    def nregions(self):
        return self._rb._codec.region_count(0)

    @property
    def region_lengths(self):
        return list(np.diff(self.offsets))

    @property
    def values(self):
        if self.level == self._rb.levels + 1:
            return self._rb._approx_array()
        if self._values is None:
            self._values = self._rb._codec.level_values(0, self.level)
        return self._values

    @values.setter
    def values(self, v):
        if self.level == self._rb.levels + 1:
            self._rb._set_approx_array(np.asarray(v, dtype=np.float64))
        else:
            self._values = np.asarray(v)

    @property
    def base_points(self):
        return tuple(self._rc(self._incoming()))

    @property
    def points(self):
        return {p: v for p, v in zip(self.base_points, self.values)}

    @property
    def subregions(self):
        if self._sub is None:
            c, L, lev = self._rb._codec, self._rb.levels, self.level
            off = self.offsets
            vals = self.values
            if lev <= L:
                pix = c.paths(0, lev)
                perm = c.perm(0, lev)
            else:
                pix, perm = self._incoming(), None
            sub = []
            for r in range(len(off) - 1):
                a, b = int(off[r]), int(off[r + 1])
                if perm is not None:
                    p = perm[a:b]
                    sub.append(RegionView(self._rc(pix[a:b]), vals[a:b][p], [int(x) for x in p] if b > a else None))
                else:
                    sub.append(RegionView(self._rc(pix[a:b]), vals[a:b], list(range(b - a))))
            self._sub = sub
        return self._sub

    def __len__(self):
        return self.nregions

    def __getitem__(self, key):
        if isinstance(key, tuple):
            return self.points[key]
        return self.subregions[key]

    def __iter__(self):
        return iter(enumerate(self.subregions))


class DecodedRegionCollectionView(RegionCollectionView):
    """What Rbepwt.decode returns (rbepwt.py:2055-2079): the level-1 regions in their INCOMING order (expand
    undoes every permutation, 1600-1608 -- row-major inside each region) carrying the DECODED values, not yet
    clipped (the clip is Image.decode_rbepwt's, 313-314).  Values are fetched from the GPU on first access."""

    def __init__(self, rb):
        super().__init__(rb, 1)

    @property
    def values(self):
        if self._values is None:
            flat = self._rb._codec.decode(clip=False)[0].ravel()
            self._values = flat[self._incoming()]
        return self._values

    @values.setter
    def values(self, v):
        self._values = np.asarray(v)

    @property
    def subregions(self):
        if self._sub is None:
            off, vals, pts = self.offsets, self.values, self._rc(self._incoming())
            self._sub = [RegionView(pts[int(off[r]):int(off[r + 1])], vals[int(off[r]):int(off[r + 1])], None)
                         for r in range(len(off) - 1)]
        return self._sub


class _LevelDict(dict):
    """{level: RegionCollectionView} for levels 1..L+1; the views are created on first access, and every way
    of looking at the dict (len, in, iteration, keys/values/items) sees all L+1 levels."""

    def __init__(self, rb):
        super().__init__()
        self._rb = rb

    def _levels(self):
        return range(1, self._rb.levels + 2)

    def __missing__(self, level):
        if not isinstance(level, (int, np.integer)) or not 1 <= level <= self._rb.levels + 1:
            raise KeyError(level)
        v = RegionCollectionView(self._rb, int(level))
        self[int(level)] = v
        return v

    def __contains__(self, level):
        return isinstance(level, (int, np.integer)) and 1 <= level <= self._rb.levels + 1

    def __len__(self):
        return self._rb.levels + 1

    def __iter__(self):
        return iter(self._levels())

    def keys(self):
        return self._levels()

    def values(self):
        return [self[lev] for lev in self._levels()]

    def items(self):
        return [(lev, self[lev]) for lev in self._levels()]

    def get(self, level, default=None):
        return self[level] if level in self else default


class Rbepwt:
    """Transform driver (reference: class Rbepwt, rbepwt.py:1974-2246)."""

    def __init__(self, img, levels, wavelet, path_type="easypath", paths_first_level=False, region_collection=None):
        if region_collection is not None:
            raise NotImplementedError("encoding a bare RegionCollection is outside the B200 hot path")
        if 2 ** levels > img.size:  # rbepwt.py:1977-1978
            raise Exception("2^levels must be smaller or equal to the number of pixels in the image")
        if type(img).__name__ != "Image":  # rbepwt.py:1979-1980
            raise Exception("First argument must be an Image instance")
        self.img = img
        self.levels = levels
        self.has_encoding = False
        self.wavelet = wavelet
        self.path_type = path_type
        self.paths_first_level = paths_first_level
        self.region_collection = None
        self._codec = None
        self._details = None  # host mirror {level: ndarray}, views of _flat until reassigned
        self._flat = None
        self._approx = None

    # -- encode / threshold / decode --------------------------------------------------------
    def encode(self, onlypaths=False, euclidean_distance=True):
        if onlypaths:
            raise NotImplementedError("use rbepwt_b200.full_decode(); the reference's onlypaths branch is broken "
                                      "under current numpy (rbepwt.py:2016, 2038-2039)")
        img = self.img
        if self.path_type != "epwt-easypath" and not img.has_segmentation:
            print("Segmenting image with default parameters...")  # rbepwt.py:2001-2003
            img.segment()
        labels = None if self.path_type == "epwt-easypath" else img.label_img
        _release_codec(self._codec)
        self._codec = _acquire_codec()
        self._codec.encode(img.img, labels, self.levels, self.wavelet, self.path_type, euclidean_distance,
                           paths_first_level=self.paths_first_level)
        self._euclid = bool(euclidean_distance)
        self._details = self._flat = self._approx = None
        self.region_collection_at_level = _LevelDict(self)
        self.has_encoding = True

    def __del__(self):
        try:
            _release_codec(self.__dict__.pop("_codec", None))
        except Exception:  # noqa: BLE001  (interpreter shutdown)
            pass

    # -- persistence (Image.save_pickle / load_pickle, rbepwt.py:447-472) ----------------------------------------
    # The state lives on the GPU; what is pickled is what determines it: the parameters and the coefficients (with any
    # edits and thresholding).  The paths are a function of the label map / image and are recomputed on load, the way
    # full_decode recomputes them (rbepwt.py:106-130).
    def __getstate__(self):
        d = dict(self.__dict__)
        codec = d.pop("_codec", None)
        d.pop("region_collection_at_level", None)
        d["_details"] = d["_flat"] = d["_approx"] = None
        d["_saved_flat"] = self.flat_wavelet().copy() if (self.has_encoding and codec is not None) else d.get("_saved_flat")
        d["_saved_euclid"] = getattr(self, "_euclid", True)
        return d

    def __setstate__(self, d):
        self.__dict__.update(d)
        self._codec = None

    def _restore_gpu_state(self):
        """After unpickling: encode again (same inputs, same paths), then put the saved coefficients back."""
        flat = self.__dict__.pop("_saved_flat", None)
        if flat is None or not self.has_encoding:
            return
        self.encode(euclidean_distance=self.__dict__.pop("_saved_euclid", True))
        self._codec.set_coefs(flat, 0)

    def _level_slices(self):
        n, off, out = self.img.size, 0, {}
        for lev in range(1, self.levels + 1):
            out[lev] = (off, off + (n >> lev))
            off += n >> lev
        out[self.levels + 1] = (off, n)
        return out

    def _materialise(self):
        if self._details is None:
            self._flat = self._codec.coefs(0)
            sl = self._level_slices()
            self._details = {lev: self._flat[a:b] for lev, (a, b) in sl.items() if lev <= self.levels}
            a, b = sl[self.levels + 1]
            self._approx = self._flat[a:b]

    @property
    def wavelet_details(self):
        self._materialise()
        return self._details

    @wavelet_details.setter
    def wavelet_details(self, d):
        self._materialise()
        self._details = d

    def _approx_array(self):
        self._materialise()
        return self._approx

    def _set_approx_array(self, v):
        self._materialise()
        self._approx = v

    def _upload_if_mirrored(self):
        """Host mirrors, once handed out, are the source of truth (callers edit them in place)."""
        if self._details is None:
            return
        flat = np.concatenate([np.asarray(self._details[lev], dtype=np.float64) for lev in range(1, self.levels + 1)]
                              + [np.asarray(self._approx, dtype=np.float64)])
        if flat.size != self.img.size:
            raise Exception("wavelet_details / approximation have the wrong total length")
        self._codec.set_coefs(flat, 0)

    def _refresh_mirror(self):
        if self._details is None:
            return
        flat = self._codec.coefs(0)
        for lev, (a, b) in self._level_slices().items():
            dst = self._details[lev] if lev <= self.levels else self._approx
            dst[...] = flat[a:b]  # in place: callers may hold references (rbepwt.py:2105-2112 writes in place)

    def threshold_coefs(self, ncoefs):
        """Sets to 0 all but the ncoefs coefficients of largest absolute value (rbepwt.py:2081-2112)."""
        self._upload_if_mirrored()
        self._codec.threshold(int(ncoefs))
        self._refresh_mirror()

    def threshold_by_percentage(self, perc):
        """Keeps only perc proportion of coefficients for each region (rbepwt.py:2120-2192)."""
        self._upload_if_mirrored()
        self._codec.threshold_by_percentage(perc)
        self._refresh_mirror()

    def decode(self):
        """Returns the decoded region collection like the reference (rbepwt.py:2055-2079); the clipped image is
        kept in `decoded_img` for Image.decode_rbepwt."""
        if not self.has_encoding:
            raise Exception("There is no saved encoding to decode")  # rbepwt.py:2057-2058
        self._upload_if_mirrored()
        self.decoded_img = self._codec.decode()[0]
        print("\n--DECODING: finished working on level 1 ")
        return DecodedRegionCollectionView(self)

    def flat_wavelet(self):
        """details[1] | ... | details[L] | approximation (rbepwt.py:2195-2204)."""
        if self._details is None:
            return self._codec.coefs(0)
        return np.concatenate([np.asarray(self._details[lev]) for lev in range(1, self.levels + 1)]
                              + [np.asarray(self._approx)])

    def wavelet_coefs_dict(self):
        out = self.wavelet_details
        out[self.levels + 1] = self._approx_array()
        return out


class Dwt:
    """The tensor-product baseline (reference: class Dwt, rbepwt.py:2249-2298): pywt.wavedec2 / waverec2 with
    mode='periodization', global top-k thresholding -- on the GPU (csrc/dwt2.cuh).  `wavelet_coefs` is PyWavelets'
    list [cA_L, (cH_L, cV_L, cD_L), ..., (cH_1, cV_1, cD_1)]: views of a host mirror of the coefficient pyramid that is
    created on first access and, once created, is the source of truth (uploaded before the next threshold / decode)."""

    def __init__(self, img, levels, wavelet):
        if 2 ** levels > img.size:
            raise Exception("2^levels must be smaller or equal to the number of pixels in the image")
        if type(img).__name__ != "Image":
            raise Exception("First argument must be an Image instance")
        self.img = img
        self.levels = levels
        self.has_encoding = False
        self.wavelet = wavelet
        self._codec = None
        self._pyr = None

    def encode(self):
        _release_codec(self._codec)
        self._codec = _acquire_codec()
        self._codec.dwt2_encode(np.asarray(self.img.img, dtype=np.float64)[None], self.levels, self.wavelet)
        self._pyr = None
        self.has_encoding = True

    def __del__(self):
        try:
            _release_codec(self.__dict__.pop("_codec", None))
        except Exception:  # noqa: BLE001
            pass

    def __getstate__(self):
        d = dict(self.__dict__)
        codec = d.pop("_codec", None)
        d["_saved_pyr"] = self._pyramid().copy() if (self.has_encoding and codec is not None) else d.get("_saved_pyr")
        d["_pyr"] = None
        return d

    def __setstate__(self, d):
        self.__dict__.update(d)
        self._codec = None

    def _restore_gpu_state(self):
        pyr = self.__dict__.pop("_saved_pyr", None)
        if pyr is None or not self.has_encoding:
            return
        self.encode()
        self._codec.set_coefs(pyr.ravel(), 0)

    def _pyramid(self):
        if self._pyr is None:
            side = self.img.shape[1]
            self._pyr = self._codec.coefs(0).reshape(-1, side)
        return self._pyr

    @property
    def wavelet_coefs(self):
        p = self._pyramid()
        out = []
        for lev in range(1, self.levels + 1):
            h = p.shape[1] >> lev
            out.append((p[h:2 * h, 0:h], p[0:h, h:2 * h], p[h:2 * h, h:2 * h]))  # cH ('da'), cV ('ad'), cD ('dd')
        h = p.shape[1] >> self.levels
        return [p[0:h, 0:h]] + out[::-1]

    def _upload_if_mirrored(self):
        if self._pyr is not None:
            self._codec.set_coefs(np.ascontiguousarray(self._pyr).ravel(), 0)

    def threshold_coefs(self, ncoefs):
        self._upload_if_mirrored()
        self._codec.threshold(int(ncoefs))
        if self._pyr is not None:
            self._pyr[...] = self._codec.coefs(0).reshape(self._pyr.shape)

    def decode(self):
        """pywt.waverec2 of the (possibly thresholded) coefficients -- not clipped (rbepwt.py:2262-2263)."""
        self._upload_if_mirrored()
        return self._codec.decode(clip=False)[0]


class Image:
    """Reference facade (rbepwt.py:190-577), hot-path subset."""

    def __init__(self):
        self.has_segmentation = False
        self.has_decoded_img = False
        self.method = None
        self.segmentation_method = None

    def __getitem__(self, idx):
        return self.img[idx]

    def read(self, filepath):
        """skimage.io.imread(filepath, as_grey=True) semantics (rbepwt.py:200-206): a grayscale file stays
        uint8, a colour file becomes float64 luminance in [0,1]."""
        import PIL.Image

        with PIL.Image.open(filepath) as im:
            if im.mode in ("L", "P", "1", "I;16", "I"):
                arr = np.array(im.convert("L") if im.mode in ("P", "1") else im)
            else:
                rgb = np.asarray(im.convert("RGB"), dtype=np.float64) / 255.0
                arr = rgb @ np.array([0.2125, 0.7154, 0.0721])
        self.read_array(arr)
        self.imgpath = filepath

    def read_array(self, array):
        self.img = array
        self.size = self.img.size
        self.shape = self.img.shape

    def segment(self, method="felzenszwalb", **args):
        """Image.segment (rbepwt.py:220-245).  method='felzenszwalb' (779-785): scikit-image's own function when it is
        installed -- the reference's dependency --, else the library's restatement of it (csrc/segment.hpp,
        rbepwt_felzenszwalb: host code, parity with scikit-image unpinned).  Other label maps: set_labels()."""
        if method != "felzenszwalb":
            raise NotImplementedError("only method='felzenszwalb' or set_labels() / load_mat_segmentation()")
        scale, sigma, min_size = args.get("scale", 200), args.get("sigma", 2), args.get("min_size", 10)
        try:
            from skimage.segmentation import felzenszwalb
            lab = felzenszwalb(self.img, scale=float(scale), sigma=float(sigma), min_size=int(min_size))
        except ImportError:
            lab = felzenszwalb_labels(self.img, scale, sigma, min_size)
        self.set_labels(lab, "felzenszwalb")
        self.felz_scale, self.felz_sigma, self.felz_min_size = scale, sigma, min_size

    def set_labels(self, label_img, segm_method="external"):
        """Inject an externally computed label map, as load_mat_segmentation does (rbepwt.py:250-258)."""
        label_img = np.asarray(label_img)
        if label_img.shape != self.img.shape:
            raise ValueError("label map must have the shape of the image")
        self.segmentation = Segmentation(self.img)
        self.segmentation_method = segm_method
        self.label_img = label_img
        self.segmentation.label_img = label_img
        self.segmentation.nlabels = int(label_img.max()) + 1
        self.segmentation.compute_label_dict()
        self.has_segmentation = True

    def load_mat_segmentation(self, filepath, offset=-1, matlabvar="labels", segm_method="tbes"):
        import scipy.io

        self.set_labels((scipy.io.loadmat(filepath)[matlabvar] + offset).astype("int"), segm_method)

    def encode_rbepwt(self, levels, wavelet, path_type="easypath", euclidean_distance=True, paths_first_level=False):
        self.method = "rbepwt"
        self.rbepwt_path_type = path_type
        if not ispowerof2(self.img.size):
            raise Exception("Image size must be a power of 2")  # rbepwt.py:301-302
        path_mode(path_type, euclidean_distance)  # validates
        self.rbepwt_levels = levels
        self.rbepwt = Rbepwt(self, levels, wavelet, path_type=path_type, paths_first_level=paths_first_level)
        self.rbepwt.encode(euclidean_distance=euclidean_distance)

    def decode_rbepwt(self):
        self.decoded_region_collection = self.rbepwt.decode()  # decoded values by region, fetched on access
        self.decoded_img = self.rbepwt.decoded_img  # float64, clipped to [0,255] on the GPU, not rounded (rbepwt.py:312-314)
        self.has_decoded_img = True

    def save_pickle(self, filepath):
        """pickle.dump(self.__dict__) like the reference (rbepwt.py:447-450); the GPU-resident encoding is saved as its
        parameters + coefficients (Rbepwt.__getstate__)."""
        import pickle

        with open(filepath, "wb") as f:
            pickle.dump(self.__dict__, f, 3)

    def load_pickle(self, filepath):
        """rbepwt.py:452-472: restore the attributes, re-link the transform object, re-create its GPU state."""
        import pickle

        with open(filepath, "rb") as f:
            tmpdict = pickle.load(f)
        self.__dict__.update(tmpdict)
        self.has_segmentation = hasattr(self, "label_img")
        self.has_decoded_img = hasattr(self, "decoded_img")
        if hasattr(self, "dwt"):
            self.dwt.img = self
            self.dwt._restore_gpu_state()
        if hasattr(self, "rbepwt"):
            self.rbepwt.img = self
            self.rbepwt._restore_gpu_state()
            if self.has_decoded_img and hasattr(self, "decoded_region_collection"):
                self.decoded_region_collection._rb = self.rbepwt

    def encode_dwt(self, levels, wavelet):
        """The 2-D DWT baseline (rbepwt.py:318-324)."""
        self.method = "dwt"
        if not ispowerof2(self.img.size):
            raise Exception("Image size must be a power of 2")
        self.dwt = Dwt(self, levels, wavelet)
        self.dwt_levels = levels
        self.dwt.encode()

    def decode_dwt(self):
        """rbepwt.py:326-333: waverec2, then the clip to [0,255]."""
        self.decoded_img = np.clip(self.dwt.decode(), 0.0, 255.0)
        self.has_decoded_img = True

    def encode_epwt(self, levels, wavelet):
        self.method = "epwt"
        self.encode_rbepwt(levels, wavelet, "epwt-easypath")  # note: leaves method == 'rbepwt' (rbepwt.py:335-337)

    def decode_epwt(self):
        self.decode_rbepwt()

    def threshold_coefs(self, ncoefs):
        if self.method in ("epwt", "rbepwt"):
            self.rbepwt.threshold_coefs(ncoefs)
        elif self.method == "dwt":
            self.dwt.threshold_coefs(ncoefs)

    def psnr(self, filtered=False):
        """PSNR of the decoded image vs. the original (rbepwt.py:361-368)."""
        if filtered:
            raise NotImplementedError("Image.filter is outside the B200 hot path")
        codec = self.dwt._codec if self.method == "dwt" else self.rbepwt._codec
        v = float(codec.psnr(np.asarray(self.img, dtype=np.float64)[None], self.decoded_img[None])[0])
        return -1 if v == -1.0 else v

    def nonzero_coefs(self):
        if self.method in ("rbepwt", "epwt"):
            self.rbepwt._upload_if_mirrored()
            return int(self.rbepwt._codec.nonzero_coefs()[0])
        self.dwt._upload_if_mirrored()  # rbepwt.py:434-439
        return int(self.dwt._codec.nonzero_coefs()[0])


def full_decode(wavelet_details_dict, wavelet_approx, label_img, wavelet, path_type="easypath",
                euclidean_distance=True):
    """Decoded image from coefficients + label map only: every path is recomputed (rbepwt.py:106-130)."""
    levels = len(wavelet_details_dict)
    flat = np.concatenate([np.asarray(wavelet_details_dict[lev], dtype=np.float64) for lev in range(1, levels + 1)]
                          + [np.asarray(wavelet_approx, dtype=np.float64)])
    label_img = np.asarray(label_img)
    if flat.size != label_img.size:
        raise Exception("coefficient count does not match the label image")
    codec = _acquire_codec()
    try:
        return codec.full_decode(flat[None], label_img[None], levels, wavelet, path_type, euclidean_distance)[0]
    finally:
        _release_codec(codec)
