"""Batched host-side driver of the CUDA path: thin Python over the C ABI (include/rbepwt_b200.h).

`BatchCodec` is what a throughput user calls: B images of one shape in, coefficients / decoded
images out, everything resident on one GPU between the three calls.  Inputs may be numpy arrays
(host path: the library copies host<->device inside the call) or torch CUDA tensors (device path:
raw data_ptr()s are passed, no copy).  The reference-shaped single-image facade is image.py.
"""
import ctypes

import numpy as np

from . import _capi
from .wavelets import filter_bank

PATH_MODES = {("easypath", True): _capi.PATH_EUCLID, ("easypath", False): _capi.PATH_CHEB}


def path_mode(path_type, euclidean_distance=True):
    """Reference arguments -> C-ABI path mode (rbepwt.py:2026-2031, 1301-1306)."""
    if path_type == "epwt-easypath":
        return _capi.PATH_EPWT  # euclidean_distance is ignored by the reference in this mode
    if path_type == "easypath":
        return PATH_MODES[("easypath", bool(euclidean_distance))]
    if path_type == "gradpath":
        # Region.grad_path (rbepwt.py:1190-1271).  Complete ties of its gradient preference (two opposite offsets, always)
        # are resolved by CPython set iteration order in the reference -- unpinned; here: first in row-major order.
        return _capi.PATH_GRAD if euclidean_distance else _capi.PATH_GRAD_CHEB
    raise ValueError("unknown path_type %r" % (path_type,))


def _is_torch_cuda(x):
    return type(x).__module__.startswith("torch") and getattr(x, "is_cuda", False)


def _ptr(x):
    if x is None:
        return None
    if isinstance(x, np.ndarray):
        return x.ctypes.data_as(ctypes.c_void_p)
    return ctypes.c_void_p(x.data_ptr())


def _labels_int32(labels):
    """Label maps as contiguous int32; integer labels that do not survive the cast are refused (no silent wrap)."""
    lab = np.asarray(labels)
    lab32 = np.ascontiguousarray(lab, dtype=np.int32)
    if lab.dtype != np.int32 and not np.array_equal(lab32, lab):
        raise ValueError("labels must be integers representable as int32")
    return lab32


class BatchCodec:
    """One GPU context: encode -> threshold -> decode for a batch of same-shape images."""

    def __init__(self, device=0, stream=None):
        self._lib = _capi.lib()
        self._ctx = ctypes.c_void_p()
        sp = ctypes.c_void_p(int(stream)) if stream else None
        _capi.check(self._lib.rbepwt_create(int(device), sp, ctypes.byref(self._ctx)))
        self.device = int(device)
        self.shape = None       # (B, H, W)
        self.levels = None
        self.wavelet = None
        self._keep = []         # keeps device inputs alive while the library references them
        self._bank_key = None

    def close(self):
        if getattr(self, "_ctx", None):
            self._lib.rbepwt_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    # -- configuration -----------------------------------------------------------------
    def set_wavelet(self, wavelet):
        dl, dh, rl, rh = filter_bank(wavelet)
        key = b"".join(f.tobytes() for f in (dl, dh, rl, rh))
        if key != self._bank_key:  # the upload synchronises the stream: skip it when nothing changed
            _capi.check(self._lib.rbepwt_set_wavelet(self._ctx, dl.size, _ptr(dl), _ptr(dh), _ptr(rl), _ptr(rh)))
            self._bank_key = key
        self.wavelet = wavelet

    def enable_timing(self, on=True):
        _capi.check(self._lib.rbepwt_enable_timing(self._ctx, int(bool(on))))

    def timings(self):
        """{stage: ms} summed over the calls since the previous timings() (CUDA events on the context's stream)."""
        n = len(_capi.T_NAMES)
        ms = (ctypes.c_float * n)()
        _capi.check(min(0, self._lib.rbepwt_get_timings(self._ctx, ms, n)))
        return dict(zip(_capi.T_NAMES, [float(v) for v in ms]))

    def stage_launches(self):
        """{stage: kernel launches} covered by the last timings() call."""
        n = len(_capi.T_NAMES)
        cnt = (ctypes.c_int64 * n)()
        self._lib.rbepwt_get_stage_launches(self._ctx, cnt, n)
        return dict(zip(_capi.T_NAMES, [int(v) for v in cnt]))

    def launch_count(self):
        return int(self._lib.rbepwt_launch_count(self._ctx))

    def sync(self):
        _capi.check(self._lib.rbepwt_sync(self._ctx))

    # -- the path --------------------------------------------------------------------------
    def _check_f64_buffer(self, x, numel, what):
        """A buffer the C ABI reads or writes `numel` doubles through: contiguous float64 of exactly that many
        elements, a numpy array (host path) or a CUDA tensor on this codec's device (device path).  Returns True
        for the device path.  The library sees a raw pointer, so everything is checked here."""
        if _is_torch_cuda(x):
            import torch
            if x.dtype != torch.float64 or not x.is_contiguous():
                raise ValueError("%s must be a contiguous float64 CUDA tensor" % what)
            if x.device.index != self.device:
                raise ValueError("%s lives on cuda:%s, the codec on cuda:%d" % (what, x.device.index, self.device))
            if x.numel() != numel:
                raise ValueError("%s must hold %d elements, not %d" % (what, numel, x.numel()))
            return True
        if not (isinstance(x, np.ndarray) and x.dtype == np.float64 and x.flags.c_contiguous):
            raise ValueError("%s must be a contiguous float64 numpy array or CUDA tensor" % what)
        if x.size != numel:
            raise ValueError("%s must hold %d elements, not %d" % (what, numel, x.size))
        return False

    def _prep(self, imgs, labels, mode):
        dev = _is_torch_cuda(imgs)
        u8 = False
        if dev:
            import torch
            if imgs.dtype != torch.float64 or not imgs.is_contiguous():
                raise ValueError("device images must be contiguous float64")
            if imgs.device.index != self.device:
                raise ValueError("images live on cuda:%s, the codec on cuda:%d" % (imgs.device.index, self.device))
            if labels is not None and (not _is_torch_cuda(labels) or labels.dtype != torch.int32
                                       or not labels.is_contiguous() or labels.device != imgs.device):
                raise ValueError("device labels must be contiguous int32 CUDA tensors on the images' device")
            shape = tuple(imgs.shape)
        else:
            imgs = np.asarray(imgs)
            u8 = imgs.dtype == np.uint8
            imgs = np.ascontiguousarray(imgs, dtype=np.float64)
            if labels is not None:
                labels = _labels_int32(labels)
            shape = imgs.shape
        if len(shape) == 2:
            shape = (1,) + tuple(shape)
        if len(shape) != 3:
            raise ValueError("images must be [H,W] or [B,H,W]")
        if mode != _capi.PATH_EPWT:
            if labels is None:
                raise ValueError("a label map is required for path_type='easypath'")
            lshape = tuple(labels.shape)
            if (lshape if len(lshape) == 3 else (1,) + lshape) != tuple(shape):
                raise ValueError("labels must have the shape of the images")
        else:
            labels = None
        return imgs, labels, shape, dev, u8

    def encode(self, imgs, labels, levels, wavelet, path_type="easypath", euclidean_distance=True,
               paths_first_level=False):
        mode = path_mode(path_type, euclidean_distance)
        imgs, labels, shape, dev, u8 = self._prep(imgs, labels, mode)
        self.set_wavelet(wavelet)
        flags = (_capi.DEVICE_PTRS if dev else 0) | (_capi.U8_WRAP if (u8 and mode == _capi.PATH_EPWT) else 0)
        flags |= _capi.PATHS_FIRST_LEVEL if paths_first_level else 0
        B, H, W = shape
        self._keep = [imgs, labels]
        _capi.check(self._lib.rbepwt_encode(self._ctx, _ptr(imgs), _ptr(labels), B, H, W, int(levels), mode, flags))
        self.shape, self.levels, self.mode = shape, int(levels), mode
        return self

    def set_option(self, streams=None, sub_batch=None, path_group=None, coop_limit=None):
        """streams: 1 (all kernels on one stream) or 2 units in flight; sub_batch / path_group: images per
        transform sub-batch / per path group (0 = auto)."""
        if path_group is not None:
            _capi.check(self._lib.rbepwt_set_option(self._ctx, _capi.OPT_PATHGROUP, int(path_group)))
        if coop_limit is not None:
            _capi.check(self._lib.rbepwt_set_option(self._ctx, _capi.OPT_COOP_LIMIT, int(coop_limit)))
        if streams is not None:
            _capi.check(self._lib.rbepwt_set_option(self._ctx, _capi.OPT_STREAMS, int(streams)))
        if sub_batch is not None:
            _capi.check(self._lib.rbepwt_set_option(self._ctx, _capi.OPT_SUBBATCH, int(sub_batch)))

    def transcode(self, imgs, labels, levels, wavelet, ncoefs, path_type="easypath", euclidean_distance=True, out=None,
                  paths_first_level=False):
        """encode -> threshold(ncoefs) -> decode in one pipelined call (rbepwt_transcode); returns the decoded
        images.  Inputs and `out` must be all numpy (host path) or all torch CUDA tensors (device path)."""
        mode = path_mode(path_type, euclidean_distance)
        imgs, labels, shape, dev, u8 = self._prep(imgs, labels, mode)
        self.set_wavelet(wavelet)
        if out is None:
            if dev:
                import torch
                out = torch.empty_like(imgs)
            else:
                out = np.empty(shape, dtype=np.float64)
        B, H, W = shape
        if self._check_f64_buffer(out, B * H * W, "out") != dev:
            raise ValueError("`out` must live where the inputs live")
        flags = (_capi.DEVICE_PTRS if dev else 0) | (_capi.U8_WRAP if (u8 and mode == _capi.PATH_EPWT) else 0)
        flags |= _capi.PATHS_FIRST_LEVEL if paths_first_level else 0
        self._keep = [imgs, labels, out]
        _capi.check(self._lib.rbepwt_transcode(self._ctx, _ptr(imgs), _ptr(labels), B, H, W, int(levels), mode, int(ncoefs),
                                               _ptr(out), flags))
        self.shape, self.levels, self.mode = shape, int(levels), mode
        return out

    def transcode_ex(self, imgs, labels, levels, wavelet, ncoefs, path_type="easypath", euclidean_distance=True,
                     paths_first_level=False, out=None, out_dtype=None, want_image=True, want_psnr=False, want_kept=False):
        """encode -> threshold(ncoefs) -> decode with narrow element types and the outputs a codec needs
        (rbepwt_transcode_ex).  Pixels may be float64, float32 or uint8, labels int32 or uint16: they cross PCIe as
        they are and are widened (exactly) on the GPU.  Returns a dict with the requested entries:
          'image'    decoded images [B,H,W] in `out_dtype` (float64 default; float32 / uint8 are conveniences the
                     reference does not compute) -- `out` may supply the destination,
          'psnr'     float64 [B], psnr(input, decoded) per image,
          'kept_idx', 'kept_val'  int32 / float64 [B, ncoefs]: the surviving coefficients, flat index ascending.
        With want_image and want_psnr both False nothing is decoded.  Inputs / `out`: all numpy, or all CUDA
        tensors on this codec's device."""
        mode = path_mode(path_type, euclidean_distance)
        dev = _is_torch_cuda(imgs)
        pix_codes = {"float64": _capi.F64, "float32": _capi.F32, "uint8": _capi.U8}
        lab_codes = {"int32": _capi.I32, "uint16": _capi.U16}

        def tname(x):
            return str(x.dtype).replace("torch.", "")

        if dev:
            if tname(imgs) not in pix_codes or not imgs.is_contiguous() or imgs.device.index != self.device:
                raise ValueError("device images must be contiguous float64 / float32 / uint8 tensors on cuda:%d" % self.device)
            if labels is not None and (not _is_torch_cuda(labels) or tname(labels) not in lab_codes
                                       or not labels.is_contiguous() or labels.device != imgs.device):
                raise ValueError("device labels must be contiguous int32 / uint16 CUDA tensors on the images' device")
        else:
            imgs = np.ascontiguousarray(imgs)
            if tname(imgs) not in pix_codes:
                imgs = np.ascontiguousarray(imgs, dtype=np.float64)
            if labels is not None:
                labels = np.ascontiguousarray(labels)
                if tname(labels) not in lab_codes:
                    labels = _labels_int32(labels)
        shape = tuple(imgs.shape)
        if len(shape) == 2:
            shape = (1,) + shape
        if len(shape) != 3:
            raise ValueError("images must be [H,W] or [B,H,W]")
        if mode == _capi.PATH_EPWT:
            labels = None
        elif labels is None:
            raise ValueError("a label map is required for path_type='easypath'")
        elif tuple(labels.shape) not in (shape, shape[1:]):
            raise ValueError("labels must have the shape of the images")
        B, H, W = shape
        self.set_wavelet(wavelet)
        res = {}
        out_code = _capi.F64
        if want_image:
            odt = np.dtype(out_dtype or (tname(out) if out is not None else "float64")).name
            if odt not in pix_codes:
                raise ValueError("out_dtype must be float64, float32 or uint8")
            out_code = pix_codes[odt]
            if out is None:
                if dev:
                    import torch
                    out = torch.empty(shape, dtype=getattr(torch, odt), device=imgs.device)
                else:
                    out = np.empty(shape, dtype=odt)
            if _is_torch_cuda(out) != dev or tname(out) != odt or int(np.prod(tuple(out.shape))) != B * H * W:
                raise ValueError("`out` must live where the inputs live, hold B*H*W elements and have dtype %s" % odt)
            if not (out.is_contiguous() if dev else out.flags.c_contiguous):
                raise ValueError("`out` must be contiguous")
            res["image"] = out
        if want_psnr:
            res["psnr"] = np.empty(B, dtype=np.float64)
        if want_kept:
            res["kept_idx"] = np.empty((B, int(ncoefs)), dtype=np.int32)
            res["kept_val"] = np.empty((B, int(ncoefs)), dtype=np.float64)
        flags = (_capi.DEVICE_PTRS if dev else 0) | (_capi.PATHS_FIRST_LEVEL if paths_first_level else 0)
        self._keep = [imgs, labels, out]
        _capi.check(self._lib.rbepwt_transcode_ex(
            self._ctx, _ptr(imgs), pix_codes[tname(imgs)], _ptr(labels), lab_codes[tname(labels)] if labels is not None else 0,
            B, H, W, int(levels), mode, int(ncoefs), _ptr(res.get("image")), out_code, _ptr(res.get("psnr")),
            _ptr(res.get("kept_idx")), _ptr(res.get("kept_val")), flags))
        self.shape, self.levels, self.mode = shape, int(levels), mode
        return res

    def threshold(self, k):
        _capi.check(self._lib.rbepwt_threshold(self._ctx, int(k)))
        return self

    def dwt2_encode(self, imgs, levels, wavelet):
        """The tensor-product baseline (class Dwt, rbepwt.py:2249-2263): pywt.wavedec2(img, wavelet, level=levels,
        mode='periodization') of a batch of square images.  threshold / decode / coefs / set_coefs / nonzero_coefs then
        work on this encoding; coefs(b) is the H x W pyramid described in include/rbepwt_b200.h."""
        dev = _is_torch_cuda(imgs)
        if dev:
            import torch
            if imgs.dtype != torch.float64 or not imgs.is_contiguous() or imgs.device.index != self.device:
                raise ValueError("device images must be contiguous float64 on cuda:%d" % self.device)
        else:
            imgs = np.ascontiguousarray(imgs, dtype=np.float64)
        shape = tuple(imgs.shape)
        if len(shape) == 2:
            shape = (1,) + shape
        if len(shape) != 3:
            raise ValueError("images must be [H,W] or [B,H,W]")
        self.set_wavelet(wavelet)
        B, H, W = shape
        self._keep = [imgs]
        _capi.check(self._lib.rbepwt_dwt2_encode(self._ctx, _ptr(imgs), B, H, W, int(levels), _capi.DEVICE_PTRS if dev else 0))
        self.shape, self.levels, self.mode = shape, int(levels), None
        return self

    def threshold_by_percentage(self, perc):
        """Rbepwt.threshold_by_percentage (rbepwt.py:2120-2192): per region, keep that proportion of its coefficients."""
        _capi.check(self._lib.rbepwt_threshold_percentage(self._ctx, float(perc)))
        return self

    def decode(self, out=None, clip=True):
        """Decoded images, float64 [B,H,W], clipped to [0,255] (Image.decode_rbepwt, rbepwt.py:313-314) unless
        `clip` is False (what Rbepwt.decode itself returns, rbepwt.py:2055-2079).  `out` may be a torch CUDA
        tensor (device path) or a numpy array (host path); default: a new numpy array."""
        if self.shape is None:
            raise Exception("There is no saved encoding to decode")
        if out is None:
            out = np.empty(self.shape, dtype=np.float64)
        B, H, W = self.shape
        dev = self._check_f64_buffer(out, B * H * W, "out")
        flags = (_capi.DEVICE_PTRS if dev else 0) | (0 if clip else _capi.NO_CLIP)
        _capi.check(self._lib.rbepwt_decode(self._ctx, _ptr(out), flags))
        return out

    def full_decode(self, coefs, labels, levels, wavelet, path_type="easypath", euclidean_distance=True,
                    paths_first_level=False):
        """Decoder side (reference full_decode, rbepwt.py:106-130): paths regenerated from labels."""
        mode = path_mode(path_type, euclidean_distance)
        labels = _labels_int32(labels)
        if labels.ndim not in (2, 3):
            raise ValueError("labels must be [H,W] or [B,H,W]")
        shape = labels.shape if labels.ndim == 3 else (1,) + labels.shape
        coefs = np.ascontiguousarray(coefs, dtype=np.float64)
        if coefs.size != labels.size:
            raise ValueError("coefs must hold B*H*W = %d values, not %d" % (labels.size, coefs.size))
        coefs = coefs.reshape(shape[0], -1)
        self.set_wavelet(wavelet)
        out = np.empty(shape, dtype=np.float64)
        B, H, W = shape
        _capi.check(self._lib.rbepwt_full_decode(self._ctx, _ptr(coefs), _ptr(labels), B, H, W, int(levels), mode,
                                                 _ptr(out), _capi.PATHS_FIRST_LEVEL if paths_first_level else 0))
        self.shape, self.levels, self.mode = shape, int(levels), mode
        return out

    def psnr(self, a, b):
        """psnr(a[i], b[i]) per image (rbepwt.py:156-162); a and b are both numpy arrays or both CUDA tensors on
        this codec's device, [B,H,W] or a single [H,W]."""
        dev = _is_torch_cuda(a)
        if dev != _is_torch_cuda(b):
            raise ValueError("psnr: a and b must both be numpy arrays or both CUDA tensors")
        if not dev:
            a = np.ascontiguousarray(a, dtype=np.float64)
            b = np.ascontiguousarray(b, dtype=np.float64)
        shape = tuple(a.shape)
        if tuple(b.shape) != shape:
            raise ValueError("psnr: a and b must have the same shape")
        B = shape[0] if len(shape) == 3 else 1
        n = int(np.prod(shape)) // B
        self._check_f64_buffer(a, B * n, "a")
        self._check_f64_buffer(b, B * n, "b")
        out = np.empty(B, dtype=np.float64)
        _capi.check(self._lib.rbepwt_psnr(self._ctx, _ptr(a), _ptr(b), B, n, _ptr(out), _capi.DEVICE_PTRS if dev else 0))
        return out

    def nonzero_coefs(self):
        out = np.empty(self.shape[0], dtype=np.int64)
        _capi.check(self._lib.rbepwt_nonzero_coefs(self._ctx, _ptr(out)))
        return out

    # -- views (host copies) ---------------------------------------------------------------
    @property
    def npix(self):
        return self.shape[1] * self.shape[2]

    def coefs(self, b=0):
        out = np.empty(self.npix, dtype=np.float64)
        _capi.check(self._lib.rbepwt_get_coefs(self._ctx, int(b), _ptr(out)))
        return out

    def set_coefs(self, flat, b=0):
        flat = np.ascontiguousarray(flat, dtype=np.float64)
        if flat.size != self.npix:
            raise ValueError("flat coefficient vector must have H*W entries")
        _capi.check(self._lib.rbepwt_set_coefs(self._ctx, int(b), _ptr(flat)))

    def region_count(self, b=0):
        r = ctypes.c_int32()
        _capi.check(self._lib.rbepwt_region_count(self._ctx, int(b), ctypes.byref(r)))
        return int(r.value)

    def region_offsets(self, b=0, level=1):
        """int32 [R+1] offsets of the regions in the level's concatenated signal."""
        off = np.empty(self.region_count(b) + 1, dtype=np.int32)
        _capi.check(self._lib.rbepwt_region_offsets(self._ctx, int(b), _ptr(off)))
        if level > 1:
            sh = level - 1
            off = ((off.astype(np.int64) + (1 << sh) - 1) >> sh).astype(np.int32)
        return off

    def region_labels(self, b=0):
        lab = np.empty(self.region_count(b), dtype=np.int32)
        _capi.check(self._lib.rbepwt_region_labels(self._ctx, int(b), _ptr(lab)))
        return lab

    def paths(self, b=0, level=1):
        """Pixel ids in path order at `level` (0: level-1 incoming order, L+1: approximation points)."""
        n = self.npix >> max(level - 1, 0)
        out = np.empty(n, dtype=np.int32)
        _capi.check(self._lib.rbepwt_get_paths(self._ctx, int(b), int(level), _ptr(out)))
        return out

    def perm(self, b=0, level=1):
        out = np.empty(self.npix >> (level - 1), dtype=np.int32)
        _capi.check(self._lib.rbepwt_get_perm(self._ctx, int(b), int(level), _ptr(out)))
        return out

    def level_values(self, b=0, level=1):
        out = np.empty(self.npix >> (level - 1), dtype=np.float64)
        _capi.check(self._lib.rbepwt_get_level_values(self._ctx, int(b), int(level), _ptr(out)))
        return out


def encode_threshold_decode(imgs, labels, levels, wavelet, ncoefs, path_type="easypath",
                            euclidean_distance=True, codec=None, out=None):
    """The whole hot path for a batch in one call (pipelined, rbepwt_transcode); returns the decoded images."""
    codec = codec or BatchCodec()
    return codec.transcode(imgs, labels, levels, wavelet, ncoefs, path_type, euclidean_distance, out)
