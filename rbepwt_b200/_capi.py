"""ctypes binding of the C ABI (include/rbepwt_b200.h).  No CPU fallback: if the CUDA library
cannot be loaded or no GPU is present, calls raise."""
import ctypes
import os

from . import build as _build

_lib = None

E_NOT_POW2, E_LEVELS, E_NO_ENCODING, E_NO_GPU = -2, -3, -4, -7
PATH_EUCLID, PATH_CHEB, PATH_EPWT, PATH_GRAD, PATH_GRAD_CHEB = 0, 1, 2, 3, 4
DEVICE_PTRS, U8_WRAP, PATHS_FIRST_LEVEL, NO_CLIP = 1, 2, 4, 8
OPT_STREAMS, OPT_SUBBATCH, OPT_PATHGROUP, OPT_COOP_LIMIT = 1, 2, 3, 4
F64, F32, U8 = 0, 1, 2   # pixel element types of rbepwt_transcode_ex
I32, U16 = 0, 1         # label element types
T_NAMES = ["h2d", "regions", "paths", "dwt", "select", "idwt", "d2h", "paths_big", "perm"]

EXPORTS = [
    "rbepwt_create", "rbepwt_destroy", "rbepwt_last_error", "rbepwt_sync", "rbepwt_set_wavelet",
    "rbepwt_felzenszwalb", "rbepwt_encode", "rbepwt_threshold", "rbepwt_threshold_percentage", "rbepwt_decode", "rbepwt_transcode", "rbepwt_transcode_ex", "rbepwt_set_option",
    "rbepwt_full_decode", "rbepwt_dwt2_encode", "rbepwt_psnr",
    "rbepwt_nonzero_coefs", "rbepwt_get_coefs", "rbepwt_set_coefs", "rbepwt_region_count",
    "rbepwt_region_offsets", "rbepwt_region_labels", "rbepwt_get_paths", "rbepwt_get_perm",
    "rbepwt_get_level_values", "rbepwt_enable_timing", "rbepwt_get_timings", "rbepwt_get_stage_launches",
    "rbepwt_launch_count",
]


class RbepwtError(Exception):
    pass


def lib():
    """The loaded CUDA library; built in-tree with nvcc if the .so is absent or stale."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB_PATH
    if _build.needs_build():
        try:
            _build.build_library()
        except Exception as e:  # noqa: BLE001
            # never load a library older than its sources: results would silently come from stale kernels
            raise RbepwtError("rbepwt_b200: the CUDA library is %s and could not be built (%s); there is no CPU "
                              "fallback" % ("stale" if os.path.isfile(path) else "missing", e))
    L = ctypes.CDLL(path)
    vp, i32, i64, u32 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_uint
    L.rbepwt_last_error.restype = ctypes.c_char_p
    L.rbepwt_create.argtypes = [i32, vp, ctypes.POINTER(vp)]
    L.rbepwt_destroy.argtypes = [vp]
    L.rbepwt_destroy.restype = None
    L.rbepwt_sync.argtypes = [vp]
    L.rbepwt_set_wavelet.argtypes = [vp, i32, vp, vp, vp, vp]
    L.rbepwt_encode.argtypes = [vp, vp, vp, i32, i32, i32, i32, i32, u32]
    L.rbepwt_threshold.argtypes = [vp, i64]
    L.rbepwt_threshold_percentage.argtypes = [vp, ctypes.c_double]
    L.rbepwt_dwt2_encode.argtypes = [vp, vp, i32, i32, i32, i32, u32]
    L.rbepwt_decode.argtypes = [vp, vp, u32]
    L.rbepwt_transcode.argtypes = [vp, vp, vp, i32, i32, i32, i32, i32, i64, vp, u32]
    L.rbepwt_transcode_ex.argtypes = [vp, vp, i32, vp, i32, i32, i32, i32, i32, i32, i64, vp, i32, vp, vp, vp, u32]
    L.rbepwt_set_option.argtypes = [vp, i32, i64]
    L.rbepwt_felzenszwalb.argtypes = [vp, i32, i32, ctypes.c_double, ctypes.c_double, i32, vp, vp]
    L.rbepwt_full_decode.argtypes = [vp, vp, vp, i32, i32, i32, i32, i32, vp, u32]
    L.rbepwt_psnr.argtypes = [vp, vp, vp, i32, i64, vp, u32]
    L.rbepwt_nonzero_coefs.argtypes = [vp, vp]
    L.rbepwt_get_coefs.argtypes = [vp, i32, vp]
    L.rbepwt_set_coefs.argtypes = [vp, i32, vp]
    L.rbepwt_region_count.argtypes = [vp, i32, ctypes.POINTER(ctypes.c_int32)]
    L.rbepwt_region_offsets.argtypes = [vp, i32, vp]
    L.rbepwt_region_labels.argtypes = [vp, i32, vp]
    L.rbepwt_get_paths.argtypes = [vp, i32, i32, vp]
    L.rbepwt_get_perm.argtypes = [vp, i32, i32, vp]
    L.rbepwt_get_level_values.argtypes = [vp, i32, i32, vp]
    L.rbepwt_enable_timing.argtypes = [vp, i32]
    L.rbepwt_get_timings.argtypes = [vp, vp, i32]
    L.rbepwt_get_stage_launches.argtypes = [vp, vp, i32]
    L.rbepwt_launch_count.argtypes = [vp]
    L.rbepwt_launch_count.restype = i64
    _lib = L
    return L


def check(rc):
    """Non-zero status -> the reference's own exception messages (rbepwt.py:302, 1978, 2058)."""
    if rc == 0:
        return
    msg = lib().rbepwt_last_error().decode("utf-8", "replace")
    if rc in (E_NOT_POW2, E_LEVELS, E_NO_ENCODING):
        raise Exception(msg)
    raise RbepwtError(msg)
