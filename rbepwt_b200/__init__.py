"""rbepwt_b200 -- B200-native RBEPWT encode -> threshold -> decode (drop-in for that path of nareto/rbepwt).

    import rbepwt_b200 as rbepwt
    im = rbepwt.Image(); im.read_array(img); im.set_labels(label_img)
    im.encode_rbepwt(16, 'bior4.4'); im.threshold_coefs(512); im.decode_rbepwt(); im.psnr()

Throughput API: rbepwt_b200.BatchCodec (batched, device-resident), rbepwt_b200.BoxCodec (all GPUs of a box).  Everything computes in
hand-written sm_100a CUDA kernels behind the C ABI in include/rbepwt_b200.h; there is no CPU fallback.
"""
from .box import BoxCodec  # noqa: F401
from .codec import BatchCodec, encode_threshold_decode, path_mode  # noqa: F401
from .roi import Roi  # noqa: F401
from .image import felzenszwalb_labels  # noqa: F401
from .image import Dwt, Image, Rbepwt, Segmentation, full_decode, ispowerof2, psnr  # noqa: F401
from .wavelets import filter_bank, wavelist  # noqa: F401

__all__ = ["Image", "Rbepwt", "Dwt", "Roi", "felzenszwalb_labels", "Segmentation", "BatchCodec", "BoxCodec", "encode_threshold_decode", "full_decode",
           "psnr", "ispowerof2", "filter_bank", "wavelist", "path_mode"]
