"""One box, several GPUs, one process: the image batch sharded by image across the GPUs (the
only natural shard of this path -- the per-level DWT couples all regions of an image, so an image is never split,
and no collective runs on the data path).  One BatchCodec (= one CUDA context, one set of streams, one pinned
staging area inside the library) and one host thread per GPU; ctypes releases the GIL during the C-ABI calls, so
the per-GPU pipelines (H2D copies, kernels, D2H copies) run concurrently.
"""
import threading

import numpy as np

from .codec import BatchCodec
from .shard import shard_range


class BoxCodec:
    """encode -> threshold -> decode of host-resident batches on several GPUs of one box.

        box = BoxCodec(devices=range(8))
        decoded = box.transcode(imgs, labels, 16, 'bior4.4', 2048)

    Image i of the batch goes to GPU shard_range(B, ngpus, g) contains i; results land in one output array, in
    batch order, bit-identical to what a single BatchCodec produces."""

    def __init__(self, devices=None):
        if devices is None:
            import torch
            devices = range(torch.cuda.device_count())
        self.devices = [int(d) for d in devices]
        if not self.devices:
            raise ValueError("BoxCodec needs at least one device")
        self.codecs = [BatchCodec(device=d) for d in self.devices]

    def close(self):
        for c in self.codecs:
            c.close()
        self.codecs = []

    def set_option(self, **kw):
        for c in self.codecs:
            c.set_option(**kw)

    def _run(self, work):
        """work(g, codec, lo, hi) on one thread per GPU; re-raises the first failure."""
        errs = [None] * len(self.codecs)

        def runner(g):
            try:
                work(g)
            except BaseException as e:  # noqa: BLE001
                errs[g] = e

        ths = [threading.Thread(target=runner, args=(g,)) for g in range(len(self.codecs))]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        for e in errs:
            if e is not None:
                raise e

    def transcode(self, imgs, labels, levels, wavelet, ncoefs, path_type="easypath", euclidean_distance=True, out=None,
                  paths_first_level=False):
        """Decoded images float64 [B,H,W] (BatchCodec.transcode on every shard).  numpy inputs (pinned host memory
        makes the copies asynchronous); `out` may supply the destination."""
        imgs = np.asarray(imgs)
        B = imgs.shape[0]
        if out is None:
            out = np.empty(imgs.shape, dtype=np.float64)
        n = len(self.codecs)

        def work(g):
            lo, hi = shard_range(B, n, g)
            if hi > lo:
                self.codecs[g].transcode(imgs[lo:hi], None if labels is None else labels[lo:hi], levels, wavelet, ncoefs,
                                         path_type, euclidean_distance, out[lo:hi], paths_first_level)

        self._run(work)
        return out

    def transcode_ex(self, imgs, labels, levels, wavelet, ncoefs, **kw):
        """BatchCodec.transcode_ex on every shard; the per-shard result arrays are concatenated in batch order."""
        imgs = np.asarray(imgs)
        B = imgs.shape[0]
        n = len(self.codecs)
        parts = [None] * n

        def work(g):
            lo, hi = shard_range(B, n, g)
            if hi > lo:
                parts[g] = self.codecs[g].transcode_ex(imgs[lo:hi], None if labels is None else labels[lo:hi], levels, wavelet,
                                                       ncoefs, **kw)

        self._run(work)
        parts = [p for p in parts if p is not None]
        return {key: np.concatenate([p[key] for p in parts]) for key in parts[0]}
