#!/usr/bin/env python
"""bench.py -- encode + threshold + decode throughput of the RBEPWT path (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]

One "step" = one pass of the hot path (rbepwt_encode -> rbepwt_threshold -> rbepwt_decode through the
C ABI) over one batch of B distinct synthetic 512x512 images per GPU (warped-Voronoi label maps of
1024 regions, per-region ramps + noise, float64), 16 levels, bior4.4, k = 2048 kept coefficients,
path_type='easypath', euclidean_distance=True.  Weak scaling: every rank holds its own B images
(image batch sharded by image, no collective on the data path -- SURVEY.md section 8e); at N = 8 and
the default B = 512 that is BASELINE.json's "4096 x 512^2" configuration.

`value`      images/s with the batch already resident in HBM (CUDA events on the context's stream,
             barrier + synchronize both sides, max over ranks).
`e2e`        the same through the public host-buffer API (BatchCodec with numpy views of pinned host
             memory): H2D of images + labels and D2H of the decoded images inside the timed region.
`roofline`   the dominant kernel (largest share of the step's device time), algorithmic bytes per launch
             / its CUDA-event duration measured over the timed steps, against MEASURED_PEAKS.json.
`cpu_baseline` the C oracle (oracle/rbepwt_oracle.c, a port of the reference's algorithm) on host cores,
             bounded sample of the same workload; rank 0, N = 1 only.
`--impl reference` times that CPU port with all host threads (the reference itself is pure Python +
             PyWavelets, cannot travel to the GPU box and needs ~1 h per 512^2 image).
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

H = W = 512
LEVELS = 16
WAVELET = "bior4.4"
NCOEFS = 2048
NSEEDS = 1024
SEED0 = 1000
METRIC = "encode+threshold+decode images/sec (512x512, 16 levels bior4.4)"
UNIT = "images/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=512, help="images per GPU per step")
    ap.add_argument("--cpu-sample", type=int, default=0, help="images in the CPU sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-inflight", type=int, default=2, help="steps in flight in the end-to-end measurement")
    ap.add_argument("--inflight", type=int, default=1,
                    help="steps in flight in the device-resident measurement (`value`): each on its own context and stream, "
                         "so that the path kernels of one step (issue-bound) run beside the transforms of the previous one "
                         "(HBM-bound) -- measured: 45.0k against 44.8k images/s, the step is bound by the issue slots of the path kernels either "
                         "way, so the default is 1 = strictly one step at a time (extra.one_step_in_flight repeats that number)")
    ap.add_argument("--sub-batch", type=int, default=0, help="images per transform sub-batch (0 = library default)")
    ap.add_argument("--path-group", type=int, default=0, help="images per path group (0 = library default)")
    ap.add_argument("--total", type=int, default=4096,
                    help="strong-scaling line: this many images in all, total/N per rank per step (0 = skip)")
    ap.add_argument("--no-extras", action="store_true", help="skip the uint8, heavy-tailed, strong-scaling and BoxCodec lines")
    ap.add_argument("--parity-images", type=int, default=16,
                    help="images of the timed batch checked against the CPU port in the cpu_baseline leg")
    ap.add_argument("--clock-period", type=float, default=0.02, help="seconds between NVML clock samples in the timed region")
    return ap.parse_args()


def workload(batch):
    return {"workload": "512x512 float64 synthetic (warped Voronoi labels, 1024 regions/image, distinct per image), "
                        "easypath euclidean, 16 levels bior4.4, keep 2048 coefs",
            "images_per_gpu_per_step": batch, "height": H, "width": W, "levels": LEVELS, "wavelet": WAVELET,
            "ncoefs": NCOEFS, "path_type": "easypath", "euclidean_distance": True,
            "parallelism": "image batch sharded by image, no collective",
            "l2": "inputs larger than L2 (%.0f MB of images+labels per step vs 126 MB L2)" % (batch * H * W * 12 / 1e6)}


# ------------------------------------------------------------------------------ clocks -----
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index, period=0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.sm_max = [], set(), None
        self._stop_ev = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:  # noqa: BLE001
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake_slowdown",
        }
        while not self._stop_ev.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:  # noqa: BLE001
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            self._stop_ev.wait(self.period)

    def finish(self):
        self._stop_ev.set()
        if self.is_alive():
            self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons),
                "samples": len(s)}


def physical_gpu_index(local):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local])
        except Exception:  # noqa: BLE001
            return local
    return local


# ------------------------------------------------------------------------------ CPU arm ----
def cpu_port_throughput(imgs, labs, threads, keep=0):
    """images/s of the C oracle (encode + threshold + decode) over the given host arrays with `threads`
    host threads (ctypes releases the GIL).  TEST-INFRASTRUCTURE code, used here only as the baseline and --
    with `keep` > 0, which also returns (paths, thresholded coefficients, decoded image) of the first `keep`
    images -- as the checker of the GPU results of the timed configuration."""
    from concurrent.futures import ThreadPoolExecutor

    from oracle import c_oracle, pywt_port

    fb = pywt_port.filter_bank(WAVELET)
    c_oracle.lib()

    kept = {}

    def one(i):
        enc = c_oracle.encode(imgs[i], labs[i], LEVELS, fb, c_oracle.MODE_EUCLID)
        th = c_oracle.threshold(enc["coefs"], NCOEFS)
        dec = c_oracle.decode(enc, th, fb)
        if i < keep:
            kept[i] = (enc["path_pix"], th, dec)
        return dec

    one(0)  # warm (page in the library, allocate)
    t0 = time.perf_counter()
    if threads == 1:
        for i in range(len(imgs)):
            one(i)
    else:
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(one, range(len(imgs))))
    v = len(imgs) / (time.perf_counter() - t0)
    return (v, kept) if keep else v


def check_against_port(codec, imgs_np, out_np, kept, npix):
    """The GPU results of the timed configuration (state left by the last rbepwt_transcode) against the CPU port,
    image by image: paths of every level and kept-coefficient index sets bit-exact, decoded pixels within
    1e-9 * 255, PSNR equal to 6 decimals.  Returns the number of images checked; raises on the first difference."""
    import numpy as np

    from oracle import c_oracle

    for b, (path_pix, th, dec) in sorted(kept.items()):
        lo = 0
        for lev in range(1, LEVELS + 1):
            n = npix >> (lev - 1)
            if not np.array_equal(codec.paths(b, lev), path_pix[lo:lo + n]):
                raise AssertionError("bench parity: paths of image %d differ from the CPU port at level %d" % (b, lev))
            lo += n
        got = codec.coefs(b)
        if not np.array_equal(np.flatnonzero(got), np.flatnonzero(th)):
            raise AssertionError("bench parity: kept coefficient indices of image %d differ from the CPU port" % b)
        if np.max(np.abs(got - th)) > 1e-9 * np.max(np.abs(th)):
            raise AssertionError("bench parity: kept coefficient values of image %d differ from the CPU port" % b)
        if np.max(np.abs(out_np[b] - dec)) > 1e-9 * 255:
            raise AssertionError("bench parity: decoded pixels of image %d differ from the CPU port" % b)
        p_gpu = float(codec.psnr(imgs_np[b:b + 1], out_np[b:b + 1])[0])
        if abs(p_gpu - c_oracle.psnr(imgs_np[b], dec)) >= 5e-7:
            raise AssertionError("bench parity: PSNR of image %d differs from the CPU port" % b)
    return len(kept)


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:  # noqa: BLE001
        return os.cpu_count() or 1


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import torch

    from rbepwt_b200 import synth

    cores = host_cores()
    # ~0.2 s per image per core: size the sample so that every step is a few seconds of all-core work
    nimg = args.cpu_sample or max(cores, min(4 * cores, 64))
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    imgs, labs = synth.torch_batch(nimg, H, W, NSEEDS, SEED0, device=dev)
    imgs, labs = imgs.cpu().numpy(), labs.cpu().numpy()
    for _ in range(args.warmup):
        cpu_port_throughput(imgs[:cores], labs[:cores], cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_port_throughput(imgs, labs, cores)
    dt = time.perf_counter() - t0
    v = nimg * args.steps / dt
    sample = ("%d distinct images of the workload per step (a bounded sample of the %d-image batch: the per-image "
              "work is identical), %d host threads, C port of the reference algorithm" % (nimg, args.batch, cores))
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload(args.batch),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "the reference itself is pure Python (+PyWavelets) and cannot run on the GPU box; measured in the "
                    "build container it needs ~1 h per 512x512 image (BASELINE.md section 2). This arm times the C "
                    "port of its algorithm (oracle/), which is orders of magnitude faster than the reference."}
    _emit(line)
    return 0


# ------------------------------------------------------------------------------ our arm ----
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import rbepwt_b200 as rb
    from rbepwt_b200 import synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; rbepwt_b200 has no CPU path")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    B, K, Wm = args.batch, args.steps, args.warmup
    N = H * W

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    imgs, labs = synth.torch_batch(B, H, W, NSEEDS, SEED0 + rank * B, device="cuda")
    torch.cuda.synchronize()
    # dedicated (non-default) streams: the library launches on them and the timing events are recorded on them.
    # F steps in flight = F contexts, each with its own stream and output buffer; step i runs on context i mod F.
    F_dev = max(1, args.inflight)
    streams = [torch.cuda.Stream() for _ in range(F_dev)]
    assert all(st.cuda_stream != 0 for st in streams)
    codecs = [rb.BatchCodec(device=local, stream=st.cuda_stream) for st in streams]
    outs = [torch.empty_like(imgs) for _ in range(F_dev)]
    for cd in codecs:
        cd.set_option(sub_batch=args.sub_batch, path_group=args.path_group)
    stream, codec, out = streams[0], codecs[0], outs[0]

    def step(i=0):  # encode -> threshold -> decode, one pipelined C-ABI call (rbepwt_transcode), device pointers
        codecs[i % F_dev].transcode(imgs, labs, LEVELS, WAVELET, NCOEFS, "easypath", True, outs[i % F_dev])

    def timed_steps(nsteps, nflight):
        """device time of nsteps steps, up to nflight in flight: CUDA events on the contexts' own streams"""
        e0 = torch.cuda.Event(enable_timing=True)
        ends = [torch.cuda.Event(enable_timing=True) for _ in range(nflight)]
        barrier()
        torch.cuda.synchronize()
        e0.record(streams[0])
        for st in streams[1:nflight]:
            st.wait_event(e0)
        for i in range(nsteps):
            step(i % nflight)
        for st, ev in zip(streams[:nflight], ends):
            ev.record(st)
        torch.cuda.synchronize()
        barrier()
        return max(e0.elapsed_time(ev) for ev in ends)

    for i in range(max(Wm, 1) * F_dev):
        step(i)
    torch.cuda.synchronize()
    # correctness guard of the measured configuration: untouched survivors + PSNR is finite
    for cd, o in zip(codecs, outs):
        nz = cd.nonzero_coefs()
        assert int(nz.min()) == NCOEFS and int(nz.max()) == NCOEFS, nz
        assert torch.equal(o, outs[0])
    psnr0 = float(codec.psnr(imgs[:1], out[:1])[0])

    sampler = ClockSampler(physical_gpu_index(local), period=args.clock_period)
    l0 = sum(cd.launch_count() for cd in codecs)
    sampler.start()
    ms = max_over_ranks(timed_steps(K, F_dev))
    clocks = sampler.finish()
    launches = sum(cd.launch_count() for cd in codecs) - l0
    value = world * B * K / (ms * 1e-3)
    one_ms = max_over_ranks(timed_steps(K, 1)) if F_dev > 1 else ms
    for cd in codecs[1:]:  # the other measurements use the first context only
        cd.close()
    del outs[1:], codecs[1:]
    F_used = F_dev
    F_dev = 1
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    # ---- per-kernel durations: the same K steps again with one sub-batch in flight (RBEPWT_OPT_STREAMS = 1),
    # so that a kernel's CUDA-event time is its own and not shared with the kernels it overlaps with above
    codec.set_option(streams=1)
    step()
    codec.enable_timing(True)
    codec.timings()  # drop events of anything before
    barrier()
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(K):
        step()
    e1.record(stream)
    torch.cuda.synchronize()
    serial_ms = e0.elapsed_time(e1)
    stage_ms = codec.timings()
    stage_n = codec.stage_launches()
    codec.enable_timing(False)
    codec.set_option(streams=2)

    # ---- e2e: the public host-buffer API, pinned host memory, copies inside the timed region.
    # `--e2e-inflight` (default 2) steps are in flight at once, each on its own context / host thread / pinned
    # buffers (a serving loop's double buffering): the output copy of one step overlaps the input copy of the
    # next.  Every step still copies its own inputs in and its own results out; 1 = strictly one step at a time.
    e2e = None
    extras = {"one_step_in_flight": {"value": world * B * K / (one_ms * 1e-3), "unit": UNIT, "ms_per_step": one_ms / K,
                                     "note": "the same K steps on ONE context, each step ordered after the previous one"}}
    if not args.no_e2e:
        import threading as _th

        F = max(1, args.e2e_inflight)
        h_img = torch.empty((B, H, W), dtype=torch.float64, pin_memory=True)
        h_lab = torch.empty((B, H, W), dtype=torch.int32, pin_memory=True)
        h_img.copy_(imgs)
        h_lab.copy_(labs)
        torch.cuda.synchronize()
        n_img, n_lab = h_img.numpy(), h_lab.numpy()
        lanes = []
        for _ in range(F):
            hc = rb.BatchCodec(device=local)
            hc.set_option(sub_batch=args.sub_batch, path_group=args.path_group)
            ho = torch.empty((B, H, W), dtype=torch.float64, pin_memory=True)
            lanes.append((hc, ho, ho.numpy()))

        def run_steps(nsteps, call=None):
            """nsteps steps, up to F in flight: lane j takes steps j, j+F, ...; a call returns after its last D2H."""
            def worker(j):
                torch.cuda.set_device(local)
                hc, _, no = lanes[j]
                for _ in range(j, nsteps, F):
                    if call is None:
                        hc.transcode(n_img, n_lab, LEVELS, WAVELET, NCOEFS, "easypath", True, no)
                    else:
                        call(j, hc)
            ths = [_th.Thread(target=worker, args=(j,)) for j in range(min(F, nsteps))]
            for t in ths:
                t.start()
            for t in ths:
                t.join()

        def timed_e2e(call):
            run_steps(max(Wm, 1) * F, call)
            barrier()
            torch.cuda.synchronize()
            t0_ = time.perf_counter()
            run_steps(K, call)
            for hc_, _, _ in lanes:
                hc_.sync()
            dt_ = max_over_ranks(time.perf_counter() - t0_)
            barrier()
            return dt_

        run_steps(max(Wm, 1) * F)
        barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        run_steps(K)
        for hc, _, _ in lanes:
            hc.sync()
        dt = max_over_ranks(time.perf_counter() - t0)
        barrier()
        want0 = out[0].cpu().numpy()
        for _, _, no in lanes[:min(F, K)]:
            assert np.array_equal(no[0], want0), "host-buffer path and device path disagree"
        e2e = {"value": world * B * K / dt, "unit": UNIT, "h2d_bytes_per_step": B * N * 12, "d2h_bytes_per_step": B * N * 8,
               "ms_per_step": 1e3 * dt / K, "steps_in_flight": F,
               "api": "rbepwt_b200.BatchCodec.transcode (rbepwt_transcode: encode+threshold+decode, host-pointer path) "
                      "with numpy views of pinned host memory; %d context(s), one host thread each" % F}
        # ---- the same end-to-end path fed the way the reference's Image.read delivers grayscale files (rbepwt.py:200-206):
        # uint8 pixels, and the ~10^3 label values of an image as uint16 -- 3 bytes per pixel over PCIe instead of 12;
        # (a) decoded images back as uint8 (1 byte per pixel instead of 8), (b) what a codec keeps: PSNR + the k
        # surviving (index, value) pairs per image, nothing else.  Same label maps, pixels rounded to uint8.
        if not args.no_extras:
            h_img8 = torch.empty((B, H, W), dtype=torch.uint8, pin_memory=True)
            h_lab16 = torch.empty((B, H, W), dtype=torch.uint16, pin_memory=True)
            h_img8.copy_(imgs.round().clamp_(0, 255).to(torch.uint8))
            h_lab16.copy_(labs.to(torch.uint16))
            torch.cuda.synchronize()
            n_img8, n_lab16 = h_img8.numpy(), h_lab16.numpy()
            out8 = [torch.empty((B, H, W), dtype=torch.uint8, pin_memory=True) for _ in range(F)]
            n_out8 = [o.numpy() for o in out8]
            res_b = [None] * F

            def call_a(j, hc):
                hc.transcode_ex(n_img8, n_lab16, LEVELS, WAVELET, NCOEFS, out=n_out8[j])

            def call_b(j, hc):
                res_b[j] = hc.transcode_ex(n_img8, n_lab16, LEVELS, WAVELET, NCOEFS, want_image=False, want_psnr=True, want_kept=True)

            dt_a = timed_e2e(call_a)
            dt_b = timed_e2e(call_b)
            # consistency: the kept pairs rebuild (through the decoder side, labels + coefficients) the uint8 output
            flat0 = np.zeros((1, N))
            flat0[0, res_b[0]["kept_idx"][0]] = res_b[0]["kept_val"][0]
            dec0 = lanes[0][0].full_decode(flat0, n_lab16[:1].astype(np.int32), LEVELS, WAVELET)
            assert np.array_equal(np.rint(dec0[0]).astype(np.uint8), n_out8[0][0]), "compact output does not rebuild the image"
            extras["e2e_uint8"] = {
                "value": world * B * K / dt_a, "unit": UNIT, "ms_per_step": 1e3 * dt_a / K,
                "h2d_bytes_per_step": B * N * 3, "d2h_bytes_per_step": B * N,
                "io": "uint8 pixels + uint16 labels in, uint8 decoded images out (rbepwt_transcode_ex)"}
            extras["e2e_uint8_compact"] = {
                "value": world * B * K / dt_b, "unit": UNIT, "ms_per_step": 1e3 * dt_b / K,
                "h2d_bytes_per_step": B * N * 3, "d2h_bytes_per_step": B * (8 + NCOEFS * 12),
                "io": "uint8 pixels + uint16 labels in; per image PSNR + the %d kept (index, value) pairs out, nothing "
                      "decoded to the host (rbepwt_transcode_ex)" % NCOEFS,
                "psnr_image0": float(res_b[0]["psnr"][0])}
        for hc, _, _ in lanes:
            hc.close()

    # ---- further lines of the same metric (reported under `extra`, never mixed into `value`)
    if not args.no_extras:
        def device_timed(fn, nsteps):
            for _ in range(max(Wm, 1)):
                fn()
            barrier()
            torch.cuda.synchronize()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record(stream)
            for _ in range(nsteps):
                fn()
            a1.record(stream)
            torch.cuda.synchronize()
            barrier()
            return max_over_ranks(a0.elapsed_time(a1))

        # (1) heavy-tailed label maps -- what felzenszwalb(scale=200) gives on natural images: median region ~90 pixels, a
        # few regions of 5-10 thousand -- 256 images per GPU (8 distinct maps x 32), resident in HBM like `value`;
        # two images checked against the CPU port.
        HB = 256
        ht_labs = np.stack([synth.heavytail_labels(H, 600, 100 + i) for i in range(8)])
        ht_imgs = np.stack([synth.piecewise_smooth_image(l, seed=3 + i) for i, l in enumerate(ht_labs)])
        t_hl = torch.from_numpy(np.concatenate([ht_labs] * (HB // 8))).cuda()
        t_hi = torch.from_numpy(np.concatenate([ht_imgs] * (HB // 8))).cuda()
        t_ho = torch.empty_like(t_hi)
        ms_ht = device_timed(lambda: codec.transcode(t_hi, t_hl, LEVELS, WAVELET, NCOEFS, "easypath", True, t_ho), K)
        ht = {"value": world * HB * K / (ms_ht * 1e-3), "unit": UNIT, "ms_per_step": ms_ht / K, "images_per_gpu_per_step": HB,
              "workload": "512x512, 600 Voronoi seeds of which 85 % crowd into three blobs (median region ~90 pixels, largest "
                          "5-10 thousand), otherwise as `config`"}
        if rank == 0 and not args.no_cpu_baseline:
            _, kept_ht = cpu_port_throughput(ht_imgs[:2], ht_labs[:2], 1, keep=2)
            ht["parity_checked_images"] = check_against_port(codec, ht_imgs[:2], t_ho[:2].cpu().numpy(), kept_ht, N)
        extras["heavytail"] = ht
        del t_hl, t_hi, t_ho

        # (2) strong scaling (BASELINE.json configs[4]: 4096 x 512^2 sharded by image over the GPUs): total/N images
        # per rank per step; the rank's 512 distinct images repeated to that count (the per-image work is what
        # counts; at N = 8 every image is distinct)
        if args.total and args.total % world == 0 and (args.total // world) % B == 0:
            per = args.total // world
            reps = per // B
            s_img = imgs.repeat(reps, 1, 1) if reps > 1 else imgs
            s_lab = labs.repeat(reps, 1, 1) if reps > 1 else labs
            s_out = torch.empty_like(s_img)
            ks = max(2, K // 4)
            ms_s = device_timed(lambda: codec.transcode(s_img, s_lab, LEVELS, WAVELET, NCOEFS, "easypath", True, s_out), ks)
            assert torch.equal(s_out[:B], out), "strong-scaling batch and the bench batch disagree"
            extras["strong_scaling"] = {"total_images": args.total, "images_per_gpu_per_step": per, "steps": ks,
                                        "value": args.total * ks / (ms_s * 1e-3), "unit": UNIT, "ms_per_step": ms_s / ks,
                                        "scaling": "strong"}
            del s_img, s_lab, s_out
        torch.cuda.empty_cache()

        # (3) the in-process driver: ONE process, one context + host thread per GPU (rbepwt_b200.BoxCodec), host buffers
        # in, decoded images out -- the same end-to-end path as `e2e`, without torchrun.  Rank 0 drives all N GPUs
        # while the other ranks wait.
        if not args.no_e2e:
            barrier()
            if rank == 0:
                ndev = min(world, torch.cuda.device_count())
                bh_img = torch.empty((ndev * B, H, W), dtype=torch.float64, pin_memory=True)
                bh_lab = torch.empty((ndev * B, H, W), dtype=torch.int32, pin_memory=True)
                bh_out = torch.empty((ndev * B, H, W), dtype=torch.float64, pin_memory=True)
                for d in range(ndev):
                    bh_img[d * B:(d + 1) * B].copy_(imgs)
                    bh_lab[d * B:(d + 1) * B].copy_(labs)
                torch.cuda.synchronize()
                box = rb.BoxCodec(range(ndev))
                bi, bl, bo = bh_img.numpy(), bh_lab.numpy(), bh_out.numpy()
                for _ in range(max(Wm, 1)):
                    box.transcode(bi, bl, LEVELS, WAVELET, NCOEFS, out=bo)
                tb = time.perf_counter()
                for _ in range(K):
                    box.transcode(bi, bl, LEVELS, WAVELET, NCOEFS, out=bo)
                dtb = time.perf_counter() - tb
                assert np.array_equal(bo[(ndev - 1) * B], out[0].cpu().numpy()), "BoxCodec and the device path disagree"
                extras["boxcodec_e2e"] = {"value": ndev * B * K / dtb, "unit": UNIT, "ms_per_step": 1e3 * dtb / K, "gpus": ndev,
                                          "api": "rbepwt_b200.BoxCodec(range(%d)).transcode: one process, one context and host "
                                                 "thread per GPU, one step in flight per GPU" % ndev}
                box.close()
                del bh_img, bh_lab, bh_out
            barrier()

    # ---- roofline of the dominant kernel (stage with the largest share of the device time)
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:  # noqa: BLE001
        pass
    peak, peak_src = (float(peaks["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs") if "hbm_gbs" in peaks else (6650.0, "fallback")
    S = 2 * N - (N >> (LEVELS - 1))  # sum of the level lengths
    # algorithmic bytes per image of each kernel family
    alg = {
        "regions": ("k0_count+k0_regions_fast+queue", 3 * 4 * N),         # three passes over the labels
        "paths": ("k1_walk", 4 * N + 4 * S),                               # labels in; pixel ids (all levels) out
        "paths_big": ("k1_paths_big", 4 * N + 4 * S),
        "perm": ("k2_perm", 4 * S + 4 * (S - N)),                          # pixel ids in; positions (levels >= 2) out
        "dwt": ("k3_dwt_level", 4 * S + 8 * S + 8 * S),                    # paths + gathered values in, cA/cD out
        "select": ("k4_select", 8 * N),                                      # one read of the coefficients (the threshold is applied by K5 on load)
        "idwt": ("k5_idwt_level", 4 * S + 8 * S + 8 * S),
    }
    kern_stages = [s for s in alg if stage_ms.get(s, 0.0) > 0.0]
    total_kernel_ms = sum(stage_ms[s] for s in kern_stages)
    dom = max(kern_stages, key=lambda s: stage_ms[s])
    # a "launch" of the paths stage is one path group: kq_slots + k1_bitmaps, then the two k1_walk instantiations side by
    # side (the windowed one takes the few large bitmaps) -- four kernel launches timed as one unit
    per_unit = {"paths": 4}
    dom_launches = max(stage_n.get(dom, 0) // per_unit.get(dom, 1), 1)
    dom_ms_per_launch = stage_ms[dom] / dom_launches
    bytes_per_launch = alg[dom][1] * B * K / dom_launches
    achieved = bytes_per_launch / (dom_ms_per_launch * 1e-3) / 1e9
    traffic, traffic_src = None, None
    try:  # DRAM bytes of the dominant kernel from the committed ncu capture, scaled to this run's launch size
        with open(os.path.join(ROOT, "profiles", "r2_traffic.json")) as f:
            tj = json.load(f)
        es = [tj[k_] for k_ in ((alg[dom][0], "k1_bitmaps") if dom == "paths" else (alg[dom][0],)) if k_ in tj]
        if es:
            traffic = sum((e["dram_bytes_read"] + e["dram_bytes_write"]) / e["images_per_launch"] for e in es) * (B * K / dom_launches)
            traffic_src = tj["source"]
    except Exception:  # noqa: BLE001
        pass
    roofline = {"bound": "hbm", "kernel": alg[dom][0], "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "avg_launch_ms": dom_ms_per_launch, "launches": dom_launches,
                "algorithmic_bytes_per_launch": bytes_per_launch,
                "share_of_step": stage_ms[dom] / total_kernel_ms,
                "note": "k1 path construction (k1_bitmaps + k1_walk, one unit per path group) is dependent chains "
                        "(issue bound, see profiles/), not an HBM-bound kernel; see `kernels` for the HBM-bound ones"}
    kernels = {}
    for s in kern_stages:
        if stage_ms[s] < 0.01 * total_kernel_ms:
            continue  # e.g. k1_paths_big on this workload: a near-empty launch (no oversized region in the batch)
        n = max(stage_n.get(s, 0), 1)
        gbs = alg[s][1] * B * K / (stage_ms[s] * 1e-3) / 1e9
        kernels[alg[s][0]] = {"ms_per_step": stage_ms[s] / K, "launches_per_step": n / K, "share": stage_ms[s] / total_kernel_ms,
                              "algorithmic_GBps": gbs, "frac_of_hbm_peak": gbs / peak}
    path_gbs = 68 * N * B * K / (ms * 1e-3) / 1e9  # per GPU
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": dict(workload(B), steps_in_flight=F_used), "clocks": clocks, "e2e": e2e,
            "gpu_launches": int(launches),
            "roofline": roofline, "kernels": kernels,
            "kernels_measured": "same K steps repeated with RBEPWT_OPT_STREAMS=1 (no overlap between kernels), "
                                "%.3f ms/step serial vs %.3f ms/step pipelined" % (serial_ms / K, ms / K),
            "whole_path": {"algorithmic_bytes_per_image": 68 * N, "achieved_GBps_per_gpu": path_gbs,
                           "frac_of_hbm_peak": path_gbs / peak},
            "psnr_image0": psnr0, "extra": extras}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        ns = args.cpu_sample or 64  # ~0.2 s per image on one core -> ~13 s
        si, sl = imgs[:ns].cpu().numpy(), labs[:ns].cpu().numpy()
        nchk = min(ns, args.parity_images)
        step()  # the extra lines above left other batches in the codec: the timed configuration's state again
        torch.cuda.synchronize()
        v, kept = cpu_port_throughput(si, sl, 1, keep=max(nchk, 1))
        # the CPU port's outputs are the checker of the timed configuration: `codec` still holds the state of the last
        # timed-style step (paths, thresholded coefficients) and `out` its decoded images
        line["parity_checked_images"] = check_against_port(codec, si, out[:ns].cpu().numpy(), kept, N) if nchk else 0
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
                                "sample": "first %d images of the step's batch, single-threaded C port of the reference "
                                          "algorithm (oracle/rbepwt_oracle.c); the Python reference itself needs ~1 h per "
                                          "512x512 image (BASELINE.md)" % ns}
    if rank == 0:
        _emit(line)
    codec.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def _emit(line):
    """The ONE JSON line goes to the real stdout; everything else this process (or NCCL) prints goes to stderr."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def main():
    global _REAL_STDOUT
    args = parse()
    world_env = os.environ.get("WORLD_SIZE")
    if args.gpus > 1 and world_env is None:
        # convenience: relaunch under torchrun, one rank per GPU
        import subprocess
        port = os.environ.get("MASTER_PORT", "29517")
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", port, os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)  # library banners (e.g. "NCCL version ...") must not precede the JSON line on stdout
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
