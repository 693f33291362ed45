"""CPU: the per-lane logic of the CUDA path walker (struct Walker in rbepwt_b200/csrc/walk.cuh -- the same
__host__ __device__ source the kernel k1_walk compiles) executed on the host, region after region, against the C
oracle: the paths of every level must be bit-exact.  This pins the step rule of the GPU kernel (5x5 window classes,
unit-step table, list mode, level transitions, survivor planes) without a GPU; the warp loop, the warp's search
beyond the 5x5 window (done here by a literal restatement of the reference's probes), the arena and the chunk table
are covered by the -m gpu tests."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
HW_DIR = os.path.join(ROOT, "tests", "host_walk")
HW_LIB = os.path.join(HW_DIR, "_build", "libhostwalk.so")
CSRC = os.path.join(ROOT, "rbepwt_b200", "csrc")

_lib = None


def hostwalk():
    global _lib
    if _lib is None:
        srcs = [os.path.join(HW_DIR, "host_walk.cu")] + [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
        if not os.path.isfile(HW_LIB) or any(os.path.getmtime(s) > os.path.getmtime(HW_LIB) for s in srcs):
            os.makedirs(os.path.dirname(HW_LIB), exist_ok=True)
            subprocess.check_call(["nvcc", "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-fmad=false",
                                   "-Xcompiler", "-fPIC", "-shared", "-o", HW_LIB, os.path.join(HW_DIR, "host_walk.cu")])
        L = ctypes.CDLL(HW_LIB)
        i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
        L.hw_walk_image.restype = ctypes.c_int
        L.hw_walk_image.argtypes = [i32p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                    i32p, i32p, np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")]
        _lib = L
    return _lib


def host_paths(lab, levels, euclid=True, widewin=False, use_lut=True):
    lab = np.ascontiguousarray(lab, dtype=np.int32)
    H, W = lab.shape
    Q = np.full(2 * H * W, -1, dtype=np.int32)
    Pm = np.full(2 * H * W, -1, dtype=np.int32)
    kinds = np.zeros(8, dtype=np.int64)
    R = hostwalk().hw_walk_image(lab, H, W, levels, 0 if euclid else 1, int(widewin), int(use_lut), Q, Pm, kinds)
    assert R > 0, "walker reported corrupt region state"
    return Q, Pm, kinds


def oracle_paths(lab, levels, euclid=True):
    """(paths as pixel ids, positions in the incoming order = region offset + generating permutation), all levels."""
    from oracle import c_oracle, pywt_port

    H, W = lab.shape
    enc = c_oracle.encode(np.zeros((H, W)), lab, levels, pywt_port.filter_bank("haar"),
                          c_oracle.MODE_EUCLID if euclid else c_oracle.MODE_CHEB)
    pos = enc["perm"].copy()
    lo = 0
    for lev in range(1, levels + 1):
        nl = (H * W) >> (lev - 1)
        roff = enc["roff"][lev - 1]
        pos[lo:lo + nl] += np.repeat(roff[:-1], np.diff(roff))
        lo += nl
    return enc["path_pix"], pos


def check(lab, levels, euclid=True, widewin=False, use_lut=True):
    Q, Pm, kinds = host_paths(lab, levels, euclid, widewin, use_lut)
    want, want_pos = oracle_paths(lab, levels, euclid)
    n = want.size
    wrote = Pm[:n] >= 0  # the walker writes positions for its list-mode levels only
    assert np.array_equal(Pm[:n][wrote], want_pos[wrote]), "list-mode positions differ from the oracle's permutation"
    if not np.array_equal(Q[:n], want):
        bad = int(np.flatnonzero(Q[:n] != want)[0])
        N, lev, lo = lab.size, 1, 0
        while bad >= lo + (N >> (lev - 1)):
            lo += N >> (lev - 1)
            lev += 1
        raise AssertionError("path differs at level %d, position %d (euclid=%s widewin=%s lut=%s)"
                             % (lev, bad - lo, euclid, widewin, use_lut))
    return kinds


def _maps():
    from rbepwt_b200 import synth

    rng = np.random.default_rng(3)
    yield "voronoi64", synth.voronoi_labels(64, 64, 40, seed=4), 12
    yield "voronoi128", synth.voronoi_labels(128, 128, 96, seed=9), 14
    yield "one_label32", np.zeros((32, 32), np.int32), 10
    yield "noise32_5", rng.integers(0, 5, size=(32, 32)).astype(np.int32), 10
    yield "noise64_20", rng.integers(0, 20, size=(64, 64)).astype(np.int32), 12
    yield "stripes64", (np.arange(64)[:, None] // 3 + 0 * np.arange(64)[None, :]).astype(np.int32), 12
    yield "wide_boxes", ((np.arange(64)[:, None] // 2) * 4 + np.arange(128)[None, :] // 40).astype(np.int32), 13
    yield "rect32x64", synth.voronoi_labels(32, 64, 30, seed=2), 11
    yield "every_pixel", rng.permutation(256).reshape(16, 16).astype(np.int32), 8
    yield "tiny2x2", np.array([[0, 1], [1, 0]], np.int32), 2
    yield "row1x8", np.array([[0, 0, 1, 0, 1, 1, 0, 0]], np.int32), 3
    ii, jj = np.meshgrid(np.arange(128), np.arange(128), indexing="ij")
    yield "interleaved128", (((ii // 5) + (jj // 3)) % 2).astype(np.int32), 14
    yield "antidiag128", ((ii + jj) % 97).astype(np.int32), 14


@pytest.mark.parametrize("name,lab,levels", list(_maps()), ids=[m[0] for m in _maps()])
@pytest.mark.parametrize("euclid", [True, False], ids=["euclid", "cheb"])
def test_host_walker_matches_oracle(name, lab, levels, euclid):
    check(lab, levels, euclid, widewin=False, use_lut=True)
    check(lab, levels, euclid, widewin=True, use_lut=True)
    check(lab, levels, euclid, widewin=False, use_lut=False)


def test_t2_hash_is_a_bijection():
    assert hostwalk().hw_t2_hash_is_perfect() == 1


def test_host_walker_bench_and_heavytail_maps():
    """512x512: one map of the benchmark generator and one heavy-tailed map, all 16 levels."""
    from rbepwt_b200 import synth

    k = check(synth.voronoi_labels(512, 512, 1024, seed=1000), 16)
    assert k[1] > 0.85 * 2 * 512 * 512  # the 5x5 step resolves the bulk of the benchmark's steps
    check(synth.heavytail_labels(512, 600, 100), 16, widewin=True)
    check(synth.heavytail_labels(512, 600, 101), 16, euclid=False)


def test_host_walker_fuzz():
    from rbepwt_b200 import synth

    rng = np.random.default_rng(77)
    for case in range(60):
        lh, lw = int(rng.integers(0, 7)), int(rng.integers(0, 8))
        if lh + lw < 2:
            lw = 2 - lh
        H, W = 1 << lh, 1 << lw
        levels = int(rng.integers(1, lh + lw + 1))
        kind = int(rng.integers(0, 4))
        if kind == 0 and min(H, W) >= 8:
            lab = synth.voronoi_labels(H, W, int(rng.integers(1, max(2, H * W // 24))), seed=case)
        elif kind == 1:
            lab = rng.integers(0, int(rng.integers(1, 9)), size=(H, W)).astype(np.int32)
        elif kind == 2:
            lab = ((np.arange(H)[:, None] // int(rng.integers(1, 4))) * 3 + (np.arange(W)[None, :] // int(rng.integers(1, 5)))).astype(np.int32)
        else:
            lab = rng.integers(-5, H * W, size=(H, W)).astype(np.int32)
        try:
            check(lab, levels, bool(rng.integers(0, 2)), bool(rng.integers(0, 2)), bool(rng.integers(0, 2)))
        except AssertionError as e:
            raise AssertionError("fuzz case %d (%dx%d L=%d kind=%d): %s" % (case, H, W, levels, kind, e))
