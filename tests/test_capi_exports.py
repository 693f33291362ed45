"""CPU: the C-ABI library builds, loads and exports every symbol include/rbepwt_b200.h declares; without
a GPU it refuses to create a context (there is no CPU fallback).  No compute call is made here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "rbepwt_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rbepwt_[a-z_0-9]+)\s*\(", src)))


def test_header_cites_the_reference_for_every_entry_point():
    src = open(os.path.join(ROOT, "include", "rbepwt_b200.h")).read()
    assert len(re.findall(r"rbepwt\.py[: ]", src)) >= 10
    assert 'extern "C"' in src
    code = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    assert "torch" not in code and "Tensor" not in code  # plain pointers and sizes only


def test_library_exports_every_declared_symbol():
    from rbepwt_b200 import _capi

    L = _capi.lib()
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), "librbepwt_b200.so does not export %s" % n
    assert sorted(_capi.EXPORTS) == names


def test_no_cpu_fallback_without_a_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from rbepwt_b200 import _capi
    import rbepwt_b200 as rb

    ctx = ctypes.c_void_p()
    rc = _capi.lib().rbepwt_create(0, None, ctypes.byref(ctx))
    assert rc == _capi.E_NO_GPU and not ctx.value
    assert b"no CPU fallback" in _capi.lib().rbepwt_last_error()
    with pytest.raises(_capi.RbepwtError):
        rb.BatchCodec()


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "rbepwt_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "liboracle" not in text, f


def test_wavelet_tables_match_the_oracle_port():
    import numpy as np

    from oracle import pywt_port
    import rbepwt_b200 as rb

    for name in ("haar", "db2", "db3", "db4", "bior4.4"):
        for a, b in zip(rb.filter_bank(name), pywt_port.filter_bank(name)):
            # the product's tables are regenerated in 60-digit arithmetic (tools/gen_wavelets.py); the
            # oracle's are literals recalled from PyWavelets, good to ~4e-12 -- both far inside 1e-9
            np.testing.assert_allclose(a, b, rtol=0, atol=1e-11)
    with pytest.raises(ValueError):
        rb.filter_bank("nosuchwavelet")


def test_path_mode_mapping_and_guards():
    import rbepwt_b200 as rb

    assert rb.path_mode("easypath", True) == 0 and rb.path_mode("easypath", False) == 1
    assert rb.path_mode("epwt-easypath", True) == 2 and rb.path_mode("epwt-easypath", False) == 2
    assert rb.path_mode("gradpath", True) == 3 and rb.path_mode("gradpath", False) == 4
    with pytest.raises(ValueError):
        rb.path_mode("zigzag")
    assert rb.ispowerof2(256 * 512) and not rb.ispowerof2(48)


def test_bench_traffic_file_matches_bench():
    """bench.py scales the dominant kernel's DRAM traffic from a committed ncu summary: the file it names must
    exist and carry the entries it reads (a rename would silently null `roofline.traffic` on the GPU box)."""
    import json
    import re

    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
    src = open(os.path.join(root, "bench.py")).read()
    names = re.findall(r'"(r\w+_traffic\.json)"', src)
    assert len(names) == 1, names
    with open(os.path.join(root, "profiles", names[0])) as f:
        tj = json.load(f)
    for k in ("k1_walk", "k1_bitmaps"):
        e = tj[k]
        assert e["dram_bytes_read"] > 0 and e["dram_bytes_write"] > 0 and e["images_per_launch"] > 0
    assert isinstance(tj["source"], str)
