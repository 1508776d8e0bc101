"""The C-ABI shared library loads and exports exactly what include/visma_b200.h declares; without a GPU the
compute entries fail loudly (no CPU fallback).  CPU only — no compute calls."""
import ctypes as C
import os
import re

import numpy as np

from conftest import ROOT


def header_symbols():
    src = open(os.path.join(ROOT, "include", "visma_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vb200_[a-z0-9_]+)\s*\(", src)))


def test_exports_match_header(vb):
    declared = header_symbols()
    assert declared == sorted(vb.lib.SYMBOLS)
    L = vb.lib.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert L.vb200_version() == 100


def test_no_oracle_in_product():
    """The product path never imports or links the oracle."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "visma_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "pyoracle" not in txt and "pyref" not in txt and "libvisma_oracle" not in txt, f
                assert not re.search(r'#include\s+"[^"]*oracle', txt), f


def test_status_strings(vb):
    L = vb.lib.lib()
    for s in range(-6, 1):
        assert len(L.vb200_strerror(s)) > 0
    assert b"no CPU fallback" in L.vb200_strerror(vb.lib.ERR_NO_DEVICE)


def test_fails_loudly_without_device(vb):
    L = vb.lib.lib()
    if L.vb200_device_count() > 0:
        import pytest
        pytest.skip("a GPU is present")
    import pytest
    with pytest.raises(vb.pkg.VismaB200Error) as e:
        vb.reg.Scene(np.zeros((10, 3)), 0.1)
    assert e.value.status == vb.lib.ERR_NO_DEVICE
    out = np.zeros((4, 4))
    rc = L.vb200_estimate(None, 0, None, None, 0, None, 0, 0, None, 0, out.ctypes.data_as(C.POINTER(C.c_double)))
    assert rc == 0 and np.array_equal(out, np.eye(4))  # corres.empty() -> Identity needs no device
    with pytest.raises(vb.pkg.VismaB200Error):
        vb.reg.VoxelDownSample(np.zeros((10, 3)), 0.1)


def test_argument_validation(vb):
    L = vb.lib.lib()
    h = C.c_void_p()
    dp = C.POINTER(C.c_double)
    x = np.zeros((4, 3))
    assert L.vb200_scene_create(x.ctypes.data_as(dp), None, 4, -1.0, 0, C.byref(h)) == vb.lib.ERR_INVALID
    assert L.vb200_scene_create(None, None, 4, 0.1, 0, C.byref(h)) == vb.lib.ERR_INVALID
    assert L.vb200_scene_destroy(None) == 0 and L.vb200_batch_destroy(None) == 0
    k = C.c_int64()
    assert L.vb200_voxel_downsample(x.ctypes.data_as(dp), None, 4, 0.0, 0, x.ctypes.data_as(dp), None,
                                    C.byref(k)) == vb.lib.ERR_INVALID
