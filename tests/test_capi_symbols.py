"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol
include/goldfish_b200.h declares (no compute without a GPU)."""
import os
import re
import ctypes
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "goldfish_b200.h")).read()
    return sorted(set(re.findall(r"\b(gf_[a-z_A-Z0-9]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(built_lib):
    from goldfish_b200 import _capi
    syms = declared_symbols()
    assert len(syms) >= 15
    raw = ctypes.CDLL(_capi.LIB_PATH)
    for s in syms:
        assert hasattr(raw, s), "missing export %s" % s
    assert sorted(_capi.SIGNATURES) == syms
    assert built_lib.gf_version() >= 100


def test_struct_layout_matches_header(built_lib):
    """Every ctypes mirror against the compiled header: size and offset of the last field (gf_abi_layout is a
    host-only export, so this runs without a GPU)."""
    from goldfish_b200 import _capi
    assert ctypes.sizeof(_capi.GfPatchDesc) == 18 * 4 + 5 * 8
    assert ctypes.sizeof(_capi.GfCsr) == 48
    _capi.check_abi(built_lib)
    for i, (T, last) in enumerate(_capi.ABI_STRUCTS):
        size, off = ctypes.c_int64(0), ctypes.c_int64(0)
        assert built_lib.gf_abi_layout(i, ctypes.byref(size), ctypes.byref(off)) == 0
        assert (size.value, off.value) == (ctypes.sizeof(T), getattr(T, last).offset), T.__name__
    assert built_lib.gf_abi_layout(99, ctypes.byref(size), ctypes.byref(off)) != 0


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from goldfish_b200 import problems, _capi
    from goldfish_b200.device_model import DeviceModel
    with pytest.raises(_capi.GoldfishError):
        DeviceModel(problems.tbeam(num_el=4))


def test_product_never_imports_oracle():
    pk = os.path.join(ROOT, "goldfish_b200")
    for dp, _, fs in os.walk(pk):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f
