"""CPU-side checks of the boundary: the library loads, exports every symbol the
header declares, and refuses to run without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from bullet_js_b200 import capi, codec

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(capi.LIB_PATH):
        import __graft_entry__ as g

        g.build()
    return capi.load()


def test_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "bullet_b200.h")).read()
    declared = set(re.findall(r"\b(bb_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(capi.EXPORTS), declared ^ set(capi.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.bb_abi_version() == capi.ABI_VERSION


def test_struct_sizes_match_header():
    assert C.sizeof(capi.BBConfig) == 64
    assert C.sizeof(capi.BBBatch) == 40 and C.sizeof(capi.BBChanges) == 56
    assert codec.ROW_DTYPE.itemsize == 128 and codec.HEAD_DTYPE.itemsize == 16


def test_no_cpu_fallback(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    cfg = capi.make_config(16)
    h = C.c_void_p()
    rc = lib.bb_create(C.byref(cfg), C.byref(h))
    assert rc == capi.ERR_CUDA and not h.value
    assert b"no CPU fallback" in lib.bb_last_error(None)


def test_bad_config_rejected(lib):
    cfg = capi.make_config(0)
    h = C.c_void_p()
    assert lib.bb_create(C.byref(cfg), C.byref(h)) == capi.ERR_ARG
    cfg = capi.make_config(16, n_fields=9)
    assert lib.bb_create(C.byref(cfg), C.byref(h)) == capi.ERR_ARG


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "bullet_js_b200")
    for dp, _dn, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dp, fn), encoding="utf-8").read()
                assert "oracle" not in src.replace("oracle/", "").lower() or fn == "__init__.py" or \
                    not re.search(r"^\s*(from|import)\s+oracle", src, re.M), fn
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), fn
                assert "bullet_oracle" not in src, fn
