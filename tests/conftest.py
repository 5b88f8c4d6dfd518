import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


# Development aid (tools/emu): BB_EMU_TESTS=1 points the ctypes loader of THIS test process at the host-emulated
# build of the kernels (fibers, no GPU) and runs the `gpu` tests against it.  It debugs kernel logic in the GPU-less
# container; it is not a product path (bullet_js_b200/ never looks for that library) and proves nothing about the
# CUDA build - the round-end `-m gpu` run on a B200 does.
EMU = bool(os.environ.get("BB_EMU_TESTS"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    if EMU:
        from bullet_js_b200 import capi

        capi.LIB_PATH = os.path.join(ROOT, "tools", "emu", "_build", "libbulletb200_emu.so")


def _has_cuda():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_cuda() or EMU:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
