"""ctypes wrapper around oracle/bullet_oracle.c (TEST INFRASTRUCTURE ONLY)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from bullet_js_b200 import capi, codec

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libbullet_oracle.so")


def build(force=False):
    src = os.path.join(HERE, "bullet_oracle.c")
    hdr = os.path.join(HERE, "..", "include", "bullet_b200.h")
    if (not force and os.path.exists(LIB)
            and os.path.getmtime(LIB) >= max(os.path.getmtime(src), os.path.getmtime(hdr))):
        return LIB
    subprocess.check_call(["make", "-C", HERE, "-B", "_build/libbullet_oracle.so"],
                          stdout=subprocess.DEVNULL)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        vp = C.c_void_p
        _lib.bo_merge_batch.argtypes = [C.POINTER(capi.BBConfig), vp, C.c_uint64,
                                        C.POINTER(capi.BBBatch), C.POINTER(capi.BBChanges)]
        _lib.bo_merge_batch.restype = C.c_int
        _lib.bo_merge_batch_mt.argtypes = _lib.bo_merge_batch.argtypes + [C.c_int]
        _lib.bo_merge_batch_mt.restype = C.c_int
        _lib.bo_materialise.argtypes = [vp, C.c_uint64]
        _lib.bo_materialise.restype = None
    return _lib


class TypedOracle:
    """Host table + sequential replay: the restated reference on typed buffers."""

    def __init__(self, cfg: capi.BBConfig):
        self.cfg = cfg
        self.table = np.zeros(int(cfg.capacity), codec.ROW_DTYPE)
        self.seq = 0

    def load(self, path_id, rows):
        self.table[np.asarray(path_id, np.int64)] = rows

    def merge(self, batch: codec.Batch, threads: int = 1, out: capi.ChangeBuffers | None = None):
        out = out or capi.ChangeBuffers(batch.n)
        bs, cs = capi.batch_struct(batch), out.struct()
        if threads > 1:
            rc = lib().bo_merge_batch_mt(C.byref(self.cfg), self.table.ctypes.data, self.seq,
                                         C.byref(bs), C.byref(cs), threads)
        else:
            rc = lib().bo_merge_batch(C.byref(self.cfg), self.table.ctypes.data, self.seq,
                                      C.byref(bs), C.byref(cs))
        if rc != 0:
            raise capi.BulletB200Error(rc, "oracle")
        self.seq += batch.n
        return out.result(batch.n)

    def read(self, path_id, materialise=False):
        ids = np.asarray(path_id, np.int64)
        if materialise:
            for i in ids:
                lib().bo_materialise(self.table[i:i + 1].ctypes.data, self.seq)
        return self.table[ids].copy()
