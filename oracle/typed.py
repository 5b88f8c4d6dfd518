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
        u64 = C.c_uint64
        _lib.bo_index_new.argtypes = [C.c_int]
        _lib.bo_index_new.restype = vp
        _lib.bo_index_free.argtypes = [vp]
        _lib.bo_index_free.restype = None
        _lib.bo_index_build.argtypes = [C.POINTER(capi.BBConfig), vp, vp]
        _lib.bo_index_build.restype = None
        _lib.bo_merge_batch_indexed.argtypes = _lib.bo_merge_batch.argtypes + [C.POINTER(vp), C.c_int]
        _lib.bo_merge_batch_indexed.restype = C.c_int
        _lib.bo_index_equals.argtypes = [vp, u64, vp, u64]
        _lib.bo_index_equals.restype = u64
        _lib.bo_index_count.argtypes = [vp, u64]
        _lib.bo_index_count.restype = u64
        _lib.bo_index_range.argtypes = [vp, C.POINTER(capi.BBBound), C.POINTER(capi.BBBound), vp, u64]
        _lib.bo_index_range.restype = u64
        _lib.bo_index_entries.argtypes = [vp]
        _lib.bo_index_entries.restype = u64
    return _lib


class TypedOracle:
    """Host table + sequential replay: the restated reference on typed buffers."""

    def __init__(self, cfg: capi.BBConfig):
        self.cfg = cfg
        self.table = np.zeros(int(cfg.capacity), codec.ROW_DTYPE)
        self.seq = 0
        self.indices: dict[int, int] = {}  # field slot -> bo_index*

    def __del__(self):
        for h in getattr(self, "indices", {}).values():
            lib().bo_index_free(h)

    # ---- BulletQuery (results in the reference's Map / Set order)
    def index_create(self, field: int):
        if field in self.indices:
            return
        h = lib().bo_index_new(field)
        lib().bo_index_build(C.byref(self.cfg), self.table.ctypes.data, h)
        self.indices[field] = h

    def query_equals(self, field: int, key: int) -> np.ndarray:
        n = int(lib().bo_index_count(self.indices[field], key))
        out = np.zeros(max(n, 1), np.uint32)
        lib().bo_index_equals(self.indices[field], key, out.ctypes.data, n)
        return out[:n]

    def query_count(self, field: int, key: int) -> int:
        return int(lib().bo_index_count(self.indices[field], key))

    def query_range(self, field: int, lo, hi) -> np.ndarray:
        bl, bh = capi.bound_struct(lo), capi.bound_struct(hi)
        n = int(lib().bo_index_range(self.indices[field], C.byref(bl), C.byref(bh), None, 0))
        out = np.zeros(max(n, 1), np.uint32)
        lib().bo_index_range(self.indices[field], C.byref(bl), C.byref(bh), out.ctypes.data, n)
        return out[:n]

    def index_entries(self, field: int) -> int:
        return int(lib().bo_index_entries(self.indices[field]))

    def load(self, path_id, rows):
        self.table[np.asarray(path_id, np.int64)] = rows

    def merge(self, batch: codec.Batch, threads: int = 1, out: capi.ChangeBuffers | None = None):
        out = out or capi.ChangeBuffers(batch.n)
        bs, cs = capi.batch_struct(batch), out.struct()
        if self.indices:
            hs = (C.c_void_p * len(self.indices))(*self.indices.values())
            rc = lib().bo_merge_batch_indexed(C.byref(self.cfg), self.table.ctypes.data, self.seq,
                                              C.byref(bs), C.byref(cs), hs, len(self.indices))
        elif threads > 1:
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
