"""ctypes view of include/bullet_b200.h and the loader for libbulletb200.so.

There is NO CPU fallback: `load()` raises if the CUDA library has not been built
(run `python -c "import __graft_entry__ as g; g.build()"` or
`make -C bullet_js_b200/csrc`).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import codec

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "csrc", "libbulletb200.so")

ABI_VERSION = 1
NO_SLOT = 0x1FFFFFFF
NCCL_ID_BYTES = 128

OK, ERR_ARG, ERR_CUDA, ERR_DOMAIN, ERR_CAPACITY, ERR_STATE = 0, -1, -2, -3, -4, -5
_ERR_NAMES = {
    ERR_ARG: "BB_ERR_ARG", ERR_CUDA: "BB_ERR_CUDA", ERR_DOMAIN: "BB_ERR_DOMAIN",
    ERR_CAPACITY: "BB_ERR_CAPACITY", ERR_STATE: "BB_ERR_STATE",
}


class BBConfig(C.Structure):
    _fields_ = [
        ("abi_version", C.c_uint32), ("device", C.c_int32), ("n_fields", C.c_uint32),
        ("local_peer", C.c_uint32), ("capacity", C.c_uint64), ("rank_object", C.c_uint64),
        ("rank_true", C.c_uint64), ("rank_false", C.c_uint64), ("rank_nan", C.c_uint64),
        ("flags", C.c_uint32), ("reserved", C.c_uint32),
    ]


class BBBatch(C.Structure):
    _fields_ = [
        ("n", C.c_uint64), ("path_id", C.c_void_p), ("head", C.c_void_p),
        ("clk", C.c_void_p), ("val", C.c_void_p),
    ]


class BBChanges(C.Structure):
    _fields_ = [
        ("cap", C.c_uint64), ("verdict", C.c_void_p), ("n_changes", C.c_void_p),
        ("idx", C.c_void_p), ("head", C.c_void_p), ("clk", C.c_void_p), ("val", C.c_void_p),
    ]


class BBBound(C.Structure):
    _fields_ = [("num", C.c_double), ("rank", C.c_uint64), ("flags", C.c_uint32), ("reserved", C.c_uint32)]


class BBHits(C.Structure):
    _fields_ = [("cap", C.c_uint64), ("node", C.c_void_p), ("n_dense", C.c_void_p), ("n_extra", C.c_void_p)]


class BBGatheredHits(C.Structure):
    """bb_gathered_hits: the all-gathered result of a sharded query (device pointer + per-rank offsets)."""
    _fields_ = [("node", C.c_void_p), ("offset", C.c_uint64 * 17), ("total", C.c_uint64)]


class BulletB200Error(RuntimeError):
    def __init__(self, code, msg=""):
        self.code = code
        super().__init__(f"{_ERR_NAMES.get(code, code)}: {msg}")


# Every symbol include/bullet_b200.h declares (tests check the .so exports them all).
EXPORTS = [
    "bb_abi_version", "bb_create", "bb_destroy", "bb_last_error",
    "bb_table_load", "bb_table_read", "bb_table_clear", "bb_reserve",
    "bb_merge_batch", "bb_merge_batch_dev", "bb_sync", "bb_epoch", "bb_sync_collect",
    "bb_index_create", "bb_index_create_fields", "bb_query_equals", "bb_query_count", "bb_query_range",
    "bb_query_equals_dev", "bb_query_range_dev", "bb_index_stats",
    "bb_route_pack_dev", "bb_router_unique_id", "bb_router_create", "bb_router_destroy",
    "bb_router_last_error", "bb_router_set_sharding", "bb_router_route_dev", "bb_router_acquire", "bb_router_release",
    "bb_router_merge_batch", "bb_router_query_reserve", "bb_router_query_range", "bb_router_query_equals", "bb_router_query_fetch",
    "bb_router_sent_bytes", "bb_router_launch_count", "bb_router_last_ms",
    "bb_launch_count", "bb_last_phase_ms", "bb_phase_ms", "bb_phase_events",
]

_lib = None


def load():
    """dlopen libbulletb200.so (no GPU needed just to load it)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA extension must be built "
            "(__graft_entry__.build()); bullet_js_b200 has no CPU fallback"
        )
    lib = C.CDLL(LIB_PATH)
    vp, u64, i32 = C.c_void_p, C.c_uint64, C.c_int
    lib.bb_abi_version.restype = i32
    lib.bb_create.argtypes = [C.POINTER(BBConfig), C.POINTER(vp)]
    lib.bb_create.restype = i32
    lib.bb_destroy.argtypes = [vp]
    lib.bb_destroy.restype = i32
    lib.bb_last_error.argtypes = [vp]
    lib.bb_last_error.restype = C.c_char_p
    lib.bb_table_load.argtypes = [vp, u64, vp, vp]
    lib.bb_table_load.restype = i32
    lib.bb_table_read.argtypes = [vp, u64, vp, vp, i32]
    lib.bb_table_read.restype = i32
    lib.bb_table_clear.argtypes = [vp]
    lib.bb_table_clear.restype = i32
    lib.bb_merge_batch.argtypes = [vp, C.POINTER(BBBatch), C.POINTER(BBChanges)]
    lib.bb_merge_batch.restype = i32
    lib.bb_merge_batch_dev.argtypes = [vp, C.POINTER(BBBatch), C.POINTER(BBChanges), vp]
    lib.bb_merge_batch_dev.restype = i32
    lib.bb_reserve.argtypes = [vp, u64, i32]
    lib.bb_reserve.restype = i32
    lib.bb_sync.argtypes = [vp, vp]
    lib.bb_sync.restype = i32
    u32 = C.c_uint32
    lib.bb_epoch.argtypes = [vp]
    lib.bb_epoch.restype = u64
    lib.bb_sync_collect.argtypes = [vp, u64, u32, u64, vp, vp, vp, C.POINTER(u64)]
    lib.bb_sync_collect.restype = i32
    lib.bb_index_create_fields.argtypes = [vp, u32, u64]
    lib.bb_index_create_fields.restype = i32
    lib.bb_index_create.argtypes = [vp, u32, u64]
    lib.bb_index_create.restype = i32
    lib.bb_query_equals.argtypes = [vp, u32, u64, C.POINTER(BBHits)]
    lib.bb_query_equals.restype = i32
    lib.bb_query_count.argtypes = [vp, u32, u64, C.POINTER(u64)]
    lib.bb_query_count.restype = i32
    lib.bb_query_range.argtypes = [vp, u32, C.POINTER(BBBound), C.POINTER(BBBound), C.POINTER(BBHits)]
    lib.bb_query_range.restype = i32
    lib.bb_query_equals_dev.argtypes = [vp, u32, u64, C.POINTER(BBHits), vp]
    lib.bb_query_equals_dev.restype = i32
    lib.bb_query_range_dev.argtypes = [vp, u32, C.POINTER(BBBound), C.POINTER(BBBound), C.POINTER(BBHits), vp]
    lib.bb_query_range_dev.restype = i32
    lib.bb_index_stats.argtypes = [vp, u32, C.POINTER(u64), C.POINTER(u64)]
    lib.bb_index_stats.restype = i32
    lib.bb_route_pack_dev.argtypes = [vp, u32, C.POINTER(BBBatch), C.POINTER(BBBatch), vp, vp]
    lib.bb_route_pack_dev.restype = i32
    lib.bb_router_unique_id.argtypes = [C.c_char_p]
    lib.bb_router_unique_id.restype = i32
    lib.bb_router_create.argtypes = [C.c_int32, u32, u32, C.c_char_p, u64, u64, C.POINTER(vp)]
    lib.bb_router_create.restype = i32
    lib.bb_router_set_sharding.argtypes = [vp, u32]
    lib.bb_router_set_sharding.restype = i32
    lib.bb_router_destroy.argtypes = [vp]
    lib.bb_router_destroy.restype = i32
    lib.bb_router_last_error.argtypes = [vp]
    lib.bb_router_last_error.restype = C.c_char_p
    lib.bb_router_route_dev.argtypes = [vp, C.POINTER(BBBatch), u32, C.POINTER(u64), vp]
    lib.bb_router_route_dev.restype = i32
    lib.bb_router_acquire.argtypes = [vp, u32, vp, C.POINTER(BBBatch)]
    lib.bb_router_acquire.restype = i32
    lib.bb_router_release.argtypes = [vp, u32, vp]
    lib.bb_router_release.restype = i32
    lib.bb_router_merge_batch.argtypes = [vp, vp, C.POINTER(BBBatch), C.POINTER(BBChanges), u32, C.POINTER(u64), vp]
    lib.bb_router_merge_batch.restype = i32
    lib.bb_router_query_reserve.argtypes = [vp, u64]
    lib.bb_router_query_reserve.restype = i32
    lib.bb_router_query_range.argtypes = [vp, vp, u32, C.POINTER(BBBound), C.POINTER(BBBound), C.POINTER(BBGatheredHits), vp]
    lib.bb_router_query_range.restype = i32
    lib.bb_router_query_equals.argtypes = [vp, vp, u32, u64, C.POINTER(BBGatheredHits), vp]
    lib.bb_router_query_equals.restype = i32
    lib.bb_router_query_fetch.argtypes = [vp, u64, u64, vp]
    lib.bb_router_query_fetch.restype = i32
    lib.bb_router_last_ms.argtypes = [vp, C.POINTER(C.c_double)]
    lib.bb_router_last_ms.restype = i32
    lib.bb_router_sent_bytes.argtypes = [vp]
    lib.bb_router_sent_bytes.restype = u64
    lib.bb_router_launch_count.argtypes = [vp]
    lib.bb_router_launch_count.restype = u64
    lib.bb_launch_count.argtypes = [vp]
    lib.bb_launch_count.restype = u64
    lib.bb_last_phase_ms.argtypes = [vp, C.c_char_p]
    lib.bb_last_phase_ms.restype = C.c_double
    lib.bb_phase_events.argtypes = [vp, i32]
    lib.bb_phase_events.restype = i32
    lib.bb_phase_ms.argtypes = [vp, C.c_char_p, C.c_uint32]
    lib.bb_phase_ms.restype = C.c_double
    _lib = lib
    return lib


def make_config(capacity, n_fields=codec.MAX_FIELDS, local_peer=0, device=0, flags=0,
                rank_object=0, rank_true=0, rank_false=0, rank_nan=0) -> BBConfig:
    return BBConfig(
        abi_version=ABI_VERSION, device=device, n_fields=n_fields, local_peer=local_peer,
        capacity=capacity, rank_object=rank_object, rank_true=rank_true,
        rank_false=rank_false, rank_nan=rank_nan, flags=flags, reserved=0,
    )


def bound_struct(b) -> BBBound:
    """codec.Schema.bound() record -> bb_bound."""
    return BBBound(num=float(b["num"]), rank=int(b["rank"]), flags=int(b["flags"]), reserved=0)


class HitBuffers:
    """Caller-owned output of bb_query_equals / bb_query_range."""

    def __init__(self, cap: int):
        self.cap = max(int(cap), 1)
        self.node = np.zeros(self.cap, np.uint32)
        self.n_dense = np.zeros(1, np.uint64)
        self.n_extra = np.zeros(1, np.uint64)

    def struct(self) -> BBHits:
        return BBHits(cap=self.cap, node=_ptr(self.node), n_dense=_ptr(self.n_dense), n_extra=_ptr(self.n_extra))

    def result(self) -> np.ndarray:
        return self.node[: int(self.n_dense[0]) + int(self.n_extra[0])].copy()


def _ptr(a: np.ndarray):
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data


def batch_struct(b: codec.Batch) -> BBBatch:
    assert b.path_id.dtype == np.uint64 and b.head.dtype == codec.HEAD_DTYPE
    assert b.clk.dtype == np.uint32 and b.val.dtype == np.uint64
    return BBBatch(n=b.n, path_id=_ptr(b.path_id), head=_ptr(b.head), clk=_ptr(b.clk),
                   val=_ptr(b.val))


class ChangeBuffers:
    """Caller-owned output buffers for bb_merge_batch (reusable across calls)."""

    def __init__(self, cap: int):
        cap = max(int(cap), 1)
        self.cap = cap
        self.verdict = np.zeros(cap, np.uint32)
        self.n_changes = np.zeros(1, np.uint64)
        self.idx = np.zeros(cap, np.uint32)
        self.head = np.zeros(cap, codec.HEAD_DTYPE)
        self.clk = np.zeros((cap, codec.MAX_PEERS), np.uint32)
        self.val = np.zeros((cap, codec.MAX_FIELDS), np.uint64)

    def struct(self) -> BBChanges:
        return BBChanges(cap=self.cap, verdict=_ptr(self.verdict), n_changes=_ptr(self.n_changes),
                         idx=_ptr(self.idx), head=_ptr(self.head), clk=_ptr(self.clk),
                         val=_ptr(self.val))

    def result(self, n: int, batch: "codec.Batch | None" = None) -> codec.Changes:
        k = int(self.n_changes[0])
        return codec.Changes.from_verdicts(self.verdict[:n], self.idx[:k], self.head[:k], self.clk[:k],
                                           self.val[:k], batch)
