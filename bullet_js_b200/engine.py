"""Thin object wrapper over the C ABI (include/bullet_b200.h).

`Engine` owns one bb_ctx == one GPU-resident shard of the graph table.  All
compute goes through libbulletb200.so; a missing library or device raises.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi, codec


class Engine:
    def __init__(self, capacity: int, *, local_peer: int = 0, n_fields: int = codec.MAX_FIELDS,
                 device: int = 0, post_getdata: bool = False, ordered_changes: bool = False, radix_sort: bool = False,
                 full_sort: bool = False, hot_keys: bool = False, compact_changes: bool = False, track_modified: bool = False, exact_order: bool = False, rank_object: int = 0, rank_true: int = 0,
                 rank_false: int = 0, rank_nan: int = 0):
        self.lib = capi.load()
        self.cfg = capi.make_config(
            capacity, n_fields=n_fields, local_peer=local_peer, device=device,
            flags=(codec.CFG_POST_GETDATA if post_getdata else 0)
            | (codec.CFG_ORDERED_CHANGES if ordered_changes else 0)
            | (codec.CFG_RADIX_SORT if radix_sort else 0)
            | (codec.CFG_FULL_SORT if full_sort else 0)
            | (codec.CFG_HOT_KEYS if hot_keys else 0)
            | (codec.CFG_COMPACT_CHANGES if compact_changes else 0)
            | (codec.CFG_TRACK_MODIFIED if track_modified else 0)
            | (codec.CFG_EXACT_ORDER if exact_order else 0), rank_object=rank_object,
            rank_true=rank_true, rank_false=rank_false, rank_nan=rank_nan)
        h = C.c_void_p()
        rc = self.lib.bb_create(C.byref(self.cfg), C.byref(h))
        if rc != 0:
            raise capi.BulletB200Error(rc, (self.lib.bb_last_error(None) or b"").decode())
        self._h = h
        self.capacity = int(capacity)
        self.device = device

    @classmethod
    def for_schema(cls, schema: codec.Schema, capacity: int, **kw):
        return cls(capacity, local_peer=schema.peers.index(schema.local_peer), **schema.config_ranks(), **kw)

    def _check(self, rc):
        if rc != 0:
            raise capi.BulletB200Error(rc, (self.lib.bb_last_error(self._h) or b"").decode())

    def close(self):
        if getattr(self, "_h", None):
            self.lib.bb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- table
    def table_load(self, path_id, rows):
        path_id = np.ascontiguousarray(path_id, np.uint64)
        rows = np.ascontiguousarray(rows, codec.ROW_DTYPE)
        assert path_id.shape[0] == rows.shape[0]
        self._check(self.lib.bb_table_load(self._h, path_id.shape[0], path_id.ctypes.data, rows.ctypes.data))

    def table_read(self, path_id, materialise=False):
        path_id = np.ascontiguousarray(path_id, np.uint64)
        rows = np.zeros(path_id.shape[0], codec.ROW_DTYPE)
        self._check(self.lib.bb_table_read(self._h, path_id.shape[0], path_id.ctypes.data, rows.ctypes.data,
                                           1 if materialise else 0))
        return rows

    def reserve(self, max_batch: int, host_entry: bool = True):
        self._check(self.lib.bb_reserve(self._h, int(max_batch), 1 if host_entry else 0))

    def table_clear(self):
        self._check(self.lib.bb_table_clear(self._h))

    # ---- merge, host buffers (the reference-facing call)
    def merge(self, batch: codec.Batch, out: capi.ChangeBuffers | None = None) -> codec.Changes:
        out = out or capi.ChangeBuffers(batch.n)
        bs, cs = capi.batch_struct(batch), out.struct()
        self._check(self.lib.bb_merge_batch(self._h, C.byref(bs), C.byref(cs)))
        self.last_emitted = int(out.n_changes[0])  # entries that crossed the link (fewer with compact_changes)
        return out.result(batch.n, batch)

    def merge_raw(self, bs: capi.BBBatch, cs: capi.BBChanges):
        """bb_merge_batch on prebuilt structs (pinned buffers, no numpy copies)."""
        self._check(self.lib.bb_merge_batch(self._h, C.byref(bs), C.byref(cs)))

    # ---- merge, device pointers (inputs already resident in HBM)
    def merge_dev(self, bs: capi.BBBatch, cs: capi.BBChanges, stream: int = 0):
        self._check(self.lib.bb_merge_batch_dev(self._h, C.byref(bs), C.byref(cs), C.c_void_p(stream)))

    def route_pack_dev(self, world: int, bs: capi.BBBatch, out: capi.BBBatch, counts_ptr: int, stream: int = 0):
        """Stable partition of a device batch by owner rank (bullet_js_b200/shard.py)."""
        self._check(self.lib.bb_route_pack_dev(self._h, world, C.byref(bs), C.byref(out), C.c_void_p(counts_ptr),
                                               C.c_void_p(stream)))

    # ---- sync producer side (src/bullet-network-sync.js:592-664)
    @property
    def epoch(self) -> int:
        """Ordinal of the most recent merge call (what an accepted update's row is stamped with)."""
        return int(self.lib.bb_epoch(self._h))

    def sync_collect(self, since_epoch: int = 0, filter_records: bool = False, cap: int | None = None):
        """-> (path ids, rows, epochs) of the rows _collectFullSyncData(since) would visit, selected on the device."""
        cap = self.capacity if cap is None else int(cap)
        ids = np.zeros(max(cap, 1), np.uint64)
        rows = np.zeros(max(cap, 1), codec.ROW_DTYPE)
        ep = np.zeros(max(cap, 1), np.uint32)
        n = C.c_uint64(0)
        self._check(self.lib.bb_sync_collect(self._h, int(since_epoch), codec.COLLECT_FILTER_RECORDS if filter_records else 0,
                                             cap, ids.ctypes.data, rows.ctypes.data, ep.ctypes.data, C.byref(n)))
        k = int(n.value)
        order = np.argsort(ids[:k], kind="stable")  # warps finish in any order: ascending path id for the caller
        return ids[:k][order], rows[:k][order], ep[:k][order]

    def sync(self, stream: int = 0):
        self._check(self.lib.bb_sync(self._h, C.c_void_p(stream)))

    # ---- indices and queries (src/bullet-query.js)
    def index_create(self, field: int, extra_capacity: int | None = None):
        if extra_capacity is None:
            extra_capacity = max(4 * self.capacity, 1 << 16)
        self._check(self.lib.bb_index_create(self._h, field, int(extra_capacity)))

    def index_create_fields(self, fields, extra_capacity: int | None = None):
        """Several indices in one pass over the table."""
        if extra_capacity is None:
            extra_capacity = max(4 * self.capacity, 1 << 16)
        mask = 0
        for f in fields:
            mask |= 1 << int(f)
        self._check(self.lib.bb_index_create_fields(self._h, mask, int(extra_capacity)))

    def _hits(self, out, field):
        return out or capi.HitBuffers(sum(self.index_stats(field)))  # every entry could match

    def query_equals(self, field: int, key: int, out: capi.HitBuffers | None = None) -> np.ndarray:
        """Node ids whose index entry has this key (codec.Schema.index_key)."""
        out = self._hits(out, field)
        hs = out.struct()
        self._check(self.lib.bb_query_equals(self._h, field, int(key), C.byref(hs)))
        return out.result()

    def query_range(self, field: int, lo, hi, out: capi.HitBuffers | None = None) -> np.ndarray:
        """lo / hi: codec.Schema.bound() records."""
        out = self._hits(out, field)
        hs = out.struct()
        bl, bh = capi.bound_struct(lo), capi.bound_struct(hi)
        self._check(self.lib.bb_query_range(self._h, field, C.byref(bl), C.byref(bh), C.byref(hs)))
        return out.result()

    def query_range_raw(self, field: int, bl: capi.BBBound, bh: capi.BBBound, hs: capi.BBHits):
        """bb_query_range on prebuilt structs (pinned hit buffer, no numpy copies)."""
        self._check(self.lib.bb_query_range(self._h, field, C.byref(bl), C.byref(bh), C.byref(hs)))

    def query_count(self, field: int, key: int) -> int:
        n = C.c_uint64(0)
        self._check(self.lib.bb_query_count(self._h, field, int(key), C.byref(n)))
        return int(n.value)

    def index_stats(self, field: int):
        a, b = C.c_uint64(0), C.c_uint64(0)
        self._check(self.lib.bb_index_stats(self._h, field, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    # ---- telemetry
    def launch_count(self) -> int:
        return int(self.lib.bb_launch_count(self._h))

    def phase_events(self, on: bool = True):
        """Record the event between the front end and the merge kernels (serialises them: measurement only)."""
        self._check(self.lib.bb_phase_events(self._h, 1 if on else 0))

    def phase_ms(self, phase: str, calls_ago: int = 0) -> float:
        return float(self.lib.bb_phase_ms(self._h, phase.encode(), calls_ago))
