"""The `native` addon of js/bullet-b200.js for tests: its typed-array surface (create / mergeBatch / tableRead,
see INTEGRATION.md) implemented in Python over any typed engine with the TypedOracle / Engine interface and handed
to the shim running inside oracle/minijs.  The buffers cross exactly as they would cross N-API: the JS side packs
(js/pack.js), this side only reinterprets bytes as bb_batch / bb_changes / bb_row."""
from __future__ import annotations

import numpy as np

from bullet_js_b200 import capi, codec
from oracle.minijs import interp as I
from oracle.minijs.builtins import to_py


class NativeBridge:
    def __init__(self, rt, make_engine):
        self.rt, self.make_engine = rt, make_engine
        self.codes: list[int] = []    # every decision the engine took, in order
        self.batches: list[int] = []  # size of every mergeBatch call
        self.last_batch = None        # codec.Batch view of the most recent call's input buffers
        self.engine = None

    def js_object(self):
        o = I.JSObject(I.OBJECT_PROTO)
        for name, fn in (("create", self._create), ("mergeBatch", self._merge_batch), ("tableRead", self._table_read),
                         ("indexCreate", self._index_create), ("queryEquals", self._query_equals),
                         ("queryCount", self._query_count), ("queryRange", self._query_range)):
            o.define(name, I.JSFunction(name, (lambda f: lambda this, a: f(*a))(fn)), enumerable=True)
        return o

    def _create(self, options):
        opt = to_py(options)
        cfg = capi.make_config(int(opt["capacity"]), n_fields=int(opt["nFields"]), local_peer=int(opt["localPeer"]),
                               flags=int(opt["flags"]), rank_object=int(opt["rankObject"]), rank_true=int(opt["rankTrue"]),
                               rank_false=int(opt["rankFalse"]), rank_nan=int(opt["rankNaN"]))
        self.engine = self.make_engine(cfg)
        return 1.0

    @staticmethod
    def batch_of(n, path_id, head, clk, val) -> codec.Batch:
        return codec.Batch(np.frombuffer(path_id.raw(), np.uint64).copy(), np.frombuffer(head.raw(), codec.HEAD_DTYPE).copy(),
                           np.frombuffer(clk.raw(), np.uint32).reshape(n, 8).copy(),
                           np.frombuffer(val.raw(), np.uint64).reshape(n, 4).copy())

    def _merge_batch(self, ctx, n, path_id, head, clk, val, out):
        n = int(n)
        batch = self.batch_of(n, path_id, head, clk, val)
        self.last_batch = batch
        ch = self.engine.merge(batch)
        self.codes.extend(ch.decision.tolist())
        self.batches.append(n)
        k = len(ch.idx)
        verdict = (ch.decision.astype(np.uint32) << 29) | np.uint32(codec.NO_SLOT)
        verdict[ch.idx.astype(np.int64)] = (ch.decision[ch.idx.astype(np.int64)].astype(np.uint32) << 29) | np.arange(k, dtype=np.uint32)

        def fill(name, arr):
            t = out.get(name)
            raw = np.ascontiguousarray(arr).tobytes()
            t.buf[t.off:t.off + len(raw)] = raw
        fill("verdict", verdict)
        fill("idx", ch.idx.astype(np.uint32))
        fill("head", ch.head)
        fill("clk", ch.clk)
        fill("val", ch.val)
        return float(k)

    def _table_read(self, ctx, path_id):
        ids = np.array([int(path_id)], np.uint64)
        row = (self.engine.table_read(ids) if hasattr(self.engine, "table_read") else self.engine.read(ids))[0]
        proto = self.rt.globals.vars["Uint32Array"].get("prototype")
        return I.JSTypedArray(proto, "Uint32Array", bytearray(row.tobytes()), "<I", 4, 0, 32)

    # ---- bb_index_create / bb_query_*
    def _u32(self, ids):
        proto = self.rt.globals.vars["Uint32Array"].get("prototype")
        raw = np.ascontiguousarray(ids, np.uint32).tobytes()
        return I.JSTypedArray(proto, "Uint32Array", bytearray(raw), "<I", 4, 0, len(raw) // 4)

    def _index_create(self, ctx, field):
        self.engine.index_create(int(field))
        return 0.0

    def _query_equals(self, ctx, field, key_lo, key_hi):
        return self._u32(self.engine.query_equals(int(field), int(key_lo) | (int(key_hi) << 32)))

    def _query_count(self, ctx, field, key_lo, key_hi):
        return float(self.engine.query_count(int(field), int(key_lo) | (int(key_hi) << 32)))

    def _query_range(self, ctx, field, lo_num, lo_rank, lo_flags, hi_num, hi_rank, hi_flags):
        def rec(num, rank, flags):
            b = np.zeros((), codec.BOUND_DTYPE)
            b["num"], b["rank"], b["flags"] = float(num), int(rank), int(flags)
            return b
        return self._u32(self.engine.query_range(int(field), rec(lo_num, lo_rank, lo_flags), rec(hi_num, hi_rank, hi_flags)))
