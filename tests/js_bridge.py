"""The `native` object of js/bullet-b200.js for tests: the N-API addon's surface (create / merge / clocks)
implemented in Python over any typed engine with the TypedOracle / Engine interface, handed to the shim running
inside oracle/minijs.  In production this is native/bullet_b200_napi.c over libbulletb200.so."""
from __future__ import annotations

import numpy as np

from bullet_js_b200 import codec
from oracle.jsvalue import UNDEFINED
from oracle.minijs import interp as I
from oracle.minijs.builtins import from_py, to_py
from tests import streamgen


class NativeBridge:
    def __init__(self, make_engine):
        self.make_engine = make_engine
        self.codes: list[int] = []   # every decision the engine took, in order
        self.batches: list[int] = []  # size of every merge call
        self.schema = None
        self.engine = None

    def js_object(self):
        o = I.JSObject(I.OBJECT_PROTO)
        for name, fn in (("create", self._create), ("merge", self._merge), ("clocks", self._clocks)):
            o.define(name, I.JSFunction(name, (lambda f: lambda this, a: f(*a))(fn)), enumerable=True)
        return o

    def _create(self, options):
        opt = to_py(options)
        self.schema = codec.Schema(streamgen.FIELDS, streamgen.PEERS, codec.StringDict(streamgen.STRINGS), opt["localPeer"])
        self.engine = self.make_engine(self.schema, int(opt.get("capacity", 64)), bool(opt.get("postGetData", False)))
        return 1.0

    def _merge(self, ctx, entries):
        ops = []
        for e in to_py(entries):
            clock = e.get("vectorClock", UNDEFINED)
            local = e["local"] or clock is UNDEFINED or clock is None
            ops.append((e["path"], e["data"], None if local else clock))
        batch = codec.encode_updates(self.schema, ops)
        ch = self.engine.merge(batch)
        self.codes.extend(ch.decision.tolist())
        self.batches.append(len(ops))
        changes = [dict(i=float(c["seq"]), value=c["value"], vectorClock=c["vectorClock"])
                   for c in codec.decode_changes(self.schema, batch, ch)]
        return from_py(dict(codes=[float(x) for x in ch.decision.tolist()], changes=changes))

    def _clocks(self, ctx, path):
        i = self.schema.paths.id(path)
        row = (self.engine.table_read(np.array([i], np.uint64)) if hasattr(self.engine, "table_read")
               else self.engine.read(np.array([i], np.uint64)))[0]
        d = codec.decode_row(self.schema, row)
        return from_py(dict(meta=d["M"] if d["M"] is not None else UNDEFINED, crt=d["V"] if d["V"] is not None else UNDEFINED))
