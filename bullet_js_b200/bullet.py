"""Host-side mirror of the reference's public API for this path: `Bullet`, `BulletNode`,
`index / equals / range / count`, the batched sync ingress - same names, argument meaning
and return shapes as the JavaScript, every decision taken on the GPU.

    reference                                            here
    new Bullet({storage:false, disableNetwork:true})     Bullet({"users": schema}, capacity=...)
    bullet.get(path).put(data) / .value() / .on(cb)      src/bullet.js:681-759
    bullet.setData(path, data, broadcast)                src/bullet.js:139-155
    sync._processSyncEntries(entries)                    src/bullet-network-sync.js:551-569
    bullet.index / equals / range / count                src/bullet.js:313-357 -> src/bullet-query.js
    bullet.on(path, cb) + ancestor notification          src/bullet.js:227-266

It is the Python twin of the JS shim in INTEGRATION.md: intern paths, pack SoA batches
(codec), call the C ABI (engine), then replay the returned change set in arrival order
through the host-side effects of `_applyUpdate` (log capped at 1000, listeners on the path
and every ancestor).  One device table per collection (`users/*`, `products/*`): paths must
be `<collection>/<key>`, one granularity per collection (SURVEY.md 8a restriction 2).
There is no CPU fallback: values outside the typed domain raise `codec.DomainError`.
"""
from __future__ import annotations

import time
from typing import Callable

import numpy as np

from . import codec
from .engine import Engine


class BulletNode:
    """src/bullet.js:681-759 (value / put / on / get)."""

    def __init__(self, bullet: "Bullet", path: str):
        self.bullet, self.path = bullet, path

    def value(self):
        return self.bullet._get_data(self.path)

    def put(self, data):
        self.bullet.setData(self.path, data)
        return self

    def on(self, callback: Callable):
        """src/bullet.js:710-720: subscribe, then call back at once with the current value."""
        self.bullet.on(self.path, callback)
        callback(self.value())
        return self

    def get(self, sub: str) -> "BulletNode":
        return BulletNode(self.bullet, f"{self.path}/{sub}")

    def off(self, callback: Callable | None = None):
        """src/bullet.js:733-746: drop one subscription, or all of this path's."""
        subs = self.bullet.listeners.get(self.path)
        if subs is not None:
            if callback is not None:
                if callback in subs:
                    subs.remove(callback)
            else:
                self.bullet.listeners[self.path] = []
        return self

    def remove(self):
        """src/bullet.js:752-755: a local put of null (the reference has no tombstones)."""
        self.bullet.setData(self.path, None)
        return self

    def __repr__(self):
        return f"BulletNode({self.path!r})"

    def __eq__(self, other):
        return isinstance(other, BulletNode) and other.path == self.path

    def __hash__(self):
        return hash(self.path)


class _Collection:
    def __init__(self, name: str, schema: codec.Schema, capacity: int, device: int, track_modified: bool = False,
                 exact_order: bool = False):
        self.name, self.schema = name, schema
        self.engine = Engine.for_schema(schema, capacity, device=device, post_getdata=True, track_modified=track_modified,
                                        exact_order=exact_order)
        self.indexed: set[int] = set()
        self.epoch_ms: list[float] = [0.0]  # Date.now() of merge call k (meta.lastModified at call granularity); [0] unused

    def slot(self, field: str) -> int:
        return self.schema.fields.index(field)


class Bullet:
    def __init__(self, collections: dict[str, codec.Schema], capacity: int = 1 << 16, device: int = 0,
                 track_modified: bool = False, clock: Callable[[], float] | None = None, exact_order: bool = False):
        """exact_order: equals / range return the reference's exact lists - buckets in Map order, paths in Set order
        (BB_CFG_EXACT_ORDER) - instead of the same nodes in device order."""
        self._c = {name: _Collection(name, schema, capacity, device, track_modified, exact_order)
                   for name, schema in collections.items()}
        self.track_modified = track_modified
        self._now = clock or (lambda: time.time() * 1000.0)  # Date.now()
        self.log: list[dict] = []          # src/bullet.js:206-215
        self.listeners: dict[str, list] = {}
        self.decisions: list[int] = []     # decision code of every setData, arrival order

    # ---- src/bullet.js:104-108, 681-759
    def get(self, path: str) -> BulletNode:
        return BulletNode(self, path)

    def on(self, path: str, callback: Callable):
        self.listeners.setdefault(path, []).append(callback)
        return self

    def close(self):
        for c in self._c.values():
            c.engine.close()

    # ---- writes
    def _split(self, path: str):
        parts = path.split("/")
        if len(parts) != 2 or not all(parts) or parts[0] not in self._c:
            raise codec.DomainError(f"path {path!r} is not <collection>/<key> of a configured collection")
        return self._c[parts[0]]

    def setData(self, path: str, data, broadcast: bool = True):
        """One local put (src/bullet.js:139-155): a batch of one."""
        self._merge([(path, data, None)])

    def process_sync_entries(self, entries):
        """BulletNetworkSync._processSyncEntries (sync:551-569): entries = [{path, data,
        vectorClock}] applied in order; objects carry their clock, primitives do not (sync:560-563)."""
        self._merge([(e["path"], e["data"], e.get("vectorClock") if isinstance(e["data"], dict) else None)
                     for e in entries])

    def _merge(self, updates):
        # consecutive runs of the same collection keep the global arrival order of effects
        i = 0
        while i < len(updates):
            col = self._split(updates[i][0])
            j = i
            while j < len(updates) and self._split(updates[j][0]) is col:
                j += 1
            run = updates[i:j]
            batch = codec.encode_updates(col.schema, run)
            ch = col.engine.merge(batch)
            while len(col.epoch_ms) <= col.engine.epoch:  # lastModified of everything this call accepted (src/bullet.js:201)
                col.epoch_ms.append(self._now())
            self.decisions.extend(ch.decision.tolist())
            for entry in codec.decode_changes(col.schema, batch, ch):   # arrival order
                self._apply_effects(entry["path"], entry["value"], entry["vectorClock"])
            i = j

    def _apply_effects(self, path, value, clock):
        """The host-visible part of _applyUpdate + _notify (src/bullet.js:206-266)."""
        self.log.append({"op": "set", "path": path, "data": value, "vectorClock": clock, "timestamp": time.time()})
        if len(self.log) > 1000:
            del self.log[: len(self.log) - 1000]
        parts = path.split("/")
        for cb in self.listeners.get(path, []):
            cb(value)
        for k in range(len(parts) - 1, -1, -1):  # every ancestor, root ("") last (src/bullet.js:238-255)
            parent = "/".join(parts[:k])
            for cb in self.listeners.get(parent, []):
                cb(self._get_data(parent) if parent else {c: self._get_data(c) for c in self._c})

    # ---- cold start / export (src/bullet-file-storage.js:96-210; bullet_js_b200/persist.py)
    def load_reference_state(self, store: dict, meta: dict):
        """What BulletFileStorage._loadData merges back at start: every configured collection of `store`
        goes into its device table with M = meta[path].vectorClock and no crt clock (V absent)."""
        from . import persist

        for name, col in self._c.items():
            if name in store:
                ids, rows = persist.import_collection(col.schema, name, store, meta)
                if len(ids):
                    col.engine.table_load(ids, rows)
        return self

    def export_reference_state(self):
        """-> (store, meta) as BulletFileStorage._saveData would stringify them (without source / lastModified)."""
        from . import persist

        store, meta = {}, {}
        for name, col in self._c.items():
            n = len(col.schema.paths)
            ids = np.arange(n, dtype=np.uint64)
            records, m = persist.export_collection(col.schema, name, ids, col.engine.table_read(ids) if n else [])
            if records:
                store[name] = records
            meta.update(m)
        return store, meta

    def collect_sync_entries(self, since: float = 0, filter_records: bool = False):
        """What a reference peer would be sent for a full sync request (src/bullet-network-sync.js:592-664),
        in chunks of 50 (`:713-723`).  With `track_modified` the rows are selected ON THE DEVICE (bb_sync_collect):
        `since` (a Date.now() value) becomes the ordinal of the first merge call at or after it, only the selected
        rows cross PCIe, and every entry carries the lastModified of the call that last wrote its path.
        `filter_records`: also filter whole records by their own lastModified (the reference never does: it looks
        meta up at the leaf path, src/bullet-network-sync.js:627-636)."""
        from . import persist

        if not self.track_modified:
            store, meta = self.export_reference_state()
            return persist.chunk_sync_data(persist.collect_full_sync_data(store, meta, since))
        store, meta = {}, {}
        for name, col in self._c.items():
            since_epoch = 0
            if since > 0:  # first call whose time is >= since; none: one past the last call
                since_epoch = next((k for k in range(1, len(col.epoch_ms)) if col.epoch_ms[k] >= since), len(col.epoch_ms))
            ids, rows, ep = col.engine.sync_collect(since_epoch, filter_records)
            records, m = persist.export_collection(col.schema, name, ids, rows)
            for pid, e in zip(ids.tolist(), ep.tolist()):
                path = col.schema.paths.name(int(pid))
                if path in m and e:
                    m[path]["lastModified"] = col.epoch_ms[e]
            if records:
                store[name] = records
            meta.update(m)
        return persist.chunk_sync_data(persist.collect_full_sync_data(store, meta, since))

    # ---- reads (src/bullet.js:115-129; materialising like the reference's _getData)
    def _get_data(self, path: str):
        parts = path.split("/")
        if len(parts) == 1 and parts[0] in self._c:
            col = self._c[parts[0]]
            n = len(col.schema.paths)
            rows = col.engine.table_read(np.arange(n, dtype=np.uint64)) if n else []
            order = sorted(range(n), key=lambda i: int(rows[i]["cseq"]))
            return {col.schema.paths.name(i).split("/")[1]: codec.decode_row(col.schema, rows[i])["value"]
                    for i in order if int(rows[i]["cseq"])}
        col = self._split(path)
        row = col.engine.table_read([col.schema.paths.id(path)], materialise=True)[0]
        return codec.decode_row(col.schema, row)["value"]

    def meta(self, path: str):
        """meta[path].vectorClock (src/bullet.js:198-203) or None."""
        col = self._split(path)
        return codec.decode_row(col.schema, col.engine.table_read([col.schema.paths.id(path)])[0])["M"]

    # ---- queries (src/bullet.js:313-357 -> src/bullet-query.js)
    def index(self, path: str, field: str):
        col = self._c[path]
        f = col.slot(field)
        if f not in col.indexed:
            col.engine.index_create(f)
            col.indexed.add(f)
        return self

    def _nodes(self, col, ids):
        return [BulletNode(self, col.schema.paths.name(i)) for i in ids]

    def equals(self, path: str, field: str, value):
        col = self._c[path]
        self.index(path, field)  # query:194-196 creates the index on first use
        key = col.schema.index_key(value)
        return [] if key is None else self._nodes(col, col.engine.query_equals(col.slot(field), key))

    def count(self, path: str, field: str, value) -> int:
        col = self._c[path]
        self.index(path, field)
        key = col.schema.index_key(value)
        return 0 if key is None else col.engine.query_count(col.slot(field), key)

    def range(self, path: str, field: str, min=None, max=None, *, min_undefined=False, max_undefined=False):
        """query:221-261.  Python has no `undefined`: pass min_undefined / max_undefined for it
        (the reference then matches nothing); None is JS null."""
        col = self._c[path]
        self.index(path, field)
        if min_undefined or max_undefined:
            return []
        return self._nodes(col, col.engine.query_range(col.slot(field), col.schema.bound(min, False),
                                                       col.schema.bound(max, True)))
