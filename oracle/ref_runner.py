"""Runs the UNMODIFIED reference sources (src/bullet.js and what it requires) inside `oracle/minijs`
and exposes the same observable surface as `oracle.js_literal.RefBullet`, so the restatements can be
diffed against the reference itself.  TEST INFRASTRUCTURE ONLY.

What is the reference's and what is ours:
  * `new Bullet({...})`, `bullet.setData`, `bullet.get(p).put(v)`, `bullet.index/equals/range/count`,
    `BulletCRT.handleUpdate/resolve`, `BulletQuery._updateIndices`, the middleware wrapper and
    `BulletNetworkSync._processSyncEntries` (src/bullet-network-sync.js:551-569) are executed from the
    files under $BULLET_REFERENCE (default /root/reference), byte for byte.
  * The harness below only *observes*: it wraps `bullet.crt.handleUpdate` and `bullet._applyUpdate`
    on the instance (the reference's own extension idiom, src/bullet-query.js:13-21) to record the
    decision of every update and the ordered change set, and overrides `bullet.id` (a random UUID in
    the reference, src/bullet.js:273-282) so traces are reproducible.
  * The network is disabled (`disableNetwork: true`); the sync driver is constructed on a stub
    EventEmitter because only its `_processSyncEntries` loop is on the path.

`tests/golden/make_golden.py` uses this module to write tests/golden/*.json; the GPU box has no
/root/reference, so everything that runs there reads the committed fixtures instead.
"""
from __future__ import annotations

import math
import os

from .jsvalue import UNDEFINED
from .minijs import interp as I
from .minijs.builtins import Runtime, from_py, to_py

REASONS = {
    "no current state": 0,
    "identical clocks and values": 1,
    "identical clocks, decided by value comparison": None,  # 2 (incoming) / 3 (current)
    "incoming vector clock dominates": 4,
    "current vector clock dominates (incoming is historical)": 5,
    "concurrent modifications, merged objects": 6,
}

_HARNESS = r"""
const bullet = new Bullet(Object.assign({
  disableNetwork: true, server: false, storage: true, storageType: "memory",
  enableMiddleware: enableMiddleware, enableIndexing: enableIndexing,
  enableValidation: true, enableSerializer: true,
}, extraOptions));
bullet.id = peerId;
const trace = { decisions: [], changes: [] };
const crt = bullet.crt;
const origHandle = crt.handleUpdate;
crt.handleUpdate = function (path, data, fromNetwork) {
  const r = origHandle.call(crt, path, data, fromNetwork);
  trace.decisions.push({ path: path, reason: r.decision.reason, incoming: !!r.decision.incoming,
                         current: !!r.decision.current, concurrent: !!r.decision.concurrent,
                         historical: !!r.decision.historical, converge: !!r.decision.converge,
                         defer: !!r.decision.defer, doUpdate: !!r.doUpdate });
  return r;
};
const origApply = bullet._applyUpdate;
bullet._applyUpdate = function (path, value, vectorClock, fromNetwork) {
  trace.changes.push({ seq: trace.decisions.length - 1, path: path, value: snapshot(value),
                       vectorClock: snapshot(vectorClock), fromNetwork: !!fromNetwork });
  return origApply.call(bullet, path, value, vectorClock, fromNetwork);
};
const net = new EventEmitter();
const sync = new BulletNetworkSync(bullet, net, {});
return { bullet: bullet, trace: trace, sync: sync };
"""


def _reference_root():
    return os.environ.get("BULLET_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(_reference_root(), "src", "bullet.js"))


def js_value(v):
    """Python oracle value (dict / float / str / bool / None / UNDEFINED, NaN and -0 included) -> JS value."""
    return from_py(v)


def py_value(v):
    return to_py(v)


class _CrtView:
    def __init__(self, owner):
        self._o = owner

    @property
    def vectorClocks(self):
        m = I.get_member(self._o.bullet.get("crt"), "vectorClocks")
        return {k[1]: to_py(v) for k, (_, v) in m.data.items()}


class JSRefBullet:
    """The reference `Bullet`, driven from Python.  Attribute names follow js_literal.RefBullet."""

    def __init__(self, peer_id: str, enable_middleware=True, enable_indexing=True, options=None, console=None,
                 files=None):
        """files: dict shared between instances = the disk behind the reference's BulletFileStorage
        (options={"storageType": "file", "storagePath": "/data", "saveInterval": 0})."""
        root = _reference_root()
        self.rt = Runtime(console=console, files=files)
        rt = self.rt
        Bullet = rt.require(os.path.join(root, "src", "bullet.js"))
        Sync = rt.require(os.path.join(root, "src", "bullet-network-sync.js"))
        snapshot = I.JSFunction("snapshot", lambda this, a: from_py(to_py(a[0])))
        r = rt.eval(_HARNESS, Bullet=Bullet, BulletNetworkSync=Sync, EventEmitter=rt.require("events"),
                    peerId=peer_id, enableMiddleware=bool(enable_middleware), enableIndexing=bool(enable_indexing),
                    extraOptions=from_py(options or {}), snapshot=snapshot)
        self.bullet = r.get("bullet")
        self._trace = r.get("trace")
        self._sync = r.get("sync")
        self.id = peer_id
        self.crt = _CrtView(self)

    # ---- observation
    @property
    def decisions(self):
        out = []
        for d in to_py(self._trace.get("decisions")):
            code = REASONS[d["reason"]]
            if code is None:
                code = 2 if d["incoming"] else 3
            out.append(dict(path=d["path"], code=code, reason=d["reason"], doUpdate=d["doUpdate"]))
        return out

    @property
    def decision_flags(self):
        """Every decision's flags exactly as the reference returned them (src/bullet-crt.js:174-184, 208-278)."""
        keys = ("reason", "incoming", "current", "concurrent", "historical", "converge", "defer", "doUpdate")
        return [{k: d[k] for k in keys} for d in to_py(self._trace.get("decisions"))]

    @property
    def changes(self):
        return to_py(self._trace.get("changes"))

    def n_changes(self) -> int:
        return len(self._trace.get("changes").items)

    def last_change(self):
        return to_py(self._trace.get("changes").items[-1])

    @property
    def store(self):
        return to_py(self.bullet.get("store"))

    @property
    def meta(self):
        """path -> {source, vectorClock} (lastModified is wall-clock, dropped)."""
        m = to_py(self.bullet.get("meta"))
        return {p: {k: v for k, v in e.items() if k != "lastModified"} for p, e in m.items()}

    def alias(self, path) -> bool:
        """Is meta[path].vectorClock the very object crt.vectorClocks holds for the path?"""
        meta = self.bullet.get("meta").get(path)
        if not isinstance(meta, I.JSObject):
            return False
        m = meta.get("vectorClock")
        vc = I.get_member(self.bullet.get("crt"), "vectorClocks").data.get(I.map_key(path))
        return vc is not None and isinstance(m, I.JSObject) and vc[1] is m

    @property
    def log(self):
        out = []
        for e in to_py(self.bullet.get("log")):
            out.append({k: v for k, v in e.items() if k != "timestamp"})
        return out

    def index_dump(self):
        """{indexKey: [[bucketKey, [paths in Set order]] in Map order]} - the reference's exact iteration order."""
        q = self.bullet.get("query")
        out = {}
        indices = q.get("indices")  # plain object of Map(bucket key -> Set(node path)), src/bullet-query.js:4,38
        for k in indices.enumerable_keys():
            idx = indices.get(k)
            out[k] = [[bk, [p for p, _ in s.data.values()]] for bk, s in idx.data.values()]
        return out

    def save(self):
        """bullet.storage.save() (src/bullet-storage.js:169-171 -> BulletFileStorage._saveData)."""
        self.rt.method(self.bullet.get("storage"), "save")

    # ---- driving (every call goes through the reference's own public entry points)
    def setData(self, path, data, broadcast=True):
        self.rt.method(self.bullet, "setData", path, from_py(data), bool(broadcast))

    def put(self, path, data):
        node = self.rt.method(self.bullet, "get", path)
        self.rt.method(node, "put", from_py(data))

    def process_sync_entries(self, entries, peer_id="remote"):
        self.rt.method(self._sync, "_processSyncEntries", from_py(entries), peer_id)

    def handle_put(self, path, data):
        """BulletNetwork._handlePut's payload shape (src/bullet-network.js:332-346); the network class itself
        needs sockets, so only its three lines of wrapping are restated here."""
        if isinstance(data, dict):
            data = {**data, "__fromNetwork": True}
        self.setData(path, data, False)

    def on(self, path, cb):
        f = I.JSFunction("listener", lambda this, a: cb(to_py(a[0]) if a else UNDEFINED) or UNDEFINED)
        node = self.rt.method(self.bullet, "get", path)
        self.rt.method(node, "on", f)

    def index(self, path, field=None):
        self.rt.method(self.bullet, "index", path, UNDEFINED if field is None else field)
        return self

    def _paths(self, nodes):
        return [n.get("path") for n in nodes.items]

    def equals(self, path, field, value):
        return self._paths(self.rt.method(self.bullet, "equals", path, field, from_py(value)))

    def range(self, path, field, mn, mx):
        return self._paths(self.rt.method(self.bullet, "range", path, field, from_py(mn), from_py(mx)))

    def count(self, path, field, value):
        # Bullet has no count(); the documented call is bullet.query.count (examples/bullet-query-example.js:94)
        return int(self.rt.method(self.bullet.get("query"), "count", path, field, from_py(value)))

    def resolve(self, key, incoming_clock, current_clock, incoming_value, current_value):
        r = self.rt.method(self.bullet.get("crt"), "resolve", key, from_py(incoming_clock), from_py(current_clock),
                           from_py(incoming_value), from_py(current_value))
        return to_py(r)


def jsonable(v):
    """Python JS-value -> JSON-safe form that keeps NaN / +-Infinity / -0 / undefined distinct."""
    if isinstance(v, dict):
        return {"%o": [[k, jsonable(x)] for k, x in v.items()]}
    if isinstance(v, list):
        return [jsonable(x) for x in v]
    if v is UNDEFINED:
        return {"%u": 1}
    if isinstance(v, bool) or v is None or isinstance(v, str):
        return v
    if isinstance(v, (int, float)):
        v = float(v)
        if v != v:
            return {"%n": "NaN"}
        if v in (math.inf, -math.inf):
            return {"%n": "Infinity" if v > 0 else "-Infinity"}
        if v == 0 and math.copysign(1, v) < 0:
            return {"%n": "-0"}
        return v
    raise TypeError(type(v))


def unjsonable(v):
    if isinstance(v, dict):
        if "%o" in v:
            return {k: unjsonable(x) for k, x in v["%o"]}
        if "%u" in v:
            return UNDEFINED
        if "%n" in v:
            return {"NaN": math.nan, "Infinity": math.inf, "-Infinity": -math.inf, "-0": -0.0}[v["%n"]]
        raise ValueError(v)
    if isinstance(v, list):
        return [unjsonable(x) for x in v]
    if isinstance(v, bool) or v is None or isinstance(v, str):
        return v
    return float(v)
