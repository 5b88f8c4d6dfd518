"""Literal oracle: a CPU restatement of the reference's hot path over JS-like
Python values (TEST INFRASTRUCTURE ONLY - never imported by the product path).

PARITY PIN: KORandi/bullet-js ships no tests, fixtures or golden vectors
and no JS engine exists in this image, so the reference's own source files are executed
unmodified by `oracle/minijs` (an ECMAScript-subset interpreter written for the purpose,
`oracle/ref_runner.py`) and the traces are committed under tests/golden/
(tests/golden/make_golden.py).  tests/test_golden.py and tests/test_minijs.py hold this
restatement to them: decisions, change sets, store, both clock maps and their aliasing,
index Maps in exact (Map, Set) order and query results, on the SURVEY 8c scenarios and
on random streams that take every branch of `resolve`.  Caveat: the interpreter is ours,
not V8 (its semantics are unit-tested, not compared with node).  The hand-derived
known-answer traces of SURVEY.md 8c remain in tests/test_oracle_kat.py.

Every method names the reference lines it follows (paths relative to
/root/reference).  Objects are Python dicts (insertion ordered == JS own-key
order for non-integer-like keys), so clock key order, the `crt.vectorClocks`
aliasing of `meta[path].vectorClock` and `_getData`'s falsy materialisation are
all reproduced by construction rather than modelled.
"""
from __future__ import annotations

import copy
from typing import Any, Callable

from .jsvalue import (
    UNDEFINED,
    JSTypeError,
    get_prop,
    greater_equal,
    json_stringify,
    less_equal,
    less_than,
    norm,
    strict_equals,
    to_number,
    to_string,
    truthy,
    typeof,
)

REASON_NO_CURRENT = "no current state"
REASON_IDENTICAL = "identical clocks and values"
REASON_VALUE = "identical clocks, decided by value comparison"
REASON_INCOMING = "incoming vector clock dominates"
REASON_HISTORICAL = "current vector clock dominates (incoming is historical)"
REASON_CONCURRENT = "concurrent modifications, merged objects"

# 7 decision codes: the value-comparison reason splits by sign.
CODE_NO_CURRENT = 0
CODE_IDENTICAL = 1
CODE_TIE_INCOMING = 2
CODE_TIE_CURRENT = 3
CODE_INCOMING = 4
CODE_HISTORICAL = 5
CODE_CONCURRENT = 6


def decision_code(d: dict) -> int:
    r = d["reason"]
    if r == REASON_NO_CURRENT:
        return CODE_NO_CURRENT
    if r == REASON_IDENTICAL:
        return CODE_IDENTICAL
    if r == REASON_VALUE:
        return CODE_TIE_INCOMING if d["incoming"] else CODE_TIE_CURRENT
    if r == REASON_INCOMING:
        return CODE_INCOMING
    if r == REASON_HISTORICAL:
        return CODE_HISTORICAL
    if r == REASON_CONCURRENT:
        return CODE_CONCURRENT
    raise ValueError(r)


class RefCRT:
    """src/bullet-crt.js (BulletCRT)."""

    def __init__(self, bullet: "RefBullet"):
        self.bullet = bullet
        self.vectorClocks: dict[str, dict] = {}  # crt:8 (a Map)
        self.compare: Callable[[Any, Any], int] = self._default_compare

    @staticmethod
    def _default_compare(incoming, existing) -> int:
        """crt:11-15."""
        if strict_equals(incoming, existing):
            return 0
        if less_than(incoming, existing):
            return -1
        return 1

    def createVectorClock(self, key):
        """crt:33-37."""
        clock = {self.bullet.id: 1.0}
        self.vectorClocks[key] = clock
        return clock

    def getVectorClock(self, key):
        """crt:44-49."""
        if key not in self.vectorClocks:
            return self.createVectorClock(key)
        return self.vectorClocks[key]

    def incrementVectorClock(self, key):
        """crt:56-60 (in place: aliases see the increment)."""
        clock = self.getVectorClock(key)
        me = self.bullet.id
        cur = clock.get(me, UNDEFINED)
        clock[me] = (cur if truthy(cur) else 0.0) + 1.0
        return clock

    def compareVectorClocks(self, clock1, clock2) -> int:
        """crt:68-95."""
        if not truthy(clock1):
            return -1
        if not truthy(clock2):
            return 1
        d1 = d2 = False
        nodes = list(dict.fromkeys(list(clock1.keys()) + list(clock2.keys())))
        for node in nodes:
            v1 = clock1.get(node, UNDEFINED)
            v1 = v1 if truthy(v1) else 0.0
            v2 = clock2.get(node, UNDEFINED)
            v2 = v2 if truthy(v2) else 0.0
            if v1 > v2:
                d1 = True
            elif v2 > v1:
                d2 = True
            if d1 and d2:
                return 0
        if d1:
            return 1
        if d2:
            return -1
        return 0

    def mergeVectorClocks(self, clock1, clock2):
        """crt:103-114."""
        if not truthy(clock1):
            return dict(clock2)
        if not truthy(clock2):
            return dict(clock1)
        result = dict(clock1)
        for node, value in clock2.items():
            r = result.get(node, UNDEFINED)
            result[node] = max(r if truthy(r) else 0.0, value)
        return result

    def mergeValues(self, incoming, current):
        """crt:122-153."""
        if (
            typeof(incoming) != "object"
            or typeof(current) != "object"
            or incoming is None
            or current is None
        ):
            return incoming if self.compare(incoming, current) >= 0 else current
        result = dict(current)
        for key, value in incoming.items():
            if key in result:
                result[key] = self.mergeValues(value, result[key])
            else:
                result[key] = value
        return result

    def resolve(self, key, incomingClock, currentClock, incomingValue, currentValue):
        """crt:164-279."""
        if not truthy(currentClock):
            clock = self.incrementVectorClock(key)
            return dict(
                defer=False, historical=False, converge=True, incoming=True, current=False,
                concurrent=False, vectorClock=clock, reason=REASON_NO_CURRENT, value=incomingValue,
            )
        comparison = self.compareVectorClocks(incomingClock, currentClock)
        merged = self.mergeVectorClocks(incomingClock, currentClock)
        self.vectorClocks[key] = merged  # crt:197
        if comparison == 0 and json_stringify(incomingClock) == json_stringify(currentClock):
            vc = self.compare(incomingValue, currentValue)
            if vc == 0:
                return dict(
                    defer=False, historical=False, converge=True, incoming=False, current=False,
                    concurrent=False, vectorClock=merged, reason=REASON_IDENTICAL, value=currentValue,
                )
            return dict(
                defer=False, historical=False, converge=True, incoming=vc > 0, current=vc < 0,
                concurrent=False, vectorClock=merged, reason=REASON_VALUE,
                value=incomingValue if vc > 0 else currentValue,
            )
        if comparison > 0:
            return dict(
                defer=False, historical=False, converge=True, incoming=True, current=False,
                concurrent=False, vectorClock=merged, reason=REASON_INCOMING, value=incomingValue,
            )
        if comparison < 0:
            return dict(
                defer=False, historical=True, converge=True, incoming=False, current=True,
                concurrent=False, vectorClock=merged, reason=REASON_HISTORICAL, value=currentValue,
            )
        mergedValue = self.mergeValues(incomingValue, currentValue)
        return dict(
            defer=False, historical=False, converge=True, incoming=False, current=False,
            concurrent=True, vectorClock=merged, reason=REASON_CONCURRENT, value=mergedValue,
        )

    def processUpdate(self, key, incomingValue, incomingClock, currentValue, currentClock):
        """crt:304-318."""
        d = self.resolve(key, incomingClock, currentClock, incomingValue, currentValue)
        return dict(value=d["value"], vectorClock=d["vectorClock"], decision=d)

    def handleUpdate(self, path, incomingData, isFromNetwork=False):
        """crt:329-385."""
        currentData = self.bullet._getData(path)
        currentMeta = self.bullet.meta.get(path) or {}
        currentClock = currentMeta.get("vectorClock", UNDEFINED)
        dataToStore = incomingData
        if (
            isFromNetwork
            and truthy(incomingData)
            and typeof(incomingData) == "object"
            and truthy(get_prop(incomingData, "__vectorClock"))
        ):
            incomingClock = incomingData["__vectorClock"]
            dataToStore = {k: v for k, v in incomingData.items() if k != "__vectorClock"}
        else:
            incomingClock = self.incrementVectorClock(path)
        result = self.resolve(path, incomingClock, currentClock, dataToStore, currentData)
        broadcastData = result["value"]
        if typeof(broadcastData) == "object" and broadcastData is not None:
            broadcastData = {**broadcastData, "__vectorClock": result["vectorClock"]}
        return dict(
            value=result["value"],
            vectorClock=result["vectorClock"],
            broadcastData=broadcastData,
            decision=result,
            doUpdate=bool(result["incoming"] or (not truthy(currentClock)) or result["concurrent"]),
        )


class RefQuery:
    """src/bullet-query.js (BulletQuery) - index / equals / range / count and the
    post-write hook.  Map -> dict, Set -> dict of path->True (both insertion ordered,
    delete + re-insert moves to the end exactly like JS)."""

    def __init__(self, bullet: "RefBullet"):
        self.bullet = bullet
        self.indices: dict[str, dict[str, dict[str, bool]]] = {}
        self.indexedPaths: dict[str, bool] = {}

    def index(self, path, field=None):
        """query:30-45."""
        indexKey = f"{path}:{field}" if truthy(field) else path
        if indexKey in self.indices:
            return self
        self.indices[indexKey] = {}
        self.indexedPaths[path] = True
        self._buildIndex(path, field)
        return self

    def _buildIndex(self, path, field):
        """query:53-73."""
        indexKey = f"{path}:{field}" if truthy(field) else path
        index = self.indices[indexKey]
        baseData = self.bullet._getData(path)
        if typeof(baseData) == "object" and baseData is not None:
            if truthy(field):
                for key, value in list(baseData.items()):
                    if typeof(value) == "object" and value is not None and field in value:
                        self._addToIndex(index, value[field], f"{path}/{key}")
            else:
                for key, value in list(baseData.items()):
                    self._addToIndex(index, value, f"{path}/{key}")

    def _addToIndex(self, index, value, nodePath):
        """query:82-94."""
        if value is None or value is UNDEFINED:
            return
        iv = self._getIndexableValue(value)
        if iv not in index:
            index[iv] = {}
        index[iv].setdefault(nodePath, True)

    def _removeFromIndex(self, index, value, nodePath):
        """query:103-118."""
        if value is None or value is UNDEFINED:
            return
        iv = self._getIndexableValue(value)
        if iv in index:
            paths = index[iv]
            paths.pop(nodePath, None)
            if len(paths) == 0:
                del index[iv]

    @staticmethod
    def _getIndexableValue(value) -> str:
        """query:126-131."""
        if typeof(value) == "object" and value is not None:
            return json_stringify(value)
        return to_string(value)

    def _updateIndices(self, path, newData):
        """query:139-176."""
        for indexedPath in list(self.indexedPaths):
            if path.startswith(indexedPath + "/"):
                relativePath = path[len(indexedPath) + 1:]
                parts = relativePath.split("/")
                for indexKey, index in list(self.indices.items()):
                    sp = indexKey.split(":")
                    basePath = sp[0]
                    field = sp[1] if len(sp) > 1 else UNDEFINED
                    if basePath != indexedPath:
                        continue
                    if truthy(field) and len(parts) == 1:
                        nodePath = f"{indexedPath}/{parts[0]}"
                        oldData = self.bullet._getData(nodePath)
                        if truthy(oldData) and truthy(get_prop(oldData, field)):
                            self._removeFromIndex(index, get_prop(oldData, field), nodePath)
                        if truthy(newData) and truthy(get_prop(newData, field)):
                            self._addToIndex(index, get_prop(newData, field), nodePath)
                    elif (not truthy(field)) and len(parts) == 1:
                        oldData = self.bullet._getData(path)
                        self._removeFromIndex(index, oldData, path)
                        self._addToIndex(index, newData, path)

    def equals(self, path, field, value):
        """query:186-210 (called with 3 args via src/bullet.js:332-334). Returns paths."""
        indexKey = f"{path}:{field}" if truthy(field) else path
        if indexKey not in self.indices:
            self.index(path, field)
        index = self.indices[indexKey]
        iv = self._getIndexableValue(value)
        return list(index[iv].keys()) if iv in index else []

    def range(self, path, field, mn, mx):
        """query:221-261. Returns paths in (Map order, Set order)."""
        indexKey = f"{path}:{field}" if truthy(field) else path
        if indexKey not in self.indices:
            self.index(path, field)
        index = self.indices[indexKey]
        results = []
        for indexValue, paths in index.items():
            value = to_number(indexValue)
            if value != value:
                value = indexValue
            if (
                mn is not UNDEFINED
                and greater_equal(value, mn)
                and mx is not UNDEFINED
                and less_equal(value, mx)
            ):
                results.extend(paths.keys())
        return results

    def count(self, path, field, value) -> int:
        """query:293-313."""
        indexKey = f"{path}:{field}" if truthy(field) else path
        if indexKey not in self.indices:
            self.index(path, field)
        index = self.indices[indexKey]
        iv = self._getIndexableValue(value)
        return len(index[iv]) if iv in index else 0


class RefBullet:
    """The slice of src/bullet.js that defines state and the change set:
    `_getData` 115-129, `setData` 139-155, `_stripNetworkFlag` 161-178,
    `_applyUpdate` 184-220, `_notify` 227-266, wrapped the way the constructor
    (37-64) wraps it: query hook (query:13-21) around middleware
    (src/bullet-middleware.js:70-135) around the prototype method.
    """

    def __init__(self, peer_id: str, enable_middleware=True, enable_indexing=True):
        self.id = peer_id
        self.store: dict = {}
        self.meta: dict = {}
        self.log: list = []
        self.listeners: dict[str, list] = {}
        self.enable_middleware = enable_middleware
        self.query = RefQuery(self) if enable_indexing else None
        self.crt = RefCRT(self)
        # instrumentation (not part of the reference): everything the parity tests compare
        self.changes: list[dict] = []    # ordered accepted updates == the emitted change set
        self.decisions: list[dict] = []  # one per setData call
        self.notifications: list[tuple] = []

    # ---- src/bullet.js:115-129
    def _getData(self, path):
        if not path:
            return self.store
        parts = [p for p in path.split("/") if p]
        current = self.store
        for part in parts:
            if not isinstance(current, dict):
                # property read on a truthy primitive gives undefined, then the
                # strict-mode assignment `current[part] = {}` throws.
                raise JSTypeError(f"Cannot create property '{part}' on {typeof(current)}")
            if not truthy(current.get(part, UNDEFINED)):
                current[part] = {}
            current = current[part]
        return current

    # ---- wrapper stack
    def setData(self, path, data, broadcast=True):
        if self.query is not None:
            self._setData_middleware(path, data, broadcast)
            self.query._updateIndices(path, data)  # query:16-20, raw argument
            return UNDEFINED
        return self._setData_middleware(path, data, broadcast)

    def _setData_middleware(self, path, data, broadcast):
        if not self.enable_middleware:
            return self._setData_proto(path, data, broadcast)
        self._getData(path)  # mw:108 oldData = originalGetData(path)
        self._setData_proto(path, data, broadcast)
        return True

    # ---- src/bullet.js:139-155
    def _setData_proto(self, path, rawData, broadcast=True):
        data, fromNetwork = self._stripNetworkFlag(rawData)
        r = self.crt.handleUpdate(path, data, fromNetwork)
        self.decisions.append(
            dict(path=path, code=decision_code(r["decision"]), reason=r["decision"]["reason"],
                 doUpdate=r["doUpdate"])
        )
        if not r["doUpdate"]:
            return r["value"]
        self._applyUpdate(path, r["value"], r["vectorClock"], fromNetwork)
        return r["value"]

    # ---- src/bullet.js:161-178
    @staticmethod
    def _stripNetworkFlag(inp):
        fromNetwork = False
        data = inp
        if truthy(inp) and typeof(inp) == "object" and truthy(get_prop(inp, "__fromNetwork")):
            fromNetwork = True
            data = {k: v for k, v in inp.items() if k != "__fromNetwork"}
        return data, fromNetwork

    # ---- src/bullet.js:184-220
    def _applyUpdate(self, path, value, vectorClock, fromNetwork):
        parts = [p for p in path.split("/") if p]
        node = self.store
        for part in parts[:-1]:
            if not isinstance(node, dict):
                raise JSTypeError(f"Cannot create property '{part}' on {typeof(node)}")
            if not truthy(node.get(part, UNDEFINED)):
                node[part] = {}
            node = node[part]
        if not parts:
            return
        key = parts[-1]
        if not isinstance(node, dict):
            raise JSTypeError(f"Cannot create property '{key}' on {typeof(node)}")
        node[key] = value
        self.meta[path] = {
            **(self.meta.get(path) or {}),
            "source": "network" if fromNetwork else "local",
            "vectorClock": vectorClock,  # the same object crt.vectorClocks holds
        }
        entry = dict(op="set", path=path, data=value, vectorClock=vectorClock)
        self.log.append(entry)
        if len(self.log) > 1000:
            del self.log[: len(self.log) - 1000]
        self.changes.append(
            dict(seq=len(self.decisions) - 1, path=path, value=copy.deepcopy(value),
                 vectorClock=dict(vectorClock), fromNetwork=fromNetwork)
        )
        self._notify(path, value)

    # ---- src/bullet.js:227-266 (listener fan-out only)
    def _notify(self, path, data):
        for cb in self.listeners.get(path, []):
            cb(data)
        parts = [p for p in path.split("/") if p]
        while parts:
            parts.pop()
            parent = "/".join(parts)
            if parent in self.listeners:
                pdata = self._getData(parent)
                for cb in self.listeners[parent]:
                    cb(pdata)

    # ---- BulletNode surface (src/bullet.js:681-759) flattened
    def put(self, path, data):
        self.setData(path, norm(data))

    def on(self, path, cb):
        self.listeners.setdefault(path, []).append(cb)
        cb(self._getData(path))

    # ---- ingress shapes
    def process_sync_entries(self, entries):
        """src/bullet-network-sync.js:551-569."""
        for entry in entries:
            path, data = entry["path"], norm(entry.get("data"))
            if entry.get("deleted"):
                self.setData(path, None, False)
                continue
            if typeof(data) == "object" and data is not None:
                networkData = {**data, "__fromNetwork": True,
                               "__vectorClock": norm_clock(entry.get("vectorClock", UNDEFINED))}
            else:
                networkData = data
            self.setData(path, networkData, False)

    def handle_put(self, path, data):
        """src/bullet-network.js:332-346 (payload shape only)."""
        data = norm(data)
        if typeof(data) == "object" and data is not None:
            data = {**data, "__fromNetwork": True}
        self.setData(path, data, False)

    # ---- query facade (src/bullet.js:313-357)
    def index(self, path, field=None):
        self.query.index(path, field)
        return self

    def equals(self, path, field, value):
        return self.query.equals(path, field, norm(value))

    def range(self, path, field, mn, mx):
        return self.query.range(path, field, norm(mn), norm(mx))

    def count(self, path, field, value):
        return self.query.count(path, field, norm(value))


def norm_clock(c):
    if c is UNDEFINED or c is None:
        return c
    return {k: float(v) for k, v in c.items()}
