"""js/bullet-b200.js - the reference-side shim - executed for real: the reference's own `Bullet` (src/*.js,
unmodified) runs in oracle/minijs with the shim installed on it, the shim's `native` addon bridged to a typed
engine (tests/js_bridge.py).  Everything around the decision stays the reference's code (setData, middleware,
query hook, _applyUpdate, listeners), so the JS-visible outcome must equal the pure reference's golden traces."""
import os

import pytest

from oracle import ref_runner
from oracle.minijs import interp as I
from oracle.minijs.builtins import Runtime, from_py, to_py
from oracle.ref_runner import unjsonable
from oracle.typed import TypedOracle
from tests import golden_io
from tests.golden_io import clock_items, same_js
from tests.js_bridge import NativeBridge
from tests.test_oracle_typed import make_cfg

pytestmark = pytest.mark.skipif(not ref_runner.available(), reason="reference sources not present")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STREAMS = golden_io.load("streams.json.gz")["cases"]

HARNESS = r"""
const bullet = new Bullet({ disableNetwork: true, server: false, storage: true, storageType: "memory",
                            enableIndexing: indexed });
bullet.id = "p0";
const shim = new BulletB200(bullet, native, { capacity: 64, postGetData: indexed });
const changes = [];
const origApply = bullet._applyUpdate;
bullet._applyUpdate = function (path, value, vectorClock, fromNetwork) {
  changes.push({ path: path, value: snapshot(value), vectorClock: snapshot(vectorClock), fromNetwork: !!fromNetwork });
  return origApply.call(bullet, path, value, vectorClock, fromNetwork);
};
const heard = [];
bullet.get("users").on((v) => heard.push(Object.keys(v).length));
return { bullet: bullet, shim: shim, changes: changes, heard: heard };
"""


def boot(indexed, make_engine):
    rt = Runtime(console=[])
    ref = ref_runner._reference_root()
    bridge = NativeBridge(make_engine)
    r = rt.eval(HARNESS, Bullet=rt.require(os.path.join(ref, "src", "bullet.js")),
                BulletB200=rt.require(os.path.join(ROOT, "js", "bullet-b200.js")), native=bridge.js_object(),
                indexed=indexed, snapshot=I.JSFunction("snapshot", lambda this, a: from_py(to_py(a[0]))))
    return rt, bridge, r


def oracle_engine(schema, capacity, post_getdata):
    return TypedOracle(make_cfg(schema, capacity, post_getdata))


@pytest.mark.parametrize("k", [0, 1, 10])
def test_reference_with_shim_equals_reference(k):
    case = STREAMS[k]
    indexed = bool(case["index_fields"]) or bool(case["late_index"])
    rt, bridge, r = boot(indexed, oracle_engine)
    bullet, shim = r.get("bullet"), r.get("shim")
    for f in case["index_fields"]:
        rt.method(bullet, "index", "users", f)
    ops = golden_io.ops_of(case)
    pending = []

    def flush():
        if pending:
            rt.method(shim, "processSyncEntries", from_py(pending))  # the batched ingress: one device call
            pending.clear()

    for i, (path, value, clock) in enumerate(ops):
        for f, at in case["late_index"].items():
            if at == i:
                flush()
                rt.method(bullet, "index", "users", f)
        if clock is not None and isinstance(value, dict):
            pending.append(dict(path=path, data=value, vectorClock=clock))
            if len(pending) == 50:  # src/bullet-network-sync.js: chunkSize
                flush()
        else:
            flush()
            rt.method(rt.method(bullet, "get", path), "put", from_py(value))  # BulletNode.put -> setData -> crt.handleUpdate
    flush()
    assert "".join(map(str, bridge.codes)) == case["codes"]
    assert max(bridge.batches) > 1 and len(bridge.batches) < len(ops)  # the sync ingress really was batched
    changes = to_py(r.get("changes"))
    assert len(changes) == len(case["changes"])
    for got, (_seq, path, value, clock, from_net) in zip(changes, case["changes"]):
        assert got["path"] == path and got["fromNetwork"] == from_net
        assert same_js(got["value"], unjsonable(value)) and clock_items(got["vectorClock"]) == clock
    assert same_js(to_py(bullet.get("store")), unjsonable(case["store"]))
    meta = to_py(bullet.get("meta"))
    assert [[p, m["source"], clock_items(m["vectorClock"])] for p, m in meta.items()] == case["meta"]
    # crt clocks live on the device: read through the shim
    for path, clock in case["vclocks"]:
        assert clock_items(to_py(rt.method(bullet.get("crt"), "getVectorClock", path))) == clock
    if indexed:  # the stock BulletQuery keeps working on the store the reference's own _applyUpdate maintains
        q = bullet.get("query")
        dump = {key: [[bk, [p for p, _ in s.data.values()]] for bk, s in q.get("indices").get(key).data.values()]
                for key in q.get("indices").enumerable_keys()}
        assert dump == case["index"]
    assert len(to_py(r.get("heard"))) == len(changes) + 1  # the `users` listener: once at subscription, then per change


def test_shim_is_plain_commonjs():
    src = open(os.path.join(ROOT, "js", "bullet-b200.js")).read()
    assert "module.exports = BulletB200" in src and "require(" not in src.split("*/", 1)[1]
