"""js/bullet-b200.js - the reference-side shim - executed for real: the reference's own `Bullet` (src/*.js,
unmodified) runs in oracle/minijs with the shim installed on it, the shim's `native` addon bridged to a typed
engine (tests/js_bridge.py).  Everything around the decision stays the reference's code (setData, middleware,
query hook, _applyUpdate, listeners), so the JS-visible outcome must equal the pure reference's golden traces."""
import os

import numpy as np
import pytest

from oracle import ref_runner
from oracle.minijs import interp as I
from oracle.minijs.builtins import Runtime, from_py, to_py
from oracle.ref_runner import unjsonable
from oracle.typed import TypedOracle
from tests import golden_io, streamgen
from tests.golden_io import clock_items, same_js
from tests.js_bridge import NativeBridge

pytestmark = pytest.mark.skipif(not ref_runner.available(), reason="reference sources not present")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STREAMS = golden_io.load("streams.json.gz")["cases"]

HARNESS = r"""
const bullet = new Bullet({ disableNetwork: true, server: false, storage: true, storageType: "memory",
                            enableIndexing: indexed });
bullet.id = "p0";
const shim = new BulletB200(bullet, native, { capacity: 64, postGetData: indexed, fields: fields, peers: peers, strings: strings,
                                                deviceQueries: deviceQueries ? "users" : null, exactOrder: exactOrder });
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


def boot(indexed, make_engine, device_queries=False, exact_order=False):
    rt = Runtime(console=[])
    ref = ref_runner._reference_root()
    bridge = NativeBridge(rt, make_engine)
    r = rt.eval(HARNESS, Bullet=rt.require(os.path.join(ref, "src", "bullet.js")),
                BulletB200=rt.require(os.path.join(ROOT, "js", "bullet-b200.js")), native=bridge.js_object(),
                indexed=indexed, deviceQueries=device_queries, exactOrder=exact_order, fields=from_py(streamgen.FIELDS), peers=from_py(streamgen.PEERS),
                strings=from_py(streamgen.STRINGS), snapshot=I.JSFunction("snapshot", lambda this, a: from_py(to_py(a[0]))))
    return rt, bridge, r


def oracle_engine(cfg):
    return TypedOracle(cfg)


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


def cuda_engine(cfg):
    """The CUDA library itself behind the shim's addon surface (bb_create with the flags the SHIM asked for)."""
    from bullet_js_b200 import codec
    from bullet_js_b200.engine import Engine

    assert cfg.flags & codec.CFG_EXACT_ORDER and cfg.flags & codec.CFG_POST_GETDATA  # js/bullet-b200.js set them
    return Engine(int(cfg.capacity), local_peer=int(cfg.local_peer), n_fields=int(cfg.n_fields), post_getdata=True,
                  exact_order=True, rank_object=int(cfg.rank_object), rank_true=int(cfg.rank_true),
                  rank_false=int(cfg.rank_false), rank_nan=int(cfg.rank_nan))


@pytest.mark.gpu
def test_device_queries_through_the_shim_on_the_cuda_library():
    """The same flow with libbulletb200.so behind the shim and `exactOrder: true` (BB_CFG_EXACT_ORDER): the unmodified
    reference + js/bullet-b200.js + the CUDA kernels return the reference's own recorded result lists, in order.  Needs
    the reference sources, so it runs where they are (this container, kernels emulated: BB_EMU_TESTS=1) and is skipped
    on the GPU boxes, where /root/reference does not exist."""
    run_device_queries(cuda_engine, exact_order=True)


def test_device_queries_through_the_shim():
    """options.deviceQueries: bullet.index / equals / range / count answered by the library's index (built and kept
    by the merge kernel's hook) - with the typed oracle behind the addon the results come in the reference's exact
    (Map, Set) order, so they must equal the golden query results; keys and bounds are computed in JavaScript."""
    run_device_queries(oracle_engine)


def run_device_queries(make_engine, exact_order=False):
    import itertools

    from tests.test_oracle_query import BOUNDS, EQ_VALUES

    case = STREAMS[1]
    rt, bridge, r = boot(True, make_engine, device_queries=True, exact_order=exact_order)
    bullet, shim = r.get("bullet"), r.get("shim")
    for f in case["index_fields"]:
        rt.method(bullet, "index", "users", f)
    ops = golden_io.ops_of(case)
    for i, (path, value, clock) in enumerate(ops):
        for f, at in case["late_index"].items():
            if at == i:
                rt.method(bullet, "index", "users", f)
        if clock is not None and isinstance(value, dict):
            rt.method(shim, "processSyncEntries", from_py([dict(path=path, data=value, vectorClock=clock)]))
        else:
            rt.method(rt.method(bullet, "get", path), "put", from_py(value))
    assert "".join(map(str, bridge.codes)) == case["codes"]
    assert not bullet.get("query").get("indices").enumerable_keys()  # no JS-side index was ever built
    paths = lambda nodes: [n.get("path") for n in nodes.items]  # noqa: E731
    for name, q in case["queries"].items():
        for v, want, cnt in zip(EQ_VALUES, q["equals"], q["count"]):
            assert paths(rt.method(bullet, "equals", "users", name, from_py(v))) == want, (name, v)
            assert rt.method(bullet.get("query"), "count", "users", name, from_py(v)) == float(cnt)
        for (lo, hi), want in zip(itertools.product(BOUNDS, BOUNDS), q["range"]):
            assert paths(rt.method(bullet, "range", "users", name, from_py(lo), from_py(hi))) == want, (name, lo, hi)
    assert paths(rt.method(bullet, "range", "users", "age", 0.0)) == []  # max undefined


def test_js_keys_and_bounds_equal_python_codec():
    from bullet_js_b200 import codec
    from tests.test_oracle_query import BOUNDS, EQ_VALUES

    rt = Runtime(console=[])
    pack = rt.require(os.path.join(ROOT, "js", "pack.js"))
    sj = rt.new(pack.get("Schema"), from_py(dict(fields=streamgen.FIELDS, peers=streamgen.PEERS, strings=streamgen.STRINGS,
                                                 localPeer="p0")))
    schema = streamgen.make_schema()
    for v in EQ_VALUES + [{"a": 1.0}]:
        got = to_py(rt.method(sj, "indexKey", from_py(v)))
        want = schema.index_key(v)
        assert (None if got is None else int(got[0]) | (int(got[1]) << 32)) == want, v
    for v in BOUNDS:
        for upper in (False, True):
            got = to_py(rt.method(sj, "bound", from_py(v), upper))
            want = schema.bound(v, upper)
            same_num = got["num"] == float(want["num"]) or (got["num"] != got["num"] and float(want["num"]) != float(want["num"]))
            assert same_num and got["rank"] == float(want["rank"]) and got["flags"] == float(want["flags"]), (v, upper, got, want)


@pytest.mark.parametrize("k", [1, 11])
def test_js_packer_equals_python_codec(k):
    """js/pack.js fills the bb_batch buffers byte for byte like bullet_js_b200/codec.py (values incl. NaN / -0 /
    +-Infinity, key orders, clocks and their key order, flavours, interned path ids), and its bb_config ranks and
    string dictionary are the codec's."""
    from bullet_js_b200 import codec
    from tests.js_bridge import NativeBridge

    case = STREAMS[k]
    ops = golden_io.ops_of(case)
    rt = Runtime(console=[])
    pack = rt.require(os.path.join(ROOT, "js", "pack.js"))
    schema_js = rt.new(pack.get("Schema"), from_py(dict(fields=streamgen.FIELDS, peers=streamgen.PEERS, strings=streamgen.STRINGS,
                                                        localPeer="p0")))
    entries = [dict(path=p, data=v, vectorClock=c, local=not (c is not None and isinstance(v, dict))) if c is not None
               else dict(path=p, data=v, local=True) for p, v, c in ops]
    b = rt.call(pack.get("packEntries"), None, schema_js, from_py(entries))
    got = NativeBridge.batch_of(len(ops), b.get("pathId"), b.get("head"), b.get("clk"), b.get("val"))
    schema = streamgen.make_schema()
    want = codec.encode_updates(schema, ops)
    want.head["user"] = np.arange(len(ops), dtype=np.uint32)  # pack.js stamps the arrival index into `user`
    assert np.array_equal(got.path_id, want.path_id) and np.array_equal(got.head, want.head)
    assert np.array_equal(got.clk, want.clk) and np.array_equal(got.val, want.val)
    assert to_py(schema_js.get("strings")) == schema.strings.strings
    ranks = to_py(rt.method(schema_js, "ranks"))
    assert {"rank_object": ranks["rankObject"], "rank_true": ranks["rankTrue"], "rank_false": ranks["rankFalse"],
            "rank_nan": ranks["rankNaN"]} == schema.config_ranks()
    with pytest.raises(Exception) as e:
        rt.call(pack.get("packEntries"), None, schema_js, from_py([dict(path="users/x", data={"age": {"deep": 1.0}}, local=True)]))
    assert "DomainError" in str(e.value)


def test_shim_is_plain_commonjs():
    src = open(os.path.join(ROOT, "js", "bullet-b200.js")).read()
    assert "module.exports = BulletB200" in src and src.split("*/", 1)[1].count("require(") == 1  # only ./pack


def test_quick_start_flow_with_the_shim():
    """docs/quick-start.md:183-209 (index first, then puts, then equals) on an unmodified Bullet with the shim and the
    device index: the documented answer, plus listeners, node.value() and a whole-collection read from the
    reference's own store mirror."""
    rt = Runtime(console=[])
    ref = ref_runner._reference_root()
    bridge = NativeBridge(rt, oracle_engine)
    out = rt.eval(r"""
        const bullet = new Bullet({ disableNetwork: true, server: false, storageType: "memory" });
        const shim = new BulletB200(bullet, native, {
          capacity: 16, fields: ["name", "email", "role"], peers: [bullet.id], postGetData: true,
          strings: ["Alice", "Bob", "alice@example.com", "bob@example.com", "admin", "user"], deviceQueries: "users" });
        const seen = [];
        bullet.get("users/alice").on((v) => seen.push(v && v.role));
        bullet.index("users", "role");
        bullet.get("users/alice").put({ name: "Alice", email: "alice@example.com", role: "admin" });
        bullet.get("users/bob").put({ name: "Bob", email: "bob@example.com", role: "user" });
        const admins = bullet.equals("users", "role", "admin");
        return { admins: admins.map((n) => n.path), value: admins[0].value(), count: bullet.query.count("users", "role", "user"),
                 all: Object.keys(bullet.get("users").value()), seen: seen, calls: shim.calls,
                 clock: bullet.meta["users/alice"].vectorClock, crt: bullet.crt.getVectorClock("users/alice") };
    """, Bullet=rt.require(os.path.join(ref, "src", "bullet.js")), BulletB200=rt.require(os.path.join(ROOT, "js", "bullet-b200.js")),
        native=bridge.js_object())
    r = to_py(out)
    assert r["admins"] == ["users/alice"] and r["value"] == {"name": "Alice", "email": "alice@example.com", "role": "admin"}
    assert r["count"] == 1.0 and r["all"] == ["alice", "bob"] and r["seen"][-1] == "admin" and r["calls"] == 2.0
    assert list(r["clock"].values()) == [3.0] and r["clock"] == r["crt"]  # a first local write: {me: 3}, aliased


FLAG_HARNESS = r"""
const seen = [];
const h = bullet.crt.handleUpdate;
bullet.crt.handleUpdate = function (path, data, fromNetwork) {
  const r = h.call(bullet.crt, path, data, fromNetwork);
  const d = r.decision;
  seen.push({ reason: d.reason, incoming: !!d.incoming, current: !!d.current, concurrent: !!d.concurrent,
              historical: !!d.historical, converge: !!d.converge, defer: !!d.defer, doUpdate: !!r.doUpdate });
  return r;
};
return seen;
"""


@pytest.mark.parametrize("k", [0, 3, 7])
def test_shim_decision_flags_equal_the_reference(k):
    """`handleUpdate(...).decision` is public surface: every flag of every update - incoming, current,
    concurrent, historical, converge, defer - and doUpdate must be what the unmodified reference returns."""
    case = STREAMS[k]
    ops = golden_io.ops_of(case)
    ref = ref_runner.JSRefBullet("p0", enable_indexing=False)
    rt, bridge, r = boot(False, oracle_engine)
    bullet = r.get("bullet")
    seen = rt.eval(FLAG_HARNESS, bullet=bullet)
    for path, value, clock in ops:
        if clock is not None and isinstance(value, dict):  # the sync ingress, entry by entry (sync:560-566)
            ref.handle_put(path, {**value, "__vectorClock": dict(clock)})
            rt.method(bullet, "setData", path, from_py({**value, "__fromNetwork": True, "__vectorClock": dict(clock)}), False)
        else:
            ref.put(path, value)
            rt.method(rt.method(bullet, "get", path), "put", from_py(value))
    want, got = ref.decision_flags, to_py(seen)
    assert len(want) == len(got) == len(ops)
    assert got == want
    assert {d["reason"] for d in want} >= {"no current state", "current vector clock dominates (incoming is historical)"}
    assert any(d["concurrent"] for d in want) and all(d["converge"] for d in want)
