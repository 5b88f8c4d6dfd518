"""The host mirror of the reference API (bullet_js_b200/bullet.py) on the GPU, driven with the
reference's own example scripts and the known-answer traces of SURVEY.md 8c; every run is
mirrored on the literal oracle and compared."""
import pytest

from bullet_js_b200 import codec
from oracle.js_literal import RefBullet
from tests.test_oracle_kat import PRODUCTS, USERS

pytestmark = pytest.mark.gpu


def js(v):
    """Python ints of the fixtures -> JS numbers."""
    if isinstance(v, dict):
        return {k: js(x) for k, x in v.items()}
    return float(v) if isinstance(v, int) and not isinstance(v, bool) else v


def schemas(me="me"):
    ustr = [u["name"] for u in USERS.values()] + ["admin", "user", "editor"]
    pstr = [p["name"] for p in PRODUCTS.values()] + ["electronics", "accessories", "furniture"]
    peers = [me, "peerA", "peerB"]
    return {
        "users": codec.Schema(["name", "age", "active", "role"], peers, codec.StringDict(ustr), me),
        "products": codec.Schema(["name", "price", "stock", "category"], peers, codec.StringDict(pstr), me),
    }


def paths(nodes):
    return sorted(n.path for n in nodes)


def test_query_example_script():
    """examples/bullet-query-example.js:17-139 (KAT-Q1), same calls on both sides."""
    from bullet_js_b200.bullet import Bullet

    db, ref = Bullet(schemas(), capacity=64), RefBullet("me")
    for k, v in USERS.items():
        db.get(f"users/{k}").put(js(v))
        ref.put(f"users/{k}", js(v))
    for k, v in PRODUCTS.items():
        db.get(f"products/{k}").put(js(v))
        ref.put(f"products/{k}", js(v))
    db.index("users", "role").index("users", "age").index("users", "active")
    db.index("products", "category").index("products", "price")
    assert paths(db.equals("users", "role", "admin")) == ["users/user1", "users/user10", "users/user6"]
    assert paths(db.range("users", "age", 30, 40)) == sorted(["users/user2", "users/user5", "users/user8", "users/user10"])
    assert [db.count("users", "role", r) for r in ("admin", "user", "editor")] == [3, 5, 2]
    assert paths(db.range("products", "price", 100, 300)) == sorted(
        ["products/prod3", "products/prod6", "products/prod7", "products/prod9"])
    for args in (("users", "active", True), ("users", "active", "true"), ("users", "age", "28"),
                 ("products", "category", "furniture"), ("products", "stock", 5), ("users", "role", "nobody")):
        assert paths(db.equals(*args)) == sorted(ref.equals(*[js(a) for a in args])), args
    for args in (("users", "age", 0, 100), ("users", "age", "30", 40), ("users", "role", "a", "f"),
                 ("products", "price", None, 100), ("users", "active", "a", "z"), ("users", "age", 40, 30)):
        assert paths(db.range(*args)) == sorted(ref.range(*[js(a) for a in args])), args
    assert db.get("users/user3").value() == js(USERS["user3"])
    assert list(db._get_data("users").keys()) == list(USERS.keys())  # Object.entries order
    db.close()


def test_index_before_puts():
    """docs/quick-start.md:183-209 (KAT-Q2): the index exists first, the hook fills it."""
    from bullet_js_b200.bullet import Bullet

    s = {"users": codec.Schema(["name", "email", "role"], ["me"],
                               codec.StringDict(["Alice", "Bob", "alice@example.com", "bob@example.com", "admin", "user"]), "me")}
    db = Bullet(s, capacity=16)
    db.index("users", "role")
    db.get("users/alice").put({"name": "Alice", "email": "alice@example.com", "role": "admin"})
    db.get("users/bob").put({"name": "Bob", "email": "bob@example.com", "role": "user"})
    assert paths(db.equals("users", "role", "admin")) == ["users/alice"]
    db.close()


def test_network_entries_and_listeners():
    """KAT-N through process_sync_entries on peer "B", with listeners and the log."""
    from bullet_js_b200.bullet import Bullet

    s = {"users": codec.Schema(["age", "role"], ["A", "B"], codec.StringDict(["user", "admin"]), "B")}
    db, ref = Bullet(s, capacity=8), RefBullet("B")
    seen, seen_ref = [], []
    db.get("users/u1").on(seen.append)      # called at once with {} (the read materialises the node)
    ref.on("users/u1", seen_ref.append)
    for x in (db, ref):
        x.index("users", "age")
        x.index("users", "role")
    entries = [
        {"path": "users/u1", "data": {"age": 30.0, "role": "user"}, "vectorClock": {"A": 3.0}},
        {"path": "users/u1", "data": {"age": 25.0, "role": "admin"}, "vectorClock": {"A": 4.0}},
        {"path": "users/u1", "data": {"age": 40.0}, "vectorClock": {"A": 5.0, "B": 2.0}},
        {"path": "users/u1", "data": {"age": 10.0}, "vectorClock": {"A": 4.0, "B": 2.0}},
        {"path": "users/u1", "data": {"age": 10.0}, "vectorClock": {"B": 2.0, "A": 5.0}},
        {"path": "users/u1", "data": {"age": 10.0}, "vectorClock": {"B": 2.0, "A": 5.0}},
    ]
    db.process_sync_entries(entries)
    ref.process_sync_entries(entries)
    assert db.decisions == [0, 6, 4, 5, 6, 2] == [d["code"] for d in ref.decisions]
    assert db.get("users/u1").value() == {"age": 10.0} == ref.store["users"]["u1"]
    assert db.meta("users/u1") == {"B": 2.0, "A": 5.0} and list(db.meta("users/u1")) == ["B", "A"]
    assert seen == seen_ref and len(seen) == 6          # 1 at subscription + 5 accepted; the rejected update notifies nobody
    assert [e["data"] for e in db.log] == [c["value"] for c in ref.changes]
    for v in (10.0, 25.0, 30.0, 40.0):
        assert paths(db.equals("users", "age", v)) == sorted(ref.equals("users", "age", v)), v
    assert paths(db.equals("users", "role", "admin")) == sorted(ref.equals("users", "role", "admin"))
    db.close()


def test_restart_export_sync_and_node_surface():
    """Cold start from a reference data directory, BulletNode.off / remove, export and the sync producer side
     through the mirror: mirrored on the literal oracle."""
    from bullet_js_b200 import persist
    from bullet_js_b200.bullet import Bullet

    store = {"users": {k: js(v) for k, v in USERS.items()}}
    meta = {f"users/{k}": {"source": "local", "vectorClock": {"peerA": float(i + 1)}} for i, k in enumerate(USERS)}
    db, ref = Bullet(schemas(), capacity=64), RefBullet("me")
    db.load_reference_state(store, meta)
    ref.store = {"users": {k: js(v) for k, v in USERS.items()}}
    ref.meta = {p: {"source": "local", "vectorClock": dict(m["vectorClock"])} for p, m in meta.items()}
    # with an index on the collection the reference's post-write hook re-reads the node (query:151,169): the
    # mirror's tables are created for that regime (a stored null becomes {} right after the write)
    db.index("users", "age")
    ref.index("users", "age")
    heard = []
    cb = heard.append
    db.get("users/user2").on(cb)
    for x in (db, ref):
        x.get("users/user2").put({"age": 36.0}) if x is db else x.put("users/user2", {"age": 36.0})
    db.get("users/user2").off(cb)
    db.get("users/user3").remove()
    ref.put("users/user3", None)
    db.process_sync_entries([{"path": "users/user1", "data": {"age": 99.0}, "vectorClock": {"peerA": 5.0}}])
    ref.process_sync_entries([{"path": "users/user1", "data": {"age": 99.0}, "vectorClock": {"peerA": 5.0}}])
    assert db.decisions == [d["code"] for d in ref.decisions]
    assert len(heard) == 2  # at subscription and for the put; nothing after off()
    got_store, got_meta = db.export_reference_state()
    assert got_store["users"] == ref.store["users"] and list(got_store["users"]) == list(ref.store["users"])
    assert {p: m["vectorClock"] for p, m in got_meta.items()} == {p: m["vectorClock"] for p, m in ref.meta.items()}
    chunks = db.collect_sync_entries()
    want = persist.collect_full_sync_data(ref.store, ref.meta)
    assert [e for c in chunks for e in c] == want and all(len(c) <= 50 for c in chunks)
    db.close()


def test_device_sync_producer_with_since():
    """SURVEY 8f-2 on the device: bb_sync_collect (k_sync_collect) selects the rows _collectFullSyncData(since) visits -
    every record, and the primitives whose lastModified (the merge call that last wrote them, BB_CFG_TRACK_MODIFIED) is
    not older than `since` - and only those cross PCIe.  Expectation: persist.collect_full_sync_data (pinned to the
    live reference, `since` included: tests/test_persist.py) over the literal oracle's store with the same stamps."""
    import numpy as np

    from bullet_js_b200 import persist
    from bullet_js_b200.bullet import Bullet

    now = [1000.0]

    def clock():
        now[0] += 10.0
        return now[0]

    sch = schemas()
    sch["settings"] = codec.Schema(["v"], ["me", "peerA", "peerB"], codec.StringDict(["dark", "light", "en", "de"]), "me")
    db, ref = Bullet(sch, capacity=64, track_modified=True, clock=clock), RefBullet("me")
    stamps = {}

    def put(path, value):
        before = len(ref.changes)
        db.get(path).put(value)
        ref.put(path, value)
        if len(ref.changes) > before:
            stamps[path] = now[0]  # the mirror's Date.now() of this call

    for k, v in USERS.items():
        put(f"users/{k}", js(v))
    put("settings/theme", "dark")
    put("settings/lang", "en")
    mid = now[0] + 1.0
    put("users/user2", {"age": 36.0})
    put("settings/theme", "light")
    put("settings/theme", "dark")  # rejected ("dark" < "light"): lastModified stays
    assert db.decisions == [d["code"] for d in ref.decisions]
    meta = {p: dict(m, lastModified=stamps[p]) if p in stamps else dict(m) for p, m in ref.meta.items()}
    for since in (0, 1.0, mid, now[0] + 1.0):
        got = [e for c in db.collect_sync_entries(since) for e in c]
        want = persist.collect_full_sync_data(ref.store, meta, since)
        assert got == want, since
    # records filtered by their own lastModified as well (not what the reference does): only what changed since `mid`
    delta = [e["path"] for c in db.collect_sync_entries(mid, filter_records=True) for e in c]
    assert delta == ["users/user2/age", "settings/theme"]
    eng = db._c["users"].engine
    ids, rows, ep = eng.sync_collect(0)
    assert len(ids) == len(USERS) and ep.max() == eng.epoch and (np.diff(ids.astype(np.int64)) > 0).all()
    with pytest.raises(Exception):
        eng.sync_collect(0, cap=1)
    db.close()


def test_query_example_exact_order():
    """KAT-Q1 and KAT-H of SURVEY 8c as EXACT lists through the mirror with exact_order=True (BB_CFG_EXACT_ORDER): the
    example script's results in the reference's own order, the duplicate of a stale entry, and delete-then-add moving a
    path to the end of its bucket - every list against the literal oracle's."""
    from bullet_js_b200.bullet import Bullet

    db, ref = Bullet(schemas(), capacity=64, exact_order=True), RefBullet("me")
    for k, v in USERS.items():
        db.get(f"users/{k}").put(js(v))
        ref.put(f"users/{k}", js(v))
    for k, v in PRODUCTS.items():
        db.get(f"products/{k}").put(js(v))
        ref.put(f"products/{k}", js(v))
    db.index("users", "role").index("users", "age").index("users", "active")
    db.index("products", "category").index("products", "price")
    for x in ("role", "age", "active"):
        ref.index("users", x)
    for x in ("category", "price"):
        ref.index("products", x)
    order = lambda nodes: [n.path for n in nodes]
    assert order(db.equals("users", "role", "admin")) == ["users/user1", "users/user6", "users/user10"]
    assert order(db.range("users", "age", 30, 40)) == ["users/user2", "users/user5", "users/user8", "users/user10"]
    assert order(db.range("products", "price", 100, 300)) == ["products/prod3", "products/prod6", "products/prod7", "products/prod9"]
    # the hook path: stale entries (a node twice), buckets emptied and re-created at the end of the Map order
    for path, value in (("users/user2", {"age": 31.0}), ("users/user5", {"age": 35.0}), ("users/user2", {"age": 35.0}),
                        ("users/user8", {"role": "admin", "age": 39.0}), ("users/user1", {"role": "user"}),
                        ("users/user1", {"role": "admin"}), ("users/user3", {"age": 42.0, "role": "admin"})):
        db.get(path).put(js(value))
        ref.put(path, js(value))
    assert db.decisions == [d["code"] for d in ref.decisions]
    for args in (("users", "role", "admin"), ("users", "role", "user"), ("users", "age", 35), ("users", "age", 31),
                 ("users", "active", True)):
        assert order(db.equals(*args)) == ref.equals(*[js(a) for a in args]), args
    for args in (("users", "age", 0, 100), ("users", "age", 30, 40), ("users", "role", "a", "z"), ("users", "age", 31, 35)):
        assert order(db.range(*args)) == ref.range(*[js(a) for a in args]), args
    assert len(order(db.range("users", "age", 0, 100))) > len(USERS)  # stale entries: some node is listed twice
    db.close()
