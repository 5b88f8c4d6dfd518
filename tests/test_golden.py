"""Parity pinned to the reference itself: tests/golden/*.json.gz hold outputs of /root/reference/src/*.js
(executed unmodified by oracle/minijs, see tests/golden/make_golden.py).  Checked here against
  * the literal oracle (oracle/js_literal.py)            - CPU
  * the typed C oracle (oracle/bullet_oracle.c)          - CPU
  * libbulletb200.so through the C ABI                   - `-m gpu`
"""
import itertools

import numpy as np
import pytest

from bullet_js_b200 import capi, codec, synth
from oracle.js_literal import RefBullet
from oracle.jsvalue import UNDEFINED
from oracle.ref_runner import unjsonable
from oracle.typed import TypedOracle
from tests import golden_io, streamgen
from tests.golden_io import clock_items, same_js
from tests.test_oracle_query import BOUNDS, EQ_VALUES
from tests.test_oracle_typed import make_cfg

STREAMS = golden_io.load("streams.json.gz")
KAT = golden_io.load("kat.json.gz")
CONFIG1 = golden_io.load("config1.json.gz")
STREAM_IDS = [f"seed{c['seed']}" for c in STREAMS["cases"]]


def test_fixtures_name_the_reference():
    for fx in (STREAMS, KAT, CONFIG1):
        ref = fx["reference"]
        assert ref["name"].endswith("bullet-js") and len(ref["sha256"]) == 64 and "src/bullet-crt.js" in ref["files"]


# ----------------------------------------------------------------------------- literal oracle
def literal_state(ref):
    return dict(
        store=ref.store,
        meta=[[p, m.get("source"), clock_items(m["vectorClock"])] for p, m in ref.meta.items()],
        vclocks=[[p, clock_items(c)] for p, c in ref.crt.vectorClocks.items()],
        alias=[p for p, m in ref.meta.items() if m["vectorClock"] is ref.crt.vectorClocks.get(p)],
        index=None if ref.query is None else
        {k: [[bk, list(s)] for bk, s in idx.items()] for k, idx in ref.query.indices.items()},
    )


def assert_state(got, want):
    assert same_js(got["store"], unjsonable(want["store"]))
    assert got["meta"] == want["meta"]
    assert got["vclocks"] == want["vclocks"]
    assert got["alias"] == want["alias"]
    if "index" in want:
        assert got["index"] == want["index"]


def replay_literal(case):
    indexed = bool(case["index_fields"]) or bool(case["late_index"])
    ref = RefBullet("p0", enable_indexing=indexed)
    for f in case["index_fields"]:
        ref.index("users", f)
    for k, op in enumerate(golden_io.ops_of(case)):
        for f, at in case["late_index"].items():
            if at == k:
                ref.index("users", f)
        streamgen.apply_op(ref, op)
    return ref


@pytest.mark.parametrize("case", STREAMS["cases"], ids=STREAM_IDS)
def test_literal_oracle_equals_reference(case):
    ref = replay_literal(case)
    assert "".join(str(d["code"]) for d in ref.decisions) == case["codes"]
    assert "".join("1" if d["doUpdate"] else "0" for d in ref.decisions) == case["do_update"]
    assert len(ref.changes) == len(case["changes"])
    for got, (seq, path, value, clock, from_net) in zip(ref.changes, case["changes"]):
        assert (got["seq"], got["path"], got["fromNetwork"]) == (seq, path, from_net)
        assert same_js(got["value"], unjsonable(value))
        assert clock_items(got["vectorClock"]) == clock
    assert_state(literal_state(ref), case)
    for name, q in case.get("queries", {}).items():
        assert [ref.equals("users", name, v) for v in EQ_VALUES] == q["equals"]
        assert [ref.count("users", name, v) for v in EQ_VALUES] == q["count"]
        assert [ref.range("users", name, lo, hi) for lo, hi in itertools.product(BOUNDS, BOUNDS)] == q["range"]
        assert [ref.range("users", name, UNDEFINED, 5.0), ref.range("users", name, 0.0, UNDEFINED)] == q["range_undefined"]


def test_streams_take_every_branch_of_resolve():
    assert set("".join(c["codes"] for c in STREAMS["cases"])) == set("0123456")
    assert all(len(set(c["codes"])) >= 6 for c in STREAMS["cases"])


def _literal_for_kat(case):
    ref = RefBullet(case["peer"], enable_indexing=case["name"] != "KAT-L")
    for entry in case["steps"]:
        step = unjsonable(entry["step"])
        if step[0] == "put":
            ref.put(step[1], step[2])
        elif step[0] == "recv":
            ref.process_sync_entries([dict(path=step[1], data=step[2], vectorClock=step[3])])
        else:
            ref.index(step[1], step[2])
        if "code" in entry:
            d = ref.decisions[-1]
            assert (d["code"], d["reason"], d["doUpdate"]) == (entry["code"], entry["reason"], entry["doUpdate"])
        assert_state(literal_state(ref), entry)
    return ref


def test_kat_traces_equal_reference():
    by_name = {c["name"]: c for c in KAT["cases"]}
    _literal_for_kat(by_name["KAT-L"])
    _literal_for_kat(by_name["KAT-N"])
    h = _literal_for_kat(by_name["KAT-H"])
    assert h.range("users", "age", 30.0, 31.0) == by_name["KAT-H"]["range_30_31"] == ["users/u1", "users/u1"]
    # the hand-derived expectations of SURVEY 8c are what the reference really does
    assert [e["code"] for e in by_name["KAT-L"]["steps"]] == [0, 3, 4, 1, 4, 2, 3, 4, 2]
    assert [e["code"] for e in by_name["KAT-N"]["steps"][2:]] == [0, 6, 4, 5, 6, 2]
    ref = RefBullet("A")
    for r in by_name["KAT-R"]["resolve"]:
        key, inc, cur, x, y = unjsonable(r["args"])
        out = ref.crt.resolve(key, inc, cur, x, y)
        assert out["reason"] == r["reason"] and same_js(out["value"], unjsonable(r["value"]))
        assert clock_items(out["vectorClock"]) == r["vectorClock"]
    q1 = by_name["KAT-Q1"]
    assert q1["equals_role_admin"] == ["users/user1", "users/user6", "users/user10"]
    assert q1["range_age_30_40"] == ["users/user2", "users/user5", "users/user8", "users/user10"]
    assert q1["count_role"] == [3, 5, 2]
    assert q1["range_price_100_300"] == ["products/prod3", "products/prod6", "products/prod7", "products/prod9"]
    assert by_name["KAT-Q2"]["equals_role_admin"] == ["users/alice"]
    lines = by_name["examples/bullet-query-example.js"]["console"]
    assert "- Alice Johnson (ID: user1)" in lines and "- Harry Taylor, 39 years old" in lines
    assert any(line.strip() == "All query examples completed." for line in lines)


def test_documented_conflict_examples():
    """docs/conflict-resolution.md:446-483: the reference does what its documentation says, and the literal
    oracle (which models nested records too) does what the reference does."""
    import copy

    runs = {c["name"]: c for c in KAT["cases"]}["docs/conflict-resolution.md"]["runs"]
    theme, bob = runs
    assert unjsonable(theme["store_a"]) == unjsonable(theme["store_b"]) == {"settings": {"theme": "light"}}
    final = unjsonable(bob["store_a"])["users"]["bob"]
    assert final["name"] == "Robert Smith" and final["age"] == 30.0 and final["location"] == "New York"
    assert final["preferences"] == {"theme": "dark", "notifications": False}
    for run in runs:
        a, b = RefBullet("peerA", enable_indexing=False), RefBullet("peerB", enable_indexing=False)
        a.put(run["path"], copy.deepcopy(unjsonable(run["a"])))
        b.put(run["path"], copy.deepcopy(unjsonable(run["b"])))
        ca, cb = a.changes[-1], b.changes[-1]
        for dst, ch in ((b, ca), (a, cb)):
            v = copy.deepcopy(ch["value"])
            dst.handle_put(run["path"], {**v, "__vectorClock": dict(ch["vectorClock"])} if isinstance(v, dict) else v)
        assert same_js(a.store, unjsonable(run["store_a"])) and same_js(b.store, unjsonable(run["store_b"]))
        assert [d["reason"] for d in a.decisions] == run["reasons_a"] and [d["reason"] for d in b.decisions] == run["reasons_b"]


def test_kat_q1_literal_index_order():
    from tests.test_oracle_kat import PRODUCTS, USERS
    b = RefBullet("me")
    for k, v in USERS.items():
        b.put(f"users/{k}", v)
    for k, v in PRODUCTS.items():
        b.put(f"products/{k}", v)
    b.index("users", "role").index("users", "age").index("users", "active")
    b.index("products", "category").index("products", "price")
    want = {c["name"]: c for c in KAT["cases"]}["KAT-Q1"]["index"]
    assert {k: [[bk, list(s)] for bk, s in idx.items()] for k, idx in b.query.indices.items()} == want


# ----------------------------------------------------------------------------- typed engines (C oracle, GPU)
FIELD_SLOT = {"age": 0, "score": 1, "role": 2, "name": 3}


def typed_replay(case, make_engine):
    """Replay a golden stream through a typed engine (same surface for TypedOracle and Engine) in uneven
    batches; returns (schema, engine, decisions, decoded change set)."""
    ops = golden_io.ops_of(case)
    schema = streamgen.make_schema()
    batch = codec.encode_updates(schema, ops)
    indexed = bool(case["index_fields"]) or bool(case["late_index"])
    eng = make_engine(schema, indexed)
    for f in case["index_fields"]:
        eng.index_create(FIELD_SLOT[f])
    cuts = sorted({0, 1, 7, 300, 301, 900, len(ops)} | set(case["late_index"].values()))
    codes, changes = [], []
    for lo, hi in zip(cuts, cuts[1:]):
        for f, at in case["late_index"].items():
            if at == lo:
                eng.index_create(FIELD_SLOT[f])
        ch = eng.merge(batch.slice(lo, hi))
        codes.extend(ch.decision.tolist())
        sub = codec.decode_changes(schema, batch.slice(lo, hi), ch)
        for c in sub:
            c["seq"] += lo
        changes.extend(sub)
    return schema, eng, codes, changes


def assert_typed_matches(case, schema, codes, changes, rows_by_path):
    assert "".join(map(str, codes)) == case["codes"]
    assert len(changes) == len(case["changes"])
    for got, (seq, path, value, clock, _from_net) in zip(changes, case["changes"]):
        assert (got["seq"], got["path"]) == (seq, path)
        assert same_js(got["value"], unjsonable(value)), (got, value)
        assert clock_items(got["vectorClock"]) == clock
    users = unjsonable(case["store"]).get("users", {})
    meta = {p: c for p, _src, c in case["meta"]}
    vclocks = dict((p, c) for p, c in case["vclocks"])
    decoded = {p: codec.decode_row(schema, r) for p, r in rows_by_path.items()}
    for path, d in decoded.items():
        key = path.split("/")[1]
        if key not in users:
            assert d["kind"] == codec.KIND_NONE
            continue
        assert same_js(users[key], d["value"]), (path, users[key], d)
        assert clock_items(d["M"]) == meta.get(path)
        assert clock_items(d["V"]) == vclocks.get(path)
        assert d["alias"] == (path in case["alias"])
    order = sorted((d["cseq"], p) for p, d in decoded.items() if d["cseq"])
    assert [p.split("/")[1] for _, p in order] == list(users.keys())  # Object.entries order of the collection


@pytest.mark.parametrize("case", STREAMS["cases"], ids=STREAM_IDS)
def test_typed_oracle_equals_reference(case):
    schema, orc, codes, changes = typed_replay(
        case, lambda schema, indexed: TypedOracle(make_cfg(schema, 64, indexed)))
    rows = {schema.paths.name(i): orc.table[i] for i in range(len(schema.paths))}
    assert_typed_matches(case, schema, codes, changes, rows)
    for name, q in case.get("queries", {}).items():
        f = FIELD_SLOT[name]
        paths = lambda ids: [schema.paths.name(i) for i in ids]  # noqa: E731
        for v, want, cnt in zip(EQ_VALUES, q["equals"], q["count"]):
            key = schema.index_key(v)
            assert ([] if key is None else paths(orc.query_equals(f, key))) == want, (name, v)
            assert (0 if key is None else orc.query_count(f, key)) == cnt
        for (lo, hi), want in zip(itertools.product(BOUNDS, BOUNDS), q["range"]):
            got = paths(orc.query_range(f, schema.bound(lo, False), schema.bound(hi, True)))
            assert got == want, (name, lo, hi)  # exact (Map order, Set order) of the reference


def config1_inputs():
    c = CONFIG1["case"]
    rng = synth.rng_for(1)
    table = synth.make_table(c["n_records"], rng)
    batch = synth.make_batch(table, c["n_updates"], rng, keys=c["keys"])
    return c, table, batch


def config1_check(c, schema, batch, ch, rows, equals_ids, range_ids, counts, ordered_queries):
    assert "".join(map(str, ch.decision.tolist())) == c["codes"]
    assert len(ch.idx) == c["n_changes"]
    assert golden_io.changes_sha256(codec.decode_changes(schema, batch, ch)) == c["changes_sha256"]

    def row_items():
        for i in range(c["n_records"]):
            d = codec.decode_row(schema, rows[i])
            yield f"users/u{i}", d["value"], d["M"], d["V"], d["alias"]
    assert golden_io.table_sha256(row_items()) == c["table_sha256"]
    if ordered_queries:
        assert list(equals_ids) == c["equals_role_admin"] and list(range_ids) == c["range_age_20_30"]
    else:
        assert sorted(equals_ids) == sorted(c["equals_role_admin"])
        assert sorted(range_ids) == sorted(c["range_age_20_30"])
    assert list(counts) == c["count_role"]


def config1_run(eng, c, table, batch, n_chunks):
    schema = synth.synth_schema(c["n_records"])
    for i in range(c["n_records"]):
        assert schema.paths.id(f"users/u{i}") == i  # dense path ids: row index == record number
    ids = np.arange(c["n_records"], dtype=np.uint64)
    eng.table_load(ids, table.rows) if hasattr(eng, "table_load") else eng.load(ids, table.rows)
    eng.index_create(2)
    eng.index_create(0)
    parts = []
    cuts = np.linspace(0, batch.n, n_chunks + 1).astype(int)
    for lo, hi in zip(cuts, cuts[1:]):
        ch = eng.merge(batch.slice(int(lo), int(hi)))
        parts.append((int(lo), ch))
    dec = np.concatenate([ch.decision for _, ch in parts])
    merged = codec.Changes(dec, np.concatenate([ch.idx + lo for lo, ch in parts]),
                           np.concatenate([ch.head for _, ch in parts]), np.concatenate([ch.clk for _, ch in parts]),
                           np.concatenate([ch.val for _, ch in parts]))
    eq = eng.query_equals(2, schema.index_key("admin"))
    rg = eng.query_range(0, schema.bound(20.0, False), schema.bound(30.0, True))
    counts = [eng.query_count(2, schema.index_key(r)) for r in ("admin", "editor", "user")]
    return schema, merged, eq, rg, counts


def test_config1_typed_oracle_equals_reference():
    """BASELINE config 1 at full size (10 000 records, 100 000 updates + equals(users, role, admin))."""
    c, table, batch = config1_inputs()
    cfg = capi.make_config(c["n_records"], local_peer=0, flags=codec.CFG_POST_GETDATA, **synth.synth_ranks(c["n_records"]))
    orc = TypedOracle(cfg)
    schema, ch, eq, rg, counts = config1_run(orc, c, table, batch, 3)
    config1_check(c, schema, batch, ch, orc.table, eq.tolist(), rg.tolist(), counts, ordered_queries=True)


# ----------------------------------------------------------------------------- GPU through the C ABI
@pytest.mark.gpu
@pytest.mark.parametrize("case", STREAMS["cases"], ids=STREAM_IDS)
def test_gpu_equals_reference(case):
    from bullet_js_b200.engine import Engine

    schema, eng, codes, changes = typed_replay(
        case, lambda schema, indexed: Engine.for_schema(schema, 64, post_getdata=indexed))
    n = len(schema.paths)
    got = eng.table_read(np.arange(n, dtype=np.uint64))
    rows = {schema.paths.name(i): got[i] for i in range(n)}
    assert_typed_matches(case, schema, codes, changes, rows)
    for name, q in case.get("queries", {}).items():
        f = FIELD_SLOT[name]
        ids = lambda paths: sorted(schema.paths.id(p) for p in paths)  # noqa: E731
        for v, want, cnt in zip(EQ_VALUES, q["equals"], q["count"]):
            key = schema.index_key(v)
            assert ([] if key is None else sorted(eng.query_equals(f, key).tolist())) == ids(want), (name, v)
            assert (0 if key is None else eng.query_count(f, key)) == cnt
        for (lo, hi), want in zip(itertools.product(BOUNDS, BOUNDS), q["range"]):
            got_ids = eng.query_range(f, schema.bound(lo, False), schema.bound(hi, True)).tolist()
            assert sorted(got_ids) == ids(want), (name, lo, hi)  # multiset, duplicates from stale entries included
    eng.close()


@pytest.mark.gpu
@pytest.mark.parametrize("case", [c for c in STREAMS["cases"] if c.get("queries")], ids=[i for c, i in zip(STREAMS["cases"], STREAM_IDS) if c.get("queries")])
def test_gpu_exact_query_order_equals_reference(case):
    """with BB_CFG_EXACT_ORDER the device returns the reference's own result LISTS - recorded from the
    reference's sources run unmodified (tests/golden/streams.json.gz): Map order of the buckets, Set order inside a
    bucket, duplicates from stale entries in place.  Exact list equality, no sorting on either side."""
    from bullet_js_b200.engine import Engine

    schema, eng, codes, changes = typed_replay(
        case, lambda schema, indexed: Engine.for_schema(schema, 64, post_getdata=indexed, exact_order=True))
    n = len(schema.paths)
    got = eng.table_read(np.arange(n, dtype=np.uint64))
    assert_typed_matches(case, schema, codes, changes, {schema.paths.name(i): got[i] for i in range(n)})
    checked = reordered = 0
    for name, q in case["queries"].items():
        f = FIELD_SLOT[name]
        ids = lambda paths: [schema.paths.id(p) for p in paths]  # noqa: E731
        for v, want in zip(EQ_VALUES, q["equals"]):
            key = schema.index_key(v)
            assert ([] if key is None else eng.query_equals(f, key).tolist()) == ids(want), (name, v)
        for (lo, hi), want in zip(itertools.product(BOUNDS, BOUNDS), q["range"]):
            got_ids = eng.query_range(f, schema.bound(lo, False), schema.bound(hi, True)).tolist()
            assert got_ids == ids(want), (name, lo, hi)
            checked += 1
            reordered += got_ids != sorted(got_ids)
    assert checked > 700 and reordered > 50
    eng.close()


@pytest.mark.gpu
def test_gpu_config1_equals_reference():
    from bullet_js_b200.engine import Engine

    c, table, batch = config1_inputs()
    eng = Engine(c["n_records"], post_getdata=True, **synth.synth_ranks(c["n_records"]))
    schema, ch, eq, rg, counts = config1_run(eng, c, table, batch, 3)
    rows = eng.table_read(np.arange(c["n_records"], dtype=np.uint64))
    config1_check(c, schema, batch, ch, rows, eq.tolist(), rg.tolist(), counts, ordered_queries=False)
    eng.close()


# ----------------------------------------------------------------------------- config 4 (index build + scans), reduced
CONFIG4 = golden_io.load("config4.json.gz")["case"]


def config4_run(eng):
    c = CONFIG4
    n = c["n_records"]
    table = synth.make_table(n, synth.rng_for(4))
    schema = synth.synth_schema(n)
    ids = np.arange(n, dtype=np.uint64)
    (eng.table_load if hasattr(eng, "table_load") else eng.load)(ids, table.rows)
    eng.index_create(0)
    eng.index_create(2)
    b = lambda v, up: schema.bound(v, up)  # noqa: E731
    return dict(
        range_age_20_30=eng.query_range(0, b(20.0, False), b(30.0, True)).tolist(),
        range_age_0_1000=len(eng.query_range(0, b(0.0, False), b(1000.0, True))),
        equals_role_admin=eng.query_equals(2, schema.index_key("admin")).tolist(),
        equals_age_25=eng.query_equals(0, schema.index_key(25.0)).tolist(),
        count_role=[eng.query_count(2, schema.index_key(r)) for r in ("admin", "editor", "user")])


def test_config4_typed_oracle_equals_reference():
    """Index BUILD from the store (_buildIndex) + range / equals / count on 20 000 nodes, exact result order."""
    cfg = capi.make_config(CONFIG4["n_records"], local_peer=0, flags=codec.CFG_POST_GETDATA,
                           **synth.synth_ranks(CONFIG4["n_records"]))
    got = config4_run(TypedOracle(cfg))
    for k, v in got.items():
        assert v == CONFIG4[k], k
    assert len(CONFIG4["range_age_20_30"]) > 1500 and sum(CONFIG4["count_role"]) == CONFIG4["n_records"]


@pytest.mark.gpu
def test_gpu_config4_equals_reference():
    from bullet_js_b200.engine import Engine

    eng = Engine(CONFIG4["n_records"], post_getdata=True, **synth.synth_ranks(CONFIG4["n_records"]))
    got = config4_run(eng)
    for k, v in got.items():
        want = CONFIG4[k]
        assert (sorted(v) == sorted(want)) if isinstance(v, list) and k != "count_role" else v == want, k
    eng.close()
