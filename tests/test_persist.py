"""cold start from a reference data directory.  tests/golden/restart.json.gz holds what the
reference's own BulletFileStorage wrote (store.json, meta.json) after 800 updates, and what a NEW instance of the
reference (fresh id) did with the next 600 after loading them.  bullet_js_b200/persist.py turns the files into
table rows ("M present, V absent"); the typed oracle and (gpu) the CUDA path must carry on exactly like the
reference, and exporting the table again must give back the reference's files."""
import json

import numpy as np
import pytest

from bullet_js_b200 import codec, persist
from oracle.js_literal import RefBullet
from oracle.ref_runner import unjsonable
from oracle.typed import TypedOracle
from tests import golden_io, streamgen
from tests.golden_io import clock_items, same_js
from tests.test_oracle_typed import make_cfg

CASE = golden_io.load("restart.json.gz")["case"]
STORE = json.loads(CASE["files"]["/data/store.json"])
META = json.loads(CASE["files"]["/data/meta.json"])


def second_half():
    return golden_io.ops_of(CASE)[CASE["cut"]:]


def test_files_are_what_the_reference_held():
    before = unjsonable(CASE["before"]["store"])
    assert same_js(persist._numbers_to_float(STORE), persist.json_value(before))
    assert [[p, clock_items(m["vectorClock"])] for p, m in META.items()] == [[p, c] for p, _s, c in CASE["before"]["meta"]]
    loaded = CASE["loaded"]
    assert loaded["vclocks"] == [] and loaded["alias"] == []  # crt.vectorClocks is not persisted


def test_literal_oracle_restart():
    ref = RefBullet("p1", enable_indexing=False)
    ref.store = persist._numbers_to_float(STORE)
    ref.meta = {p: {"source": m["source"], "vectorClock": {k: float(v) for k, v in m["vectorClock"].items()}}
                for p, m in META.items()}
    for op in second_half():
        streamgen.apply_op(ref, op)
    assert "".join(str(d["code"]) for d in ref.decisions) == CASE["codes"]
    assert same_js(ref.store, unjsonable(CASE["store"]))
    assert [[p, clock_items(m["vectorClock"])] for p, m in ref.meta.items()] == [[p, c] for p, _s, c in CASE["meta"]]
    assert "6" in CASE["codes"]  # a reloaded path written locally under the fresh id is a concurrent merge


def run_typed(make_engine, read_rows):
    schema = codec.Schema(streamgen.FIELDS, streamgen.PEERS, codec.StringDict(streamgen.STRINGS), "p1")
    ids, rows = persist.import_collection(schema, "users", STORE, META)
    ops = second_half()
    batch = codec.encode_updates(schema, ops)  # interns the paths the second half creates
    eng = make_engine(schema)
    (eng.table_load if hasattr(eng, "table_load") else eng.load)(ids, rows)
    codes, changes = [], []
    cuts = [0, 1, 9, 250, len(ops)]
    for lo, hi in zip(cuts, cuts[1:]):
        ch = eng.merge(batch.slice(lo, hi))
        codes.extend(ch.decision.tolist())
        sub = codec.decode_changes(schema, batch.slice(lo, hi), ch)
        for c in sub:
            c["seq"] += lo
        changes.extend(sub)
    assert "".join(map(str, codes)) == CASE["codes"]
    assert len(changes) == len(CASE["changes"])
    for got, (seq, path, value, clock, _f) in zip(changes, CASE["changes"]):
        assert (got["seq"], got["path"]) == (seq, path)
        assert same_js(got["value"], unjsonable(value)) and clock_items(got["vectorClock"]) == clock
    n = len(schema.paths)
    all_ids = np.arange(n, dtype=np.uint64)
    table = read_rows(eng, all_ids)
    # export == the reference's state after the second half (store in own-key order, meta clocks)
    records, meta = persist.export_collection(schema, "users", all_ids, table)
    want_store = unjsonable(CASE["store"])["users"]
    assert same_js(records, want_store)
    assert {p: clock_items(m["vectorClock"]) for p, m in meta.items()} == {p: c for p, _s, c in CASE["meta"]}
    vclocks = dict((p, c) for p, c in CASE["vclocks"])
    for i in range(n):
        d = codec.decode_row(schema, table[i])
        path = schema.paths.name(i)
        assert clock_items(d["V"]) == vclocks.get(path)
        assert d["alias"] == (path in CASE["alias"])
    return eng


def test_typed_oracle_restart():
    run_typed(lambda schema: TypedOracle(make_cfg(schema, 32, False)), lambda e, ids: e.table[ids.astype(np.int64)])


def test_import_rejects_what_the_table_cannot_hold():
    schema = streamgen.make_schema()
    with pytest.raises(codec.DomainError):
        persist.import_collection(schema, "users", {"users": {"u1": {"age": {"nested": 1}}}}, {})
    with pytest.raises(codec.DomainError):
        persist.import_collection(schema, "users", {"users": {"u1": {"age": 1}}}, {"users/u1/age": {"vectorClock": {"p0": 1}}})
    ids, rows = persist.import_collection(schema, "users", {"users": {"u1": 5, "u2": None}}, {})
    assert [codec.decode_row(schema, r)["value"] for r in rows] == [5.0, None] and not rows["flags"].any()


@pytest.mark.gpu
def test_gpu_restart():
    from bullet_js_b200.engine import Engine

    eng = run_typed(lambda schema: Engine.for_schema(schema, 32), lambda e, ids: e.table_read(ids))
    eng.close()


@pytest.mark.skipif(not __import__("oracle.ref_runner", fromlist=["x"]).available(), reason="reference sources not present")
def test_sync_producer_equals_reference():
    """persist.collect_full_sync_data / chunk_sync_data == BulletNetworkSync._collectFullSyncData(0) /
    _chunkSyncData of a live reference instance (records, primitives and nulls in the store)."""
    from oracle import ref_runner
    from oracle.minijs.builtins import to_py

    js = ref_runner.JSRefBullet("p0", enable_indexing=False)
    for op in golden_io.ops_of(CASE)[:500]:
        streamgen.apply_op(js, op)
    js.put("settings/theme", "dark")
    want = to_py(js.rt.method(js._sync, "_collectFullSyncData", 0.0))
    for e in want:
        e["lastModified"] = 0  # wall clock in the reference, not kept on the device
    got = persist.collect_full_sync_data(js.store, js.meta)
    assert len(got) == len(want) > 10
    for g, w in zip(got, want):
        assert g["path"] == w["path"] and same_js(g["data"], w["data"]) and g["deleted"] is w["deleted"] is False
        assert clock_items(g["vectorClock"]) == clock_items(w["vectorClock"])
    assert any(e["vectorClock"] for e in got) and any(not e["vectorClock"] and "/" in e["path"][6:] for e in got)
    chunks = to_py(js.rt.method(js._sync, "_chunkSyncData", js.rt.method(js._sync, "_collectFullSyncData", 0.0)))
    assert [len(c) for c in persist.chunk_sync_data(got)] == [len(c) for c in chunks]
    # the `since` filter (:602, :633): meta is looked up at the LEAF path, so it only ever drops paths that were written
    # as leaves; the leaves of whole records have no lastModified and always travel (the interpreter's Date.now()
    # advances 1 ms per call, so lastModified is a total order)
    js.put("settings/lang", "en")
    raw = to_py(js.bullet.get("meta"))
    stamps = sorted(m["lastModified"] for m in raw.values() if m.get("lastModified"))
    assert len(stamps) > 10
    for since in (stamps[0], stamps[len(stamps) // 2], stamps[-1], stamps[-1] + 1.0):
        want = to_py(js.rt.method(js._sync, "_collectFullSyncData", float(since)))
        got = persist.collect_full_sync_data(js.store, raw, since)
        assert [(e["path"], e["lastModified"]) for e in got] == [(e["path"], e["lastModified"]) for e in want]
    dropped = {e["path"] for e in persist.collect_full_sync_data(js.store, raw)} - {e["path"] for e in got}
    # only primitives written at their own path are ever filtered; every leaf of a whole record still travels
    assert {"settings/theme", "settings/lang"} <= dropped and all(p in raw and p.count("/") == 1 for p in dropped)
    assert any(e["path"].count("/") == 2 for e in got)
