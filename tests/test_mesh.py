"""BASELINE config 5 (mesh-topology sync replay) at test size.

tests/golden/mesh.json.gz: 4 instances of the reference in a full mesh (tests/golden/make_golden.py, tests/meshsim.py).
  * the literal oracle, run through the same mesh simulator, must produce the same per-peer logs, decisions,
    change sets and replicas (order-dependent rule: "converged" means "equals the reference replay")
  * the typed C oracle and (gpu) the CUDA path replay every peer's log in uneven batches
  * (gpu) 8 peers x ~10 k log entries each, one replica per engine, bit-exact against the typed oracle
"""
import numpy as np
import pytest

from bullet_js_b200 import codec
from oracle.js_literal import RefBullet
from oracle.ref_runner import unjsonable
from oracle.typed import TypedOracle
from tests import golden_io, meshsim, streamgen
from tests.golden_io import clock_items, same_js
from tests.test_oracle_typed import make_cfg

MESH = golden_io.load("mesh.json.gz")["case"]


def log_of(peer):
    return [(p, unjsonable(v), None if c is None else {k: float(x) for k, x in c}) for p, v, c in peer["log"]]


def test_literal_mesh_equals_reference_mesh():
    peers, logs = meshsim.run_mesh(lambda i: RefBullet(i, enable_indexing=False), MESH["n_peers"], MESH["n_ops"],
                                   MESH["n_paths"], MESH["seed"])
    seen = set()
    for ref, log, want in zip(peers, logs, MESH["peers"]):
        wl = log_of(want)
        assert len(log) == len(wl)
        for (p, v, c), (wp, wv, wc) in zip(log, wl):
            assert p == wp and same_js(v, wv) and clock_items(c) == clock_items(wc)
        assert "".join(str(d["code"]) for d in ref.decisions) == want["codes"]
        assert len(ref.changes) == len(want["changes"])
        for got, (seq, path, value, clock, from_net) in zip(ref.changes, want["changes"]):
            assert (got["seq"], got["path"], got["fromNetwork"]) == (seq, path, from_net)
            assert same_js(got["value"], unjsonable(value)) and clock_items(got["vectorClock"]) == clock
        assert same_js(ref.store, unjsonable(want["store"]))
        assert [[p, clock_items(m["vectorClock"])] for p, m in ref.meta.items()] == [[p, c] for p, _s, c in want["meta"]]
        assert [[p, clock_items(c)] for p, c in ref.crt.vectorClocks.items()] == want["vclocks"]
        seen |= set(want["codes"])
    assert len(seen) >= 6
    stores = [unjsonable(p["store"]) for p in MESH["peers"]]
    assert not all(same_js(stores[0], s) for s in stores[1:])  # the rule is order-dependent: replicas differ


def replay_peer(want, make_engine):
    ops = log_of(want)
    schema = codec.Schema(streamgen.FIELDS, streamgen.PEERS, codec.StringDict(streamgen.STRINGS), want["id"])
    batch = codec.encode_updates(schema, ops)
    eng = make_engine(schema)
    cuts = sorted({0, 1, 5, 130, 131, 400, len(ops)})
    codes, changes = [], []
    for lo, hi in zip(cuts, cuts[1:]):
        ch = eng.merge(batch.slice(lo, hi))
        codes.extend(ch.decision.tolist())
        sub = codec.decode_changes(schema, batch.slice(lo, hi), ch)
        for c in sub:
            c["seq"] += lo
        changes.extend(sub)
    assert "".join(map(str, codes)) == want["codes"]
    assert len(changes) == len(want["changes"])
    for got, (seq, path, value, clock, _f) in zip(changes, want["changes"]):
        assert (got["seq"], got["path"]) == (seq, path)
        assert same_js(got["value"], unjsonable(value)) and clock_items(got["vectorClock"]) == clock
    return schema, eng


def check_rows(schema, rows, want):
    users = unjsonable(want["store"]).get("users", {})
    meta = {p: c for p, _s, c in want["meta"]}
    vclocks = dict((p, c) for p, c in want["vclocks"])
    for i in range(len(schema.paths)):
        path = schema.paths.name(i)
        d = codec.decode_row(schema, rows[i])
        assert same_js(users[path.split("/")[1]], d["value"])
        assert clock_items(d["M"]) == meta.get(path) and clock_items(d["V"]) == vclocks.get(path)
        assert d["alias"] == (path in want["alias"])


@pytest.mark.parametrize("k", range(len(MESH["peers"])))
def test_typed_oracle_replays_reference_peer_log(k):
    want = MESH["peers"][k]
    schema, orc = replay_peer(want, lambda schema: TypedOracle(make_cfg(schema, 32, False)))
    check_rows(schema, orc.table, want)


@pytest.mark.gpu
@pytest.mark.parametrize("k", range(len(MESH["peers"])))
def test_gpu_replays_reference_peer_log(k):
    from bullet_js_b200.engine import Engine

    want = MESH["peers"][k]
    schema, eng = replay_peer(want, lambda schema: Engine.for_schema(schema, 32))
    check_rows(schema, eng.table_read(np.arange(len(schema.paths), dtype=np.uint64)), want)
    eng.close()


@pytest.mark.gpu
def test_gpu_mesh_of_8_replicas():
    """8 simulated peers, one replica (engine) each: every peer's ~10 k-entry log replayed in batches of 1 000,
    change sets and replicas bit-exact against the typed oracle (peer i <-> GPU i is the natural mapping at
    full size; the replicas are independent, so one device serves them in turn here)."""
    from bullet_js_b200.engine import Engine

    _peers, logs = meshsim.run_mesh(lambda i: RefBullet(i, enable_indexing=False), 8, 1500, 40, 7)
    for i, log in enumerate(logs):
        schema = codec.Schema(streamgen.FIELDS, streamgen.PEERS, codec.StringDict(streamgen.STRINGS), f"p{i}")
        batch = codec.encode_updates(schema, log)
        eng = Engine.for_schema(schema, 64)
        orc = TypedOracle(eng.cfg)
        for lo in range(0, batch.n, 1000):
            b = batch.slice(lo, min(lo + 1000, batch.n))
            assert eng.merge(b).same_as(orc.merge(b)), (i, lo)
        ids = np.arange(64, dtype=np.uint64)
        assert np.array_equal(eng.table_read(ids), orc.read(ids))
        eng.close()
