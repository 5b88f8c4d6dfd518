#!/usr/bin/env python
"""Fuzz campaign (not collected by pytest): the reference run live in oracle/minijs against BOTH restatements on
many seeds and mixes.   python tests/fuzz_reference.py [--seeds 40] [--ops 800]
Prints one line per seed; exits 1 at the first divergence (with the seed and parameters to reproduce)."""
import argparse
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from bullet_js_b200 import codec  # noqa: E402
from oracle import ref_runner  # noqa: E402
from oracle.typed import TypedOracle  # noqa: E402
from tests import streamgen  # noqa: E402
from tests.golden_io import clock_items, same_js  # noqa: E402
from tests.test_oracle_typed import make_cfg  # noqa: E402


def one(seed, n_ops):
    rng = random.Random(seed)
    n_paths = rng.choice([3, 8, 20])
    p_local, p_prim = rng.choice([0.1, 0.3, 0.7]), rng.choice([0.0, 0.2, 0.6])
    fields = rng.sample(["age", "score", "role", "name"], rng.randint(0, 3))
    late = {f: rng.randrange(1, n_ops) for f in rng.sample([f for f in ["age", "score", "role", "name"] if f not in fields],
                                                           rng.randint(0, 1))}
    ops, lit = streamgen.generate(seed, n_ops, n_paths, index_fields=tuple(fields), late_index=late, p_local=p_local, p_prim=p_prim)
    indexed = bool(fields) or bool(late)
    js = ref_runner.JSRefBullet("p0", enable_indexing=indexed)
    for f in fields:
        js.index("users", f)
    schema = streamgen.make_schema()
    batch = codec.encode_updates(schema, ops)
    orc = TypedOracle(make_cfg(schema, 64, indexed))
    slot = {"age": 0, "score": 1, "role": 2, "name": 3}
    for f in fields:
        orc.index_create(slot[f])
    cuts = sorted({0, len(ops)} | set(late.values()))
    codes = []
    for lo, hi in zip(cuts, cuts[1:]):
        for f, at in late.items():
            if at == lo:
                orc.index_create(slot[f])
        codes += orc.merge(batch.slice(lo, hi)).decision.tolist()
    for k, op in enumerate(ops):
        for f, at in late.items():
            if at == k:
                js.index("users", f)
        streamgen.apply_op(js, op)
    want = [d["code"] for d in js.decisions]
    assert [d["code"] for d in lit.decisions] == want, "literal decisions"
    assert codes == want, "typed decisions"
    assert same_js(lit.store, js.store), "literal store"
    meta, vc = js.meta, js.crt.vectorClocks
    assert {p: clock_items(m["vectorClock"]) for p, m in lit.meta.items()} == {p: clock_items(m["vectorClock"]) for p, m in meta.items()}
    users = js.store.get("users", {})
    for i in range(len(schema.paths)):
        path = schema.paths.name(i)
        d = codec.decode_row(schema, orc.table[i])
        assert same_js(users[path.split("/")[1]], d["value"]), ("typed value", path)
        assert clock_items(d["M"]) == clock_items(meta[path]["vectorClock"]), ("typed M", path)
        assert clock_items(d["V"]) == clock_items(vc.get(path)) and d["alias"] == js.alias(path), ("typed V", path)
    if indexed:
        dump = js.index_dump()
        assert dump == {k: [[bk, list(s)] for bk, s in idx.items()] for k, idx in lit.query.indices.items()}, "literal index"
        for f in list(fields) + list(late):
            for lo, hi in ((0.0, 99.0), (-1e308, 1e308), ("", "zzzz"), (25.0, 25.0)):
                got = [schema.paths.name(i) for i in orc.query_range(slot[f], schema.bound(lo, False), schema.bound(hi, True))]
                assert got == js.range("users", f, lo, hi), ("typed range", f, lo, hi)
    return f"paths={n_paths} p_local={p_local} p_prim={p_prim} fields={fields} late={late} codes={sorted(set(want))}"


def one_mesh(seed, n_ops):
    """A full mesh of reference instances against the same mesh of literal oracles (tests/meshsim.py), then every
    peer's log replayed by the typed oracle."""
    from oracle.js_literal import RefBullet
    from tests import meshsim
    rng = random.Random(seed)
    n_peers, n_paths = rng.choice([2, 3, 5, 8]), rng.choice([2, 6, 15])
    kw = dict(p_prim=rng.choice([0.0, 0.1, 0.4]), mean_delay=rng.choice([0.5, 3.0, 20.0]))
    jp, jlogs = meshsim.run_mesh(lambda i: ref_runner.JSRefBullet(i, enable_indexing=False), n_peers, n_ops, n_paths, seed, **kw)
    lp, llogs = meshsim.run_mesh(lambda i: RefBullet(i, enable_indexing=False), n_peers, n_ops, n_paths, seed, **kw)
    for i, (a, b, la, lb) in enumerate(zip(jp, lp, jlogs, llogs)):
        assert len(la) == len(lb) and all(x[0] == y[0] and same_js(x[1], y[1]) and x[2] == y[2] for x, y in zip(la, lb)), ("log", i)
        want = [d["code"] for d in a.decisions]
        assert [d["code"] for d in b.decisions] == want, ("literal decisions", i)
        assert same_js(a.store, b.store), ("literal store", i)
        schema = codec.Schema(streamgen.FIELDS, streamgen.PEERS, codec.StringDict(streamgen.STRINGS), f"p{i}")
        orc = TypedOracle(make_cfg(schema, 32, False))
        assert orc.merge(codec.encode_updates(schema, la)).decision.tolist() == want, ("typed decisions", i)
        users = a.store.get("users", {})
        for k in range(len(schema.paths)):
            d = codec.decode_row(schema, orc.table[k])
            path = schema.paths.name(k)
            assert same_js(users[path.split("/")[1]], d["value"]), ("typed value", i, path)
            assert clock_items(d["M"]) == clock_items(a.meta[path]["vectorClock"]), ("typed M", i, path)
    return f"mesh peers={n_peers} paths={n_paths} {kw} log sizes={[len(l) for l in jlogs]}"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mesh", action="store_true", help="fuzz the mesh replay (BASELINE config 5) instead of single streams")
    ap.add_argument("--seeds", type=int, default=40)
    ap.add_argument("--first", type=int, default=90_000)
    ap.add_argument("--ops", type=int, default=800)
    a = ap.parse_args()
    for seed in range(a.first, a.first + a.seeds):
        try:
            print(seed, (one_mesh if a.mesh else one)(seed, a.ops), flush=True)
        except AssertionError as e:
            print("DIVERGENCE at seed", seed, e.args, flush=True)
            sys.exit(1)
    print("no divergence on", a.seeds, "seeds")


if __name__ == "__main__":
    main()
