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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, default=40)
    ap.add_argument("--first", type=int, default=90_000)
    ap.add_argument("--ops", type=int, default=800)
    a = ap.parse_args()
    for seed in range(a.first, a.first + a.seeds):
        try:
            print(seed, one(seed, a.ops), flush=True)
        except AssertionError as e:
            print("DIVERGENCE at seed", seed, e.args, flush=True)
            sys.exit(1)
    print("no divergence on", a.seeds, "seeds")


if __name__ == "__main__":
    main()
