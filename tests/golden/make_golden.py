#!/usr/bin/env python
"""Writes tests/golden/*.json.gz from the REFERENCE ITSELF: /root/reference/src/*.js executed,
unmodified, by oracle/minijs (this image has no node).  Run here (the GPU box has no
/root/reference):   python tests/golden/make_golden.py [--only kat|streams|config1]

Fixtures (all inputs are regenerated deterministically or stored alongside the outputs):
  kat.json.gz      step-by-step traces of the SURVEY 8c scenarios (KAT-L/N/H/R/Q1/Q2) + the console
                   output of examples/bullet-query-example.js run as a script.
  streams.json.gz  random JS-level update streams (tests/streamgen.py) replayed through
                   `bullet.setData` / `BulletNetworkSync._processSyncEntries`: per-update decision,
                   ordered change set, final store / meta clocks / crt.vectorClocks / aliasing, the
                   index Maps in their exact (Map, Set) order, equals / count / range results.
  mesh.json.gz     BASELINE config 5 in miniature: 4 peers in a full mesh (tests/meshsim.py), every peer an
                   instance of the reference; per-peer logs (own puts interleaved with received broadcasts),
                   decisions, change sets and final replicas.
  restart.json.gz  a peer saves through the reference's BulletFileStorage (in-memory disk), a new instance loads
                   the files and carries on: the files, the loaded state and the second half's results.
  config4.json.gz  BASELINE config 4 reduced to 20 000 nodes: index build from the store + range / equals / count.
  config1.json.gz  BASELINE config 1 at full size (10 000 records, 100 000 updates of the synthetic
                   typed schema, then equals(users, role, admin)): decisions, SHA-256 of the change
                   set and of the final table in a canonical text form, and the query results.
"""
from __future__ import annotations

import argparse
import gzip
import hashlib
import itertools
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_runner  # noqa: E402
from oracle.ref_runner import JSRefBullet, jsonable  # noqa: E402
from oracle.jsvalue import UNDEFINED  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def write(name, obj):
    os.makedirs(GOLDEN, exist_ok=True)
    path = os.path.join(GOLDEN, name)
    raw = json.dumps(obj, separators=(",", ":"), ensure_ascii=True).encode()
    with open(path, "wb") as f:
        with gzip.GzipFile(fileobj=f, mode="wb", mtime=0) as g:  # mtime=0: byte-reproducible
            g.write(raw)
    print(f"{name}: {len(raw)} B json, {os.path.getsize(path)} B gz")


def reference_identity():
    root = ref_runner._reference_root()
    h = hashlib.sha256()
    files = ["src/bullet.js", "src/bullet-crt.js", "src/bullet-query.js", "src/bullet-middleware.js",
             "src/bullet-network-sync.js", "src/bullet-storage.js"]
    for f in files:
        with open(os.path.join(root, f), "rb") as fh:
            h.update(fh.read())
    with open(os.path.join(root, "package.json")) as fh:
        pkg = json.load(fh)
    return dict(name=pkg.get("name"), version=pkg.get("version"), files=files, sha256=h.hexdigest(),
                engine="oracle/minijs (ECMAScript subset interpreter); reference sources unmodified")


def snapshot_state(js, with_index=True):
    meta = js.meta
    out = dict(
        store=jsonable(js.store),
        meta=[[p, m.get("source"), list(map(list, m["vectorClock"].items()))] for p, m in meta.items()],
        vclocks=[[p, list(map(list, c.items()))] for p, c in js.crt.vectorClocks.items()],
        alias=[p for p in meta if js.alias(p)],
    )
    if with_index and js.bullet.get("query") is not UNDEFINED:
        out["index"] = js.index_dump()
    return out


# ----------------------------------------------------------------------------- KATs
def kat_cases():
    from tests.test_oracle_kat import USERS, PRODUCTS
    cases = []

    def trace(name, js, steps):
        rec = []
        for step in steps:
            kind = step[0]
            if kind == "put":
                js.put(step[1], step[2])
            elif kind == "recv":
                js.process_sync_entries([dict(path=step[1], data=step[2], vectorClock=step[3])])
            elif kind == "index":
                js.index(step[1], step[2])
            entry = dict(step=jsonable(list(step)))
            if kind in ("put", "recv"):
                d = js.decisions[-1]
                entry.update(code=d["code"], reason=d["reason"], doUpdate=d["doUpdate"])
            entry.update(snapshot_state(js))
            rec.append(entry)
        cases.append(dict(name=name, peer=js.id, steps=rec))
        return js

    trace("KAT-L", JSRefBullet("A", enable_indexing=False),
          [("put", "k/v", float(x)) for x in (5, 3, 3, 3, 7, 9, 0, 0, 4)])
    p = "users/u1"
    trace("KAT-N", JSRefBullet("B"), [
        ("index", "users", "age"), ("index", "users", "role"),
        ("recv", p, {"age": 30.0, "role": "user"}, {"A": 3.0}),
        ("recv", p, {"age": 25.0, "role": "admin"}, {"A": 4.0}),
        ("recv", p, {"age": 40.0}, {"A": 5.0, "B": 2.0}),
        ("recv", p, {"age": 10.0}, {"A": 4.0, "B": 2.0}),
        ("recv", p, {"age": 10.0}, {"B": 2.0, "A": 5.0}),
        ("recv", p, {"age": 10.0}, {"B": 2.0, "A": 5.0}),
    ])
    js = trace("KAT-H", JSRefBullet("A"), [
        ("index", "users", "age"), ("put", p, {"age": 30.0}), ("put", p, {"age": 31.0}), ("put", p, {"age": 0.0})])
    cases[-1]["range_30_31"] = js.range("users", "age", 30.0, 31.0)

    js = JSRefBullet("A")
    r = []
    for args in [("k", {"peerA": 1.0}, {"peerB": 1.0}, "light", "dark"), ("k", {"peerB": 1.0}, {"peerA": 1.0}, "dark", "light"),
                 ("k", {}, {"A": 1.0}, "x", "y"), ("k", {}, {}, 3.0, 3.0), ("k", {"A": 2.0}, {"A": 1.0}, 1.0, 2.0),
                 ("k", {"A": 1.0, "B": 1.0}, {"B": 1.0, "A": 1.0}, {"a": 1.0, "b": "x"}, {"a": 2.0, "c": True})]:
        out = js.resolve(*args)
        r.append(dict(args=jsonable(list(args)), reason=out["reason"], value=jsonable(out["value"]),
                      vectorClock=list(map(list, out["vectorClock"].items())),
                      flags={k: bool(out.get(k)) for k in ("incoming", "current", "concurrent", "historical")}))
    cases.append(dict(name="KAT-R", resolve=r))

    js = JSRefBullet("me")
    for k, v in USERS.items():
        js.put(f"users/{k}", {a: (float(b) if isinstance(b, int) and not isinstance(b, bool) else b) for a, b in v.items()})
    for k, v in PRODUCTS.items():
        js.put(f"products/{k}", {a: (float(b) if isinstance(b, int) and not isinstance(b, bool) else b) for a, b in v.items()})
    for path, field in [("users", "role"), ("users", "age"), ("users", "active"), ("products", "category"), ("products", "price")]:
        js.index(path, field)
    cases.append(dict(
        name="KAT-Q1",
        equals_role_admin=js.equals("users", "role", "admin"),
        range_age_30_40=js.range("users", "age", 30.0, 40.0),
        count_role=[js.count("users", "role", r) for r in ("admin", "user", "editor")],
        range_price_100_300=js.range("products", "price", 100.0, 300.0),
        equals_active_true=js.equals("users", "active", True),
        index=js.index_dump()))

    js = JSRefBullet("me")
    js.index("users", "role")
    js.put("users/alice", {"name": "Alice", "email": "alice@example.com", "role": "admin"})
    js.put("users/bob", {"name": "Bob", "email": "bob@example.com", "role": "user"})
    cases.append(dict(name="KAT-Q2", equals_role_admin=js.equals("users", "role", "admin"), index=js.index_dump()))

    # docs/conflict-resolution.md:446-483 (the documented examples that agree with the code): two peers write
    # concurrently and exchange their broadcasts; the documentation states the final values
    import copy

    def two_peers(path, va, vb):
        a, b = JSRefBullet("peerA", enable_indexing=False), JSRefBullet("peerB", enable_indexing=False)
        a.put(path, copy.deepcopy(va))
        b.put(path, copy.deepcopy(vb))
        ca, cb = a.last_change(), b.last_change()
        for dst, ch in ((b, ca), (a, cb)):
            v = copy.deepcopy(ch["value"])
            dst.handle_put(path, {**v, "__vectorClock": dict(ch["vectorClock"])} if isinstance(v, dict) else v)
        return dict(path=path, a=jsonable(va), b=jsonable(vb), store_a=jsonable(a.store), store_b=jsonable(b.store),
                    reasons_a=[d["reason"] for d in a.decisions], reasons_b=[d["reason"] for d in b.decisions])
    cases.append(dict(name="docs/conflict-resolution.md", runs=[
        two_peers("settings/theme", "dark", "light"),
        two_peers("users/bob", {"name": "Bob Smith", "age": 30.0, "preferences": {"theme": "dark"}},
                  {"name": "Robert Smith", "location": "New York", "preferences": {"notifications": False}})]))

    # the reference's own example, run as a script
    from oracle.minijs.builtins import Runtime
    rt = Runtime(console=[])
    rt.require(os.path.join(ref_runner._reference_root(), "examples", "bullet-query-example.js"))
    rt.run_microtasks()
    rt.run_timers(60000)
    lines = [t for k, t in rt.console if "initialized with ID" not in t and "closed" not in t]
    cases.append(dict(name="examples/bullet-query-example.js", console=lines))
    return cases


# ----------------------------------------------------------------------------- random JS-level streams
def stream_case(seed, n_ops, n_paths, index_fields, late_index, **gen):
    from tests import streamgen
    from tests.test_oracle_query import BOUNDS, EQ_VALUES
    ops, _ref = streamgen.generate(seed, n_ops, n_paths, index_fields=index_fields, late_index=late_index, **gen)
    indexed = bool(index_fields) or bool(late_index)
    js = JSRefBullet("p0", enable_indexing=indexed)
    for f in index_fields:
        js.index("users", f)
    for k, op in enumerate(ops):
        for f, at in (late_index or {}).items():
            if at == k:
                js.index("users", f)
        streamgen.apply_op(js, op)
    case = dict(
        seed=seed, n_ops=n_ops, n_paths=n_paths, index_fields=list(index_fields), late_index=late_index or {},
        ops=[[p, jsonable(v), None if c is None else list(map(list, c.items()))] for p, v, c in ops],
        codes="".join(str(d["code"]) for d in js.decisions),
        do_update="".join("1" if d["doUpdate"] else "0" for d in js.decisions),
        changes=[[c["seq"], c["path"], jsonable(c["value"]), list(map(list, c["vectorClock"].items())),
                  c["fromNetwork"]] for c in js.changes],
    )
    case.update(snapshot_state(js))
    if indexed:
        fields = sorted(set(index_fields) | set(late_index or {}))
        q = {}
        for name in fields:
            q[name] = dict(
                equals=[js.equals("users", name, v) for v in EQ_VALUES],
                count=[js.count("users", name, v) for v in EQ_VALUES],
                range=[js.range("users", name, lo, hi) for lo, hi in itertools.product(BOUNDS, BOUNDS)],
                range_undefined=[js.range("users", name, UNDEFINED, 5.0), js.range("users", name, 0.0, UNDEFINED)],
            )
        case["queries"] = q
    return case


def stream_cases():
    out = []
    for seed in range(4):
        t = time.time()
        out.append(stream_case(1000 + seed, 1200, 12, (), None))
        out.append(stream_case(2000 + seed, 1200, 12, ("age", "role"), {"score": 600}))
        print(f"  stream seed {seed}: {time.time() - t:.1f} s")
    # one longer stream over more paths: what the GPU tests slice into uneven batches
    out.append(stream_case(3000, 4000, 37, ("age",), None))
    # other mixes: mostly local puts (the M-alias / tie-by-value branches), mostly primitives written over records
    # and back (kind changes, falsy materialisation, null = node.remove()), both with the index hook installed
    out.append(stream_case(4000, 1500, 10, ("age", "name"), None, p_local=0.85))
    out.append(stream_case(4001, 1500, 10, ("role",), {"age": 500}, p_prim=0.7))
    out.append(stream_case(4002, 1500, 6, (), None, p_local=0.6, p_prim=0.5))
    return out


# ----------------------------------------------------------------------------- config 5 in miniature: a mesh of peers
def mesh_case(n_peers=4, n_ops=600, n_paths=10, seed=42):
    from tests import meshsim
    peers, logs = meshsim.run_mesh(lambda i: JSRefBullet(i, enable_indexing=False), n_peers, n_ops, n_paths, seed)
    out = dict(n_peers=n_peers, n_ops=n_ops, n_paths=n_paths, seed=seed, peers=[])
    for js, log in zip(peers, logs):
        d = js.decisions
        entry = dict(
            id=js.id,
            log=[[p, jsonable(v), None if c is None else list(map(list, c.items()))] for p, v, c in log],
            codes="".join(str(x["code"]) for x in d),
            changes=[[c["seq"], c["path"], jsonable(c["value"]), list(map(list, c["vectorClock"].items())),
                      c["fromNetwork"]] for c in js.changes])
        entry.update(snapshot_state(js, with_index=False))
        out["peers"].append(entry)
    return out


# ----------------------------------------------------------------------------- restart from the reference's own files
def restart_case(seed=5000, n_ops=1400, cut=800, n_paths=12):
    """Peer p0 (file storage on an in-memory disk) takes ops[:cut] and saves; a new instance with a fresh id
    (p1, as src/bullet.js:33 gives every start) loads the files and takes ops[cut:]."""
    from tests import streamgen
    ops, _ref = streamgen.generate(seed, n_ops, n_paths)
    files = {}
    opt = {"storageType": "file", "storagePath": "/data", "saveInterval": 0}
    a = JSRefBullet("p0", enable_indexing=False, options=opt, files=files)
    for op in ops[:cut]:
        streamgen.apply_op(a, op)
    a.save()
    saved = {k: files[k] for k in ("/data/store.json", "/data/meta.json")}
    b = JSRefBullet("p1", enable_indexing=False, options=opt, files=files)
    loaded = snapshot_state(b, with_index=False)
    for op in ops[cut:]:
        streamgen.apply_op(b, op)
    case = dict(seed=seed, cut=cut, n_paths=n_paths, files=saved, loaded=loaded,
                before=snapshot_state(a, with_index=False),
                ops=[[p, jsonable(v), None if c is None else list(map(list, c.items()))] for p, v, c in ops],
                codes="".join(str(d["code"]) for d in b.decisions),
                changes=[[c["seq"], c["path"], jsonable(c["value"]), list(map(list, c["vectorClock"].items())),
                          c["fromNetwork"]] for c in b.changes])
    case.update(snapshot_state(b, with_index=False))
    return case


# ----------------------------------------------------------------------------- config 1 (typed synthetic schema)
def canonical_value(v):
    return json.dumps(jsonable(v), separators=(",", ":"), ensure_ascii=True)


def config1_case(n_records=10_000, n_updates=100_000, chunk=50):
    import numpy as np
    from bullet_js_b200 import codec, synth
    from oracle.minijs import interp as I
    from oracle.minijs.builtins import from_py

    rng = synth.rng_for(1)
    table = synth.make_table(n_records, rng)
    batch = synth.make_batch(table, n_updates, rng, keys="zipf")
    schema = synth.synth_schema(n_records)

    js = JSRefBullet("p0")
    # initial graph: records, meta clocks and crt clocks (aliased, as after an accepted write: src/bullet.js:198-203)
    rt = js.rt
    store_users = from_py({})
    js.bullet.get("store").put_own("users", store_users)
    meta = js.bullet.get("meta")
    vclocks = I.get_member(js.bullet.get("crt"), "vectorClocks")
    t0 = time.time()
    for i in range(n_records):
        d = codec.decode_row(schema, table.rows[i])
        store_users.put_own(f"u{i}", from_py(d["value"]))
        clock = from_py(d["M"])
        m = from_py({"source": "network"})
        m.put_own("vectorClock", clock)
        meta.put_own(f"users/u{i}", m)
        rt.method(vclocks, "set", f"users/u{i}", clock)
    js.index("users", "role")
    js.index("users", "age")
    print(f"  config1: table loaded in {time.time() - t0:.1f} s")
    t0 = time.time()
    pending = []

    def flush():
        if pending:
            js.process_sync_entries(pending)
            pending.clear()

    for k in range(n_updates):
        hdr = int(batch.head["hdr"][k])
        value = schema.dec_value(hdr, batch.val[k])
        path = f"users/u{int(batch.path_id[k])}"
        if hdr & 1:  # network flavour, delivered through the sync driver in chunks of 50 (sync:713-723)
            clock = schema.dec_clock(batch.clk[k], int(batch.head["clk_order"][k]))
            pending.append(dict(path=path, data=value, vectorClock=clock))
            if len(pending) == chunk:
                flush()
        else:
            flush()
            js.put(path, value)
        if k % 20000 == 0:
            print(f"  config1: {k} updates, {time.time() - t0:.1f} s")
    flush()
    decisions = js.decisions
    h = hashlib.sha256()
    for c in js.changes:
        h.update(f"{int(c['seq'])}|{c['path']}|{canonical_value(c['value'])}|{canonical_value(c['vectorClock'])}\n".encode())
    changes_sha = h.hexdigest()
    h = hashlib.sha256()
    store = js.store["users"]
    meta_py = js.meta
    vc = js.crt.vectorClocks
    for i in range(n_records):
        p = f"users/u{i}"
        h.update(f"{p}|{canonical_value(store[f'u{i}'])}|{canonical_value(meta_py[p]['vectorClock'])}|"
                 f"{canonical_value(vc[p])}|{int(js.alias(p))}\n".encode())
    table_sha = h.hexdigest()

    def ids(paths):
        return [int(p.split("/u")[1]) for p in paths]
    return dict(
        n_records=n_records, n_updates=n_updates, keys="zipf", rng="synth.rng_for(1)", chunk=chunk,
        codes="".join(str(d["code"]) for d in decisions),
        n_changes=len(js.changes), changes_sha256=changes_sha, table_sha256=table_sha,
        equals_role_admin=ids(js.equals("users", "role", "admin")),
        count_role=[js.count("users", "role", r) for r in ("admin", "editor", "user")],
        range_age_20_30=ids(js.range("users", "age", 20.0, 30.0)),
        index_entries={k: sum(len(b[1]) for b in v) for k, v in js.index_dump().items()},
    )


# ----------------------------------------------------------------------------- config 4 (index build + scans), reduced
def config4_case(n_records=20_000):
    """BASELINE config 4 at a size the interpreter finishes in seconds: the synthetic table is placed in the
    reference's store, then index('users','age') / index('users','role') are BUILT from it (_buildIndex,
    src/bullet-query.js:53-73) and range(20,30) / equals(role,'admin') / count run."""
    from bullet_js_b200 import codec, synth
    from oracle.minijs.builtins import from_py

    table = synth.make_table(n_records, synth.rng_for(4))
    schema = synth.synth_schema(n_records)
    js = JSRefBullet("p0")
    users = from_py({})
    js.bullet.get("store").put_own("users", users)
    for i in range(n_records):
        users.put_own(f"u{i}", from_py(codec.decode_row(schema, table.rows[i])["value"]))
    js.index("users", "age")
    js.index("users", "role")

    def ids(paths):
        return [int(p.split("/u")[1]) for p in paths]
    return dict(n_records=n_records, rng="synth.rng_for(4)",
                range_age_20_30=ids(js.range("users", "age", 20.0, 30.0)),
                range_age_0_1000=len(js.range("users", "age", 0.0, 1000.0)),
                equals_role_admin=ids(js.equals("users", "role", "admin")),
                equals_age_25=ids(js.equals("users", "age", 25.0)),
                count_role=[js.count("users", "role", r) for r in ("admin", "editor", "user")],
                buckets={k: len(v) for k, v in js.index_dump().items()})


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", choices=["kat", "streams", "config1", "mesh", "restart", "config4"])
    ap.add_argument("--config1-updates", type=int, default=100_000)
    args = ap.parse_args()
    if not ref_runner.available():
        sys.exit("reference sources not found (set BULLET_REFERENCE)")
    ident = reference_identity()
    if args.only in (None, "kat"):
        write("kat.json.gz", dict(reference=ident, cases=kat_cases()))
    if args.only in (None, "streams"):
        write("streams.json.gz", dict(reference=ident, cases=stream_cases()))
    if args.only in (None, "restart"):
        write("restart.json.gz", dict(reference=ident, case=restart_case()))
    if args.only in (None, "mesh"):
        write("mesh.json.gz", dict(reference=ident, case=mesh_case()))
    if args.only in (None, "config4"):
        write("config4.json.gz", dict(reference=ident, case=config4_case()))
    if args.only in (None, "config1"):
        write("config1.json.gz", dict(reference=ident, case=config1_case(n_updates=args.config1_updates)))


if __name__ == "__main__":
    main()
