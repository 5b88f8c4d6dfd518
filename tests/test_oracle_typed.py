"""Typed C oracle (oracle/bullet_oracle.c) == literal oracle (oracle/js_literal.py)."""
import math

import numpy as np
import pytest

from bullet_js_b200 import capi, codec
from oracle.typed import TypedOracle
from tests import streamgen


def same_js(a, b):
    if isinstance(a, dict) and isinstance(b, dict):
        return list(a.keys()) == list(b.keys()) and all(same_js(a[k], b[k]) for k in a)
    if isinstance(a, float) and isinstance(b, float):
        if math.isnan(a) or math.isnan(b):
            return math.isnan(a) and math.isnan(b)
        return a == b and math.copysign(1, a) == math.copysign(1, b)
    return type(a) is type(b) and a == b


def check_against_literal(schema, ops, ref, decisions, changes_decoded, rows_by_path):
    assert [d["code"] for d in ref.decisions] == list(decisions)
    assert len(ref.changes) == len(changes_decoded)
    for want, got in zip(ref.changes, changes_decoded):
        assert want["seq"] == got["seq"] and want["path"] == got["path"]
        assert same_js(want["value"], got["value"]), (want, got)
        assert list(want["vectorClock"].items()) == list(got["vectorClock"].items()), (want, got)
    users = ref.store.get("users", {})
    for path, row in rows_by_path.items():
        key = path.split("/")[1]
        d = codec.decode_row(schema, row)
        if key not in users:
            assert d["kind"] == codec.KIND_NONE
            continue
        assert same_js(users[key], d["value"]), (path, users[key], d)
        m = (ref.meta.get(path) or {}).get("vectorClock")
        v = ref.crt.vectorClocks.get(path)
        assert (None if m is None else list(m.items())) == (None if d["M"] is None else list(d["M"].items()))
        assert (None if v is None else list(v.items())) == (None if d["V"] is None else list(d["V"].items()))
        assert d["alias"] == (m is not None and m is v)
    order = sorted((codec.decode_row(schema, r)["cseq"], p) for p, r in rows_by_path.items()
                   if codec.decode_row(schema, r)["cseq"])
    assert [p.split("/")[1] for _, p in order] == list(users.keys())


def make_cfg(schema, capacity, post_getdata):
    return capi.make_config(capacity, local_peer=schema.peers.index(schema.local_peer),
                            flags=codec.CFG_POST_GETDATA if post_getdata else 0,
                            **schema.config_ranks())


@pytest.mark.parametrize("seed", range(6))
@pytest.mark.parametrize("indexed", [False, True])
def test_typed_equals_literal(seed, indexed):
    n_paths = 12
    ops, ref = streamgen.generate(seed, 1500, n_paths, index_fields=("age",) if indexed else ())
    schema = streamgen.make_schema()
    batch = codec.encode_updates(schema, ops)
    orc = TypedOracle(make_cfg(schema, 64, indexed))
    # several batches of uneven size: state must carry across calls
    cuts = [0, 1, 7, 300, 301, 900, len(ops)]
    decs, changes = [], []
    for lo, hi in zip(cuts, cuts[1:]):
        ch = orc.merge(batch.slice(lo, hi))
        decs.extend(ch.decision.tolist())
        sub = codec.decode_changes(schema, batch.slice(lo, hi), ch)
        for c in sub:
            c["seq"] += lo
        changes.extend(sub)
    rows = {schema.paths.name(i): orc.table[i] for i in range(len(schema.paths))}
    check_against_literal(schema, ops, ref, decs, changes, rows)
    codes = set(decs)
    assert codes == set(range(7)), codes  # every branch of resolve() was taken


def test_typed_mt_equals_sequential():
    ops, _ = streamgen.generate(99, 4000, 40)
    schema = streamgen.make_schema()
    batch = codec.encode_updates(schema, ops)
    a = TypedOracle(make_cfg(schema, 64, False))
    b = TypedOracle(make_cfg(schema, 64, False))
    ca, cb = a.merge(batch), b.merge(batch, threads=4)
    assert ca.same_as(cb)
    assert np.array_equal(a.table, b.table)


def test_kat_l_typed():
    """SURVEY.md 8c KAT-L through the typed path (falsy 0 materialises to {})."""
    schema = codec.Schema(["v"], ["A"], codec.StringDict([]), "A")
    ops = [("k/v", float(x), None) for x in (5, 3, 3, 3, 7, 9, 0, 0, 4)]
    batch = codec.encode_updates(schema, ops)
    orc = TypedOracle(make_cfg(schema, 4, False))
    ch = orc.merge(batch)
    assert ch.decision.tolist() == [0, 3, 4, 1, 4, 2, 3, 4, 2]
    assert ch.idx.tolist() == [0, 2, 4, 5, 7, 8]
    d = codec.decode_row(schema, orc.table[0])
    assert d["value"] == 4.0 and d["M"] == {"A": 11.0} and d["alias"]


def test_incoming_wins_entries_repeat_the_incoming_update():
    """The invariant a compact change set would rest on: when a NETWORK update wins outright
    (codes 2 and 4: identical clocks / incoming dominates) the emitted value and clock are the update's own, bit for
    bit, key orders included (crt:103-114: merging a dominating clock adds no key), so only concurrent merges,
    first writes and local puts need their entry shipped back to the host."""
    from bullet_js_b200 import synth

    n_rec = 5000
    rng = synth.rng_for(1, salt=77)
    table = synth.make_table(n_rec, rng)
    for keys in ("uniform", "zipf"):
        orc = TypedOracle(capi.make_config(n_rec, **synth.synth_ranks(n_rec)))
        orc.load(np.arange(n_rec, dtype=np.uint64), table.rows)  # the batch's clocks are relative to this image
        b = synth.make_batch(table, 60_000, rng, keys=keys)
        ch = orc.merge(b)
        idx = ch.idx.astype(np.int64)
        code = ch.decision[idx]
        net = (b.head["hdr"][idx] & codec.HDR_FLAVOUR_NET).astype(bool)
        outright = net & ((code == codec.DEC_TIE_INCOMING) | (code == codec.DEC_INCOMING))
        assert outright.sum() > 1000
        assert (ch.val[outright] == b.val[idx][outright]).all()
        assert ((ch.head["hdr"][outright] | 1) == (b.head["hdr"][idx][outright] | 1)).all()
        assert (ch.clk[outright] == b.clk[idx][outright]).all()
        assert (ch.head["clk_order"][outright] == b.head["clk_order"][idx][outright]).all()
        merged = code == codec.DEC_CONCURRENT
        assert not (ch.clk[merged] == b.clk[idx][merged]).all(axis=1).all()  # those do need their entry
