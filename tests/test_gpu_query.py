"""GPU parity for BulletQuery: index build, the post-write hook fused into the merge kernel, and
the equals / range / count scans, through the C ABI, against the typed oracle (multiset equality:
the device returns hits in node order, the reference in Map / Set order)."""
import itertools

import numpy as np
import pytest

from bullet_js_b200 import capi, codec, synth
from oracle.typed import TypedOracle
from tests import streamgen
from tests.test_oracle_query import BOUNDS, EQ_VALUES

pytestmark = pytest.mark.gpu


def pair(schema, capacity, **kw):
    from bullet_js_b200.engine import Engine

    if schema is not None:
        eng = Engine.for_schema(schema, capacity, post_getdata=True, **kw)
    else:
        eng = Engine(capacity, post_getdata=True, **kw)
    return eng, TypedOracle(eng.cfg)


def same_rows(eng, orc, n):
    ids = np.arange(n, dtype=np.uint64)
    got, want = eng.table_read(ids), orc.read(ids)
    got["xcnt"] = 0  # device-private: entries the node has in the index overflow sets
    bad = np.nonzero(got != want)[0]
    assert bad.size == 0, (bad[:5], got[bad[:2]], want[bad[:2]])


def check_queries(schema, eng, orc, fields, bounds=BOUNDS):
    for f in fields:
        nd, nx = eng.index_stats(f)
        assert nd + nx == orc.index_entries(f)
        for v in EQ_VALUES:
            key = schema.index_key(v)
            if key is None:
                continue
            got, want = eng.query_equals(f, key), orc.query_equals(f, key)
            assert np.array_equal(np.sort(got), np.sort(want)), (f, v)
            assert eng.query_count(f, key) == len(want)
        for lo, hi in itertools.product(bounds, bounds):
            bl, bh = schema.bound(lo, False), schema.bound(hi, True)
            got, want = eng.query_range(f, bl, bh), orc.query_range(f, bl, bh)
            assert np.array_equal(np.sort(got), np.sort(want)), (f, lo, hi)


@pytest.mark.parametrize("seed", range(3))
def test_random_js_streams_with_indices(seed):
    ops, _ref = streamgen.generate(300 + seed, 4000, 37, index_fields=("age", "role"), late_index={"score": 1500})
    schema = streamgen.make_schema()
    batch = codec.encode_updates(schema, ops)
    eng, orc = pair(schema, 64)
    for x in (eng, orc):
        x.index_create(0)
        x.index_create(2)
    for lo, hi, late in ((0, 1, False), (1, 1500, False), (1500, 1501, True), (1501, 4000, False)):
        if late:
            eng.index_create(1)
            orc.index_create(1)
        got, want = eng.merge(batch.slice(lo, hi)), orc.merge(batch.slice(lo, hi))
        assert got.same_as(want), (seed, lo, hi)
    same_rows(eng, orc, 64)
    check_queries(schema, eng, orc, (0, 1, 2))
    eng.close()


@pytest.mark.parametrize("keys", ["uniform", "zipf"])
def test_synthetic_schema_with_indices(keys):
    n_rec = 50_000
    rng = synth.rng_for(4, salt=2)
    table = synth.make_table(n_rec, rng)
    eng, orc = pair(None, n_rec, **synth.synth_ranks(n_rec))
    ids = np.arange(n_rec, dtype=np.uint64)
    eng.table_load(ids, table.rows)
    orc.load(ids, table.rows)
    schema = synth.synth_schema(n_rec)
    for f in (0, 2):  # age, role; every update can leave one more stale entry behind (query:151-167)
        eng.index_create(f, extra_capacity=1 << 20)
        orc.index_create(f)
    small = [20.0, 30.0, 0.0, 99.0, "admin", "user", "a", "zzz", "20", None]
    check_queries(schema, eng, orc, (0, 2), small)  # build path
    for _ in range(2):
        b = synth.make_batch(table, 200_000, rng, keys=keys)
        got, want = eng.merge(b), orc.merge(b)
        assert got.same_as(want)
    same_rows(eng, orc, n_rec)
    check_queries(schema, eng, orc, (0, 2), small)  # build + hook
    nd, nx = eng.index_stats(0)
    assert nx > 0  # stale entries piled up in the overflow set, as in the reference
    eng.close()


def test_query_errors():
    from bullet_js_b200.engine import Engine

    schema = streamgen.make_schema()
    eng = Engine.for_schema(schema, 64, post_getdata=True)
    with pytest.raises(capi.BulletB200Error) as e:
        eng.query_count(0, 5)
    assert e.value.code == capi.ERR_STATE
    plain = Engine.for_schema(schema, 64)
    with pytest.raises(capi.BulletB200Error) as e:
        plain.index_create(0)
    assert e.value.code == capi.ERR_STATE
    # hit buffer too small
    ops = [(f"users/u{i}", {"age": 30.0}, None) for i in range(10)]
    eng.merge(codec.encode_updates(schema, ops))
    eng.index_create(0)
    with pytest.raises(capi.BulletB200Error) as e:
        eng.query_equals(0, schema.index_key(30.0), capi.HitBuffers(4))
    assert e.value.code == capi.ERR_CAPACITY
    assert len(eng.query_equals(0, schema.index_key(30.0))) == 10
    # overflow set too small: 2000 distinct stale values for one node
    tiny = Engine.for_schema(schema, 64, post_getdata=True)
    tiny.index_create(0, extra_capacity=16)
    ops = [("users/u1", {"age": float(i + 1)}, None) for i in range(2000)]
    with pytest.raises(capi.BulletB200Error) as e:
        tiny.merge(codec.encode_updates(schema, ops))
    assert e.value.code == capi.ERR_CAPACITY
    for x in (eng, plain, tiny):
        x.close()


def test_kat_h_on_gpu():
    """SURVEY 8c KAT-H: hook staleness - u1 sits in buckets "30" and "31", {age:0} changes nothing."""
    from bullet_js_b200.engine import Engine

    schema = codec.Schema(["age", "role"], ["A", "B"], codec.StringDict(["admin", "user"]), "B")
    eng = Engine.for_schema(schema, 8, post_getdata=True)
    eng.index_create(0)
    ops = [("users/u1", {"age": 30.0}, None), ("users/u1", {"age": 31.0}, None), ("users/u1", {"age": 0.0}, None)]
    eng.merge(codec.encode_updates(schema, ops))
    assert eng.query_range(0, schema.bound(30.0, False), schema.bound(31.0, True)).tolist() == [0, 0]
    assert eng.index_stats(0) == (1, 1)
    eng.close()


def test_ordered_hits():
    """BB_CFG_ORDERED_CHANGES: the dense column's hits come out in ascending node id."""
    from bullet_js_b200.engine import Engine

    n_rec = 300_000
    rng = synth.rng_for(4, salt=5)
    table = synth.make_table(n_rec, rng)
    eng = Engine(n_rec, post_getdata=True, ordered_changes=True, **synth.synth_ranks(n_rec))
    eng.table_load(np.arange(n_rec, dtype=np.uint64), table.rows)
    eng.index_create(0, extra_capacity=1 << 12)
    lo, hi = np.zeros((), codec.BOUND_DTYPE), np.zeros((), codec.BOUND_DTYPE)
    lo["num"], hi["num"] = 20.0, 30.0
    hb = capi.HitBuffers(n_rec)
    got = eng.query_range(0, lo, hi, hb)
    ages = table.rows["val"][:, 0].view(np.float64)
    want = np.nonzero((ages >= 20) & (ages <= 30))[0]
    assert int(hb.n_extra[0]) == 0 and np.array_equal(got, want.astype(np.uint32))
    eng.close()


def test_router_queries_single_rank_and_multi_field_build():
    """bb_router_query_* on a one-rank router (the same kernels as the multi-GPU path: scan, counts through the
    control block, push into the result buffer, epoch barrier) and bb_index_create_fields (one pass, several
    indices) against the oracle, overflow entries included."""
    import torch

    from bullet_js_b200 import shard

    ops, _ref = streamgen.generate(911, 6000, 53, index_fields=("age", "role"))
    schema = streamgen.make_schema()
    batch = codec.encode_updates(schema, ops)
    eng, orc = pair(schema, 64)
    eng.index_create_fields((0, 2))
    orc.index_create(0)
    orc.index_create(2)
    assert eng.merge(batch).same_as(orc.merge(batch))
    router = shard.Router(1, 0, 1024, 0)
    stream = torch.cuda.Stream().cuda_stream
    router.query_reserve(4096)
    for lo, hi in itertools.product(BOUNDS[:6], BOUNDS[:6]):
        bl, bh = schema.bound(lo, False), schema.bound(hi, True)
        g = router.query_range(eng, 0, capi.bound_struct(bl), capi.bound_struct(bh), stream)
        want = orc.query_range(0, bl, bh)
        assert g.total == len(want) and g.offset[1] == g.total
        assert np.array_equal(np.sort(router.query_fetch(0, g.total)), np.sort(want)), (lo, hi)
    for v in EQ_VALUES:
        key = schema.index_key(v)
        if key is None:
            continue
        g = router.query_equals(eng, 2, key, stream)
        assert np.array_equal(np.sort(router.query_fetch(0, g.total)), np.sort(orc.query_equals(2, key))), v
    # a result larger than the reserved capacity is reported, not truncated
    small = shard.Router(1, 0, 1024, 0)
    small.query_reserve(1)
    with pytest.raises(capi.BulletB200Error):
        small.query_range(eng, 0, capi.bound_struct(schema.bound(-1e9, False)), capi.bound_struct(schema.bound(1e9, True)), stream)
    small.close()
    router.close()
    eng.close()


def exact_pair(schema, capacity, **kw):
    from bullet_js_b200.engine import Engine

    eng = (Engine.for_schema(schema, capacity, post_getdata=True, exact_order=True, **kw) if schema is not None
           else Engine(capacity, post_getdata=True, exact_order=True, **kw))
    return eng, TypedOracle(eng.cfg)


def check_exact_queries(schema, eng, orc, fields, bounds, stride=1):
    """equals / range as EXACT lists: the reference's (Map order of the buckets, Set order inside a bucket)."""
    n, unsorted = 0, 0
    for f in fields:
        for v in EQ_VALUES:
            key = schema.index_key(v)
            if key is None:
                continue
            got, want = eng.query_equals(f, key), orc.query_equals(f, key)
            assert got.tolist() == want.tolist(), (f, v)
        for lo, hi in list(itertools.product(bounds, bounds))[::stride]:
            bl, bh = schema.bound(lo, False), schema.bound(hi, True)
            got, want = eng.query_range(f, bl, bh), orc.query_range(f, bl, bh)
            assert got.tolist() == want.tolist(), (f, lo, hi)
            n += 1
            unsorted += len(want) > 1 and want.tolist() != sorted(want.tolist())
    return n, unsorted


@pytest.mark.parametrize("seed", range(2))
def test_exact_map_and_set_order(seed):
    """SURVEY 8f-1 on the device (BB_CFG_EXACT_ORDER): entry tags + the hook's effective add / remove events + the
    per-bucket replay after every batch give bb_query_equals / bb_query_range the reference's exact result LISTS -
    buckets in creation order (a bucket that emptied is re-created at the end), paths in insertion order, a node twice
    when it sits in two matching buckets - on the JS-semantics streams: indices created before the first put, a late
    index built from the store, batches of 1 to 1800 updates."""
    ops, _ref = streamgen.generate(320 + seed, 3000, 37, index_fields=("age", "role"), late_index={"score": 1200})
    schema = streamgen.make_schema()
    batch = codec.encode_updates(schema, ops)
    eng, orc = exact_pair(schema, 64)
    for x in (eng, orc):
        x.index_create(0)
        x.index_create(2)
    for lo, hi, late in ((0, 1, False), (1, 1200, False), (1200, 1201, True), (1201, 3000, False)):
        if late:
            eng.index_create(1)
            orc.index_create(1)
        assert eng.merge(batch.slice(lo, hi)).same_as(orc.merge(batch.slice(lo, hi))), (seed, lo, hi)
    same_rows(eng, orc, 64)
    n, unsorted = check_exact_queries(schema, eng, orc, (0, 1, 2), BOUNDS, stride=3)
    assert n > 500 and unsorted > 50  # the reference's order is not node order: the test would notice a plain multiset
    eng.close()


def test_exact_order_synthetic_hot_keys_and_clear():
    """The same on the synthetic schema: index build over 3 000 existing records (creation order), Zipf batches whose hot
    nodes pile up stale entries, then bb_table_clear and a fresh start."""
    n_rec = 3000
    rng = synth.rng_for(4, salt=9)
    table = synth.make_table(n_rec, rng)
    eng, orc = exact_pair(None, n_rec, **synth.synth_ranks(n_rec))
    ids = np.arange(n_rec, dtype=np.uint64)
    perm = rng.permutation(n_rec)  # creation order != path id order
    rows = table.rows.copy()
    rows["cseq"][perm] = 1 + np.arange(n_rec, dtype=np.uint64)
    eng.table_load(ids, rows)
    orc.load(ids, rows)
    schema = synth.synth_schema(n_rec)
    for f in (0, 2):
        eng.index_create(f, extra_capacity=1 << 18)
        orc.index_create(f)
    small = [20.0, 30.0, 0.0, 99.0, "admin", "user", "a", "zzz", None]
    check_exact_queries(schema, eng, orc, (0, 2), small)
    for keys in ("uniform", "zipf"):
        b = synth.make_batch(table, 30_000, rng, keys=keys)
        assert eng.merge(b).same_as(orc.merge(b))
        n, unsorted = check_exact_queries(schema, eng, orc, (0, 2), small)
        assert unsorted > 5
    eng.close()
