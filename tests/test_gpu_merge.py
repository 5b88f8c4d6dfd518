"""GPU parity: libbulletb200.so (through the C ABI) == the typed oracle, bit for bit."""
import numpy as np
import pytest

from bullet_js_b200 import capi, codec, synth
from oracle.typed import TypedOracle
from tests import streamgen

pytestmark = pytest.mark.gpu


def engine_and_oracle(schema=None, capacity=256, post_getdata=False, ordered=False, radix=False, **kw):
    """radix: False (default: grouping front end, hot keys handed to k_merge_hot) | True (radix sort) | "full" (counting
    sort by path id) | "hot" (the round-1 opt-in flag BB_CFG_HOT_KEYS: accepted and ignored)."""
    from bullet_js_b200.engine import Engine

    sort = dict(radix_sort=radix is True, full_sort=radix == "full", hot_keys=radix == "hot")
    if schema is not None:
        eng = Engine.for_schema(schema, capacity, post_getdata=post_getdata, ordered_changes=ordered, **sort)
    else:
        eng = Engine(capacity, post_getdata=post_getdata, ordered_changes=ordered, **sort, **kw)
    return eng, TypedOracle(eng.cfg)


def assert_same_table(eng, orc, n):
    ids = np.arange(n, dtype=np.uint64)
    got, want = eng.table_read(ids), orc.read(ids)
    bad = np.nonzero(got != want)[0]
    assert bad.size == 0, (bad[:5], got[bad[:2]], want[bad[:2]])


@pytest.mark.parametrize("seed", range(4))
@pytest.mark.parametrize("indexed", [False, True])
@pytest.mark.parametrize("radix", [False, True, "full", "hot"])
def test_random_js_streams(seed, indexed, radix):
    ops, _ref = streamgen.generate(100 + seed, 4000, 37, index_fields=("age",) if indexed else ())
    schema = streamgen.make_schema()
    batch = codec.encode_updates(schema, ops)
    eng, orc = engine_and_oracle(schema, 64, post_getdata=indexed, radix=radix)
    cuts = [0, 1, 2, 35, 36, 1000, 1001, 3000, len(ops)]
    seen = set()
    for lo, hi in zip(cuts, cuts[1:]):
        got, want = eng.merge(batch.slice(lo, hi)), orc.merge(batch.slice(lo, hi))
        assert got.same_as(want), (seed, lo, hi)
        seen |= set(got.decision.tolist())
    assert seen == set(range(7))
    assert_same_table(eng, orc, 64)
    eng.close()


def test_kat_l_on_gpu():
    schema = codec.Schema(["v"], ["A"], codec.StringDict([]), "A")
    ops = [("k/v", float(x), None) for x in (5, 3, 3, 3, 7, 9, 0, 0, 4)]
    batch = codec.encode_updates(schema, ops)
    eng, _ = engine_and_oracle(schema, 4)
    ch = eng.merge(batch)
    assert ch.decision.tolist() == [0, 3, 4, 1, 4, 2, 3, 4, 2]
    assert ch.idx.tolist() == [0, 2, 4, 5, 7, 8]
    d = codec.decode_row(schema, eng.table_read([0])[0])
    assert d["value"] == 4.0 and d["M"] == {"A": 11.0} and d["alias"]
    eng.close()


@pytest.mark.parametrize("mode", ["default", "ordered", "radix", "full", "hot"])
@pytest.mark.parametrize("keys", ["uniform", "zipf"])
def test_synthetic_schema_stream(keys, mode):
    """SURVEY 8d schema at a size the oracle replays in a second: 50k records, 3 x 200k updates."""
    n_rec = 50_000
    rng = synth.rng_for(2, salt=1)
    table = synth.make_table(n_rec, rng)
    ordered = mode == "ordered"
    eng, orc = engine_and_oracle(None, n_rec, ordered=ordered, radix={"radix": True, "full": "full", "hot": "hot"}.get(mode, False), **synth.synth_ranks(n_rec))
    ids = np.arange(n_rec, dtype=np.uint64)
    eng.table_load(ids, table.rows)
    orc.load(ids, table.rows)
    assert_same_table(eng, orc, n_rec)
    for _ in range(3):
        b = synth.make_batch(table, 200_000, rng, keys=keys)
        out = capi.ChangeBuffers(b.n)
        got, want = eng.merge(b, out), orc.merge(b)
        assert got.same_as(want)
        if ordered:  # BB_CFG_ORDERED_CHANGES: raw buffers are path-major, arrival order within a path
            raw = out.idx[: len(got.idx)].astype(np.int64)
            k = b.path_id[raw].astype(np.int64) * b.n + raw
            assert (np.diff(k) > 0).all()
        assert len(set(got.decision.tolist())) >= 4
    assert_same_table(eng, orc, n_rec)
    eng.close()


def test_fresh_table_and_local_only():
    """Every update lands on a path that was never written (M absent) or is a local put."""
    n_rec = 10_000
    rng = synth.rng_for(1, salt=3)
    table = synth.make_table(n_rec, rng)
    eng, orc = engine_and_oracle(None, n_rec, local_peer=3, **synth.synth_ranks(n_rec))
    for mix in (synth.MIX, dict(synth.MIX, local=0.7, dominating=0.1, historical=0.05, concurrent=0.05)):
        tot = sum(mix.values())
        mix = {k: v / tot for k, v in mix.items()}
        b = synth.make_batch(table, 50_000, rng, mix=mix)
        assert eng.merge(b).same_as(orc.merge(b))
    assert_same_table(eng, orc, n_rec)
    eng.close()


def test_empty_and_single():
    eng, orc = engine_and_oracle(None, 16)
    e = codec.Batch.empty(0)
    got = eng.merge(e)
    assert got.decision.size == 0 and got.idx.size == 0
    table = synth.make_table(16, synth.rng_for(0))
    b = synth.make_batch(table, 1, synth.rng_for(0, 1))
    assert eng.merge(b).same_as(orc.merge(b))
    eng.close()


def test_out_of_range_path_rejects_batch_and_keeps_table():
    n_rec = 1000
    rng = synth.rng_for(0, salt=9)
    table = synth.make_table(n_rec, rng)
    eng, orc = engine_and_oracle(None, n_rec)
    ids = np.arange(n_rec, dtype=np.uint64)
    eng.table_load(ids, table.rows)
    b = synth.make_batch(table, 5000, rng)
    b.path_id[1234] = n_rec  # one bad id
    with pytest.raises(capi.BulletB200Error) as ei:
        eng.merge(b)
    assert ei.value.code == capi.ERR_CAPACITY
    assert np.array_equal(eng.table_read(ids), table.rows)
    b.path_id[1234] = 0
    orc.load(ids, table.rows)
    assert eng.merge(b).same_as(orc.merge(b))  # the context stays usable
    eng.close()


def test_change_buffer_too_small():
    n_rec = 100
    rng = synth.rng_for(0, salt=5)
    table = synth.make_table(n_rec, rng)
    eng, _ = engine_and_oracle(None, n_rec)
    b = synth.make_batch(table, 2000, rng)
    out = capi.ChangeBuffers(2000)
    out.cap = 10
    with pytest.raises(capi.BulletB200Error) as ei:
        eng.merge(b, out)
    assert ei.value.code == capi.ERR_CAPACITY
    eng.close()


def test_table_read_materialise_matches_getdata():
    schema = streamgen.make_schema()
    ops = [("users/a", 0.0, None), ("users/b", {"age": 1.0}, None), ("users/c", False, None)]
    batch = codec.encode_updates(schema, ops)
    schema.paths.id("users/never")
    eng, orc = engine_and_oracle(schema, 8)
    assert eng.merge(batch).same_as(orc.merge(batch))
    ids = np.arange(4, dtype=np.uint64)
    got, want = eng.table_read(ids, materialise=True), orc.read(ids, materialise=True)
    kinds = [codec.decode_row(schema, r)["kind"] for r in got]
    assert kinds == [codec.KIND_OBJ] * 4
    assert [codec.decode_row(schema, r)["value"] for r in got] == [{}, {"age": 1.0}, {}, {}]
    assert np.array_equal(got["hdr"], want["hdr"]) and np.array_equal(got["val"], want["val"])
    eng.close()


def test_device_pointer_entry_matches_host_entry():
    import ctypes as C

    import torch

    n_rec = 20_000
    rng = synth.rng_for(2, salt=7)
    table = synth.make_table(n_rec, rng)
    eng, orc = engine_and_oracle(None, n_rec)
    ids = np.arange(n_rec, dtype=np.uint64)
    eng.table_load(ids, table.rows)
    orc.load(ids, table.rows)
    b = synth.make_batch(table, 100_000, rng, keys="zipf")
    want = orc.merge(b)
    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(a.view(np.uint8).reshape(-1)).to(dev)
    d_path, d_head, d_clk, d_val = t(b.path_id), t(b.head), t(b.clk), t(b.val)
    n = b.n
    o_ver = torch.zeros(n, dtype=torch.int32, device=dev)
    o_n = torch.zeros(1, dtype=torch.int64, device=dev)
    o_idx = torch.zeros(n, dtype=torch.int32, device=dev)
    o_head = torch.zeros(n * 16, dtype=torch.uint8, device=dev)
    o_clk = torch.zeros(n * 32, dtype=torch.uint8, device=dev)
    o_val = torch.zeros(n * 32, dtype=torch.uint8, device=dev)
    bs = capi.BBBatch(n=n, path_id=d_path.data_ptr(), head=d_head.data_ptr(), clk=d_clk.data_ptr(),
                      val=d_val.data_ptr())
    cs = capi.BBChanges(cap=n, verdict=o_ver.data_ptr(), n_changes=o_n.data_ptr(), idx=o_idx.data_ptr(),
                        head=o_head.data_ptr(), clk=o_clk.data_ptr(), val=o_val.data_ptr())
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    stream = side.cuda_stream
    eng.phase_events(True)  # the event between the front end and the merge (off by default: it serialises them)
    eng.merge_dev(bs, cs, stream)
    eng.sync(stream)
    k = int(o_n.item())
    got = codec.Changes.from_verdicts(
        o_ver.cpu().numpy().view(np.uint32), o_idx[:k].cpu().numpy().view(np.uint32),
        o_head[: k * 16].cpu().numpy().view(codec.HEAD_DTYPE),
        o_clk[: k * 32].cpu().numpy().view(np.uint32).reshape(k, 8),
        o_val[: k * 32].cpu().numpy().view(np.uint64).reshape(k, 4))
    assert got.same_as(want)
    assert_same_table(eng, orc, n_rec)
    assert eng.launch_count() > 0 and eng.phase_ms("merge") > 0 and eng.phase_ms("sort") > 0
    eng.close()


@pytest.mark.parametrize("n_rec", [40, 3000])
def test_hot_keys_kernel(n_rec):
    """BB_CFG_HOT_KEYS: segments of thousands (n_rec = 40) or tens (3000) of updates, local puts in between,
    replayed round by round by k_merge_hot - decisions, change set and rows as the sequential oracle."""
    rng = synth.rng_for(2, salt=21)
    table = synth.make_table(n_rec, rng)
    eng, orc = engine_and_oracle(None, n_rec, radix="hot", **synth.synth_ranks(n_rec))
    ids = np.arange(n_rec, dtype=np.uint64)
    eng.table_load(ids, table.rows)
    orc.load(ids, table.rows)
    mix = dict(synth.MIX, local=0.15, dominating=0.3)
    tot = sum(mix.values())
    mix = {k: v / tot for k, v in mix.items()}
    for keys in ("uniform", "zipf"):
        b = synth.make_batch(table, 120_000, rng, keys=keys, mix=mix)
        out = capi.ChangeBuffers(b.n)
        assert eng.merge(b, out).same_as(orc.merge(b))
    assert_same_table(eng, orc, n_rec)
    eng.close()


def test_chunked_host_call_rejects_whole_batch():
    """bb_merge_batch pipelines big batches in chunks; one bad path id in the LAST chunk must still
    leave the table untouched (include/bullet_b200.h: batch rejected whole)."""
    n_rec = 1000
    rng = synth.rng_for(2, salt=7)
    table = synth.make_table(n_rec, rng)
    eng, orc = engine_and_oracle(None, n_rec, **synth.synth_ranks(n_rec))
    ids = np.arange(n_rec, dtype=np.uint64)
    eng.table_load(ids, table.rows)
    orc.load(ids, table.rows)
    b = synth.make_batch(table, 300_000, rng)
    b.path_id[-1] = n_rec
    with pytest.raises(capi.BulletB200Error) as e:
        eng.merge(b)
    assert e.value.code == capi.ERR_CAPACITY
    assert_same_table(eng, orc, n_rec)
    b.path_id[-1] = 0
    assert eng.merge(b).same_as(orc.merge(b))  # the ctx is still usable, and chunking changes nothing
    assert_same_table(eng, orc, n_rec)
    eng.close()


@pytest.mark.parametrize("world", [1, 2, 8])
def test_route_pack_is_a_stable_partition(world):
    """bb_route_pack_dev (send side of the shard routing) == numpy stable partition by id % world."""
    import ctypes as C

    import torch

    from bullet_js_b200.engine import Engine

    n_rec, n = 4000, 70_001
    rng = synth.rng_for(2, salt=9)
    table = synth.make_table(n_rec, rng)
    b = synth.make_batch(table, n, rng, keys="zipf")
    eng = Engine(n_rec, **synth.synth_ranks(n_rec))
    dev = torch.device("cuda", 0)
    to_dev = lambda a: torch.from_numpy(a.view(np.uint8).reshape(-1).copy()).to(dev)
    src = [to_dev(x) for x in (b.path_id, b.head, b.clk, b.val)]
    dst = [torch.zeros_like(t) for t in src]
    counts = torch.zeros(world, dtype=torch.int64, device=dev)
    mk = lambda ts: capi.BBBatch(n=n, path_id=ts[0].data_ptr(), head=ts[1].data_ptr(), clk=ts[2].data_ptr(),
                                 val=ts[3].data_ptr())
    torch.cuda.synchronize()
    eng.route_pack_dev(world, mk(src), mk(dst), counts.data_ptr())
    eng.sync()
    owner = (b.path_id % np.uint64(world)).astype(np.int64)
    order = np.argsort(owner, kind="stable")
    assert counts.cpu().tolist() == np.bincount(owner, minlength=world).tolist()
    got = [t.cpu().numpy() for t in dst]
    assert np.array_equal(got[0].view(np.uint64), b.path_id[order] // np.uint64(world))
    assert np.array_equal(got[1].view(codec.HEAD_DTYPE), b.head[order])
    assert np.array_equal(got[2].view(np.uint32).reshape(n, 8), b.clk[order])
    assert np.array_equal(got[3].view(np.uint64).reshape(n, 4), b.val[order])
    eng.close()


@pytest.mark.parametrize("p2p", [True, False])
def test_native_router_single_rank(p2p, monkeypatch):
    """bb_router_* with world == 1 (its own NCCL communicator): route -> acquire -> merge -> release
    gives what a direct merge gives, for both receive slots; fused peer-store path and the
    pack + ncclSend/ncclRecv path."""
    import ctypes as C

    if not p2p:
        monkeypatch.setenv("BB_ROUTER_NO_P2P", "1")

    import torch

    from bullet_js_b200.engine import Engine

    n_rec, n = 4000, 50_000
    rng = synth.rng_for(2, salt=11)
    table = synth.make_table(n_rec, rng)
    ids = np.arange(n_rec, dtype=np.uint64)
    eng, orc = engine_and_oracle(None, n_rec, **synth.synth_ranks(n_rec))
    eng.table_load(ids, table.rows)
    orc.load(ids, table.rows)
    lib = capi.load()
    idbuf = C.create_string_buffer(capi.NCCL_ID_BYTES)
    assert lib.bb_router_unique_id(idbuf) == 0, lib.bb_router_last_error(None)
    h = C.c_void_p()
    assert lib.bb_router_create(0, 1, 0, idbuf.raw, n, 0, C.byref(h)) == 0, lib.bb_router_last_error(None)
    dev = torch.device("cuda", 0)
    st = torch.cuda.Stream(device=dev)
    for step in range(3):
        b = synth.make_batch(table, n, rng, keys="zipf")
        t = [torch.from_numpy(a.view(np.uint8).reshape(-1).copy()).to(dev) for a in (b.path_id, b.head, b.clk, b.val)]
        torch.cuda.synchronize()
        bs = capi.BBBatch(n=n, path_id=t[0].data_ptr(), head=t[1].data_ptr(), clk=t[2].data_ptr(), val=t[3].data_ptr())
        got_n = C.c_uint64(0)
        slot = step % 2
        assert lib.bb_router_route_dev(h, C.byref(bs), slot, C.byref(got_n), None) == 0, lib.bb_router_last_error(h)
        # peer-store path: the route is asynchronous, the received count is known at acquire (2**64 - 1 until then)
        assert got_n.value == (2 ** 64 - 1 if p2p else n)
        rb = capi.BBBatch()
        assert lib.bb_router_acquire(h, slot, C.c_void_p(st.cuda_stream), C.byref(rb)) == 0
        assert rb.n == n
        o = [torch.zeros(n * w, dtype=torch.uint8, device=dev) for w in (4, 8, 4, 16, 32, 32)]
        cs = capi.BBChanges(cap=n, verdict=o[0].data_ptr(), n_changes=o[1].data_ptr(), idx=o[2].data_ptr(),
                            head=o[3].data_ptr(), clk=o[4].data_ptr(), val=o[5].data_ptr())
        eng.merge_dev(rb, cs, st.cuda_stream)
        assert lib.bb_router_release(h, slot, C.c_void_p(st.cuda_stream)) == 0
        eng.sync(st.cuda_stream)
        want = orc.merge(b)
        k = int(o[1].cpu().numpy().view(np.uint64)[0])
        got = codec.Changes.from_verdicts(o[0].cpu().numpy().view(np.uint32), o[2].cpu().numpy().view(np.uint32)[:k],
                                          o[3].cpu().numpy().view(codec.HEAD_DTYPE)[:k],
                                          o[4].cpu().numpy().view(np.uint32).reshape(n, 8)[:k],
                                          o[5].cpu().numpy().view(np.uint64).reshape(n, 4)[:k])
        assert got.same_as(want)
    assert_same_table(eng, orc, n_rec)
    assert lib.bb_router_sent_bytes(h) == 0
    lib.bb_router_destroy(h)
    eng.close()


@pytest.mark.parametrize("indexed", [False, True])
def test_compact_change_set(indexed):
    """BB_CFG_COMPACT_CHANGES: an accepted update whose stored (value, clock) is the update itself gets slot
    BB_SLOT_ECHO and no entry; rebuilt from the caller's own input (codec.Changes.from_verdicts) the change set is the
    oracle's, entry for entry - JS streams with every decision code, then big uniform + Zipf batches (hot-key kernel,
    segments past their tile, chunked host call)."""
    from bullet_js_b200.engine import Engine

    ops, _ref = streamgen.generate(140, 4000, 37, index_fields=("age",) if indexed else ())
    schema = streamgen.make_schema()
    batch = codec.encode_updates(schema, ops)
    eng = Engine.for_schema(schema, 64, post_getdata=indexed, compact_changes=True)
    orc = TypedOracle(eng.cfg)
    if indexed:
        eng.index_create(0)
        orc.index_create(0)
    got, want = eng.merge(batch), orc.merge(batch)
    assert got.same_as(want)
    assert 0 < eng.last_emitted < len(want.idx)  # something was echoed, something was not
    assert_same_table_x(eng, orc, 64)
    eng.close()

    n_rec = 3000
    rng = synth.rng_for(2, salt=33)
    table = synth.make_table(n_rec, rng)
    eng = Engine(n_rec, post_getdata=indexed, compact_changes=True, **synth.synth_ranks(n_rec))
    orc = TypedOracle(eng.cfg)
    ids = np.arange(n_rec, dtype=np.uint64)
    eng.table_load(ids, table.rows)
    orc.load(ids, table.rows)
    if indexed:
        eng.index_create(0, extra_capacity=1 << 20)
        orc.index_create(0)
    for keys in ("uniform", "zipf"):
        b = synth.make_batch(table, 150_000, rng, keys=keys)
        got, want = eng.merge(b), orc.merge(b)
        assert got.same_as(want), keys
        assert eng.last_emitted <= len(want.idx) and (keys == "zipf" or eng.last_emitted < len(want.idx))
    assert_same_table_x(eng, orc, n_rec)
    eng.close()


def assert_same_table_x(eng, orc, n):
    ids = np.arange(n, dtype=np.uint64)
    got, want = eng.table_read(ids), orc.read(ids)
    got["xcnt"] = 0  # device-private index bookkeeping
    want["xcnt"] = 0
    assert np.array_equal(got, want)


def test_compact_needs_the_default_pipeline():
    from bullet_js_b200.engine import Engine

    with pytest.raises(capi.BulletB200Error):
        Engine(16, compact_changes=True, ordered_changes=True)


@pytest.mark.parametrize("compact", [False, True])
def test_router_host_entry_single_rank(compact):
    """bb_router_merge_batch on a one-rank router: host buffers in, pieces pipelined through both receive slots,
    verdicts + entries out in replay order (piece by piece) - the oracle merging the pieces one after the other."""
    from bullet_js_b200 import shard
    from bullet_js_b200.engine import Engine

    n_rec, n, pieces = 3000, 200_000, 3
    rng = synth.rng_for(2, salt=35)
    table = synth.make_table(n_rec, rng)
    eng = Engine(n_rec, compact_changes=compact, **synth.synth_ranks(n_rec))
    orc = TypedOracle(eng.cfg)
    ids = np.arange(n_rec, dtype=np.uint64)
    eng.table_load(ids, table.rows)
    orc.load(ids, table.rows)
    router = shard.Router(1, 0, n, 0)
    chunk = -(-n // pieces)
    for keys in ("uniform", "zipf"):
        b = synth.make_batch(table, n, rng, keys=keys)
        out = capi.ChangeBuffers(n)
        m, counts = router.merge_batch(eng, capi.batch_struct(b), out.struct(), pieces)
        assert m == n and counts.ravel().tolist() == [chunk, chunk, n - 2 * chunk]
        got = out.result(m, b)
        want = [orc.merge(b.slice(j * chunk, min((j + 1) * chunk, n))) for j in range(pieces)]
        assert np.array_equal(got.decision, np.concatenate([w.decision for w in want]))
        assert np.array_equal(got.idx, np.concatenate([w.idx + j * chunk for j, w in enumerate(want)]))
        for f in ("head", "clk", "val"):
            assert np.array_equal(getattr(got, f), np.concatenate([getattr(w, f) for w in want])), f
    assert_same_table(eng, orc, n_rec)
    # too small a change buffer is reported
    small = capi.ChangeBuffers(1000)
    with pytest.raises(capi.BulletB200Error):
        router.merge_batch(eng, capi.batch_struct(b), small.struct(), pieces)
    router.close()
    eng.close()


@pytest.mark.parametrize("radix", [False, True, "full"])
def test_far_out_of_range_id_in_a_smaller_second_batch(radix):
    """A rejected batch must not make the merge kernels load through a stale / half-written item list (advisor,
    round 1): a first batch leaves item lists and scratch behind, a SMALLER second batch carries path id 2**31 (far
    outside the table, hundreds of GB away as a row address) - documented outcome: BB_ERR_CAPACITY, table unchanged,
    context usable.  Every front end."""
    n_rec = 2000
    rng = synth.rng_for(0, salt=19)
    table = synth.make_table(n_rec, rng)
    eng, orc = engine_and_oracle(None, n_rec, radix=radix, **synth.synth_ranks(n_rec))
    ids = np.arange(n_rec, dtype=np.uint64)
    eng.table_load(ids, table.rows)
    orc.load(ids, table.rows)
    b1 = synth.make_batch(table, 40_000, rng)
    assert eng.merge(b1).same_as(orc.merge(b1))
    b2 = synth.make_batch(table, 3_000, rng)
    for bad in (2 ** 31, 2 ** 40 + 7, 2 ** 63):
        b2.path_id[777] = bad
        with pytest.raises(capi.BulletB200Error) as ei:
            eng.merge(b2)
        assert ei.value.code == capi.ERR_CAPACITY
        assert np.array_equal(eng.table_read(ids), orc.read(ids))
    b2.path_id[777] = 5
    assert eng.merge(b2).same_as(orc.merge(b2))
    assert_same_table(eng, orc, n_rec)
    eng.close()
