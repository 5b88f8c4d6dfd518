"""Multi-rank routing on CPU: world_size 2 and 3 over gloo (127.0.0.1).  The exchange logic of
bullet_js_b200/shard.py (counts all-to-all, variable-size payload all-to-alls, owner = id % world,
local row = id // world, replay order = (source rank, arrival index)) is run with numpy packing
and the typed oracle as the per-shard merge, and compared with ONE oracle replaying rank 0's batch,
then rank 1's, ... into the unsharded table."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from bullet_js_b200 import capi, codec, shard, synth
from oracle.typed import TypedOracle

N_REC, N_UPD = 3000, 5000


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _shard_ids(rank, world, key_bits):
    """Path ids owned by `rank`, and their local rows."""
    all_ids = np.arange(N_REC, dtype=np.uint64)
    ids = all_ids[shard.owner_of(all_ids, world, key_bits) == rank]
    return ids.astype(np.int64), shard.local_row(ids, world, key_bits)


def _worker(rank, world, port, q, key_bits):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        table = synth.make_table(N_REC, synth.rng_for(3))
        batch = synth.make_batch(table, N_UPD, synth.rng_for(3, salt=10 + rank), keys="zipf")
        ids, rows = _shard_ids(rank, world, key_bits)
        cfg = capi.make_config(shard.shard_capacity(world, key_bits, N_REC) + 1, **synth.synth_ranks(N_REC))
        orc = TypedOracle(cfg)
        orc.load(rows, table.rows[ids])
        ch, got = shard.route_on_host(world, rank, batch, dist, orc.merge, key_bits=key_bits)
        q.put((rank, orc.table[rows.astype(np.int64)].copy(), ch.decision.copy(), got.path_id.copy(), got.head["user"].copy()))
    finally:
        dist.destroy_process_group()


KEY_BITS = int(np.ceil(np.log2(N_REC)))


@pytest.mark.parametrize("key_bits", [0, KEY_BITS])
@pytest.mark.parametrize("world", [2, 3])
def test_sharded_replay_equals_single_peer(world, key_bits):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, key_bits)) for r in range(world)]
    for p in procs:
        p.start()
    results = {}
    for _ in range(world):
        r = q.get(timeout=120)
        results[r[0]] = r[1:]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0

    # the single peer: same table, rank 0's batch then rank 1's ...
    table = synth.make_table(N_REC, synth.rng_for(3))
    ref = TypedOracle(capi.make_config(N_REC, **synth.synth_ranks(N_REC)))
    ref.load(np.arange(N_REC), table.rows)
    batches = [synth.make_batch(table, N_UPD, synth.rng_for(3, salt=10 + r), keys="zipf") for r in range(world)]
    dec = [ref.merge(b).decision for b in batches]
    for r in range(world):
        rows, decision, lpath, user = results[r]
        ids, _ = _shard_ids(r, world, key_bits)
        assert np.array_equal(rows, ref.table[ids]), f"shard {r} differs from the single-peer table"
        # what the shard received, in order: source-rank-major, arrival order inside a source
        want_dec, want_path, want_user = [], [], []
        for src in range(world):
            mine = np.nonzero(shard.owner_of(batches[src].path_id, world, key_bits) == r)[0]
            want_dec.append(dec[src][mine])
            want_path.append(shard.local_row(batches[src].path_id[mine], world, key_bits))
            want_user.append(batches[src].head["user"][mine])
        assert np.array_equal(lpath, np.concatenate(want_path))
        assert np.array_equal(user, np.concatenate(want_user))
        assert np.array_equal(decision, np.concatenate(want_dec))


def _worker_pieces(rank, world, port, q, key_bits, pieces):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        table = synth.make_table(N_REC, synth.rng_for(3))
        batch = synth.make_batch(table, N_UPD, synth.rng_for(3, salt=10 + rank), keys="zipf")
        ids, rows = _shard_ids(rank, world, key_bits)
        orc = TypedOracle(capi.make_config(shard.shard_capacity(world, key_bits, N_REC) + 1, **synth.synth_ranks(N_REC)))
        orc.load(rows, table.rows[ids])
        res = shard.merge_batch_on_host(world, rank, batch, dist, orc.merge, pieces, key_bits)
        q.put((rank, orc.table[rows.astype(np.int64)].copy(), np.concatenate([ch.decision for ch, _ in res]),
               np.concatenate([got.head["user"] for _, got in res])))
    finally:
        dist.destroy_process_group()


def test_piecewise_host_entry_order():
    """The replay order of bb_router_merge_batch (piece by piece, inside a piece by source rank), on 2 ranks over gloo:
    equal to ONE oracle replaying rank 0's piece 0, rank 1's piece 0, rank 0's piece 1, ..."""
    world, pieces, key_bits = 2, 3, KEY_BITS
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_pieces, args=(r, world, port, q, key_bits, pieces)) for r in range(world)]
    for p in procs:
        p.start()
    results = {}
    for _ in range(world):
        r = q.get(timeout=120)
        results[r[0]] = r[1:]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    table = synth.make_table(N_REC, synth.rng_for(3))
    ref = TypedOracle(capi.make_config(N_REC, **synth.synth_ranks(N_REC)))
    ref.load(np.arange(N_REC), table.rows)
    batches = [synth.make_batch(table, N_UPD, synth.rng_for(3, salt=10 + r), keys="zipf") for r in range(world)]
    chunk = -(-N_UPD // pieces)
    want = {r: ([], []) for r in range(world)}
    for j in range(pieces):
        for src in range(world):
            b = batches[src].slice(j * chunk, min((j + 1) * chunk, N_UPD))
            dec = ref.merge(b).decision
            for r in range(world):
                mine = np.nonzero(shard.owner_of(b.path_id, world, key_bits) == r)[0]
                want[r][0].append(dec[mine])
                want[r][1].append(b.head["user"][mine])
    for r in range(world):
        rows, decision, user = results[r]
        ids, _ = _shard_ids(r, world, key_bits)
        assert np.array_equal(rows, ref.table[ids])
        assert np.array_equal(decision, np.concatenate(want[r][0])) and np.array_equal(user, np.concatenate(want[r][1]))
