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


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        table = synth.make_table(N_REC, synth.rng_for(3))
        batch = synth.make_batch(table, N_UPD, synth.rng_for(3, salt=10 + rank), keys="zipf")
        ids = np.arange(rank, N_REC, world)
        cfg = capi.make_config(len(ids) + 1, **synth.synth_ranks(N_REC))
        orc = TypedOracle(cfg)
        orc.load(ids // world, table.rows[ids])
        ch, got = shard.route_on_host(world, rank, batch, dist, orc.merge)
        q.put((rank, orc.table[: len(ids)].copy(), ch.decision.copy(), got.path_id.copy(), got.head["user"].copy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_replay_equals_single_peer(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
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
        ids = np.arange(r, N_REC, world)
        assert np.array_equal(rows, ref.table[ids]), f"shard {r} differs from the single-peer table"
        # what the shard received, in order: source-rank-major, arrival order inside a source
        want_dec, want_path, want_user = [], [], []
        for src in range(world):
            mine = np.nonzero(batches[src].path_id % world == r)[0]
            want_dec.append(dec[src][mine])
            want_path.append(batches[src].path_id[mine] // world)
            want_user.append(batches[src].head["user"][mine])
        assert np.array_equal(lpath, np.concatenate(want_path))
        assert np.array_equal(user, np.concatenate(want_user))
        assert np.array_equal(decision, np.concatenate(want_dec))
