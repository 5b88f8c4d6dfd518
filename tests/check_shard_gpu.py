#!/usr/bin/env python
"""Multi-GPU parity of the sharded path (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tests/check_shard_gpu.py

Every rank owns the rows whose hashed id (bullet_js_b200/shard.py: shard_mix, KEY_BITS bits; KEY_BITS=0 in the
environment selects id % N) falls on it, submits its own seeded batches through the library's native router (fused
pack + all-to-all over NVLink) and merges what it receives on its GPU.  The rounds are PIPELINED the way the
bench pipelines them - batch r + 1 is routed while batch r is still being merged, through both receive slots - and
rank 1 is slowed down artificially, so that a peer that overwrote a receive slot a slow shard is still merging
would be caught.  Each rank then checks ITS shard bit for bit against one typed-oracle replay of the whole job
(round by round: rank 0's batch, then rank 1's, ... into the unsharded table - the order the router guarantees):
table rows, decisions and change entries in received order.  Exit code 0 only if every rank agrees.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from bullet_js_b200 import capi, codec, shard, synth  # noqa: E402
from bullet_js_b200.engine import Engine  # noqa: E402
from oracle.typed import TypedOracle  # noqa: E402

N_REC, N_UPD, ROUNDS = 200_000, 300_000, 4


def check_queries(rank, world, local, kb, table, ids, rows_local, router, stream):
    """Sharded equals / range through the library (bb_router_query_*): every rank scans its shard and pushes its hit
    ids into every rank's result buffer over NVLink.  Each rank checks the WHOLE gathered result - every rank's run -
    against a numpy evaluation of the table image (multisets per run; src/bullet-query.js:186-210, 221-261)."""
    all_ids = np.arange(N_REC, dtype=np.uint64)
    own = shard.owner_of(all_ids, world, kb)
    loc = shard.local_row(all_ids, world, kb)
    cap = shard.shard_capacity(world, kb, N_REC) + 1
    qe = Engine(cap, device=local, post_getdata=True, **synth.synth_ranks(N_REC))
    qe.table_load(rows_local, table.rows[ids])
    qe.index_create_fields((0, 2), extra_capacity=1 << 12)
    router.query_reserve(N_REC + 64)
    age = table.rows["val"][:, 0].view(np.float64)
    role = table.rows["val"][:, 2]
    ok = True
    cases = [("range(age,20,30)", (age >= 20.0) & (age <= 30.0), lambda: router.query_range(
                  qe, 0, capi.BBBound(num=20.0, rank=0, flags=0, reserved=0), capi.BBBound(num=30.0, rank=0, flags=0, reserved=0), stream)),
             ("range(age,200,300)", np.zeros(N_REC, bool), lambda: router.query_range(
                  qe, 0, capi.BBBound(num=200.0, rank=0, flags=0, reserved=0), capi.BBBound(num=300.0, rank=0, flags=0, reserved=0), stream)),
             ("equals(role,admin)", role == 0, lambda: router.query_equals(qe, 2, codec.KEY_STR | 0, stream)),
             ("range(age,1,99)", (age >= 1.0) & (age <= 99.0), lambda: router.query_range(
                  qe, 0, capi.BBBound(num=1.0, rank=0, flags=0, reserved=0), capi.BBBound(num=99.0, rank=0, flags=0, reserved=0), stream))]
    for name, sel, fn in cases:
        for rep in range(2):  # twice: the second call reuses the buffers and the epoch flags
            g = fn()
            got = router.query_fetch(0, g.total)
            good = True
            for q in range(world):
                run = np.sort(got[g.offset[q]: g.offset[q + 1]].astype(np.int64))
                want = np.sort(loc[sel & (own == q)].astype(np.int64))
                good = good and np.array_equal(run, want)
            good = good and g.total == int(sel.sum())
        print(f"[rank {rank}] sharded {name}: {g.total} hits gathered from {world} ranks: {'OK' if good else 'MISMATCH'}", flush=True)
        ok = ok and good
    qe.close()
    return ok


def check_host_entry(rank, world, local, kb, table, ids, rows_local, router, batches):
    """bb_router_merge_batch: host buffers in, this shard's verdicts + change entries out, pieces pipelined; a ctx with
    BB_CFG_COMPACT_CHANGES so that the echoed entries are rebuilt from what the shard RECEIVED.  Expectation: one oracle
    replaying rank 0's piece 0, rank 1's piece 0, ..., rank 0's piece 1, ... (include/bullet_b200.h)."""
    CH = 3
    eng = Engine(shard.shard_capacity(world, kb, N_REC) + 1, device=local, compact_changes=True, **synth.synth_ranks(N_REC))
    eng.table_load(rows_local, table.rows[ids])
    ref = TypedOracle(capi.make_config(N_REC, **synth.synth_ranks(N_REC)))
    ref.load(np.arange(N_REC), table.rows)
    cap = N_UPD * world
    ok = True
    for rnd in range(2):
        mine = batches[rnd][rank]
        out = capi.ChangeBuffers(cap)
        m, counts = router.merge_batch(eng, capi.batch_struct(mine), out.struct(), CH)
        chunk = -(-N_UPD // CH)
        recv, want_dec, want_entries = [], [], []
        for j in range(CH):
            for src in range(world):
                b = batches[rnd][src].slice(j * chunk, min((j + 1) * chunk, N_UPD))
                ch = ref.merge(b)
                sel = np.nonzero(shard.owner_of(b.path_id, world, kb) == rank)[0]
                good_count = int(counts[j, src]) == sel.size
                ok = ok and good_count
                recv.append(codec.Batch(shard.local_row(b.path_id[sel], world, kb), b.head[sel], b.clk[sel], b.val[sel]))
                want_dec.append(ch.decision[sel])
                pos = {int(i): t for t, i in enumerate(ch.idx.tolist())}
                for i in sel.tolist():
                    t = pos.get(i)
                    if t is not None:
                        want_entries.append((ch.head[t].tobytes(), ch.clk[t].tobytes(), ch.val[t].tobytes()))
        rb = codec.Batch(*(np.concatenate([getattr(x, f) for x in recv]) for f in ("path_id", "head", "clk", "val")))
        want_dec = np.concatenate(want_dec)
        good = m == rb.n == len(want_dec)
        if good:
            got = out.result(m, rb)  # echoed entries rebuilt from the received updates
            good = np.array_equal(got.decision, want_dec) and len(got.idx) == len(want_entries)
            good = good and [(got.head[t].tobytes(), got.clk[t].tobytes(), got.val[t].tobytes()) for t in range(len(got.idx))] == want_entries
            good = good and int(out.n_changes[0]) < len(want_entries)  # something was echoed
        print(f"[rank {rank}] host entry round {rnd}: received {m} in {CH} pieces, {int(out.n_changes[0])} entries shipped of "
              f"{len(want_entries)} accepted: {'OK' if good else 'MISMATCH'}", flush=True)
        ok = ok and good
    rows = eng.table_read(rows_local)
    good = np.array_equal(rows, ref.table[ids])
    print(f"[rank {rank}] shard table after the host-entry rounds: {'OK' if good else 'MISMATCH'}", flush=True)
    eng.close()
    return ok and good


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    kb = int(os.environ.get("KEY_BITS", str(int(np.ceil(np.log2(N_REC))))))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    table = synth.make_table(N_REC, synth.rng_for(3))
    all_ids = np.arange(N_REC, dtype=np.uint64)
    ids = all_ids[shard.owner_of(all_ids, world, kb) == rank].astype(np.int64)
    rows_local = shard.local_row(ids, world, kb)
    eng = Engine(shard.shard_capacity(world, kb, N_REC) + 1, device=local, **synth.synth_ranks(N_REC))
    eng.table_load(rows_local, table.rows[ids])
    eng.reserve(N_UPD * world, host_entry=False)
    router = shard.Router(world, rank, N_UPD, local, key_bits=kb)
    side = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(side)
    stream = side.cuda_stream

    keys = ["uniform", "zipf", "uniform", "zipf"]
    batches = [[synth.make_batch(table, N_UPD, synth.rng_for(3, salt=100 * rnd + 10 + r), keys=keys[rnd]) for r in range(world)]
               for rnd in range(ROUNDS)]
    d_in, outs = [], []
    cap = N_UPD * world
    for rnd in range(ROUNDS):
        mine = batches[rnd][rank]
        d = [torch.from_numpy(x.view(np.uint8).reshape(-1).copy()).to(dev) for x in (mine.path_id, mine.head, mine.clk, mine.val)]
        d_in.append((d, capi.BBBatch(n=mine.n, path_id=d[0].data_ptr(), head=d[1].data_ptr(), clk=d[2].data_ptr(), val=d[3].data_ptr())))
        o = {k: torch.zeros(cap * w, dtype=torch.uint8, device=dev) for k, w in (("ver", 4), ("idx", 4), ("head", 16), ("clk", 32), ("val", 32))}
        o["n"] = torch.zeros(1, dtype=torch.int64, device=dev)
        o["cs"] = capi.BBChanges(cap=cap, verdict=o["ver"].data_ptr(), n_changes=o["n"].data_ptr(), idx=o["idx"].data_ptr(),
                                 head=o["head"].data_ptr(), clk=o["clk"].data_ptr(), val=o["val"].data_ptr())
        outs.append(o)

    # pipelined: nothing is synchronised between the rounds; rank 1's compute stream is held back before every merge
    received = []
    router.route(d_in[0][1], 0)
    for rnd in range(ROUNDS):
        if rank == 1:
            torch.cuda._sleep(2_000_000_000 // 2)  # ~0.5 s of spinning on the merge stream: a slow shard
        received.append(router.merge(eng, rnd % 2, outs[rnd]["cs"], stream))
        if rnd + 1 < ROUNDS:
            router.route(d_in[rnd + 1][1], (rnd + 1) % 2)
    eng.sync(stream)
    torch.cuda.synchronize()

    ref = TypedOracle(capi.make_config(N_REC, **synth.synth_ranks(N_REC)))
    ref.load(np.arange(N_REC), table.rows)
    ok = True
    for rnd in range(ROUNDS):
        o, m = outs[rnd], received[rnd]
        k = int(o["n"].item())
        want_dec, want_entries = [], []
        for src in range(world):  # the single peer: rank 0's batch, then rank 1's, ...
            b = batches[rnd][src]
            ch = ref.merge(b)
            sel = np.nonzero(shard.owner_of(b.path_id, world, kb) == rank)[0]
            want_dec.append(ch.decision[sel])
            pos = {int(i): j for j, i in enumerate(ch.idx.tolist())}
            for i in sel.tolist():
                j = pos.get(i)
                if j is not None:
                    want_entries.append((ch.head[j].tobytes(), ch.clk[j].tobytes(), ch.val[j].tobytes()))
        want_dec = np.concatenate(want_dec)
        good = m == len(want_dec) and k == len(want_entries)
        if good:
            got = codec.Changes.from_verdicts(
                o["ver"].cpu().numpy().view(np.uint32)[:m], o["idx"].cpu().numpy().view(np.uint32)[:k],
                o["head"].cpu().numpy().view(codec.HEAD_DTYPE)[:k], o["clk"].cpu().numpy().view(np.uint32).reshape(-1, 8)[:k],
                o["val"].cpu().numpy().view(np.uint64).reshape(-1, 4)[:k])
            good = np.array_equal(got.decision, want_dec)
            # `user` of an entry is the source's own arrival index: identical; entries come in received order
            good = good and [(got.head[j].tobytes(), got.clk[j].tobytes(), got.val[j].tobytes()) for j in range(k)] == want_entries
        print(f"[rank {rank}] round {rnd} ({keys[rnd]}): received {m}, accepted {k}: {'OK' if good else 'MISMATCH'}", flush=True)
        ok = ok and good
    rows = eng.table_read(rows_local)
    good = np.array_equal(rows, ref.table[ids])
    print(f"[rank {rank}] shard table after {ROUNDS} pipelined rounds: {'OK' if good else 'MISMATCH'}", flush=True)
    ok = ok and good
    ok = check_queries(rank, world, local, kb, table, ids, rows_local, router, stream) and ok
    ok = check_host_entry(rank, world, local, kb, table, ids, rows_local, router, batches) and ok
    t = torch.tensor([0 if ok else 1], device=dev)
    dist.all_reduce(t)
    router.close()
    eng.close()
    dist.destroy_process_group()
    if int(t.item()):
        sys.exit(1)
    if rank == 0:
        print(f"sharded parity OK on {world} GPUs (key_bits {kb}, {ROUNDS} pipelined rounds, rank 1 slowed down)")


if __name__ == "__main__":
    main()
