#!/usr/bin/env python
"""Multi-GPU parity of the sharded path (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tests/check_shard_gpu.py

Every rank owns the rows `id % N == rank`, submits its own seeded batch through the library's native
router (fused pack + all-to-all over NVLink) and merges what it receives on its GPU.  Each rank then
checks ITS shard bit for bit against one typed-oracle replay of the whole job (rank 0's batch, then
rank 1's, ... into the unsharded table - the order the router guarantees): table rows, decisions and
change entries in received order.  Two rounds through both receive slots, second one zipf-keyed.
Exit code 0 only if every rank agrees.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from bullet_js_b200 import capi, codec, synth  # noqa: E402
from bullet_js_b200.engine import Engine  # noqa: E402
from bullet_js_b200.shard import Router  # noqa: E402
from oracle.typed import TypedOracle  # noqa: E402

N_REC, N_UPD = 200_000, 300_000


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    table = synth.make_table(N_REC, synth.rng_for(3))
    ids = np.arange(rank, N_REC, world)
    eng = Engine(len(ids) + 1, device=local, **synth.synth_ranks(N_REC))
    eng.table_load((ids // world).astype(np.uint64), table.rows[ids])
    eng.reserve(N_UPD * world, host_entry=False)
    router = Router(world, rank, N_UPD, local)
    side = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(side)
    stream = side.cuda_stream

    ref = TypedOracle(capi.make_config(N_REC, **synth.synth_ranks(N_REC)))
    ref.load(np.arange(N_REC), table.rows)
    ok = True
    for rnd, keys in enumerate(("uniform", "zipf")):
        batches = [synth.make_batch(table, N_UPD, synth.rng_for(3, salt=100 * rnd + 10 + r), keys=keys) for r in range(world)]
        mine = batches[rank]
        d = [torch.from_numpy(x.view(np.uint8).reshape(-1).copy()).to(dev) for x in (mine.path_id, mine.head, mine.clk, mine.val)]
        bs = capi.BBBatch(n=mine.n, path_id=d[0].data_ptr(), head=d[1].data_ptr(), clk=d[2].data_ptr(), val=d[3].data_ptr())
        cap = N_UPD * world
        out = {k: torch.zeros(cap * w, dtype=torch.uint8, device=dev) for k, w in (("ver", 4), ("idx", 4), ("head", 16), ("clk", 32), ("val", 32))}
        o_n = torch.zeros(1, dtype=torch.int64, device=dev)
        cs = capi.BBChanges(cap=cap, verdict=out["ver"].data_ptr(), n_changes=o_n.data_ptr(), idx=out["idx"].data_ptr(),
                            head=out["head"].data_ptr(), clk=out["clk"].data_ptr(), val=out["val"].data_ptr())
        slot = rnd % 2
        router.route(bs, slot, stream)
        m = router.merge(eng, slot, cs, stream)
        eng.sync(stream)
        k = int(o_n.item())
        got = codec.Changes.from_verdicts(
            out["ver"].cpu().numpy().view(np.uint32)[:m], out["idx"].cpu().numpy().view(np.uint32)[:k],
            out["head"].cpu().numpy().view(codec.HEAD_DTYPE)[:k], out["clk"].cpu().numpy().view(np.uint32).reshape(-1, 8)[:k],
            out["val"].cpu().numpy().view(np.uint64).reshape(-1, 4)[:k])

        # the single peer: rank 0's batch, then rank 1's, ...
        want_dec, want_entries = [], []
        for src in range(world):
            ch = ref.merge(batches[src])
            sel = np.nonzero(batches[src].path_id % world == rank)[0]
            want_dec.append(ch.decision[sel])
            pos = {int(i): j for j, i in enumerate(ch.idx.tolist())}
            for i in sel.tolist():
                j = pos.get(i)
                if j is not None:
                    want_entries.append((ch.head[j].tobytes(), ch.clk[j].tobytes(), ch.val[j].tobytes()))
        want_dec = np.concatenate(want_dec)
        good = m == len(want_dec) and np.array_equal(got.decision, want_dec) and k == len(want_entries)
        if good:
            have = [(got.head[j].tobytes(), got.clk[j].tobytes(), got.val[j].tobytes()) for j in range(k)]
            # `user` of an entry is the source's own arrival index: identical; entries come in received order
            good = have == want_entries
        rows = eng.table_read((ids // world).astype(np.uint64))
        good = good and np.array_equal(rows, ref.table[ids])
        print(f"[rank {rank}] round {rnd} ({keys}): received {m}, accepted {k}: {'OK' if good else 'MISMATCH'}", flush=True)
        ok = ok and good
    t = torch.tensor([0 if ok else 1], device=dev)
    dist.all_reduce(t)
    router.close()
    eng.close()
    dist.destroy_process_group()
    if int(t.item()):
        sys.exit(1)
    if rank == 0:
        print(f"sharded parity OK on {world} GPUs")


if __name__ == "__main__":
    main()
