"""bench.py --config mesh: BASELINE config 5 (mesh-topology sync replay, SURVEY 8d / 8e "replicas only").

8 simulated peers in a full mesh, each peer = one replica table; the per-peer logs (own puts interleaved with the other
seven peers' broadcasts, --mesh-ops entries each, 10 M by default) come from the typed oracle's mesh simulator
(tests/meshgen.py, seeded delivery schedule).  Peer p is replayed on GPU p % N in batches of 1 M updates - independent
replicas, no collective.  Every batch is checked against the oracle's replay of the same log: decision histogram,
number of change entries and an order-independent checksum of the entries; at the end every row of every replica's
table, bit for bit.  ("Converged" means "equals the reference replay": the rule is order-dependent, replicas differ.)

`value` = field-merges/s over all replicas with the logs resident in HBM; `e2e` = the same through bb_merge_batch from
pinned host memory (compact change set).
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np

F = 4


def run(args, rank, world, local_rank, helpers):
    import torch

    from bullet_js_b200 import capi, codec, synth
    from bullet_js_b200.engine import Engine
    from tests import meshgen

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    P, B = 8, args.mesh_batch
    rounds = 10
    local_per_round = max(1, args.mesh_ops // (rounds * P))  # a log entry = 1 own put or 1 of 7 peers' broadcasts
    image = synth.make_table(args.records, synth.rng_for(5))
    mine = [p for p in range(P) if p % world == rank]
    t0 = time.perf_counter()
    mesh = meshgen.run_mesh_rounds(image, P, rounds, local_per_round, seed=5, keep_logs=mine, batch=B,
                                   threads=max(1, (os.cpu_count() or 8) // max(world, 1)))
    gen_s = time.perf_counter() - t0
    ranks_kw = synth.synth_ranks(image.n)
    ids = np.arange(image.n, dtype=np.uint64)
    side = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(side)
    stream = side.cuda_stream
    sampler = helpers.ClockSampler(local_rank)

    def load(e):
        for o in range(0, image.n, 1 << 20):
            e.table_load(ids[o:o + (1 << 20)], image.rows[o:o + (1 << 20)])

    out = helpers.DevOut(torch, dev, B)
    parity_bad, dev_ms, ops, launches, acc = [], 0.0, 0, 0, 0
    e2e_s, e2e_ops, e2e_d2h = 0.0, 0, 0
    distinct = []
    sampler.start()
    for p in mine:
        log, expect = mesh["logs"][p], mesh["expect"][p]
        d = helpers.dev_batch(torch, dev, log)[0]  # the whole log resident in HBM (88 B per entry)
        for compact, timed_on_device in ((False, True), (True, False)):
            eng = Engine(image.n, device=local_rank, local_peer=p, compact_changes=compact, **ranks_kw)
            load(eng)
            eng.reserve(B, host_entry=not timed_on_device)
            l0 = eng.launch_count()
            if timed_on_device:
                # warm-up on a scratch replica, then the timed replay of the whole log, batch by batch
                warm = Engine(image.n, device=local_rank, local_peer=p, **ranks_kw)
                load(warm)
                for j in range(min(3, len(expect))):
                    warm.merge_dev(_slice(capi, d, j * B, expect[j][0]), out.cs, stream)
                warm.sync(stream)
                warm.close()
                torch.cuda.synchronize()
                for j, (n, hist, k, cs) in enumerate(expect):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    eng.merge_dev(_slice(capi, d, j * B, n), out.cs, stream)
                    e1.record()
                    eng.sync(stream)
                    dev_ms += e0.elapsed_time(e1)
                    ops += n
                    got = out.changes(n)  # (not timed) this batch against the oracle's replay
                    ok = np.bincount(got.decision, minlength=7)[:7].tolist() == hist and len(got.idx) == k and \
                        meshgen.entries_checksum(got.idx, got.head, got.clk, got.val) == cs
                    if not ok:
                        parity_bad.append(f"peer {p} batch {j}")
                    acc += k
                    distinct.append(np.unique(log.path_id[j * B: j * B + n]).size)
                launches += eng.launch_count() - l0
            else:
                pin = lambda a: torch.from_numpy(a.view(np.uint8).reshape(-1)).pin_memory()
                h = [pin(x) for x in (log.path_id, log.head, log.clk, log.val)]
                hp = lambda nbytes: torch.zeros(nbytes, dtype=torch.uint8).pin_memory()
                h_ver, h_n, h_idx, h_head, h_clk, h_val = hp(4 * B), hp(8), hp(4 * B), hp(16 * B), hp(32 * B), hp(32 * B)
                hcs = capi.BBChanges(cap=B, verdict=h_ver.data_ptr(), n_changes=h_n.data_ptr(), idx=h_idx.data_ptr(),
                                     head=h_head.data_ptr(), clk=h_clk.data_ptr(), val=h_val.data_ptr())
                torch.cuda.synchronize()
                for j, (n, hist, k, cs) in enumerate(expect):
                    o = j * B
                    hb = capi.BBBatch(n=n, path_id=h[0].data_ptr() + 8 * o, head=h[1].data_ptr() + 16 * o,
                                      clk=h[2].data_ptr() + 32 * o, val=h[3].data_ptr() + 32 * o)
                    t0 = time.perf_counter()
                    eng.merge_raw(hb, hcs)
                    e2e_s += time.perf_counter() - t0
                    kk = int(h_n.view(torch.int64)[0])
                    e2e_ops += n
                    e2e_d2h += 4 * n + 8 + 84 * kk
                    got = codec.Changes.from_verdicts(h_ver.numpy().view(np.uint32)[:n], h_idx.numpy().view(np.uint32)[:kk],
                                                      h_head.numpy().view(codec.HEAD_DTYPE)[:kk],
                                                      h_clk.numpy().view(np.uint32).reshape(-1, 8)[:kk],
                                                      h_val.numpy().view(np.uint64).reshape(-1, 4)[:kk], log.slice(o, o + n))
                    ok = np.bincount(got.decision, minlength=7)[:7].tolist() == hist and len(got.idx) == k and \
                        meshgen.entries_checksum(got.idx, got.head, got.clk, got.val) == cs
                    if not ok:
                        parity_bad.append(f"peer {p} batch {j} (host entry)")
            rows = eng.table_read(ids)
            if not np.array_equal(rows, mesh["tables"][p]):
                parity_bad.append(f"peer {p} table" + ("" if timed_on_device else " (host entry)"))
            eng.close()
        del d
    clocks = sampler.stop()
    t = torch.tensor([dev_ms, float(ops), e2e_s, float(e2e_ops), float(e2e_d2h), float(len(parity_bad)), float(launches), float(acc)],
                     device=dev, dtype=torch.float64)
    tmax, tsum = t.clone(), t.clone()
    if dist is not None:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
    steps = int(float(tsum[1]) // B)
    peak, peak_src = helpers.peaks()
    acc_frac = float(tsum[7]) / max(float(tsum[1]), 1.0)
    dn = float(np.mean(distinct)) / B if distinct else 0.0
    bpu = 84 + 68 * acc_frac + 256 * dn
    if rank == 0:
        line = {
            "metric": "crdt_field_merges_per_sec", "value": float(tsum[1]) * F / (float(tmax[0]) * 1e-3), "unit": "field-merges/s",
            "n_gpus": world, "steps": steps, "warmup": 3, "ms_per_step": float(tmax[0]) / max(1, len(mine) * len(mesh["expect"][mine[0]])),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32+f64", "data": "synthetic",
            "config": {"workload": f"config5: mesh-topology sync replay, {P} peers (full mesh) = {P} replicas of a {image.n}-record table, "
                                   f"{mesh['logs'][mine[0]].n}-entry log per peer (own puts + the other 7 peers' broadcasts, seeded delivery "
                                   f"schedule), replayed in batches of {B}; replicas only, peer p on GPU p % {world}",
                       "records": image.n, "batch": B, "peers": P, "log_entries_per_peer": int(mesh["logs"][mine[0]].n),
                       "l2": "320 MB table + 88 MB batch per step > 126 MB L2; consecutive batches of a log touch different rows"},
            "ops_per_sec_per_replica_gpu": float(tsum[1]) / world / (float(tmax[0]) * 1e-3),
            "e2e": {"value": float(tsum[3]) * F / float(tmax[2]), "unit": "field-merges/s", "h2d_bytes_per_step": B * 88,
                    "d2h_bytes_per_step": int(float(tsum[4]) / max(1.0, float(tsum[3]) / B)), "ms_per_step": float(tmax[2]) * 1e3 / max(1, len(mine) * len(mesh["expect"][mine[0]])),
                    "steps": steps, "api": "bb_merge_batch (pinned host log, synchronous; ctx with BB_CFG_COMPACT_CHANGES)"},
            "gpu_launches": int(float(tsum[6])),
            "parity": {"mesh": "ok" if float(tsum[5]) == 0 else "MISMATCH", "checked": f"{steps} batches on device + {steps} through the host "
                       f"entry: decision histogram, entry count, entry checksum; {P} final tables bit for bit, twice"},
            "roofline": {"bound": "hbm", "kernel": "k_merge_stage (+ front end: whole step)", "peak": peak, "unit": "GB/s", "peak_source": peak_src,
                         "bytes_per_update": bpu, "accepted_frac": acc_frac, "distinct_paths_per_update": dn,
                         "achieved": bpu * float(tsum[1]) / world / (float(tmax[0]) * 1e-3) / 1e9,
                         "frac": bpu * float(tsum[1]) / world / (float(tmax[0]) * 1e-3) / 1e9 / peak, "traffic": None},
            "cpu_baseline": {"value": mesh["cpu_updates"] * F / mesh["cpu_seconds"], "unit": "field-merges/s", "cores": mesh["threads"], "kind": "port",
                             "sample": f"the mesh simulation itself: {mesh['cpu_updates']} updates through oracle/bullet_oracle.c in {mesh['cpu_seconds']:.1f} s",
                             "reference_kind": helpers.REFERENCE_KIND},
            "clocks": clocks, "generation_seconds": gen_s, "reference_kind": helpers.REFERENCE_KIND,
        }
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    if parity_bad:
        print(f"[bench mesh] PARITY FAILURE: {parity_bad[:10]}", file=sys.stderr)
        sys.exit(1)


def _slice(capi, d, o, n):
    return capi.BBBatch(n=n, path_id=d[0].data_ptr() + 8 * o, head=d[1].data_ptr() + 16 * o, clk=d[2].data_ptr() + 32 * o,
                        val=d[3].data_ptr() + 32 * o)
