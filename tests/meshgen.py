"""BASELINE config 5 at full size: the per-peer logs of a full mesh, produced by the typed oracle (TEST INFRASTRUCTURE:
it runs oracle/bullet_oracle.c, the restated reference, as the simulated peers - SURVEY 8d "as produced by the oracle's
mesh simulator with a seeded delivery schedule").

tests/meshsim.py drives one event at a time (and can drive the reference itself: tests/golden/mesh.json.gz); ten-million-
entry logs need batches.  Same model, in rounds: in every round each peer makes `local_per_round` local puts (record
values, `BulletNode.put`, src/bullet.js:700-703); a put that is accepted is broadcast as the STORED value with its
clock (src/bullet.js:150-152, src/bullet-crt.js:371-376) and reaches every other peer in the next round (80 %) or the one
after (20 %, decided per receiver)
(`BulletNetwork._handlePut`, src/bullet-network.js:332-346), interleaved with that peer's own puts by a seeded random
merge that keeps every sender's order.  A peer's LOG - what it processed, in order - is what a B200 replica replays;
the oracle's decisions, change entries and final table for it are the expectation.
"""
from __future__ import annotations

import time

import numpy as np

from bullet_js_b200 import capi, codec, synth
from oracle.typed import TypedOracle

_K = np.array([0x9E3779B97F4A7C15, 0xBF58476D1CE4E5B9, 0x94D049BB133111EB, 0xD6E8FEB86659FD93, 0xC2B2AE3D27D4EB4F,
               0x165667B19E3779F9, 0x27D4EB2F165667C5, 0xFF51AFD7ED558CCD, 0xC4CEB9FE1A85EC53, 0x2545F4914F6CDD1D,
               0x9FB21C651E98DF25, 0xD1B54A32D192ED03], np.uint64)


def entries_checksum(idx, head, clk, val) -> int:
    """Order-independent 64-bit checksum of a set of change entries (idx included: which update an entry belongs to)."""
    k = len(idx)
    if k == 0:
        return 0
    with np.errstate(over="ignore"):
        w = np.zeros(k, np.uint64)
        w += np.asarray(idx).astype(np.uint64) * _K[0]
        h = np.asarray(head)
        w += h["hdr"] * _K[1] + h["clk_order"].astype(np.uint64) * _K[2]
        c = np.asarray(clk).reshape(k, 4, 2).astype(np.uint64)
        w += ((c[:, :, 0] | (c[:, :, 1] << np.uint64(32))) * _K[3:7][None, :]).sum(axis=1)
        w += (np.asarray(val).reshape(k, 4) * _K[7:11][None, :]).sum(axis=1)
        w ^= w >> np.uint64(29)
        w *= _K[11]
        w ^= w >> np.uint64(32)
        return int(w.sum())


def local_puts(image, n, rng) -> codec.Batch:
    """n local record puts on uniform paths (flavour local: no clock travels with a put)."""
    return synth.make_batch(image, n, rng, mix={**{k: 0.0 for k in synth.MIX}, "local": 1.0})


def _interleave(rng, streams):
    """Random merge of update streams that keeps each stream's own order."""
    keys = np.concatenate([np.sort(rng.random(s.n)) for s in streams])
    order = np.argsort(keys, kind="stable")
    cat = codec.Batch(*(np.concatenate([getattr(s, f) for s in streams]) for f in ("path_id", "head", "clk", "val")))
    out = codec.Batch(cat.path_id[order], cat.head[order], cat.clk[order], cat.val[order])
    origin = np.concatenate([np.full(s.n, i, np.int8) for i, s in enumerate(streams)])[order]
    return out, origin


def run_mesh_rounds(image, n_peers, rounds, local_per_round, seed, keep_logs=None, batch=1_000_000, threads=1):
    """-> dict: logs[p] (codec.Batch, only for p in keep_logs / all), expect[p] = per-`batch` slices of the log:
    [(n, decision histogram[7], n_entries, entries checksum)], tables[p] (final rows), cpu_seconds, cpu_updates."""
    n_rec = image.n
    ranks = synth.synth_ranks(n_rec)
    ids = np.arange(n_rec, dtype=np.uint64)
    peers = []
    for p in range(n_peers):
        o = TypedOracle(capi.make_config(n_rec, local_peer=p, **ranks))
        o.load(ids, image.rows)
        peers.append(o)
    keep = set(range(n_peers)) if keep_logs is None else set(keep_logs)
    pieces = {p: [] for p in keep}
    dec_all = {p: [] for p in range(n_peers)}
    ent_all = {p: [] for p in range(n_peers)}
    pending = [None] * n_peers  # what peer q broadcast in the previous round ...
    older = [None] * n_peers    # ... and in the round before: a fifth of every broadcast takes two rounds to arrive
    cpu_s, cpu_n = 0.0, 0
    for rnd in range(rounds):
        produced = []
        for p in range(n_peers):
            rng = synth.rng_for(5, salt=seed * 4096 + rnd * 64 + p)
            streams = [local_puts(image, local_per_round, rng)]
            for q in range(n_peers):  # late deliveries first: a sender's later put may overtake an earlier one (-> historical)
                for age, src in ((2, older[q]), (1, pending[q])):
                    if q == p or src is None or src.n == 0:
                        continue
                    slow = synth.rng_for(5, salt=seed * 4096 + (rnd - age) * 64 + q + 1_000_000 * (p + 1)).random(src.n) < 0.2
                    sel = np.nonzero(slow if age == 2 else ~slow)[0]
                    if sel.size:
                        streams.append(codec.Batch(src.path_id[sel], src.head[sel], src.clk[sel], src.val[sel]))
            log, origin = _interleave(rng, streams)
            log.head["user"] = np.arange(log.n, dtype=np.uint32)
            t0 = time.perf_counter()
            ch = peers[p].merge(log, threads=threads)
            cpu_s += time.perf_counter() - t0
            cpu_n += log.n
            # accepted LOCAL puts go out: stored value + stored clock, network flavour (records are objects)
            mine = origin[ch.idx.astype(np.int64)] == 0
            hd = ch.head[mine].copy()
            hd["hdr"] |= np.uint64(codec.HDR_FLAVOUR_NET)
            produced.append(codec.Batch(log.path_id[ch.idx[mine].astype(np.int64)].copy(), hd, ch.clk[mine].copy(), ch.val[mine].copy()))
            if p in keep:
                pieces[p].append(log)
            dec_all[p].append(ch.decision)
            ent_all[p].append((ch.idx, ch.head, ch.clk, ch.val, log.n))
        older, pending = pending, produced
    out = {"logs": {}, "expect": {}, "tables": {p: peers[p].table for p in range(n_peers)}, "cpu_seconds": cpu_s,
           "cpu_updates": cpu_n, "threads": threads}
    for p in range(n_peers):
        # expectation per replay batch: the log cut every `batch` updates, independent of the round structure
        dec = np.concatenate(dec_all[p])
        off, gi, gh, gc, gv = 0, [], [], [], []
        for idx, head, clk, val, n in ent_all[p]:
            gi.append(idx.astype(np.int64) + off)
            gh.append(head)
            gc.append(clk)
            gv.append(val)
            off += n
        gi, gh, gc, gv = np.concatenate(gi), np.concatenate(gh), np.concatenate(gc), np.concatenate(gv)
        exp = []
        for lo in range(0, len(dec), batch):
            hi = min(lo + batch, len(dec))
            a, b = np.searchsorted(gi, lo), np.searchsorted(gi, hi)
            exp.append((hi - lo, np.bincount(dec[lo:hi], minlength=7)[:7].tolist(), int(b - a),
                        entries_checksum((gi[a:b] - lo).astype(np.uint32), gh[a:b], gc[a:b], gv[a:b])))
        out["expect"][p] = exp
        if p in keep:
            ps = pieces[p]
            lg = codec.Batch(*(np.concatenate([getattr(s, f) for s in ps]) for f in ("path_id", "head", "clk", "val")))
            lg.head["user"] = np.arange(lg.n, dtype=np.uint32)  # the entry's `user` word: position in the whole log
            out["logs"][p] = lg
    return out
