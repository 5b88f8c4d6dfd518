"""BASELINE config 5 in miniature: a full mesh of simulated peers (SURVEY 8d "Config 5").

Every peer is one replica (`RefBullet`, or the reference itself through `oracle.ref_runner.JSRefBullet`).
A local `node.put` that is accepted is broadcast (`src/bullet.js:150-152`: `network.broadcast(path,
broadcastData)`, the stored value plus `__vectorClock`, `src/bullet-crt.js:371-376`) and arrives at every
other peer after a random delay as `BulletNetwork._handlePut` would apply it (`src/bullet-network.js:332-346`;
relays are not modelled: in a full mesh every peer hears the originator directly).  The per-peer LOG - its own
puts interleaved with what it received, in processing order - is what a B200 replica replays in batches.
"""
from __future__ import annotations

import copy
import heapq
import random

from tests import streamgen


def _n_changes(p):
    return p.n_changes() if hasattr(p, "n_changes") else len(p.changes)


def _last_change(p):
    return p.last_change() if hasattr(p, "last_change") else p.changes[-1]


def run_mesh(make_peer, n_peers, n_ops, n_paths, seed, p_prim=0.1, mean_delay=3.0):
    """-> (peers, logs): logs[i] = [(path, value, clock | None)] in the order peer i processed them
    (the op format of tests/streamgen.py: replay with streamgen.apply_op or codec.encode_updates)."""
    rng = random.Random(seed)
    ids = [f"p{i}" for i in range(n_peers)]
    peers = [make_peer(i) for i in ids]
    logs = [[] for _ in peers]
    heap, t = [], 0.0
    for k in range(n_ops):
        t += rng.random()
        path = f"users/u{rng.randrange(n_paths)}"
        value = streamgen.rand_prim(rng) if rng.random() < p_prim else streamgen.rand_record(rng)
        heapq.heappush(heap, (t, k, rng.randrange(n_peers), "local", path, value, None))
    seq = n_ops
    while heap:
        when, _, i, kind, path, value, clock = heapq.heappop(heap)
        peer = peers[i]
        if kind == "local":
            before = _n_changes(peer)
            peer.put(path, copy.deepcopy(value))  # BulletNode.put -> setData(path, data) with broadcast = true
            logs[i].append((path, value, None))
            if _n_changes(peer) > before:  # doUpdate: the stored value and its clock go out to everybody
                ch = _last_change(peer)
                bval = copy.deepcopy(ch["value"])
                bclock = {k: float(v) for k, v in ch["vectorClock"].items()}
                for j in range(n_peers):
                    if j != i:
                        heapq.heappush(heap, (when + rng.expovariate(1.0 / mean_delay), seq, j, "recv", path, bval, bclock))
                        seq += 1
        else:
            if isinstance(value, dict):
                peer.handle_put(path, {**copy.deepcopy(value), "__vectorClock": dict(clock)})
                logs[i].append((path, value, clock))
            else:  # primitives travel without a clock and are applied like a local put (network:339-342)
                peer.handle_put(path, value)
                logs[i].append((path, value, None))
    return peers, logs
