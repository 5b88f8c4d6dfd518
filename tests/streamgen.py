"""Random JS-level update streams that exercise every branch of the resolver.

Ops are generated in lock-step with the literal oracle so that incoming clocks can
be built *relative to the path's current clock* (dominating / historical /
concurrent / identical / identical-but-permuted), the mix SURVEY.md 8d asks for.
"""
from __future__ import annotations

import math
import random

from bullet_js_b200 import codec
from oracle.js_literal import RefBullet

FIELDS = ["age", "score", "role", "name"]
PEERS = ["p0", "p1", "p2", "p3", "p4", "p5", "p6", "p7"]
STRINGS = ["admin", "editor", "user", "Zed", "abc", "zzz", "[a", "été", "\U0001F600x", "～q"]
NUMS = [0, -0.0, 1, 2, 25, 30, 40, 99, -5, 3.5, 1e21, 1e-7, math.inf, -math.inf, math.nan, 123456.78]


def make_schema(local_peer="p0"):
    return codec.Schema(FIELDS, PEERS, codec.StringDict(STRINGS), local_peer)


def rand_prim(rng):
    r = rng.random()
    if r < 0.45:
        return float(rng.choice(NUMS))
    if r < 0.75:
        return rng.choice(STRINGS)
    if r < 0.9:
        return rng.random() < 0.5
    return None


def rand_record(rng):
    ks = rng.sample(FIELDS, rng.randint(0, len(FIELDS)))
    return {k: rand_prim(rng) for k in ks}


def rand_clock(rng, peers=PEERS, hi=4):
    ks = rng.sample(peers, rng.randint(0, min(4, len(peers))))
    return {k: float(rng.randint(1, hi)) for k in ks}


def relative_clock(rng, cur, mode):
    cur = dict(cur)
    if mode == "identical":
        return cur
    if mode == "permuted":
        items = list(cur.items())
        rng.shuffle(items)
        return dict(items)
    if mode == "dominating":
        c = dict(cur)
        k = rng.choice(PEERS)
        c[k] = c.get(k, 0.0) + float(rng.randint(1, 2))
        return c
    if mode == "historical":
        c = {k: v for k, v in cur.items()}
        ks = [k for k, v in c.items()]
        if not ks:
            return c
        k = rng.choice(ks)
        if c[k] > 1:
            c[k] -= 1.0
        else:
            del c[k]
        return c
    if mode == "concurrent":
        c = relative_clock(rng, cur, "historical")
        k = rng.choice(PEERS)
        c[k] = cur.get(k, 0.0) + 1.0
        return c
    raise ValueError(mode)


MODES = ["identical", "permuted", "dominating", "historical", "concurrent", "random"]


def generate(seed, n_ops, n_paths, local_peer="p0", index_fields=(), p_local=0.3, p_prim=0.2, late_index=None):
    """-> (ops, literal RefBullet after replay). ops = [(path, value, clock|None)].
    late_index: {field: op number} - index("users", field) is called just before that op."""
    rng = random.Random(seed)
    late_index = late_index or {}
    ref = RefBullet(local_peer, enable_indexing=bool(index_fields) or bool(late_index))
    for f in index_fields:
        ref.index("users", f)
    ops = []
    for k in range(n_ops):
        for f, at in late_index.items():
            if at == k:
                ref.index("users", f)
        path = f"users/u{rng.randrange(n_paths)}"
        is_prim = rng.random() < p_prim
        value = rand_prim(rng) if is_prim else rand_record(rng)
        clock = None
        if not is_prim and rng.random() > p_local:
            mode = rng.choice(MODES)
            cur = (ref.meta.get(path) or {}).get("vectorClock")
            clock = rand_clock(rng) if (mode == "random" or cur is None) else relative_clock(rng, cur, mode)
        ops.append((path, value, clock))
        apply_op(ref, ops[-1])
    return ops, ref


def apply_op(ref, op):
    path, value, clock = op
    if clock is not None and isinstance(value, dict):
        ref.process_sync_entries([dict(path=path, data=value, vectorClock=clock)])
    else:
        # local put, or a primitive arriving from the network (same flavour, sync:560-563)
        ref.setData(path, value, False)
