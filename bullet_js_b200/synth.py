"""Seeded synthetic tables and update streams (SURVEY.md 8d schema), vectorised.

Collection `users`, record path `users/u<id>` interned to the dense path id `id`;
F=4 fields: age (int 1..99 as f64), score (f64, 2 decimals), role (admin 10 % /
editor 20 % / user 70 %), name ("name%07d").  P=8 peers "p0".."p7", local peer p0.
String ids follow UTF-16 order: admin=0, editor=1, name0000000..=2+i, user=2+N.

Incoming clocks are built relative to the clock the path holds in the *initial*
table: 40 % dominating, 20 % historical, 25 % concurrent, 5 % identical (same key
order), 5 % identical counts with permuted key order, 5 % local puts.
Everything is canonical in the sense of include/bullet_b200.h (absent slots and
unused order nibbles are zero).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import codec

P, F = codec.MAX_PEERS, codec.MAX_FIELDS
SEED_BASE = 0xB2000000

FULL_RECORD_HDR = (
    (codec.KIND_OBJ << codec.HDR_KIND_SHIFT)
    | (codec.TAG_NUM << (codec.HDR_TAG_SHIFT + 0))
    | (codec.TAG_NUM << (codec.HDR_TAG_SHIFT + 3))
    | (codec.TAG_STR << (codec.HDR_TAG_SHIFT + 6))
    | (codec.TAG_STR << (codec.HDR_TAG_SHIFT + 9))
    | (0x3210 << codec.HDR_ORDER_SHIFT)
)

MIX = dict(dominating=0.40, historical=0.20, concurrent=0.25, identical=0.05, permuted=0.05, local=0.05)


def rng_for(config_index: int, salt: int = 0) -> np.random.Generator:
    return np.random.Generator(np.random.PCG64(SEED_BASE + config_index + (salt << 8)))


def _pack_order(slots_by_pos: np.ndarray, nk: np.ndarray) -> np.ndarray:
    pos = np.arange(P, dtype=np.uint64)[None, :]
    active = pos < nk[:, None].astype(np.uint64)
    return ((slots_by_pos.astype(np.uint64) * active) << (4 * pos)).sum(axis=1).astype(np.uint32)


def _unpack_order(order: np.ndarray) -> np.ndarray:
    pos = np.arange(P, dtype=np.uint32)[None, :]
    return ((order[:, None] >> (4 * pos)) & 0xF).astype(np.int64)


def random_clocks(rng, n, min_keys=1, max_keys=4, lo=2, hi=20):
    """-> cnt u32[n,8], order u32[n], nk (number of keys)."""
    nk = rng.integers(min_keys, max_keys + 1, n)
    perm = np.argsort(rng.random((n, P)), axis=1)
    active = np.arange(P)[None, :] < nk[:, None]
    by_pos = rng.integers(lo, hi + 1, (n, P)).astype(np.uint32) * active
    cnt = np.zeros((n, P), np.uint32)
    np.put_along_axis(cnt, perm, by_pos, axis=1)
    return cnt, _pack_order(perm, nk), nk


def record_values(rng, n, n_records):
    """Random full records -> val u64[n,4]."""
    val = np.zeros((n, F), np.uint64)
    val[:, 0] = rng.integers(1, 100, n).astype(np.float64).view(np.uint64)
    val[:, 1] = (rng.integers(0, 100_000_000, n).astype(np.float64) / 100.0).view(np.uint64)
    r = rng.random(n)
    val[:, 2] = np.where(r < 0.1, 0, np.where(r < 0.3, 1, 2 + n_records)).astype(np.uint64)
    val[:, 3] = (2 + rng.integers(0, n_records, n)).astype(np.uint64)
    return val


@dataclass
class TableImage:
    rows: np.ndarray  # ROW_DTYPE[n_records], row i == path id i

    @property
    def n(self):
        return int(self.rows.shape[0])


def make_table(n_records: int, rng) -> TableImage:
    rows = np.zeros(n_records, codec.ROW_DTYPE)
    rows["val"] = record_values(rng, n_records, n_records)
    rows["val"][:, 3] = 2 + np.arange(n_records, dtype=np.uint64)
    cnt, order, _ = random_clocks(rng, n_records)
    rows["m_cnt"] = cnt
    rows["v_cnt"] = cnt
    rows["m_order"] = order
    rows["v_order"] = order
    rows["hdr"] = FULL_RECORD_HDR
    rows["flags"] = codec.ROW_M_PRESENT | codec.ROW_V_PRESENT | codec.ROW_ALIAS
    rows["cseq"] = 1 + np.arange(n_records, dtype=np.uint64)
    return TableImage(rows)


def zipf_ids(rng, n, n_records, s=0.8):
    """Bounded Zipf(s) over ranks, ranks scattered over path ids by a fixed permutation."""
    w = 1.0 / np.power(np.arange(1, n_records + 1, dtype=np.float64), s)
    cdf = np.cumsum(w)
    cdf /= cdf[-1]
    ranks = np.searchsorted(cdf, rng.random(n), side="left")
    a = 0x9E3779B1  # odd multiplier: a bijection mod 2^k; fold into range by rejection-free modulo of a permutation
    perm_seed = np.random.Generator(np.random.PCG64(12345))
    perm = perm_seed.permutation(n_records) if n_records <= 50_000_000 else None
    if perm is None:
        return ((ranks.astype(np.uint64) * np.uint64(a)) % np.uint64(n_records)).astype(np.uint64)
    return perm[ranks].astype(np.uint64)


def make_batch(table: TableImage, n: int, rng, keys="uniform", partial_records=0.3, mix=MIX) -> codec.Batch:
    n_records = table.n
    if keys == "uniform":
        pid = rng.integers(0, n_records, n).astype(np.uint64)
    elif keys == "zipf":
        pid = zipf_ids(rng, n, n_records)
    else:
        raise ValueError(keys)
    rows = table.rows
    ipid = pid.astype(np.int64)
    cnt = rows["m_cnt"][ipid].copy()
    order = rows["m_order"][ipid].copy()
    slots = _unpack_order(order)
    nk = (cnt != 0).sum(axis=1)

    names = list(mix)
    mode = rng.choice(len(names), n, p=[mix[k] for k in names])
    is_ = {k: mode == i for i, k in enumerate(names)}
    ar = np.arange(n)

    # historical / concurrent: decrement one existing key (table counts are >= 2)
    dec = is_["historical"] | is_["concurrent"]
    j = (rng.random(n) * nk).astype(np.int64)
    dslot = slots[ar, j]
    cnt[ar[dec], dslot[dec]] -= 1
    # dominating / concurrent: increment some slot (other than the decremented one); append if new
    inc = is_["dominating"] | is_["concurrent"]
    islot = rng.integers(0, P, n)
    clash = inc & dec & (islot == dslot)
    islot[clash] = (islot[clash] + 1) % P
    new_key = inc & (cnt[ar, islot] == 0)
    cnt[ar[inc], islot[inc]] += 1
    slots[ar[new_key], nk[new_key]] = islot[new_key]
    nk = nk + new_key
    # permuted: rotate the key order by one position
    rot = np.take_along_axis(slots, (np.arange(P)[None, :] + 1) % np.maximum(nk, 1)[:, None], axis=1)
    slots = np.where(is_["permuted"][:, None], rot, slots)
    order = _pack_order(slots, nk)

    b = codec.Batch.empty(n)
    b.path_id[:] = pid
    b.val[:] = record_values(rng, n, n_records)
    hdr = np.full(n, FULL_RECORD_HDR, np.uint64)
    # partial records: keep a random non-empty prefix-free subset of fields, canonical key order
    part = rng.random(n) < partial_records
    if part.any():
        keep = rng.random((n, F)) < 0.5
        keep[:, 0] |= ~keep.any(axis=1)
        keep |= ~part[:, None]
        tags = np.array([codec.TAG_NUM, codec.TAG_NUM, codec.TAG_STR, codec.TAG_STR], np.uint64)
        tagbits = ((keep * tags[None, :]) << (codec.HDR_TAG_SHIFT + 3 * np.arange(F, dtype=np.uint64))[None, :]).sum(axis=1)
        # order nibbles: kept slots in ascending slot order
        rank = np.maximum(np.cumsum(keep, axis=1) - 1, 0)
        ordbits = ((keep * np.arange(F, dtype=np.uint64)[None, :]) << (4 * rank.astype(np.uint64))).sum(axis=1)
        hdr = (np.uint64(codec.KIND_OBJ << codec.HDR_KIND_SHIFT) | tagbits.astype(np.uint64)
               | (ordbits.astype(np.uint64) << np.uint64(codec.HDR_ORDER_SHIFT)))
        b.val[:] = b.val * keep
    net = ~is_["local"]
    b.head["hdr"] = hdr | net.astype(np.uint64)
    b.head["clk_order"] = order * net
    b.head["user"] = np.arange(n, dtype=np.uint32)
    b.clk[:] = cnt * net[:, None]
    return b


def synth_ranks(n_records: int) -> dict:
    """rank_* of bb_config for the synthetic dictionary: in UTF-16 order
    "NaN" < "[object Object]" < "admin" < "editor" < "false" < "name..." < "true" < "user"."""
    return dict(rank_object=0, rank_nan=0, rank_false=2, rank_true=2 + n_records)


def synth_schema(n_records: int) -> codec.Schema:
    """The codec.Schema of the synthetic collection (materialises the dictionary: use for
    tests and small tables; the bench builds keys and bounds from the id layout directly)."""
    strings = ["admin", "editor", "user"] + ["name%07d" % i for i in range(n_records)]
    return codec.Schema(["age", "score", "role", "name"], ["p%d" % i for i in range(P)],
                        codec.StringDict(strings), "p0")
