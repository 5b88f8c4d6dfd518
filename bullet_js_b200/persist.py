"""The reference's persisted state <-> device table rows.

`BulletFileStorage` (src/bullet-file-storage.js:170-210) writes `store.json` = JSON.stringify(bullet.store),
`meta.json` = JSON.stringify(bullet.meta) (`{path: {source, vectorClock, lastModified}}`, src/bullet.js:198-203)
and `log.json`; on start it deep-merges them back (`:96-163`).  `crt.vectorClocks` is never persisted, so a
restarted peer holds every path in the state "M present, V absent" (and has a fresh id, src/bullet.js:33).

    import_collection  store.json / meta.json (already parsed) -> (path ids, bb_row image) for bb_table_load:
                       values typed by the schema, M = meta[path].vectorClock, V absent, creation sequence in
                       the collection's own-key order (what `_buildIndex`'s Object.entries will see)
    export_collection  rows read with bb_table_read -> the two dicts the reference would have stringified
                       (`source` / `lastModified` are not kept on the device: callers that need them keep them
                       on the host)

    collect_full_sync_data / chunk_sync_data   the sync PRODUCER side: what
                       `BulletNetworkSync._collectFullSyncData` / `_chunkSyncData` (src/bullet-network-sync.js:592-664,
                       713-723) would ship for the exported state, entry for entry, so a B200 peer can answer a
                       reference peer's sync request

JSON has already flattened what it cannot hold (NaN / +-Infinity -> null, -0 -> 0) when the reference wrote the
files: the import takes the files as they are.  Paths outside `<collection>/<key>`, nested records and keys
outside the schema raise `codec.DomainError` (no CPU fallback: such collections stay with the stock storage).
"""
from __future__ import annotations

import json
import os

import numpy as np

from . import codec


def import_collection(schema: codec.Schema, collection: str, store: dict, meta: dict):
    """-> (path_id u64[n], rows ROW_DTYPE[n]); interns `<collection>/<key>` in the collection's key order."""
    records = store.get(collection, {})
    if not isinstance(records, dict):
        raise codec.DomainError(f"store[{collection!r}] is not a collection")
    n = len(records)
    ids = np.zeros(n, np.uint64)
    rows = np.zeros(n, codec.ROW_DTYPE)
    for k, (key, value) in enumerate(records.items()):
        path = f"{collection}/{key}"
        ids[k] = schema.paths.id(path)
        hdr, val = schema.enc_value(_numbers_to_float(value))
        row = rows[k]
        row["hdr"] = hdr
        row["val"] = val
        flags = 0
        m = (meta.get(path) or {}).get("vectorClock")
        if m is not None:
            cnt, order = schema.enc_clock(m)
            row["m_cnt"] = cnt
            row["m_order"] = order
            flags |= codec.ROW_M_PRESENT
        row["flags"] = flags  # V absent, no alias: crt.vectorClocks starts empty after a restart
        row["cseq"] = k + 1
    for path in meta:
        if path.startswith(collection + "/") and path.split("/", 1)[1] not in records:
            raise codec.DomainError(f"meta entry {path!r} has no record (finer write granularity than the collection's)")
    return ids, rows


def export_collection(schema: codec.Schema, collection: str, path_ids, rows):
    """-> (store[collection], {path: {"vectorClock": ...}}) in creation order (the reference's own-key order)."""
    decoded = []
    for pid, row in zip(path_ids, rows):
        d = codec.decode_row(schema, row)
        if d["kind"] != codec.KIND_NONE:
            decoded.append((d["cseq"], schema.paths.name(int(pid)), d))
    decoded.sort(key=lambda t: t[0])
    records, meta = {}, {}
    for _, path, d in decoded:
        records[path.split("/", 1)[1]] = d["value"]
        if d["M"] is not None:
            meta[path] = {"vectorClock": d["M"]}
    return records, meta


def _numbers_to_float(v):
    if isinstance(v, dict):
        return {k: _numbers_to_float(x) for k, x in v.items()}
    if isinstance(v, bool) or v is None or isinstance(v, str):
        return v
    if isinstance(v, (int, float)):
        return float(v)
    raise codec.DomainError(f"unsupported persisted value {v!r}")


def read_dir(path: str):
    """store.json / meta.json of an (unencrypted) reference data directory -> (store, meta)."""
    def load(name):
        p = os.path.join(path, name)
        if not os.path.exists(p):
            return {}
        with open(p, encoding="utf-8") as f:
            return json.load(f)
    return load("store.json"), load("meta.json")


def json_value(v):
    """A value as JSON.stringify leaves it (store.json): non-finite numbers become null, -0 becomes 0."""
    if isinstance(v, dict):
        return {k: json_value(x) for k, x in v.items()}
    if isinstance(v, float):
        if v != v or v in (float("inf"), float("-inf")):
            return None
        return 0.0 if v == 0 else v
    return v


def collect_full_sync_data(store: dict, meta: dict, since: float = 0):
    """BulletNetworkSync._collectFullSyncData (src/bullet-network-sync.js:592-664) over (store, meta) dicts as
    `export_collection` returns them.  The reference walks the store down to its LEAVES and looks the clock up in
    `meta[<leaf path>]` - which only exists for paths that were written as such, so a collection written record
    by record ships `users/u1/age`-style entries with `{}` clocks, while a primitive written at `users/u3`
    carries its own.  Reproduced as is; `lastModified` is not kept on the device (0 here), and `meta.deleted`
    is never set by the reference either, so there are no tombstone entries."""
    entries = []

    def leaf(path, data):
        m = meta.get(path) or {}
        last = m.get("lastModified") or 0
        if since > 0 and last and last < since:
            return
        entries.append({"path": path, "data": data, "vectorClock": m.get("vectorClock") or {},
                        "lastModified": last, "deleted": False})

    def traverse(obj, path):
        if not isinstance(obj, dict):
            leaf(path[1:], obj)
            return
        for key, value in obj.items():
            if isinstance(value, dict):
                traverse(value, path + "/" + key)
            else:
                leaf((path + "/" + key)[1:], value)

    traverse(store, "")
    return entries


def chunk_sync_data(entries, chunk_size: int = 50):
    """BulletNetworkSync._chunkSyncData (src/bullet-network-sync.js:713-723)."""
    return [entries[i:i + chunk_size] for i in range(0, len(entries), chunk_size)]
