"""Loading and comparing the committed golden fixtures (tests/golden/*.json.gz).

The fixtures are OUTPUTS OF THE REFERENCE ITSELF (src/*.js executed by oracle/minijs,
tests/golden/make_golden.py); nothing here touches /root/reference, so these helpers also run on the GPU box.
"""
from __future__ import annotations

import gzip
import hashlib
import json
import math
import os

from oracle.ref_runner import jsonable, unjsonable

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    with gzip.open(os.path.join(GOLDEN, name), "rb") as f:
        return json.loads(f.read().decode())


def ops_of(case):
    return [(p, unjsonable(v), None if c is None else {k: float(x) for k, x in c}) for p, v, c in case["ops"]]


def same_js(a, b):
    """Deep equality that keeps own-key order, NaN == NaN and +0 != -0."""
    if isinstance(a, dict) and isinstance(b, dict):
        return list(a.keys()) == list(b.keys()) and all(same_js(a[k], b[k]) for k in a)
    if isinstance(a, list) and isinstance(b, list):
        return len(a) == len(b) and all(same_js(x, y) for x, y in zip(a, b))
    if isinstance(a, float) and isinstance(b, float):
        if math.isnan(a) or math.isnan(b):
            return math.isnan(a) and math.isnan(b)
        return a == b and math.copysign(1, a) == math.copysign(1, b)
    return type(a) is type(b) and a == b


def clock_items(c):
    return None if c is None else [[k, float(v)] for k, v in c.items()]


def canonical_value(v):
    return json.dumps(jsonable(v), separators=(",", ":"), ensure_ascii=True)


def changes_sha256(changes):
    """changes: iterable of dict(seq, path, value, vectorClock) in change-set order (tests/golden/make_golden.py)."""
    h = hashlib.sha256()
    for c in changes:
        h.update(f"{int(c['seq'])}|{c['path']}|{canonical_value(c['value'])}|{canonical_value(c['vectorClock'])}\n".encode())
    return h.hexdigest()


def table_sha256(rows):
    """rows: iterable of (path, value, M, V, alias)."""
    h = hashlib.sha256()
    for p, value, m, v, alias in rows:
        h.update(f"{p}|{canonical_value(value)}|{canonical_value(m)}|{canonical_value(v)}|{int(alias)}\n".encode())
    return h.hexdigest()
