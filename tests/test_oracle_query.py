"""BulletQuery: typed C oracle (oracle/bullet_oracle.c) == literal oracle (oracle/js_literal.py),
results compared in the reference's exact (Map order, Set order); plus the host codec that turns
JS query arguments into keys and bounds (bullet_js_b200/codec.py)."""
import itertools
import math

import numpy as np
import pytest

from bullet_js_b200 import capi, codec
from oracle.jsvalue import UNDEFINED
from oracle.typed import TypedOracle
from tests import streamgen
from tests.test_oracle_typed import make_cfg

EQ_VALUES = ([float(x) for x in streamgen.NUMS] + streamgen.STRINGS
             + [True, False, None, "25", "25.0", "2.5e1", "0", "-0", "true", "false", "NaN", "null",
                "Infinity", "1e+21", "1e21", "1e-7", "nosuch", "3.5", " 25"])
BOUNDS = [0.0, 1.0, 25.0, 30.0, 99.0, -math.inf, math.inf, math.nan, "25", "40", "", "a", "admin", "b", "user",
          "zzz", "Zed", "[", "NaN", "true", "false", "g", "￿", True, False, None, -5.0, 3.5]


def paths_of(schema, ids):
    return [schema.paths.name(i) for i in ids]


def replay(seed, n_ops=1500, n_paths=12):
    ops, ref = streamgen.generate(seed, n_ops, n_paths, index_fields=("age", "role"), late_index={"score": 700})
    schema = streamgen.make_schema()
    batch = codec.encode_updates(schema, ops)
    orc = TypedOracle(make_cfg(schema, 64, True))
    orc.index_create(0)
    orc.index_create(2)
    orc.merge(batch.slice(0, 700))
    orc.index_create(1)
    orc.merge(batch.slice(700, n_ops))
    return schema, ref, orc


@pytest.mark.parametrize("seed", range(4))
def test_equals_and_count_exact_order(seed):
    schema, ref, orc = replay(seed)
    nonempty = 0
    for f, name in ((0, "age"), (1, "score"), (2, "role")):
        for v in EQ_VALUES:
            want = ref.equals("users", name, v)
            key = schema.index_key(v)
            got = [] if key is None else paths_of(schema, orc.query_equals(f, key))
            assert got == want, (name, v, got, want)
            assert (0 if key is None else orc.query_count(f, key)) == ref.count("users", name, v)
            nonempty += bool(want)
    assert nonempty > 10


@pytest.mark.parametrize("seed", range(4))
def test_range_exact_order(seed):
    schema, ref, orc = replay(seed)
    nonempty = 0
    for f, name in ((0, "age"), (1, "score"), (2, "role")):
        for lo, hi in itertools.product(BOUNDS, BOUNDS):
            want = ref.range("users", name, lo, hi)
            got = paths_of(schema, orc.query_range(f, schema.bound(lo, False), schema.bound(hi, True)))
            assert got == want, (name, lo, hi, got, want)
            nonempty += bool(want)
        assert ref.range("users", name, UNDEFINED, 5.0) == [] and ref.range("users", name, 0.0, UNDEFINED) == []
    assert nonempty > 50


def test_index_holds_stale_entries_like_the_reference():
    """SURVEY 8c KAT-H / KAT-N on the typed path: the hook never removes the pre-update value."""
    schema = codec.Schema(["age", "role"], ["A", "B"], codec.StringDict(["admin", "user"]), "B")
    orc = TypedOracle(make_cfg(schema, 8, True))
    orc.index_create(0)
    ops = [("users/u1", {"age": 30.0}, None), ("users/u1", {"age": 31.0}, None), ("users/u1", {"age": 0.0}, None)]
    orc.merge(codec.encode_updates(schema, ops))
    lo, hi = schema.bound(30.0, False), schema.bound(31.0, True)
    assert orc.query_range(0, lo, hi).tolist() == [0, 0]  # u1 twice: buckets "30" and "31"
    assert orc.index_entries(0) == 2


def test_codec_keys():
    schema = streamgen.make_schema()
    k = schema.index_key
    assert k(25.0) == k("25") != k("25.0") and k("25.0") is None
    assert k(-0.0) == k(0.0) == k("0") == 0 and k("-0") is None
    assert k(math.nan) == k("NaN") == codec.KEY_NAN
    assert k(True) == k("true") == codec.KEY_BOOL | 1 and k(False) == k("false")
    assert k("admin") == codec.KEY_STR | schema.strings.id("admin")
    assert k(None) is None and k("nosuch") is None
    assert k(1e21) == k("1e+21") and k("1e21") is None
    b = schema.bound("b", False)
    assert int(b["rank"]) == 4 and int(b["flags"]) == codec.BOUND_IS_STRING | codec.BOUND_TRUE | codec.BOUND_FALSE
    assert math.isnan(float(b["num"]))
