"""Host codec (bullet_js_b200/codec.py): what goes into the struct-of-arrays buffers comes back out, for every
value / clock shape of the typed domain (property-based), and what lies outside it is refused."""
import math

import numpy as np
import pytest
from hypothesis import given, settings
from hypothesis import strategies as st

from bullet_js_b200 import codec
from tests import streamgen
from tests.golden_io import same_js

SCHEMA = streamgen.make_schema()
prims = st.one_of(st.floats(allow_nan=True, allow_infinity=True), st.sampled_from(streamgen.STRINGS), st.booleans(), st.none())
records = st.lists(st.sampled_from(streamgen.FIELDS), unique=True, max_size=4).flatmap(
    lambda ks: st.tuples(*[prims for _ in ks]).map(lambda vs: dict(zip(ks, vs))))
clocks = st.lists(st.sampled_from(streamgen.PEERS), unique=True, max_size=8).flatmap(
    lambda ks: st.tuples(*[st.integers(1, 2 ** 32 - 1) for _ in ks]).map(lambda vs: {k: float(v) for k, v in zip(ks, vs)}))


@settings(max_examples=300, deadline=None)
@given(st.one_of(prims, records))
def test_value_roundtrip_keeps_key_order_nan_and_signed_zero(v):
    hdr, val = SCHEMA.enc_value(v)
    back = SCHEMA.dec_value(hdr, np.array(val, np.uint64))
    assert same_js(back, {k: (float(x) if isinstance(x, (int, float)) and not isinstance(x, bool) else x) for k, x in v.items()}
                   if isinstance(v, dict) else (float(v) if isinstance(v, (int, float)) and not isinstance(v, bool) else v))


@settings(max_examples=300, deadline=None)
@given(clocks)
def test_clock_roundtrip_keeps_key_order(c):
    cnt, order = SCHEMA.enc_clock(c)
    assert list(SCHEMA.dec_clock(np.array(cnt, np.uint32), order).items()) == list(c.items())


@settings(max_examples=200, deadline=None)
@given(st.lists(st.tuples(st.integers(0, 30), st.one_of(prims, records), st.one_of(st.none(), clocks)), max_size=40))
def test_batch_roundtrip(updates):
    schema = streamgen.make_schema()
    ops = [(f"users/u{i}", v, c) for i, v, c in updates]
    b = codec.encode_updates(schema, ops)
    assert b.n == len(ops)
    for k, (path, v, c) in enumerate(ops):
        assert schema.paths.name(int(b.path_id[k])) == path
        hdr = int(b.head["hdr"][k])
        net = bool(hdr & codec.HDR_FLAVOUR_NET)
        assert net == (c is not None and isinstance(v, dict))  # primitives never carry a clock (sync:560-563)
        got = schema.dec_value(hdr & ~codec.HDR_FLAVOUR_NET, b.val[k])
        want = ({kk: (float(x) if isinstance(x, (int, float)) and not isinstance(x, bool) else x) for kk, x in v.items()}
                if isinstance(v, dict) else (float(v) if isinstance(v, (int, float)) and not isinstance(v, bool) else v))
        assert same_js(got, want)
        if net:
            assert list(schema.dec_clock(b.clk[k], int(b.head["clk_order"][k])).items()) == list(c.items())
        else:
            assert not b.clk[k].any() and int(b.head["clk_order"][k]) == 0  # canonical: absent slots are zero


@pytest.mark.parametrize("bad", [{"age": {"nested": 1.0}}, {"nosuchfield": 1.0}, {"age": [1.0]}, "not in the dictionary"])
def test_values_outside_the_typed_domain_are_refused(bad):
    with pytest.raises(codec.DomainError):
        SCHEMA.enc_value(bad)


@pytest.mark.parametrize("bad", [{"p0": 0.0}, {"p0": -1.0}, {"stranger": 1.0}, {"p0": 2.0 ** 32}])
def test_clocks_outside_the_typed_domain_are_refused(bad):
    with pytest.raises(codec.DomainError):
        SCHEMA.enc_clock(bad)


def test_string_dictionary_is_utf16_ordered():
    d = codec.StringDict(streamgen.STRINGS)
    by_id = [d.string(i) for i in range(len(d))]
    assert by_id == sorted(streamgen.STRINGS, key=lambda s: s.encode("utf-16-be", "surrogatepass"))
    assert math.isnan(codec.js_string_to_number("abc")) and codec.js_string_to_number(" 12 ") == 12.0


def test_from_verdicts_rebuilds_echoed_entries():
    """BB_CFG_COMPACT_CHANGES (include/bullet_b200.h): slot BB_SLOT_ECHO = "the entry is the update itself" - idx = i,
    head = the update's header without the flavour bit, clk / val = the update's."""
    import numpy as np

    from bullet_js_b200 import codec

    b = codec.Batch.empty(4)
    b.head["hdr"] = np.array([0x10 | codec.HDR_FLAVOUR_NET, 0x21, 0x30 | codec.HDR_FLAVOUR_NET, 0x40], np.uint64)
    b.head["clk_order"] = [7, 8, 9, 10]
    b.head["user"] = [100, 101, 102, 103]
    b.clk[:] = np.arange(32, dtype=np.uint32).reshape(4, 8)
    b.val[:] = np.arange(16, dtype=np.uint64).reshape(4, 4)
    verdict = np.array([4 << 29 | codec.SLOT_ECHO, 5 << 29 | codec.NO_SLOT, 6 << 29 | 0, 2 << 29 | codec.SLOT_ECHO], np.uint32)
    e_head = np.zeros(1, codec.HEAD_DTYPE)
    e_head["hdr"], e_head["clk_order"], e_head["user"] = 0x99, 5, 102
    e_clk, e_val = np.full((1, 8), 3, np.uint32), np.full((1, 4), 4, np.uint64)
    ch = codec.Changes.from_verdicts(verdict, np.array([2], np.uint32), e_head, e_clk, e_val, b)
    assert ch.decision.tolist() == [4, 5, 6, 2] and ch.idx.tolist() == [0, 2, 3]
    assert ch.head["hdr"].tolist() == [0x10, 0x99, 0x40] and ch.head["user"].tolist() == [100, 102, 103]
    assert ch.head["clk_order"].tolist() == [7, 5, 10]
    assert np.array_equal(ch.clk[0], b.clk[0]) and np.array_equal(ch.clk[1], e_clk[0]) and np.array_equal(ch.val[2], b.val[3])
    import pytest

    with pytest.raises(ValueError):
        codec.Changes.from_verdicts(verdict, np.array([2], np.uint32), e_head, e_clk, e_val)
