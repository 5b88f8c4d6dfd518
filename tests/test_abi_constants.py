"""include/bullet_b200.h is the contract: the Python host codec must agree with every constant it mirrors."""
import os
import re

from bullet_js_b200 import codec

HEADER = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "bullet_b200.h")).read()


def defines(prefix):
    out = {}
    for name, val in re.findall(r"#define\s+(" + prefix + r"\w+)\s+\(?([0-9xXa-fA-F]+)u?(?:ll|ull)?\)?", HEADER):
        out[name] = int(val, 0)
    return out


def test_config_flags_match_the_header():
    h = defines("BB_CFG_")
    mirrored = {k: v for k, v in vars(codec).items() if k.startswith("CFG_")}
    assert mirrored, "codec mirrors no flags?"
    for k, v in mirrored.items():
        assert h["BB_" + k] == v, k
    assert set("BB_" + k for k in mirrored) == set(h), (sorted(h), sorted(mirrored))
    assert len(set(h.values())) == len(h) and all(v & (v - 1) == 0 for v in h.values())  # distinct single bits


def test_row_and_batch_sizes_match_the_header():
    assert codec.ROW_DTYPE.itemsize == 128 and codec.HEAD_DTYPE.itemsize == 16
    assert codec.MAX_PEERS == defines("BB_MAX_")["BB_MAX_PEERS"] and codec.MAX_FIELDS == defines("BB_MAX_")["BB_MAX_FIELDS"]
    row = defines("BB_ROW_")
    assert (row["BB_ROW_M_PRESENT"], row["BB_ROW_V_PRESENT"], row["BB_ROW_ALIAS"]) == (
        codec.ROW_M_PRESENT, codec.ROW_V_PRESENT, codec.ROW_ALIAS)


def test_slots_and_collect_flags_match_the_header():
    slots = defines("BB_SLOT_") | defines("BB_NO_")
    assert slots["BB_NO_SLOT"] == codec.NO_SLOT and slots["BB_SLOT_ECHO"] == codec.SLOT_ECHO == codec.NO_SLOT - 1
    assert defines("BB_COLLECT_")["BB_COLLECT_FILTER_RECORDS"] == codec.COLLECT_FILTER_RECORDS
    # the JS shim mirrors the flags it sets
    js = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "js", "bullet-b200.js")).read()
    for name in ("BB_CFG_POST_GETDATA", "BB_CFG_EXACT_ORDER"):
        m = re.search(r"const\s+" + name + r"\s*=\s*(\d+)", js)
        assert m and int(m.group(1)) == defines("BB_CFG_")[name], name
