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
