"""Known-answer traces of SURVEY.md 8c against the literal oracle.

The reference ships no tests; these traces were hand-derived from its source
(file:line in SURVEY.md 8c) independently of `oracle/js_literal.py`.
"""
from oracle.js_literal import (
    CODE_CONCURRENT, CODE_HISTORICAL, CODE_IDENTICAL, CODE_INCOMING, CODE_NO_CURRENT,
    CODE_TIE_CURRENT, CODE_TIE_INCOMING, RefBullet,
)
from oracle.jsvalue import UNDEFINED, number_to_string, string_to_number, less_than, norm

USERS = {
    "user1": {"name": "Alice Johnson", "age": 28, "active": True, "role": "admin"},
    "user2": {"name": "Bob Smith", "age": 35, "active": True, "role": "user"},
    "user3": {"name": "Carol Davis", "age": 42, "active": False, "role": "user"},
    "user4": {"name": "Dave Wilson", "age": 23, "active": True, "role": "editor"},
    "user5": {"name": "Eve Brown", "age": 31, "active": True, "role": "user"},
    "user6": {"name": "Frank Miller", "age": 47, "active": False, "role": "admin"},
    "user7": {"name": "Grace Lee", "age": 29, "active": True, "role": "editor"},
    "user8": {"name": "Harry Taylor", "age": 39, "active": True, "role": "user"},
    "user9": {"name": "Irene Clark", "age": 26, "active": False, "role": "user"},
    "user10": {"name": "Jack Roberts", "age": 33, "active": True, "role": "admin"},
}
PRODUCTS = {
    "prod1": {"name": "Laptop", "price": 1200, "stock": 15, "category": "electronics"},
    "prod2": {"name": "Smartphone", "price": 800, "stock": 25, "category": "electronics"},
    "prod3": {"name": "Headphones", "price": 150, "stock": 50, "category": "accessories"},
    "prod4": {"name": "Mouse", "price": 30, "stock": 100, "category": "accessories"},
    "prod5": {"name": "Keyboard", "price": 80, "stock": 40, "category": "accessories"},
    "prod6": {"name": "Monitor", "price": 300, "stock": 20, "category": "electronics"},
    "prod7": {"name": "Desk Chair", "price": 250, "stock": 10, "category": "furniture"},
    "prod8": {"name": "Desk", "price": 400, "stock": 5, "category": "furniture"},
    "prod9": {"name": "Printer", "price": 200, "stock": 8, "category": "electronics"},
    "prod10": {"name": "Camera", "price": 600, "stock": 12, "category": "electronics"},
}


def test_kat_q1_query_example():
    """examples/bullet-query-example.js:17-139."""
    b = RefBullet("me")
    for k, v in USERS.items():
        b.put(f"users/{k}", v)
    for k, v in PRODUCTS.items():
        b.put(f"products/{k}", v)
    b.index("users", "role").index("users", "age").index("users", "active")
    b.index("products", "category").index("products", "price")
    assert b.equals("users", "role", "admin") == ["users/user1", "users/user6", "users/user10"]
    assert b.range("users", "age", 30, 40) == [
        "users/user2", "users/user5", "users/user8", "users/user10"]
    assert [b.count("users", "role", r) for r in ("admin", "user", "editor")] == [3, 5, 2]
    assert b.range("products", "price", 100, 300) == [
        "products/prod3", "products/prod6", "products/prod7", "products/prod9"]
    assert list(b.query.indices["users:age"].keys()) == [
        "28", "35", "42", "23", "31", "47", "29", "39", "26", "33"]


def test_kat_q2_index_first():
    """docs/quick-start.md:183-209."""
    b = RefBullet("me")
    b.index("users", "role")
    b.put("users/alice", {"name": "Alice", "email": "alice@example.com", "role": "admin"})
    b.put("users/bob", {"name": "Bob", "email": "bob@example.com", "role": "user"})
    assert b.equals("users", "role", "admin") == ["users/alice"]


def _state(b, path):
    m = b.meta.get(path, {}).get("vectorClock")
    v = b.crt.vectorClocks.get(path)
    return m, (m is v)


def test_kat_l_local_leaf():
    b = RefBullet("A", enable_indexing=False)
    p = "k/v"
    expect = [
        (5, CODE_NO_CURRENT, True, 5.0, {"A": 3.0}, True),
        (3, CODE_TIE_CURRENT, False, 5.0, {"A": 4.0}, False),
        (3, CODE_INCOMING, True, 3.0, {"A": 5.0}, True),
        (3, CODE_IDENTICAL, False, 3.0, {"A": 6.0}, False),
        (7, CODE_INCOMING, True, 7.0, {"A": 7.0}, True),
        (9, CODE_TIE_INCOMING, True, 9.0, {"A": 8.0}, True),
        (0, CODE_TIE_CURRENT, False, 9.0, {"A": 9.0}, False),
        (0, CODE_INCOMING, True, 0.0, {"A": 10.0}, True),
        (4, CODE_TIE_INCOMING, True, 4.0, {"A": 11.0}, True),
    ]
    for x, code, do, s, m, alias in expect:
        b.put(p, x)
        d = b.decisions[-1]
        assert (d["code"], d["doUpdate"]) == (code, do), (x, d)
        assert b.store["k"]["v"] == s or (s == 0.0 and b.store["k"]["v"] in (0.0, {}))
        mm, a = _state(b, p)
        assert mm == m and a == alias, (x, mm, a)


def _recv(b, path, data, clock):
    b.process_sync_entries([dict(path=path, data=data, vectorClock=clock)])
    return b.decisions[-1]


def test_kat_n_network_objects():
    b = RefBullet("B")
    b.index("users", "age").index("users", "role")
    p = "users/u1"
    age = b.query.indices["users:age"]
    role = b.query.indices["users:role"]

    d = _recv(b, p, {"age": 30, "role": "user"}, {"A": 3})
    assert d["code"] == CODE_NO_CURRENT and b.meta[p]["vectorClock"] == {"B": 2.0}
    assert {k: list(v) for k, v in age.items()} == {"30": [p]}
    assert {k: list(v) for k, v in role.items()} == {"user": [p]}

    d = _recv(b, p, {"age": 25, "role": "admin"}, {"A": 4})
    assert d["code"] == CODE_CONCURRENT and d["doUpdate"]
    assert b.store["users"]["u1"] == {"age": 30.0, "role": "user"}
    assert list(b.meta[p]["vectorClock"].items()) == [("A", 4.0), ("B", 2.0)]
    assert {k: list(v) for k, v in age.items()} == {"25": [p]}
    assert {k: list(v) for k, v in role.items()} == {"admin": [p]}

    d = _recv(b, p, {"age": 40}, {"A": 5, "B": 2})
    assert d["code"] == CODE_INCOMING and b.store["users"]["u1"] == {"age": 40.0}
    assert list(b.meta[p]["vectorClock"].items()) == [("A", 5.0), ("B", 2.0)]
    assert list(age.keys()) == ["25", "40"]

    d = _recv(b, p, {"age": 10}, {"A": 4, "B": 2})
    assert d["code"] == CODE_HISTORICAL and not d["doUpdate"]
    assert b.store["users"]["u1"] == {"age": 40.0}
    assert b.crt.vectorClocks[p] == {"A": 5.0, "B": 2.0}
    assert b.crt.vectorClocks[p] is not b.meta[p]["vectorClock"]
    assert list(age.keys()) == ["25", "10"]

    d = _recv(b, p, {"age": 10}, {"B": 2, "A": 5})
    assert d["code"] == CODE_CONCURRENT and b.store["users"]["u1"] == {"age": 40.0}
    assert list(b.meta[p]["vectorClock"].items()) == [("B", 2.0), ("A", 5.0)]
    assert list(age.keys()) == ["25", "10"]

    d = _recv(b, p, {"age": 10}, {"B": 2, "A": 5})
    assert d["code"] == CODE_TIE_INCOMING and b.store["users"]["u1"] == {"age": 10.0}


def test_kat_r_resolve_direct():
    """docs/conflict-resolution.md:446-452 ('light' beats 'dark')."""
    b = RefBullet("me", enable_indexing=False)
    r = b.crt.resolve("k", {"peerA": 1.0}, {"peerB": 1.0}, "light", "dark")
    assert r["concurrent"] and r["value"] == "light"
    r = b.crt.resolve("k", {"peerB": 1.0}, {"peerA": 1.0}, "dark", "light")
    assert r["concurrent"] and r["value"] == "light"
    r = b.crt.resolve("k", {}, {"A": 1.0}, 1.0, 2.0)
    assert r["historical"]
    r = b.crt.resolve("k", {}, {}, 3.0, 3.0)
    assert r["reason"] == "identical clocks and values"
    # docs/conflict-resolution.md:458-483 deep merge
    r = b.crt.resolve("k", {"a": 1.0}, {"b": 1.0},
                      norm({"name": "Bob Smith", "age": 31}), norm({"name": "Robert Smith", "age": 30}))
    assert r["value"] == {"name": "Robert Smith", "age": 31.0}


def test_kat_h_hook_staleness():
    b = RefBullet("me")
    b.index("users", "age")
    p = "users/u1"
    age = b.query.indices["users:age"]
    b.put(p, {"age": 30})
    assert {k: list(v) for k, v in age.items()} == {"30": [p]}
    b.put(p, {"age": 31})
    assert b.decisions[-1]["code"] == CODE_TIE_INCOMING
    assert {k: list(v) for k, v in age.items()} == {"30": [p], "31": [p]}
    assert b.range("users", "age", 30, 31) == [p, p]
    b.put(p, {"age": 0})
    assert {k: list(v) for k, v in age.items()} == {"30": [p], "31": [p]}


def test_js_number_to_string():
    cases = {
        0.0: "0", -0.0: "0", 25.0: "25", 1e21: "1e+21", 1e20: "100000000000000000000",
        123456.78: "123456.78", 0.000001: "0.000001", 1e-7: "1e-7", 1.5e-7: "1.5e-7",
        -3.25: "-3.25", 5e-324: "5e-324", 1.7976931348623157e308: "1.7976931348623157e+308",
        float("nan"): "NaN", float("inf"): "Infinity", 0.1 + 0.2: "0.30000000000000004",
    }
    for x, s in cases.items():
        assert number_to_string(x) == s, (x, number_to_string(x))


def test_js_relational_quirks():
    assert string_to_number("") == 0.0 and string_to_number(" 12 ") == 12.0
    assert string_to_number("0x1A") == 26.0 and string_to_number("1e3") == 1000.0
    assert string_to_number("abc") != string_to_number("abc")
    assert less_than("Zed", {}) and not less_than("abc", {})     # "[object Object]"
    assert less_than({}, "abc") and not less_than({}, 5.0)
    assert less_than(None, 5.0) and less_than(False, True) and not less_than(UNDEFINED, 1.0)
