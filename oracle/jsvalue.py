"""ECMAScript value semantics needed by the literal oracle (TEST INFRASTRUCTURE ONLY).

The reference (KORandi/bullet-js) is plain JavaScript and there is no JS engine in
this image, so the literal oracle (`oracle/js_literal.py`) restates the reference's
code over Python objects.  This module restates the handful of ECMAScript abstract
operations that code leans on (ECMA-262 section numbers in each docstring):

    ===            IsStrictlyEqual            7.2.16
    <  >=  <=      IsLessThan                 7.2.14
    ToNumber / StringToNumber                 7.1.4
    ToString / Number::toString               7.1.17 / 6.1.6.1.20
    ToBoolean (truthiness)                    7.1.2
    JSON.stringify (objects / primitives)     25.5.2

JS value model used everywhere in `oracle/`:
    number -> Python float (ints are accepted and converted)
    string -> str, boolean -> bool, null -> None, undefined -> UNDEFINED
    object -> dict (insertion ordered, like JS own string keys that are not
              integer-like; identity is Python identity)
Arrays and nested objects are outside the accelerated domain (SURVEY.md 8a
"Domain restrictions" 6) but objects nest fine here.

Only `tests/`, `__graft_entry__.smoke()` and bench.py's cpu_baseline leg may import
anything under `oracle/`.
"""
from __future__ import annotations

import math
import re
from decimal import Decimal


class _Undefined:
    __slots__ = ()

    def __repr__(self):
        return "undefined"

    def __bool__(self):
        return False


UNDEFINED = _Undefined()


def norm(v):
    """Normalise a Python literal into the JS value model (ints -> float)."""
    if isinstance(v, bool) or v is None or v is UNDEFINED or isinstance(v, (str, float)):
        return v
    if isinstance(v, int):
        return float(v)
    if isinstance(v, dict):
        return {k: norm(x) for k, x in v.items()}
    raise TypeError(f"value outside the modelled JS domain: {v!r}")


def is_object(v) -> bool:
    return isinstance(v, dict)


def typeof(v) -> str:
    if v is UNDEFINED:
        return "undefined"
    if v is None or isinstance(v, dict):
        return "object"
    if isinstance(v, bool):
        return "boolean"
    if isinstance(v, float):
        return "number"
    if isinstance(v, str):
        return "string"
    raise TypeError(v)


def truthy(v) -> bool:
    """ToBoolean (7.1.2): false for undefined, null, false, +-0, NaN, ""."""
    if v is UNDEFINED or v is None:
        return False
    if isinstance(v, bool):
        return v
    if isinstance(v, float):
        return not (v == 0.0 or v != v)
    if isinstance(v, str):
        return len(v) > 0
    return True  # every object, including {}


_DEC_RE = re.compile(r"^[+-]?(\d+\.?\d*([eE][+-]?\d+)?|\.\d+([eE][+-]?\d+)?)$")
_WS = "\t\n\v\f\r \u00a0\u1680\u2000\u2001\u2002\u2003\u2004\u2005\u2006\u2007\u2008\u2009\u200a\u2028\u2029\u202f\u205f\u3000\ufeff"


def string_to_number(s: str) -> float:
    """StringToNumber (7.1.4.1.1)."""
    t = s.strip(_WS)
    if t == "":
        return 0.0
    if t in ("Infinity", "+Infinity"):
        return math.inf
    if t == "-Infinity":
        return -math.inf
    low = t[:2].lower()
    try:
        if low == "0x":
            return float(int(t[2:], 16)) if re.fullmatch(r"[0-9a-fA-F]+", t[2:]) else math.nan
        if low == "0o":
            return float(int(t[2:], 8)) if re.fullmatch(r"[0-7]+", t[2:]) else math.nan
        if low == "0b":
            return float(int(t[2:], 2)) if re.fullmatch(r"[01]+", t[2:]) else math.nan
    except ValueError:
        return math.nan
    if _DEC_RE.match(t):
        return float(t)
    return math.nan


def to_primitive(v):
    """ToPrimitive, hint number, for plain objects: valueOf() returns the object,
    so Object.prototype.toString() is used -> "[object Object]"."""
    if isinstance(v, dict):
        return "[object Object]"
    return v


def to_number(v) -> float:
    v = to_primitive(v)
    if v is UNDEFINED:
        return math.nan
    if v is None:
        return 0.0
    if isinstance(v, bool):
        return 1.0 if v else 0.0
    if isinstance(v, float):
        return v
    if isinstance(v, str):
        return string_to_number(v)
    raise TypeError(v)


def strict_equals(a, b) -> bool:
    """=== (7.2.16). Objects compare by identity; NaN !== NaN; +0 === -0."""
    ta, tb = typeof(a), typeof(b)
    if ta != tb:
        return False
    if isinstance(a, dict) or isinstance(b, dict):
        return a is b
    if a is None or a is UNDEFINED:
        return a is b
    if isinstance(a, float):
        return a == b
    return a == b


def utf16_key(s: str) -> bytes:
    """Sort key whose bytewise order is UTF-16 code-unit order."""
    return s.encode("utf-16-be", "surrogatepass")


def _is_less_than(x, y):
    """IsLessThan (7.2.14) -> True / False / None(undefined)."""
    px, py = to_primitive(x), to_primitive(y)
    if isinstance(px, str) and isinstance(py, str):
        return utf16_key(px) < utf16_key(py)
    nx, ny = to_number(px), to_number(py)
    if nx != nx or ny != ny:
        return None
    return nx < ny


def less_than(a, b) -> bool:  # a < b
    return _is_less_than(a, b) is True


def greater_equal(a, b) -> bool:  # a >= b
    r = _is_less_than(a, b)
    return r is False


def less_equal(a, b) -> bool:  # a <= b
    r = _is_less_than(b, a)
    return r is False


def number_to_string(x: float) -> str:
    """Number::toString(x, 10) (6.1.6.1.20): shortest round-trip digits."""
    if x != x:
        return "NaN"
    if x == 0.0:
        return "0"
    if x < 0.0:
        return "-" + number_to_string(-x)
    if x == math.inf:
        return "Infinity"
    _sign, digs, exp = Decimal(repr(x)).as_tuple()
    digs = list(digs)
    while len(digs) > 1 and digs[-1] == 0:
        digs.pop()
        exp += 1
    s = "".join(map(str, digs))
    k = len(s)
    n = k + exp  # x = 0.s * 10**n
    if k <= n <= 21:
        return s + "0" * (n - k)
    if 0 < n <= 21:
        return s[:n] + "." + s[n:]
    if -6 < n <= 0:
        return "0." + "0" * (-n) + s
    e = n - 1
    es = ("+" if e > 0 else "-") + str(abs(e))
    if k == 1:
        return s + "e" + es
    return s[0] + "." + s[1:] + "e" + es


def to_string(v) -> str:
    """String(v)."""
    if v is UNDEFINED:
        return "undefined"
    if v is None:
        return "null"
    if isinstance(v, bool):
        return "true" if v else "false"
    if isinstance(v, float):
        return number_to_string(v)
    if isinstance(v, str):
        return v
    return "[object Object]"


def _quote(s: str) -> str:
    out = ['"']
    for ch in s:
        o = ord(ch)
        if ch == '"':
            out.append('\\"')
        elif ch == "\\":
            out.append("\\\\")
        elif ch == "\b":
            out.append("\\b")
        elif ch == "\f":
            out.append("\\f")
        elif ch == "\n":
            out.append("\\n")
        elif ch == "\r":
            out.append("\\r")
        elif ch == "\t":
            out.append("\\t")
        elif o < 0x20 or 0xD800 <= o <= 0xDFFF:
            out.append("\\u%04x" % o)
        else:
            out.append(ch)
    out.append('"')
    return "".join(out)


def json_stringify(v):
    """JSON.stringify(v) for the modelled domain. Returns UNDEFINED for undefined."""
    if v is UNDEFINED:
        return UNDEFINED
    if v is None:
        return "null"
    if isinstance(v, bool):
        return "true" if v else "false"
    if isinstance(v, float):
        return number_to_string(v) if math.isfinite(v) else "null"
    if isinstance(v, str):
        return _quote(v)
    parts = []
    for k, x in v.items():
        sx = json_stringify(x)
        if sx is UNDEFINED:
            continue
        parts.append(_quote(k) + ":" + sx)
    return "{" + ",".join(parts) + "}"


def get_prop(obj, key):
    """obj[key] for a value that is not null/undefined (property read on a
    primitive yields undefined for the names used on this path)."""
    if isinstance(obj, dict):
        return obj.get(key, UNDEFINED)
    if obj is None or obj is UNDEFINED:
        raise JSTypeError(f"Cannot read properties of {to_string(obj)} (reading '{key}')")
    return UNDEFINED


class JSTypeError(Exception):
    """The strict-mode TypeError the reference throws when it writes through a
    truthy primitive (src/bullet.js:122-124, 188-191)."""
