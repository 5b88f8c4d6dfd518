"""A small ECMAScript interpreter: runs the reference's own sources (src/bullet.js, bullet-crt.js,
bullet-query.js, bullet-middleware.js, ...) UNMODIFIED, because this image has no JS engine.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): it exists to pin `oracle/js_literal.py` and
`oracle/bullet_oracle.c` to outputs of the reference itself (tests/golden/make_golden.py writes
tests/golden/*.json from it).  Nothing under bullet_js_b200/ may import it.

Design: the AST of parser.py is compiled once into Python closures `f(env) -> value`; statements
return a completion signal (None | BREAK | CONTINUE | ("return", v)).  All code is run as strict
mode (the reference's hot path lives in class bodies, which are always strict).

Value model (deliberately separate from oracle/jsvalue.py's dict model so the two restatements of
ECMA-262 check each other; only Number::toString / StringToNumber are shared):
    undefined -> UNDEFINED, null -> None, boolean -> bool, number -> float, string -> str,
    object -> JSObject (ordered own properties: integer-like keys ascending, then insertion order).
"""
from __future__ import annotations

import math
import os
import re

from ..jsvalue import UNDEFINED, number_to_string, string_to_number
from .parser import parse

BREAK = ("break",)
CONTINUE = ("continue",)


class JSThrow(Exception):
    def __init__(self, value):
        super().__init__(value)
        self.value = value

    def __str__(self):
        v = self.value
        if isinstance(v, JSObject):
            return f"{to_str(v.get('name'))}: {to_str(v.get('message'))}"
        return f"uncaught {v!r}"


class Accessor:
    __slots__ = ("get", "set")

    def __init__(self, get=None, set=None):
        self.get, self.set = get, set


_INT_KEY = re.compile(r"^(0|[1-9]\d*)$")


def _is_index(k) -> bool:
    return isinstance(k, str) and _INT_KEY.match(k) is not None and int(k) < 4294967295


class JSObject:
    __slots__ = ("props", "proto", "cls", "nonenum", "frozen", "has_index")

    def __init__(self, proto=None, cls="Object"):
        self.props = {}
        self.proto = proto
        self.cls = cls
        self.nonenum = None
        self.frozen = False
        self.has_index = False

    # -- own property primitives
    def own_keys(self):
        if not self.has_index:
            return list(self.props)
        idx = sorted((k for k in self.props if _is_index(k)), key=int)
        return idx + [k for k in self.props if not _is_index(k)]

    def enumerable_keys(self):
        ks = self.own_keys()
        if self.nonenum:
            ks = [k for k in ks if k not in self.nonenum]
        return ks

    def has_own(self, k):
        return k in self.props

    def lookup(self, k):
        o = self
        while o is not None:
            if k in o.props:
                return o.props[k]
            if o.__class__ is not JSObject and o.has_own(k):
                return o.get_own(k)
            o = o.proto
        return _MISSING

    def get_own(self, k):
        return self.props.get(k, UNDEFINED)

    def get(self, k, this=None):
        v = self.lookup(k)
        if v is _MISSING:
            return UNDEFINED
        if v.__class__ is Accessor:
            return call(v.get, self if this is None else this, []) if v.get is not None else UNDEFINED
        return v

    def set(self, k, v, this=None):
        cur = self.lookup(k)
        if cur.__class__ is Accessor:
            if cur.set is None:
                raise throw_type(f"Cannot set property {k} which has only a getter")
            call(cur.set, self if this is None else this, [v])
            return
        if self.frozen:
            raise throw_type(f"Cannot assign to read only property '{k}' of object")
        self.put_own(k, v)

    def put_own(self, k, v):
        if not self.has_index and _is_index(k):
            self.has_index = True
        self.props[k] = v

    def define(self, k, v, enumerable=False):
        self.put_own(k, v)
        if not enumerable:
            if self.nonenum is None:
                self.nonenum = set()
            self.nonenum.add(k)

    def delete(self, k):
        if self.frozen:
            raise throw_type(f"Cannot delete property '{k}'")
        self.props.pop(k, None)
        return True

    def __repr__(self):
        return f"<{self.cls} {list(self.props)[:6]}>"


_MISSING = object()


class JSArray(JSObject):
    __slots__ = ("items",)

    def __init__(self, items=None):
        JSObject.__init__(self, ARRAY_PROTO, "Array")
        self.items = items if items is not None else []

    def own_keys(self):
        return [number_to_string(float(i)) for i in range(len(self.items))] + list(self.props)

    def enumerable_keys(self):
        return self.own_keys()

    def has_own(self, k):
        if k == "length":
            return True
        if _is_index(k):
            return int(k) < len(self.items)
        return k in self.props

    def get_own(self, k):
        if k == "length":
            return float(len(self.items))
        if _is_index(k):
            i = int(k)
            return self.items[i] if i < len(self.items) else UNDEFINED
        return self.props.get(k, UNDEFINED)

    def put_own(self, k, v):
        if k == "length":
            n = int(to_num(v))
            if n < len(self.items):
                del self.items[n:]
            else:
                self.items.extend([UNDEFINED] * (n - len(self.items)))
            return
        if _is_index(k):
            i = int(k)
            if i >= len(self.items):
                self.items.extend([UNDEFINED] * (i + 1 - len(self.items)))
            self.items[i] = v
            return
        self.props[k] = v

    def delete(self, k):
        if _is_index(k):
            i = int(k)
            if i < len(self.items):
                self.items[i] = UNDEFINED
            return True
        return JSObject.delete(self, k)


class JSFunction(JSObject):
    __slots__ = ("native", "node", "env", "this_val", "home", "is_arrow", "is_ctor", "fields", "parent", "name",
                 "body", "params", "is_async", "derived")

    def __init__(self, name="", native=None):
        JSObject.__init__(self, FUNCTION_PROTO, "Function")
        self.native = native
        self.name = name or ""
        self.node = self.env = self.this_val = self.home = self.body = self.params = None
        self.is_arrow = self.is_ctor = self.is_async = self.derived = False
        self.fields = None
        self.parent = None

    def has_own(self, k):
        return k in ("name", "length") or k in self.props

    def get_own(self, k):
        if k in self.props:
            return self.props[k]
        if k == "name":
            return self.name
        if k == "length":
            return float(len(self.params[0])) if self.params else 0.0
        return UNDEFINED


class JSMapObj(JSObject):
    __slots__ = ("data", "seqs", "counter")  # data: normalised key -> (key, value), in insertion order

    def __init__(self, proto, cls):
        JSObject.__init__(self, proto, cls)
        self.data = {}
        self.seqs = {}     # normalised key -> insertion number (a delete + re-add gets a new one)
        self.counter = 0

    def insert(self, nk, pair):
        """set / add: a new key goes to the end, an existing one keeps its place."""
        if nk not in self.data:
            self.counter += 1
            self.seqs[nk] = self.counter
        self.data[nk] = pair

    def remove(self, nk) -> bool:
        self.seqs.pop(nk, None)
        return self.data.pop(nk, None) is not None

    def clear(self):
        self.data.clear()
        self.seqs.clear()


class JSTypedArray(JSObject):
    """Uint8Array / Uint32Array / Int32Array / Float64Array ... over a (shared) bytearray: what js/pack.js fills."""
    __slots__ = ("buf", "fmt", "size", "off", "n")

    def __init__(self, proto, cls, buf, fmt, size, off, n):
        JSObject.__init__(self, proto, cls)
        self.buf, self.fmt, self.size, self.off, self.n = buf, fmt, size, off, n

    def own_keys(self):
        return [str(i) for i in range(self.n)] + list(self.props)

    def enumerable_keys(self):
        return self.own_keys()

    def has_own(self, k):
        if k in ("length", "byteLength", "byteOffset", "buffer", "BYTES_PER_ELEMENT"):
            return True
        if _is_index(k):
            return int(k) < self.n
        return k in self.props

    def get_own(self, k):
        if _is_index(k):
            i = int(k)
            if i >= self.n:
                return UNDEFINED
            import struct
            return float(struct.unpack_from(self.fmt, self.buf, self.off + i * self.size)[0])
        if k == "length":
            return float(self.n)
        if k == "byteLength":
            return float(self.n * self.size)
        if k == "byteOffset":
            return float(self.off)
        if k == "BYTES_PER_ELEMENT":
            return float(self.size)
        if k == "buffer":
            return self.props.get("%buffer", UNDEFINED)
        return self.props.get(k, UNDEFINED)

    def put_own(self, k, v):
        if _is_index(k):
            i = int(k)
            if i < self.n:
                import struct
                x = to_num(v)
                if self.fmt == "<d":
                    val = x
                elif self.fmt == "<f":
                    val = x
                else:
                    bits = self.size * 8
                    iv = 0 if (x != x or x in (math.inf, -math.inf)) else int(x) & ((1 << bits) - 1)
                    if self.fmt in ("<b", "<h", "<i") and iv >= 1 << (bits - 1):
                        iv -= 1 << bits
                    val = iv
                struct.pack_into(self.fmt, self.buf, self.off + i * self.size, val)
            return
        self.props[k] = v

    def raw(self) -> bytes:
        return bytes(self.buf[self.off:self.off + self.n * self.size])


class Env:
    __slots__ = ("vars", "parent", "fn")

    def __init__(self, parent=None, fn=False):
        self.vars = {}
        self.parent = parent
        self.fn = fn  # function-level scope (target of `var`)

    def lookup(self, name):
        e = self
        while e is not None:
            if name in e.vars:
                return e
            e = e.parent
        return None

    def declare(self, name, value):
        self.vars[name] = value

    def fn_scope(self):
        e = self
        while not e.fn and e.parent is not None:
            e = e.parent
        return e


# ----------------------------------------------------------------------------- conversions (ECMA-262 7.1, 7.2)
def typeof(v) -> str:
    if v is UNDEFINED:
        return "undefined"
    if v is None:
        return "object"
    if v is True or v is False:
        return "boolean"
    c = v.__class__
    if c is float:
        return "number"
    if c is str:
        return "string"
    if c is JSFunction:
        return "function"
    if c is JSSymbol:
        return "symbol"
    return "object"


def truthy(v) -> bool:
    if v is UNDEFINED or v is None or v is False:
        return False
    if v is True:
        return True
    c = v.__class__
    if c is float:
        return not (v == 0.0 or v != v)
    if c is str:
        return len(v) > 0
    return True


def is_obj(v) -> bool:
    return isinstance(v, JSObject)


def to_primitive(v, hint="default"):
    if not isinstance(v, JSObject):
        return v
    order = ("toString", "valueOf") if hint == "string" else ("valueOf", "toString")
    for name in order:
        f = v.get(name)
        if isinstance(f, JSFunction):
            r = call(f, v, [])
            if not isinstance(r, JSObject):
                return r
    raise throw_type("Cannot convert object to primitive value")


def to_num(v) -> float:
    c = v.__class__
    if c is float:
        return v
    if v is UNDEFINED:
        return math.nan
    if v is None or v is False:
        return 0.0
    if v is True:
        return 1.0
    if c is str:
        return string_to_number(v)
    if c is int:
        return float(v)
    if c is JSSymbol:
        raise throw_type("Cannot convert a Symbol value to a number")
    return to_num(to_primitive(v, "number"))


def to_str(v) -> str:
    c = v.__class__
    if c is str:
        return v
    if c is float:
        return number_to_string(v)
    if v is UNDEFINED:
        return "undefined"
    if v is None:
        return "null"
    if v is True:
        return "true"
    if v is False:
        return "false"
    if c is int:
        return number_to_string(float(v))
    if c is JSSymbol:
        raise throw_type("Cannot convert a Symbol value to a string")
    return to_str(to_primitive(v, "string"))


def to_key(v):
    if v.__class__ is str:
        return v
    if v.__class__ is JSSymbol:
        return v
    return to_str(v)


def to_int(v) -> int:
    n = to_num(v)
    if n != n:
        return 0
    if n in (math.inf, -math.inf):
        return 10 ** 18 if n > 0 else -10 ** 18
    return int(n)


def to_int32(v) -> int:
    n = to_num(v)
    if n != n or n in (math.inf, -math.inf):
        return 0
    i = int(n) & 0xFFFFFFFF
    return i - 0x100000000 if i >= 0x80000000 else i


def to_uint32(v) -> int:
    return to_int32(v) & 0xFFFFFFFF


def to_object(v):
    if v is UNDEFINED or v is None:
        raise throw_type("Cannot convert undefined or null to object")
    return v


def strict_eq(a, b) -> bool:
    ca, cb = a.__class__, b.__class__
    if ca is float or cb is float:
        return ca is cb and a == b
    if ca is str or cb is str:
        return ca is cb and a == b
    return a is b  # undefined, null, booleans (singletons), objects, symbols


def loose_eq(a, b) -> bool:
    ta, tb = typeof(a), typeof(b)
    if a is None:
        ta = "null"
    if b is None:
        tb = "null"
    if ta == tb:
        return strict_eq(a, b)
    if ta in ("null", "undefined") and tb in ("null", "undefined"):
        return True
    if ta in ("null", "undefined") or tb in ("null", "undefined"):
        return False
    if ta == "number" and tb == "string":
        return a == to_num(b)
    if ta == "string" and tb == "number":
        return to_num(a) == b
    if ta == "boolean":
        return loose_eq(to_num(a), b)
    if tb == "boolean":
        return loose_eq(a, to_num(b))
    if ta in ("object", "function") and tb in ("number", "string", "symbol"):
        return loose_eq(to_primitive(a), b)
    if tb in ("object", "function") and ta in ("number", "string", "symbol"):
        return loose_eq(a, to_primitive(b))
    return False


def _utf16(s: str) -> bytes:
    return s.encode("utf-16-be", "surrogatepass")


def _less(x, y, left_first=True):
    """IsLessThan(x, y) -> True | False | None (undefined)."""
    if left_first:
        px = to_primitive(x, "number")
        py = to_primitive(y, "number")
    else:
        py = to_primitive(y, "number")
        px = to_primitive(x, "number")
    if px.__class__ is str and py.__class__ is str:
        return _utf16(px) < _utf16(py)
    nx, ny = to_num(px), to_num(py)
    if nx != nx or ny != ny:
        return None
    return nx < ny


def same_value_zero(a, b) -> bool:
    if a.__class__ is float and b.__class__ is float:
        return a == b or (a != a and b != b)
    return strict_eq(a, b)


_NAN_KEY = ("nan",)


def map_key(v):
    """Normalised dict key implementing SameValueZero for Map / Set."""
    c = v.__class__
    if c is float:
        if v != v:
            return _NAN_KEY
        return ("n", v + 0.0 if v != 0 else 0.0)
    if c is str:
        return ("s", v)
    if v is True or v is False:
        return ("b", v)
    if v is None:
        return ("null",)
    if v is UNDEFINED:
        return ("undef",)
    return ("o", id(v))


# ----------------------------------------------------------------------------- errors
def make_error(kind: str, msg: str):
    proto = ERROR_PROTOS.get(kind, ERROR_PROTOS["Error"])
    e = JSObject(proto, "Error")
    e.define("message", msg)
    e.define("stack", f"{kind}: {msg}\n    at <minijs>")
    return e


def throw_type(msg: str) -> JSThrow:
    return JSThrow(make_error("TypeError", msg))


def throw_ref(msg: str) -> JSThrow:
    return JSThrow(make_error("ReferenceError", msg))


# ----------------------------------------------------------------------------- property access on any value
def get_member(obj, key):
    c = obj.__class__
    if c is str:
        if key == "length":
            return float(len(_utf16(obj)) // 2) if not obj.isascii() else float(len(obj))
        if _is_index(key):
            i = int(key)
            return obj[i] if i < len(obj) else UNDEFINED
        return STRING_PROTO.get(key, obj)
    if isinstance(obj, JSObject):
        if c is JSObject:
            v = obj.lookup(key)
            if v is _MISSING:
                return UNDEFINED
            if v.__class__ is Accessor:
                return call(v.get, obj, []) if v.get is not None else UNDEFINED
            return v
        return obj.get(key)
    if c is float:
        return NUMBER_PROTO.get(key, obj)
    if obj is True or obj is False:
        return BOOLEAN_PROTO.get(key, obj)
    if c is JSSymbol:
        if key == "description":
            return obj.desc
        return SYMBOL_PROTO.get(key, obj)
    raise throw_type(f"Cannot read properties of {to_str(obj)} (reading '{to_str(key) if key.__class__ is not JSSymbol else 'Symbol()'}')")


def set_member(obj, key, val):
    if isinstance(obj, JSObject):
        if obj.__class__ is JSObject and not obj.frozen:
            cur = obj.lookup(key)
            if cur.__class__ is Accessor:
                obj.set(key, val)
            else:
                obj.put_own(key, val)
        else:
            obj.set(key, val)
        return
    if obj is UNDEFINED or obj is None:
        raise throw_type(f"Cannot set properties of {to_str(obj)} (setting '{to_str(key)}')")
    # strict mode: creating a property on a primitive throws (src/bullet.js:122-124 relies on this)
    raise throw_type(f"Cannot create property '{to_str(key)}' on {typeof(obj)} '{to_str(obj)}'")


def has_property(obj, key) -> bool:
    if not isinstance(obj, JSObject):
        raise throw_type(f"Cannot use 'in' operator to search for '{to_str(key)}' in {to_str(obj)}")
    return obj.lookup(key) is not _MISSING


def instance_of(v, ctor) -> bool:
    if not isinstance(ctor, JSFunction):
        raise throw_type("Right-hand side of 'instanceof' is not callable")
    if not isinstance(v, JSObject):
        return False
    proto = ctor.get("prototype")
    o = v.proto
    while o is not None:
        if o is proto:
            return True
        o = o.proto
    return False


# ----------------------------------------------------------------------------- calling
class Interp:
    """Holds the global environment, the module cache and the (manual) timer / microtask queues."""
    current = None


def call(f, this, args):
    if not isinstance(f, JSFunction):
        raise throw_type(f"{to_str(f) if not isinstance(f, JSObject) else 'object'} is not a function")
    if f.native is not None:
        return f.native(this, args)
    if f.is_ctor:
        raise throw_type(f"Class constructor {f.name} cannot be invoked without 'new'")
    return _invoke(f, this, args, None)


def _bind_params(f, env, args):
    ps, rest = f.params
    n = len(args)
    for i, (target, default) in enumerate(ps):
        v = args[i] if i < n else UNDEFINED
        if v is UNDEFINED and default is not None:
            v = default(env)
        target(env, v)
    if rest is not None:
        rest(env, JSArray(list(args[len(ps):])))


def _invoke(f, this, args, new_target):
    env = Env(f.env, fn=True)
    if not f.is_arrow:
        env.vars["this"] = this
        env.vars["%home"] = f.home
        env.vars["%fn"] = f
        env.vars["%newtarget"] = new_target
        env.vars["arguments"] = _LazyArgs(args)
    _bind_params(f, env, args)
    if f.is_async:
        return _run_async(f, env)
    r = f.body(env)
    if r is not None and r is not BREAK and r is not CONTINUE:
        return r[1]
    return UNDEFINED


class _LazyArgs:
    __slots__ = ("args",)

    def __init__(self, args):
        self.args = args


def _run_async(f, env):
    p = new_promise()
    try:
        r = f.body(env)
        resolve_promise(p, r[1] if r is not None and r is not BREAK and r is not CONTINUE else UNDEFINED)
    except JSThrow as e:
        reject_promise(p, e.value)
    return p


def construct(f, args, new_target=None):
    if not isinstance(f, JSFunction):
        raise throw_type("not a constructor")
    nt = new_target or f
    if f.native is not None:
        ctor = f.props.get("%construct")
        if ctor is None:
            raise throw_type(f"{f.name} is not a constructor")
        return ctor(args, nt)
    if f.is_arrow:
        raise throw_type("not a constructor")
    if f.derived:
        env_this = _UNINIT
    else:
        proto = nt.get("prototype")
        env_this = JSObject(proto if isinstance(proto, JSObject) else OBJECT_PROTO)
        _init_fields(f, env_this)
    env = Env(f.env, fn=True)
    env.vars["this"] = env_this
    env.vars["%home"] = f.home
    env.vars["%fn"] = f
    env.vars["%newtarget"] = nt
    env.vars["arguments"] = _LazyArgs(args)
    _bind_params(f, env, args)
    r = f.body(env)
    if r is not None and r is not BREAK and r is not CONTINUE and isinstance(r[1], JSObject):
        return r[1]
    this = env.vars["this"]
    if this is _UNINIT:
        raise throw_ref("Must call super constructor in derived class before accessing 'this'")
    return this


_UNINIT = object()


def _init_fields(f, obj):
    if f.fields:
        env = Env(f.env, fn=True)
        env.vars["this"] = obj
        env.vars["%home"] = f.home
        for key, init in f.fields:
            k = key(env) if callable(key) else key
            obj.put_own(k, init(env) if init is not None else UNDEFINED)


# ----------------------------------------------------------------------------- promises (synchronous-friendly)
def new_promise():
    p = JSObject(PROMISE_PROTO, "Promise")
    p.define("%state", "pending")
    p.define("%value", UNDEFINED)
    p.define("%cbs", [])
    return p


def _settle(p, state, value):
    if p.props["%state"] != "pending":
        return
    p.props["%state"] = state
    p.props["%value"] = value
    cbs = p.props["%cbs"]
    p.props["%cbs"] = []
    for cb in cbs:
        Interp.current.microtasks.append(lambda cb=cb: cb(state, value))


def resolve_promise(p, value):
    if isinstance(value, JSObject) and value.cls == "Promise":
        _subscribe(value, lambda st, v: _settle(p, st, v))
        return
    if isinstance(value, JSObject):
        then = value.get("then")
        if isinstance(then, JSFunction):
            res = JSFunction("resolve", lambda t, a: resolve_promise(p, a[0] if a else UNDEFINED) or UNDEFINED)
            rej = JSFunction("reject", lambda t, a: reject_promise(p, a[0] if a else UNDEFINED) or UNDEFINED)
            Interp.current.microtasks.append(lambda: call(then, value, [res, rej]))
            return
    _settle(p, "fulfilled", value)


def reject_promise(p, reason):
    _settle(p, "rejected", reason)


def _subscribe(p, cb):
    if p.props["%state"] == "pending":
        p.props["%cbs"].append(cb)
    else:
        st, v = p.props["%state"], p.props["%value"]
        Interp.current.microtasks.append(lambda: cb(st, v))


def promise_then(p, on_ok, on_err):
    out = new_promise()

    def cb(state, value):
        h = on_ok if state == "fulfilled" else on_err
        if not isinstance(h, JSFunction):
            (_settle(out, state, value))
            return
        try:
            resolve_promise(out, call(h, UNDEFINED, [value]))
        except JSThrow as e:
            reject_promise(out, e.value)

    _subscribe(p, cb)
    return out


def await_value(v):
    """`await v` inside an async function that we run synchronously: drain queues until v settles."""
    if not (isinstance(v, JSObject) and v.cls == "Promise"):
        if isinstance(v, JSObject) and isinstance(v.get("then"), JSFunction):
            p = new_promise()
            resolve_promise(p, v)
            v = p
        else:
            return v
    it = Interp.current
    guard = 0
    while v.props["%state"] == "pending":
        if not it.run_microtasks() and not it.run_next_timer():
            raise JSThrow(make_error("Error", "await on a promise that never settles (minijs)"))
        guard += 1
        if guard > 100000:
            raise JSThrow(make_error("Error", "await did not settle"))
    if v.props["%state"] == "rejected":
        raise JSThrow(v.props["%value"])
    return v.props["%value"]


# ----------------------------------------------------------------------------- symbols
class JSSymbol:
    __slots__ = ("desc",)

    def __init__(self, desc):
        self.desc = desc

    def __repr__(self):
        return f"Symbol({self.desc})"


SYM_ITERATOR = JSSymbol("Symbol.iterator")


# ----------------------------------------------------------------------------- iteration
def iterate(v):
    """Python iterator over a JS iterable (arrays, strings, Map, Set, iterator objects)."""
    if v.__class__ is JSArray:
        i = 0
        while i < len(v.items):  # live, like ArrayIterator
            yield v.items[i]
            i += 1
        return
    if v.__class__ is str:
        yield from v
        return
    if v.__class__ is JSTypedArray:
        for i in range(v.n):
            yield v.get_own(str(i))
        return
    if v.__class__ is JSMapObj:
        if v.cls == "Map":
            for nk in _live_keys(v):
                k, val = v.data[nk]
                yield JSArray([k, val])
        else:
            for nk in _live_keys(v):
                yield v.data[nk][0]
        return
    if isinstance(v, JSObject):
        if v.cls == "%PyIter":
            yield from v.props["%it"]
            return
        f = v.get(SYM_ITERATOR)
        if isinstance(f, JSFunction):
            it = call(f, v, [])
            nxt = it.get("next")
            while True:
                r = call(nxt, it, [])
                if truthy(r.get("done")):
                    return
                yield r.get("value")
    raise throw_type(f"{typeof(v)} is not iterable")


def _live_keys(m):
    """Map / Set iteration is live (ECMA-262 24.1.5.1): entries are visited in insertion order, an entry deleted
    before it is reached is skipped, entries added during the iteration are visited, a deleted and re-added
    entry is visited again at its new place."""
    last = 0
    while True:
        pending = []
        for nk in reversed(m.data):  # the entries inserted after `last` are at the end
            seq = m.seqs[nk]
            if seq <= last:
                break
            pending.append((nk, seq))
        if not pending:
            return
        for nk, seq in reversed(pending):
            if m.seqs.get(nk) == seq:
                last = seq
                yield nk


def py_iter_object(gen):
    o = JSObject(ITER_PROTO, "%PyIter")
    o.define("%it", gen)
    return o


# ----------------------------------------------------------------------------- compiler: AST -> closures
class Compiler:
    def __init__(self, filename="<js>"):
        self.filename = filename

    # ---- statements
    def stmt(self, n):
        return getattr(self, "s_" + n[0])(n)

    def block_body(self, stmts):
        """Compile a statement list; function declarations are hoisted to the top of the block."""
        hoisted = [(s[1], self.func(s[2])) for s in stmts if s[0] == "funcdecl"]
        body = [self.stmt(s) for s in stmts if s[0] != "funcdecl"]
        var_names = []
        for s in stmts:
            _collect_vars(s, var_names)
        return hoisted, body, var_names

    def s_block(self, n, new_scope=True):
        hoisted, body, _ = self.block_body(n[1])

        def run(env):
            if new_scope:
                env = Env(env)
            for name, mk in hoisted:
                env.vars[name] = mk(env)
            for s in body:
                r = s(env)
                if r is not None:
                    return r
            return None
        return run

    def s_empty(self, n):
        return lambda env: None

    def s_expr(self, n):
        e = self.expr(n[1])

        def run(env):
            e(env)
            return None
        return run

    def s_var(self, n):
        kind = n[1]
        decls = [(self.pattern(t, declare=kind), self.expr(i) if i is not None else None, t) for t, i in n[2]]

        def run(env):
            for target, init, raw in decls:
                if init is None:
                    if kind == "var" and raw[0] == "id" and env.lookup(raw[1]) is not None:
                        continue
                    target(env, UNDEFINED)
                else:
                    v = init(env)
                    if v.__class__ is JSFunction and not v.name and raw[0] == "id":
                        v.name = raw[1]
                    target(env, v)
            return None
        return run

    def s_funcdecl(self, n):
        mk = self.func(n[2])
        name = n[1]

        def run(env):
            env.vars[name] = mk(env)
            return None
        return run

    def s_classdecl(self, n):
        mk = self.class_(n[2])
        name = n[1]

        def run(env):
            env.vars[name] = mk(env)
            return None
        return run

    def s_return(self, n):
        if n[1] is None:
            r = ("return", UNDEFINED)
            return lambda env: r
        e = self.expr(n[1])
        return lambda env: ("return", e(env))

    def s_if(self, n):
        test, cons = self.expr(n[1]), self.stmt(n[2])
        alt = self.stmt(n[3]) if n[3] is not None else None

        def run(env):
            if truthy(test(env)):
                return cons(env)
            if alt is not None:
                return alt(env)
            return None
        return run

    def s_for(self, n):
        init = self.stmt(n[1]) if n[1] is not None else None
        test = self.expr(n[2]) if n[2] is not None else None
        update = self.expr(n[3]) if n[3] is not None else None
        body = self.stmt(n[4])
        per_iter = n[1] is not None and n[1][0] == "var" and n[1][1] in ("let", "const")

        def run(env):
            loop = Env(env)
            if init is not None:
                init(loop)
            while True:
                if test is not None and not truthy(test(loop)):
                    return None
                r = body(loop)
                if r is not None:
                    if r is BREAK:
                        return None
                    if r is not CONTINUE:
                        return r
                if per_iter:  # fresh binding per iteration (closures capture the iteration's value)
                    nxt = Env(env)
                    nxt.vars.update(loop.vars)
                    loop = nxt
                if update is not None:
                    update(loop)
        return run

    def s_while(self, n):
        test, body = self.expr(n[1]), self.stmt(n[2])

        def run(env):
            while truthy(test(env)):
                r = body(env)
                if r is not None:
                    if r is BREAK:
                        return None
                    if r is not CONTINUE:
                        return r
            return None
        return run

    def s_dowhile(self, n):
        body, test = self.stmt(n[1]), self.expr(n[2])

        def run(env):
            while True:
                r = body(env)
                if r is not None:
                    if r is BREAK:
                        return None
                    if r is not CONTINUE:
                        return r
                if not truthy(test(env)):
                    return None
        return run

    def _for_each(self, n, keys_of):
        target = self.pattern(n[2], declare=n[1])
        right, body = self.expr(n[3]), self.stmt(n[4])

        def run(env):
            for item in keys_of(right(env)):
                loop = Env(env)
                target(loop, item)
                r = body(loop)
                if r is not None:
                    if r is BREAK:
                        return None
                    if r is not CONTINUE:
                        return r
            return None
        return run

    def s_forof(self, n):
        return self._for_each(n, iterate)

    def s_forin(self, n):
        def keys(v):
            if v is UNDEFINED or v is None:
                return
            if v.__class__ is str:
                for i in range(len(v)):
                    yield number_to_string(float(i))
                return
            if not isinstance(v, JSObject):
                return
            seen = set()
            o = v
            while o is not None:
                for k in o.own_keys():
                    if k.__class__ is JSSymbol or k in seen:
                        continue
                    seen.add(k)
                    if o.nonenum and k in o.nonenum:
                        continue
                    if o.has_own(k):  # deleted during iteration -> skipped
                        yield k
                o = o.proto
        return self._for_each(n, keys)

    def s_break(self, n):
        return lambda env: BREAK

    def s_continue(self, n):
        return lambda env: CONTINUE

    def s_throw(self, n):
        e = self.expr(n[1])

        def run(env):
            raise JSThrow(e(env))
        return run

    def s_try(self, n):
        blk = self.stmt(n[1])
        param = self.pattern(n[2], declare="let") if n[2] is not None else None
        handler = self.stmt(n[3]) if n[3] is not None else None
        final = self.stmt(n[4]) if n[4] is not None else None

        def run(env):
            try:
                try:
                    r = blk(env)
                except JSThrow as e:
                    if handler is None:
                        raise
                    henv = Env(env)
                    if param is not None:
                        param(henv, e.value)
                    r = handler(henv)
                except RecursionError:
                    if handler is None:
                        raise
                    henv = Env(env)
                    if param is not None:
                        param(henv, make_error("RangeError", "Maximum call stack size exceeded"))
                    r = handler(henv)
            finally:
                if final is not None:
                    fr = final(env)
                    if fr is not None:
                        return fr  # noqa: B012 - JS semantics: finally's completion overrides
            return r
        return run

    def s_switch(self, n):
        disc = self.expr(n[1])
        cases = [(self.expr(t) if t is not None else None, [self.stmt(s) for s in body]) for t, body in n[2]]

        def run(env):
            v = disc(env)
            env2 = Env(env)
            start = None
            for i, (t, _) in enumerate(cases):
                if t is not None and strict_eq(v, t(env2)):
                    start = i
                    break
            if start is None:
                for i, (t, _) in enumerate(cases):
                    if t is None:
                        start = i
                        break
            if start is None:
                return None
            for _, body in cases[start:]:
                for s in body:
                    r = s(env2)
                    if r is not None:
                        if r is BREAK:
                            return None
                        return r
            return None
        return run

    # ---- binding / assignment patterns -> f(env, value)
    def pattern(self, p, declare=None):
        k = p[0]
        if k == "id":
            name = p[1]
            if declare == "var":
                def bind(env, v):
                    env.fn_scope().vars[name] = v
            elif declare:
                def bind(env, v):
                    env.vars[name] = v
            else:
                def bind(env, v):
                    e = env.lookup(name)
                    if e is None:
                        raise throw_ref(f"{name} is not defined")
                    e.vars[name] = v
            return bind
        if k == "objpat":
            props = []
            for key, target, default in p[1]:
                kf = self.expr(key[1]) if isinstance(key, tuple) else key
                props.append((kf, self.pattern(target, declare), self.expr(default) if default is not None else None))
            rest = self.pattern(p[2], declare) if p[2] is not None else None

            def bind(env, v):
                if v is UNDEFINED or v is None:
                    raise throw_type(f"Cannot destructure '{to_str(v)}' as it is {to_str(v)}.")
                used = []
                for kf, target, default in props:
                    key = to_key(kf(env)) if callable(kf) else kf
                    used.append(key)
                    x = get_member(v, key)
                    if x is UNDEFINED and default is not None:
                        x = default(env)
                    target(env, x)
                if rest is not None:
                    o = JSObject(OBJECT_PROTO)
                    if isinstance(v, JSObject):
                        for key in v.enumerable_keys():
                            if key not in used:
                                o.put_own(key, v.get(key))
                    rest(env, o)
            return bind
        if k == "arrpat":
            elems = [None if e is None else (self.pattern(e[0], declare), self.expr(e[1]) if e[1] is not None else None)
                     for e in p[1]]
            rest = self.pattern(p[2], declare) if p[2] is not None else None

            def bind(env, v):
                it = iterate(v)
                for e in elems:
                    x = next(it, UNDEFINED)
                    if e is None:
                        continue
                    if x is UNDEFINED and e[1] is not None:
                        x = e[1](env)
                    e[0](env, x)
                if rest is not None:
                    rest(env, JSArray(list(it)))
            return bind
        if k == "member":
            obj, key = self.expr(p[1]), self.expr(p[2])

            def bind(env, v):
                set_member(obj(env), to_key(key(env)), v)
            return bind
        if k == "name":
            return self.pattern(("id", p[1]), declare)
        raise SyntaxError(f"bad assignment target {k}")

    # ---- functions / classes
    def func(self, n, home=None):
        _, name, ps, rest, body, is_arrow, is_async = n
        params = ([(self.pattern(t, "let"), self.expr(d) if d is not None else None) for t, d in ps],
                  self.pattern(rest, "let") if rest is not None else None)
        cbody = self.s_block(body, new_scope=False)

        def make(env, home_obj=None):
            f = JSFunction(name if isinstance(name, str) else "")
            f.env, f.body, f.params = env, cbody, params
            f.is_arrow, f.is_async = is_arrow, is_async
            f.home = home_obj
            if not is_arrow:
                proto = JSObject(OBJECT_PROTO)
                proto.define("constructor", f)
                f.define("prototype", proto)
            return f
        return make

    def class_(self, n):
        _, name, parent, members = n
        parent_e = self.expr(parent) if parent is not None else None
        ctor_node = None
        methods, fields, sfields = [], [], []
        for key, val, is_static, kind in members:
            kf = self.expr(key[1]) if isinstance(key, tuple) else key
            if kind == "field":
                (sfields if is_static else fields).append((kf, self.expr(val) if val is not None else None))
            elif key == "constructor" and not is_static:
                ctor_node = val
            else:
                methods.append((kf, self.func(val), is_static, kind))
        if ctor_node is None:
            if parent is not None:
                ctor_node = ("function", name, [], ("id", "args"),
                             ("block", [("expr", ("call", ("super",), [("spread", ("name", "args"))], False))]),
                             False, False)
            else:
                ctor_node = ("function", name, [], None, ("block", []), False, False)
        mk_ctor = self.func(ctor_node)

        def make(env):
            cenv = Env(env)
            parent_v = parent_e(env) if parent_e is not None else None
            proto = JSObject(OBJECT_PROTO)
            if parent_e is not None:
                if parent_v is None:
                    proto.proto = None
                else:
                    pp = parent_v.get("prototype")
                    proto.proto = pp if isinstance(pp, JSObject) else None
            ctor = mk_ctor(cenv, proto)
            ctor.name = name or ""
            ctor.is_ctor = True
            ctor.derived = parent_e is not None
            ctor.parent = parent_v
            if isinstance(parent_v, JSFunction):
                ctor.proto = parent_v
            ctor.fields = fields
            ctor.define("prototype", proto)
            proto.define("constructor", ctor)
            if name:
                cenv.vars[name] = ctor
            for kf, mk, is_static, kind in methods:
                target = ctor if is_static else proto
                key = to_key(kf(cenv)) if callable(kf) else kf
                f = mk(cenv, target)
                f.props.pop("prototype", None)
                if kind == "method":
                    f.name = key if isinstance(key, str) else ""
                    target.define(key, f)
                else:
                    acc = target.props.get(key)
                    if acc.__class__ is not Accessor:
                        acc = Accessor()
                        target.define(key, acc)
                    setattr(acc, kind, f)
            if sfields:
                senv = Env(cenv, fn=True)
                senv.vars["this"] = ctor
                for kf, init in sfields:
                    key = to_key(kf(senv)) if callable(kf) else kf
                    ctor.put_own(key, init(senv) if init is not None else UNDEFINED)
            return ctor
        return make

    # ---- expressions
    def expr(self, n):
        f = getattr(self, "e_" + n[0])(n)
        if n[0] in ("member", "call") and _has_optional(n):
            # an optional link short-circuits the WHOLE chain: one guard at the top of the chain
            def guarded(env):
                try:
                    return f(env)
                except _ShortCircuit:
                    return UNDEFINED
            return guarded
        return f

    def sub(self, n):
        """Object / callee position inside a member-call chain: no guard of its own."""
        if n[0] in ("member", "call"):
            return getattr(self, "e_" + n[0])(n)
        return self.expr(n)

    def e_num(self, n):
        v = float(n[1])
        return lambda env: v

    def e_str(self, n):
        v = n[1]
        return lambda env: v

    def e_bool(self, n):
        v = n[1]
        return lambda env: v

    def e_null(self, n):
        return lambda env: None

    def e_regex(self, n):
        body, flags = n[1], n[2]
        return lambda env: make_regexp(body, flags)

    def e_template(self, n):
        parts = [self.expr(p) for p in n[1]]

        def run(env):
            out = []
            for p in parts:
                v = p(env)
                out.append(v if v.__class__ is str else to_str(v))
            return "".join(out)
        return run

    def e_name(self, n):
        name = n[1]
        if name == "undefined":
            return lambda env: UNDEFINED

        def run(env):
            e = env
            while e is not None:
                vs = e.vars
                if name in vs:
                    v = vs[name]
                    if v.__class__ is _LazyArgs:
                        v = vs[name] = _make_arguments(v.args)
                    return v
                e = e.parent
            raise throw_ref(f"{name} is not defined")
        return run

    def e_this(self, n):
        def run(env):
            e = env
            while e is not None:
                if "this" in e.vars:
                    v = e.vars["this"]
                    if v is _UNINIT:
                        raise throw_ref("Must call super constructor before accessing 'this'")
                    return v
                e = e.parent
            return UNDEFINED
        return run

    def e_super(self, n):
        raise SyntaxError("'super' keyword unexpected here")

    def e_seq(self, n):
        es = [self.expr(x) for x in n[1]]

        def run(env):
            v = UNDEFINED
            for e in es:
                v = e(env)
            return v
        return run

    def e_function(self, n):
        mk = self.func(n)
        name = n[1]
        if name and not n[5]:
            def run(env):  # named function expression: its own name is in scope
                e = Env(env)
                f = mk(e)
                e.vars[name] = f
                return f
            return run
        return lambda env: mk(env, _home_of(env) if n[5] else None)

    def e_class(self, n):
        return self.class_(n)

    def e_array(self, n):
        elems = [None if x is None else (("s", self.expr(x[1])) if x[0] == "spread" else ("e", self.expr(x)))
                 for x in n[1]]

        def run(env):
            out = []
            for e in elems:
                if e is None:
                    out.append(UNDEFINED)
                elif e[0] == "s":
                    out.extend(iterate(e[1](env)))
                else:
                    out.append(e[1](env))
            return JSArray(out)
        return run

    def e_object(self, n):
        props = []
        for p in n[1]:
            if p[0] == "spread":
                props.append(("spread", self.expr(p[1])))
            elif p[0] == "accessor":
                props.append(("accessor", p[1], p[2], self.func(p[3])))
            else:
                key = p[1]
                kf = self.expr(key[1]) if isinstance(key, tuple) else key
                val = p[2]
                is_method = val[0] == "function" and not val[5] and isinstance(key, str) and val[1] == key
                props.append(("prop", kf, self.func(val) if is_method else self.expr(val), is_method))

        def run(env):
            o = JSObject(OBJECT_PROTO)
            for p in props:
                if p[0] == "prop":
                    key = to_key(p[1](env)) if callable(p[1]) else p[1]
                    if p[3]:
                        v = p[2](env, o)
                        v.props.pop("prototype", None)
                    else:
                        v = p[2](env)
                        if v.__class__ is JSFunction and not v.name and isinstance(key, str):
                            v.name = key
                    if key == "__proto__" and not callable(p[1]):
                        if isinstance(v, JSObject) or v is None:
                            o.proto = v
                        continue
                    o.put_own(key, v)
                elif p[0] == "spread":
                    src = p[1](env)
                    if isinstance(src, JSObject):
                        for k in src.enumerable_keys():
                            o.put_own(k, src.get(k))
                    elif src.__class__ is str:
                        for i, ch in enumerate(src):
                            o.put_own(str(i), ch)
                else:
                    _, kind, key, mk = p
                    acc = o.props.get(key)
                    if acc.__class__ is not Accessor:
                        acc = Accessor()
                        o.put_own(key, acc)
                    setattr(acc, kind, mk(env, o))
            return o
        return run

    def e_member(self, n):
        _, obj_n, key_n, optional = n
        if obj_n[0] == "super":
            key = self.expr(key_n)

            def run_super(env):
                home = _home_of(env)
                this = _this_of(env)
                p = home.proto if home is not None else None
                if p is None:
                    return UNDEFINED
                return p.get(to_key(key(env)), this)
            return run_super
        obj = self.sub(obj_n)
        if key_n[0] == "str":
            k = key_n[1]

            def run_static(env):
                o = obj(env)
                if optional and (o is UNDEFINED or o is None):
                    raise _ShortCircuit
                return get_member(o, k)
            return run_static
        key = self.expr(key_n)

        def run(env):
            o = obj(env)
            if optional and (o is UNDEFINED or o is None):
                raise _ShortCircuit
            kv = key(env)
            return get_member(o, kv if kv.__class__ is str else to_key(kv))
        return run

    def e_call(self, n):
        _, callee_n, args_n, optional = n
        args = [("s", self.expr(a[1])) if a[0] == "spread" else ("e", self.expr(a)) for a in args_n]
        simple = all(a[0] == "e" for a in args)
        arg_fs = [a[1] for a in args]

        def eval_args(env):
            if simple:
                return [f(env) for f in arg_fs]
            out = []
            for kind, f in args:
                if kind == "s":
                    out.extend(iterate(f(env)))
                else:
                    out.append(f(env))
            return out

        if callee_n[0] == "super":
            def run_super_call(env):
                fn = _lookup_special(env, "%fn")
                nt = _lookup_special(env, "%newtarget")
                parent = fn.parent
                a = eval_args(env)
                this = construct(parent, a, nt) if isinstance(parent, JSFunction) else JSObject(OBJECT_PROTO)
                e = env
                while e is not None:
                    if "this" in e.vars:
                        e.vars["this"] = this
                        break
                    e = e.parent
                _init_fields(fn, this)
                return UNDEFINED
            return run_super_call

        if callee_n[0] == "member":
            _, obj_n, key_n, mem_opt = callee_n
            is_super = obj_n[0] == "super"
            obj = self.sub(obj_n) if not is_super else None
            key = self.expr(key_n)
            static_key = key_n[1] if key_n[0] == "str" else None

            def run_method(env):
                if is_super:
                    this = _this_of(env)
                    home = _home_of(env)
                    k = static_key if static_key is not None else to_key(key(env))
                    f = home.proto.get(k, this) if home is not None and home.proto is not None else UNDEFINED
                else:
                    this = obj(env)
                    if mem_opt and (this is UNDEFINED or this is None):
                        raise _ShortCircuit
                    k = static_key if static_key is not None else to_key(key(env))
                    f = get_member(this, k)
                if optional and (f is UNDEFINED or f is None):
                    raise _ShortCircuit
                a = eval_args(env)
                if f.__class__ is not JSFunction:
                    raise throw_type(f"{_describe(callee_n)} is not a function")
                if f.native is not None:
                    return f.native(this, a)
                if f.is_ctor:
                    raise throw_type(f"Class constructor {f.name} cannot be invoked without 'new'")
                return _invoke(f, this, a, None)
            return run_method

        callee = self.sub(callee_n)

        def run(env):
            f = callee(env)
            if optional and (f is UNDEFINED or f is None):
                raise _ShortCircuit
            a = eval_args(env)
            if f.__class__ is not JSFunction:
                raise throw_type(f"{_describe(callee_n)} is not a function")
            return call(f, UNDEFINED, a)
        return run

    def e_new(self, n):
        callee = self.expr(n[1])
        args = [("s", self.expr(a[1])) if a[0] == "spread" else ("e", self.expr(a)) for a in n[2]]

        def run(env):
            f = callee(env)
            out = []
            for kind, g in args:
                if kind == "s":
                    out.extend(iterate(g(env)))
                else:
                    out.append(g(env))
            if f.__class__ is not JSFunction:
                raise throw_type(f"{_describe(n[1])} is not a constructor")
            return construct(f, out)
        return run

    def e_cond(self, n):
        t, a, b = self.expr(n[1]), self.expr(n[2]), self.expr(n[3])
        return lambda env: a(env) if truthy(t(env)) else b(env)

    def e_logical(self, n):
        op, a, b = n[1], self.expr(n[2]), self.expr(n[3])
        if op == "&&":
            def run(env):
                v = a(env)
                return b(env) if truthy(v) else v
        elif op == "||":
            def run(env):
                v = a(env)
                return v if truthy(v) else b(env)
        else:
            def run(env):
                v = a(env)
                return b(env) if (v is UNDEFINED or v is None) else v
        return run

    def e_await(self, n):
        e = self.expr(n[1])
        return lambda env: await_value(e(env))

    def e_unary(self, n):
        op = n[1]
        if op == "typeof":
            if n[2][0] == "name":
                name = n[2][1]

                def run_typeof_name(env):
                    e = env.lookup(name)
                    return "undefined" if e is None else typeof(e.vars[name])
                return run_typeof_name
            a = self.expr(n[2])
            return lambda env: typeof(a(env))
        if op == "delete":
            t = n[2]
            if t[0] != "member":
                return lambda env: True
            obj, key = self.expr(t[1]), self.expr(t[2])

            def run_delete(env):
                o = obj(env)
                k = to_key(key(env))
                if isinstance(o, JSObject):
                    return o.delete(k)
                if o is UNDEFINED or o is None:
                    raise throw_type(f"Cannot convert undefined or null to object")
                return True
            return run_delete
        a = self.expr(n[2])
        if op == "!":
            return lambda env: not truthy(a(env))
        if op == "-":
            return lambda env: -to_num(a(env))
        if op == "+":
            return lambda env: to_num(a(env))
        if op == "~":
            return lambda env: float(~to_int32(a(env)))
        if op == "void":
            def run_void(env):
                a(env)
                return UNDEFINED
            return run_void
        raise SyntaxError(op)

    def e_update(self, n):
        _, op, prefix, target = n
        delta = 1.0 if op == "++" else -1.0
        get = self.expr(target)
        put = self.pattern(self._as_target(target))

        def run(env):
            old = to_num(get(env))
            put(env, old + delta)
            return old + delta if prefix else old
        return run

    def _as_target(self, e):
        if e[0] == "name":
            return ("id", e[1])
        return e

    def e_assign(self, n):
        _, op, left, right = n
        rhs = self.expr(right)
        if op == "=":
            if left[0] == "member" and left[1][0] != "super":
                obj, key_n = self.expr(left[1]), left[2]
                if key_n[0] == "str":
                    k = key_n[1]

                    def run_member_static(env):
                        o = obj(env)
                        v = rhs(env)
                        set_member(o, k, v)
                        return v
                    return run_member_static
                key = self.expr(key_n)

                def run_member(env):
                    o = obj(env)
                    kv = key(env)
                    k = kv if kv.__class__ is str else to_key(kv)
                    v = rhs(env)
                    set_member(o, k, v)
                    return v
                return run_member
            put = self.pattern(left)
            is_name = left[0] == "id"

            def run(env):
                v = rhs(env)
                if is_name and v.__class__ is JSFunction and not v.name:
                    v.name = left[1]
                put(env, v)
                return v
            return run
        get = self.expr(left)
        put = self.pattern(self._as_target(left))
        if op in ("&&=", "||=", "??="):
            def run_logical(env):
                cur = get(env)
                if op == "&&=":
                    go = truthy(cur)
                elif op == "||=":
                    go = not truthy(cur)
                else:
                    go = cur is UNDEFINED or cur is None
                if not go:
                    return cur
                v = rhs(env)
                put(env, v)
                return v
            return run_logical
        bop = BINOPS[op[:-1]]

        def run_compound(env):
            v = bop(get(env), rhs(env))
            put(env, v)
            return v
        return run_compound

    def e_binary(self, n):
        op, a, b = n[1], self.expr(n[2]), self.expr(n[3])
        f = BINOPS[op]
        if op == "===":
            def run_seq(env):
                x, y = a(env), b(env)
                cx, cy = x.__class__, y.__class__
                if cx is float or cy is float or cx is str or cy is str:
                    return cx is cy and x == y
                return x is y
            return run_seq
        return lambda env: f(a(env), b(env))

    def e_spread(self, n):
        raise SyntaxError("spread element outside call / array / object")


class _ShortCircuit(Exception):
    pass


def _has_optional(n) -> bool:
    """Is there an optional link in this member/call chain (at or below n)?"""
    while n[0] in ("member", "call"):
        if n[3]:
            return True
        n = n[1]
    return False


def _describe(n) -> str:
    if n[0] == "name":
        return n[1]
    if n[0] == "member":
        k = n[2][1] if n[2][0] == "str" else "[...]"
        return f"{_describe(n[1])}.{k}"
    if n[0] == "this":
        return "this"
    return "expression"


def _lookup_special(env, name):
    e = env
    while e is not None:
        if name in e.vars:
            return e.vars[name]
        e = e.parent
    return None


def _home_of(env):
    return _lookup_special(env, "%home")


def _this_of(env):
    v = _lookup_special(env, "this")
    if v is _UNINIT:
        raise throw_ref("Must call super constructor before accessing 'this'")
    return v


def _make_arguments(args):
    a = JSArray(list(args))
    a.cls = "Arguments"
    return a


def _collect_vars(s, out):
    return  # `var` is bound at execution time in the function scope (hoisting reads are not used by the reference)


# ----------------------------------------------------------------------------- operators
def _add(a, b):
    if a.__class__ is float and b.__class__ is float:
        return a + b
    if a.__class__ is str and b.__class__ is str:
        return a + b
    pa, pb = to_primitive(a), to_primitive(b)
    if pa.__class__ is str or pb.__class__ is str:
        return to_str(pa) + to_str(pb)
    return to_num(pa) + to_num(pb)


def _div(a, b):
    x, y = to_num(a), to_num(b)
    if y == 0:
        if x != x or x == 0:
            return math.nan
        neg = (math.copysign(1, x) < 0) != (math.copysign(1, y) < 0)
        return -math.inf if neg else math.inf
    return x / y


def _mod(a, b):
    x, y = to_num(a), to_num(b)
    if x != x or y != y or x in (math.inf, -math.inf) or y == 0:
        return math.nan
    if y in (math.inf, -math.inf):
        return x
    return math.fmod(x, y)


def _pow(a, b):
    x, y = to_num(a), to_num(b)
    try:
        return float(x ** y)
    except (OverflowError, ZeroDivisionError):
        return math.inf
    except TypeError:
        return math.nan


def _mul(a, b):
    x, y = to_num(a), to_num(b)
    try:
        return x * y
    except OverflowError:
        return math.inf


def _lt(a, b):
    r = _less(a, b)
    return r is True


def _gt(a, b):
    r = _less(b, a, left_first=False)
    return r is True


def _le(a, b):
    r = _less(b, a, left_first=False)
    return r is False


def _ge(a, b):
    r = _less(a, b)
    return r is False


BINOPS = {
    "+": _add,
    "-": lambda a, b: to_num(a) - to_num(b),
    "*": _mul,
    "/": _div,
    "%": _mod,
    "**": _pow,
    "===": strict_eq,
    "!==": lambda a, b: not strict_eq(a, b),
    "==": loose_eq,
    "!=": lambda a, b: not loose_eq(a, b),
    "<": _lt, ">": _gt, "<=": _le, ">=": _ge,
    "in": lambda a, b: has_property(b, to_key(a)),
    "instanceof": instance_of,
    "&": lambda a, b: float(to_int32(a) & to_int32(b)),
    "|": lambda a, b: float(to_int32(a) | to_int32(b)),
    "^": lambda a, b: float(to_int32(a) ^ to_int32(b)),
    "<<": lambda a, b: float(to_int32(to_int32(a) << (to_uint32(b) & 31))),
    ">>": lambda a, b: float(to_int32(a) >> (to_uint32(b) & 31)),
    ">>>": lambda a, b: float(to_uint32(a) >> (to_uint32(b) & 31)),
}


# ----------------------------------------------------------------------------- RegExp (translated to Python's re)
def make_regexp(body: str, flags: str):
    r = JSObject(REGEXP_PROTO, "RegExp")
    pyflags = 0
    if "i" in flags:
        pyflags |= re.IGNORECASE
    if "m" in flags:
        pyflags |= re.MULTILINE
    if "s" in flags:
        pyflags |= re.DOTALL
    src = re.sub(r"\(\?<([A-Za-z_]\w*)>", r"(?P<\1>", body)
    src = src.replace("\\d", "[0-9]") if "u" not in flags else src
    try:
        r.define("%re", re.compile(src, pyflags))
    except re.error as e:
        raise JSThrow(make_error("SyntaxError", f"Invalid regular expression: /{body}/: {e}"))
    r.define("source", body)
    r.define("flags", flags)
    r.define("global", "g" in flags)
    r.define("lastIndex", 0.0)
    return r


def _expand_replacement(tpl: str, m) -> str:
    out, i = [], 0
    while i < len(tpl):
        c = tpl[i]
        if c == "$" and i + 1 < len(tpl):
            d = tpl[i + 1]
            if d == "$":
                out.append("$")
                i += 2
                continue
            if d == "&":
                out.append(m.group(0))
                i += 2
                continue
            if d.isdigit():
                j = i + 2 if not (i + 2 < len(tpl) and tpl[i + 2].isdigit() and int(tpl[i + 1:i + 3]) <= m.re.groups) else i + 3
                g = int(tpl[i + 1:j])
                if 1 <= g <= m.re.groups:
                    out.append(m.group(g) or "")
                    i = j
                    continue
        out.append(c)
        i += 1
    return "".join(out)


# ----------------------------------------------------------------------------- JSON
def json_quote(s: str) -> str:
    out = ['"']
    for ch in s:
        o = ord(ch)
        if ch == '"':
            out.append('\\"')
        elif ch == "\\":
            out.append("\\\\")
        elif ch == "\n":
            out.append("\\n")
        elif ch == "\r":
            out.append("\\r")
        elif ch == "\t":
            out.append("\\t")
        elif ch == "\b":
            out.append("\\b")
        elif ch == "\f":
            out.append("\\f")
        elif o < 0x20 or 0xD800 <= o <= 0xDFFF:
            out.append(f"\\u{o:04x}")
        else:
            out.append(ch)
    out.append('"')
    return "".join(out)


def json_stringify(value, replacer=UNDEFINED, space=UNDEFINED):
    if space.__class__ is float:
        gap = " " * max(0, min(10, int(space)))
    elif space.__class__ is str:
        gap = space[:10]
    else:
        gap = ""
    rep_fn = replacer if isinstance(replacer, JSFunction) else None
    allow = None
    if replacer.__class__ is JSArray:
        allow = [to_str(x) for x in replacer.items]
    stack = []

    def ser(holder, key, v, indent):
        if isinstance(v, JSObject):
            tj = v.get("toJSON")
            if isinstance(tj, JSFunction):
                v = call(tj, v, [key])
        if rep_fn is not None:
            v = call(rep_fn, holder, [key, v])
        if v is None:
            return "null"
        if v is True:
            return "true"
        if v is False:
            return "false"
        c = v.__class__
        if c is str:
            return json_quote(v)
        if c is float:
            return number_to_string(v) if math.isfinite(v) else "null"
        if v is UNDEFINED or c is JSFunction or c is JSSymbol:
            return None
        if any(v is s for s in stack):
            raise throw_type("Converting circular structure to JSON")
        stack.append(v)
        inner = indent + gap
        if c is JSArray:
            parts = []
            for i, x in enumerate(v.items):
                s = ser(v, str(i), x, inner)
                parts.append("null" if s is None else s)
            if not parts:
                out = "[]"
            elif gap:
                out = "[\n" + inner + (",\n" + inner).join(parts) + "\n" + indent + "]"
            else:
                out = "[" + ",".join(parts) + "]"
        else:
            parts = []
            keys = allow if allow is not None else [k for k in v.enumerable_keys() if k.__class__ is str]
            for k in keys:
                s = ser(v, k, v.get(k), inner)
                if s is not None:
                    parts.append(json_quote(k) + (": " if gap else ":") + s)
            if not parts:
                out = "{}"
            elif gap:
                out = "{\n" + inner + (",\n" + inner).join(parts) + "\n" + indent + "}"
            else:
                out = "{" + ",".join(parts) + "}"
        stack.pop()
        return out

    wrapper = JSObject(OBJECT_PROTO)
    wrapper.put_own("", value)
    r = ser(wrapper, "", value, "")
    return UNDEFINED if r is None else r


def json_parse(text: str, reviver=UNDEFINED):
    import json

    def conv(x):
        if isinstance(x, dict):
            o = JSObject(OBJECT_PROTO)
            for k, v in x.items():
                o.put_own(k, conv(v))
            return o
        if isinstance(x, list):
            return JSArray([conv(v) for v in x])
        if isinstance(x, bool) or x is None or isinstance(x, str):
            return x
        return float(x)

    try:
        v = conv(json.loads(text))
    except (ValueError, TypeError) as e:
        raise JSThrow(make_error("SyntaxError", f"Unexpected token in JSON: {e}"))
    if isinstance(reviver, JSFunction):
        def walk(holder, key):
            val = holder.get(key)
            if isinstance(val, JSObject):
                for k in list(val.enumerable_keys()):
                    nv = walk(val, k)
                    if nv is UNDEFINED:
                        val.delete(k)
                    else:
                        val.put_own(k, nv)
            return call(reviver, holder, [key, val])
        root = JSObject(OBJECT_PROTO)
        root.put_own("", v)
        return walk(root, "")
    return v


# ----------------------------------------------------------------------------- prototypes (filled by builtins.py)
OBJECT_PROTO = JSObject(None)
FUNCTION_PROTO = JSObject(OBJECT_PROTO, "Function")
ARRAY_PROTO = JSObject(OBJECT_PROTO, "Array")
STRING_PROTO = JSObject(OBJECT_PROTO, "String")
NUMBER_PROTO = JSObject(OBJECT_PROTO, "Number")
BOOLEAN_PROTO = JSObject(OBJECT_PROTO, "Boolean")
SYMBOL_PROTO = JSObject(OBJECT_PROTO, "Symbol")
REGEXP_PROTO = JSObject(OBJECT_PROTO, "RegExp")
PROMISE_PROTO = JSObject(OBJECT_PROTO, "Promise")
MAP_PROTO = JSObject(OBJECT_PROTO, "Map")
SET_PROTO = JSObject(OBJECT_PROTO, "Set")
DATE_PROTO = JSObject(OBJECT_PROTO, "Date")
TYPED_PROTO = JSObject(OBJECT_PROTO, "TypedArray")
BUFFER_PROTO = JSObject(OBJECT_PROTO, "ArrayBuffer")
ITER_PROTO = JSObject(OBJECT_PROTO, "Iterator")
ERROR_PROTOS = {}


def compile_program(src: str, filename="<js>"):
    ast = parse(src)
    return Compiler(filename).s_block(ast, new_scope=False)


def find_reference_root():
    return os.environ.get("BULLET_REFERENCE", "/root/reference")
