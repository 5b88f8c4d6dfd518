"""Global objects for the minijs interpreter: the subset of the ECMAScript standard library and of
Node's module system that the reference's sources touch.  TEST INFRASTRUCTURE ONLY.

Time and randomness are deterministic and under the caller's control (`Runtime.now`, `Runtime.random`):
`Date.now()` advances by 1 ms per call so `lastModified` ordering is total, `Math.random()` is an LCG.
Timers never fire on their own; `Runtime.run_timers()` fires what is due.
"""
from __future__ import annotations

import math
import os

from . import interp as I
from .interp import (UNDEFINED, JSObject, JSArray, JSFunction, JSMapObj, JSTypedArray, JSThrow, JSSymbol, Env, Accessor,
                     call, construct, to_str, to_num, to_key, to_int, truthy, typeof, strict_eq, same_value_zero,
                     map_key, iterate, py_iter_object, make_error, throw_type, number_to_string)


def native(name, fn, nargs=0):
    f = JSFunction(name, fn)
    f.define("length", float(nargs))
    return f


def method(proto, name, nargs=0):
    def deco(fn):
        proto.define(name, native(name, fn, nargs))
        return fn
    return deco


def arg(a, i):
    return a[i] if i < len(a) else UNDEFINED


def new_object():
    return JSObject(I.OBJECT_PROTO)


def from_py(x):
    if isinstance(x, dict):
        o = new_object()
        for k, v in x.items():
            o.put_own(str(k), from_py(v))
        return o
    if isinstance(x, (list, tuple)):
        return JSArray([from_py(v) for v in x])
    if isinstance(x, bool) or x is None or isinstance(x, str) or x is UNDEFINED:
        return x
    if isinstance(x, (int, float)):
        return float(x)
    return x


def to_py(v, _depth=0):
    """JS value -> plain Python (dict keeps JS own-key order; numbers stay float; undefined -> UNDEFINED)."""
    if v.__class__ is JSArray:
        return [to_py(x, _depth + 1) for x in v.items]
    if v.__class__ is JSMapObj:
        if v.cls == "Map":
            return {"%Map": [[to_py(k, _depth + 1), to_py(x, _depth + 1)] for k, x in v.data.values()]}
        return {"%Set": [to_py(k, _depth + 1) for k, _ in v.data.values()]}
    if isinstance(v, JSFunction):
        return f"%function {v.name}"
    if isinstance(v, JSObject):
        return {k: to_py(v.get(k), _depth + 1) for k in v.enumerable_keys() if k.__class__ is str}
    return v


# ----------------------------------------------------------------------------- Object
def _setup_object(g):
    OP = I.OBJECT_PROTO

    @method(OP, "hasOwnProperty", 1)
    def _(this, a):
        if not isinstance(this, JSObject):
            return False
        return this.has_own(to_key(arg(a, 0)))

    @method(OP, "toString")
    def _(this, a):
        if this is UNDEFINED:
            return "[object Undefined]"
        if this is None:
            return "[object Null]"
        if isinstance(this, JSObject):
            cls = this.cls if this.cls in ("Array", "Function", "Error", "Date", "RegExp", "Arguments") else "Object"
            return f"[object {cls}]"
        return f"[object {typeof(this).capitalize()}]"

    @method(OP, "toLocaleString")
    def _(this, a):
        return to_str(this)

    @method(OP, "valueOf")
    def _(this, a):
        return this

    @method(OP, "isPrototypeOf", 1)
    def _(this, a):
        o = arg(a, 0)
        if not isinstance(o, JSObject):
            return False
        o = o.proto
        while o is not None:
            if o is this:
                return True
            o = o.proto
        return False

    @method(OP, "propertyIsEnumerable", 1)
    def _(this, a):
        k = to_key(arg(a, 0))
        return isinstance(this, JSObject) and this.has_own(k) and not (this.nonenum and k in this.nonenum)

    OP.define("__proto__", Accessor(native("get __proto__", lambda t, a: t.proto if isinstance(t, JSObject) else None),
                                    native("set __proto__", lambda t, a: _set_proto(t, arg(a, 0)))))

    def object_ctor(this, a):
        v = arg(a, 0)
        if v is UNDEFINED or v is None:
            return new_object()
        return v
    O = native("Object", object_ctor, 1)
    O.define("%construct", lambda a, nt: object_ctor(None, a))
    O.define("prototype", OP)
    OP.define("constructor", O)

    def keys(this, a):
        o = arg(a, 0)
        if o.__class__ is str:
            return JSArray([str(i) for i in range(len(o))])
        if not isinstance(o, JSObject):
            if o is UNDEFINED or o is None:
                raise throw_type("Cannot convert undefined or null to object")
            return JSArray([])
        return JSArray([k for k in o.enumerable_keys() if k.__class__ is str])

    def values(this, a):
        o = arg(a, 0)
        return JSArray([I.get_member(o, k) for k in keys(None, [o]).items])

    def entries(this, a):
        o = arg(a, 0)
        return JSArray([JSArray([k, I.get_member(o, k)]) for k in keys(None, [o]).items])

    def assign(this, a):
        t = arg(a, 0)
        if not isinstance(t, JSObject):
            raise throw_type("Cannot convert undefined or null to object")
        for s in a[1:]:
            if isinstance(s, JSObject):
                for k in s.enumerable_keys():
                    I.set_member(t, k, s.get(k))
        return t

    def from_entries(this, a):
        o = new_object()
        for e in iterate(arg(a, 0)):
            o.put_own(to_key(I.get_member(e, "0")), I.get_member(e, "1"))
        return o

    def create(this, a):
        p = arg(a, 0)
        if not (isinstance(p, JSObject) or p is None):
            raise throw_type("Object prototype may only be an Object or null")
        o = JSObject(p)
        props = arg(a, 1)
        if isinstance(props, JSObject):
            for k in props.enumerable_keys():
                _define_property(o, k, props.get(k))
        return o

    def freeze(this, a):
        o = arg(a, 0)
        if isinstance(o, JSObject):
            o.frozen = True
        return o

    def get_proto(this, a):
        o = arg(a, 0)
        if isinstance(o, JSObject):
            return o.proto
        if o.__class__ is str:
            return I.STRING_PROTO
        if o.__class__ is float:
            return I.NUMBER_PROTO
        if o is True or o is False:
            return I.BOOLEAN_PROTO
        raise throw_type("Cannot convert undefined or null to object")

    def define_property(this, a):
        o = arg(a, 0)
        if not isinstance(o, JSObject):
            raise throw_type("Object.defineProperty called on non-object")
        _define_property(o, to_key(arg(a, 1)), arg(a, 2))
        return o

    def get_own_names(this, a):
        o = arg(a, 0)
        return JSArray([k for k in o.own_keys() if k.__class__ is str]) if isinstance(o, JSObject) else JSArray([])

    def get_own_desc(this, a):
        o, k = arg(a, 0), to_key(arg(a, 1))
        if not isinstance(o, JSObject) or not o.has_own(k):
            return UNDEFINED
        d = new_object()
        v = o.props.get(k, _NOPE)
        if v.__class__ is Accessor:
            d.put_own("get", v.get if v.get is not None else UNDEFINED)
            d.put_own("set", v.set if v.set is not None else UNDEFINED)
        else:
            d.put_own("value", o.get_own(k))
            d.put_own("writable", not o.frozen)
        d.put_own("enumerable", not (o.nonenum and k in o.nonenum))
        d.put_own("configurable", not o.frozen)
        return d

    for name, fn, n in [("keys", keys, 1), ("values", values, 1), ("entries", entries, 1), ("assign", assign, 2),
                        ("fromEntries", from_entries, 1), ("create", create, 2), ("freeze", freeze, 1),
                        ("getPrototypeOf", get_proto, 1), ("defineProperty", define_property, 3),
                        ("getOwnPropertyNames", get_own_names, 1), ("getOwnPropertyDescriptor", get_own_desc, 2),
                        ("isFrozen", lambda t, a: isinstance(arg(a, 0), JSObject) and arg(a, 0).frozen, 1),
                        ("seal", lambda t, a: arg(a, 0), 1),
                        ("setPrototypeOf", lambda t, a: _set_proto(arg(a, 0), arg(a, 1)) or arg(a, 0), 2),
                        ("is", lambda t, a: _same_value(arg(a, 0), arg(a, 1)), 2)]:
        O.define(name, native(name, fn, n))
    g["Object"] = O


_NOPE = object()


def _same_value(a, b):
    if a.__class__ is float and b.__class__ is float:
        if a != a and b != b:
            return True
        return a == b and math.copysign(1, a) == math.copysign(1, b)
    return strict_eq(a, b)


def _set_proto(o, p):
    if isinstance(o, JSObject) and (isinstance(p, JSObject) or p is None):
        o.proto = p
    return UNDEFINED


def _define_property(o, k, desc):
    if not isinstance(desc, JSObject):
        raise throw_type("Property description must be an object")
    enumerable = truthy(desc.get("enumerable"))
    if desc.lookup("get") is not I._MISSING or desc.lookup("set") is not I._MISSING:
        g, s = desc.get("get"), desc.get("set")
        o.define(k, Accessor(g if isinstance(g, JSFunction) else None, s if isinstance(s, JSFunction) else None),
                 enumerable)
    else:
        o.define(k, desc.get("value"), enumerable)
    if enumerable and o.nonenum:
        o.nonenum.discard(k)


# ----------------------------------------------------------------------------- Function
def _setup_function(g):
    FP = I.FUNCTION_PROTO

    @method(FP, "call", 1)
    def _(this, a):
        return call(this, arg(a, 0), list(a[1:]))

    @method(FP, "apply", 2)
    def _(this, a):
        arr = arg(a, 1)
        return call(this, arg(a, 0), list(iterate(arr)) if isinstance(arr, JSObject) else [])

    @method(FP, "bind", 1)
    def _(this, a):
        target, bound_this, pre = this, arg(a, 0), list(a[1:])
        f = native(f"bound {getattr(target, 'name', '')}", lambda t, args: call(target, bound_this, pre + list(args)))
        f.define("%construct", lambda args, nt: construct(target, pre + list(args)))
        return f

    @method(FP, "toString")
    def _(this, a):
        return f"function {getattr(this, 'name', '')}() {{ [native code] }}"

    F = native("Function", lambda t, a: (_ for _ in ()).throw(throw_type("Function constructor is not supported")))
    F.define("prototype", FP)
    FP.define("constructor", F)
    g["Function"] = F


# ----------------------------------------------------------------------------- Array
def _arr(this):
    if this.__class__ is not JSArray:
        raise throw_type("Array.prototype method called on a non-array")
    return this.items


def _rel_index(v, n, default):
    if v is UNDEFINED:
        return default
    i = to_int(v)
    return max(n + i, 0) if i < 0 else min(i, n)


def _default_compare(x, y):
    if x is UNDEFINED:
        return 0 if y is UNDEFINED else 1
    if y is UNDEFINED:
        return -1
    sx, sy = I._utf16(to_str(x)), I._utf16(to_str(y))
    return -1 if sx < sy else (1 if sx > sy else 0)


def _setup_array(g):
    AP = I.ARRAY_PROTO
    import functools

    @method(AP, "push", 1)
    def _(this, a):
        items = _arr(this)
        items.extend(a)
        return float(len(items))

    @method(AP, "pop")
    def _(this, a):
        items = _arr(this)
        return items.pop() if items else UNDEFINED

    @method(AP, "shift")
    def _(this, a):
        items = _arr(this)
        return items.pop(0) if items else UNDEFINED

    @method(AP, "unshift", 1)
    def _(this, a):
        items = _arr(this)
        items[0:0] = list(a)
        return float(len(items))

    @method(AP, "slice", 2)
    def _(this, a):
        items = _arr(this)
        n = len(items)
        return JSArray(items[_rel_index(arg(a, 0), n, 0):_rel_index(arg(a, 1), n, n)])

    @method(AP, "splice", 2)
    def _(this, a):
        items = _arr(this)
        n = len(items)
        if not a:
            return JSArray([])
        start = _rel_index(a[0], n, 0)
        cnt = n - start if len(a) < 2 else max(0, min(to_int(a[1]), n - start))
        removed = items[start:start + cnt]
        items[start:start + cnt] = list(a[2:])
        return JSArray(removed)

    @method(AP, "concat", 1)
    def _(this, a):
        out = list(_arr(this))
        for x in a:
            if x.__class__ is JSArray:
                out.extend(x.items)
            else:
                out.append(x)
        return JSArray(out)

    @method(AP, "join", 1)
    def _(this, a):
        sep = "," if arg(a, 0) is UNDEFINED else to_str(a[0])
        return sep.join("" if (x is UNDEFINED or x is None) else to_str(x) for x in _arr(this))

    @method(AP, "toString")
    def _(this, a):
        return ",".join("" if (x is UNDEFINED or x is None) else to_str(x) for x in _arr(this))

    @method(AP, "reverse")
    def _(this, a):
        _arr(this).reverse()
        return this

    @method(AP, "indexOf", 1)
    def _(this, a):
        items, x = _arr(this), arg(a, 0)
        for i in range(_rel_index(arg(a, 1), len(items), 0), len(items)):
            if strict_eq(items[i], x):
                return float(i)
        return -1.0

    @method(AP, "lastIndexOf", 1)
    def _(this, a):
        items, x = _arr(this), arg(a, 0)
        for i in range(len(items) - 1, -1, -1):
            if strict_eq(items[i], x):
                return float(i)
        return -1.0

    @method(AP, "includes", 1)
    def _(this, a):
        x = arg(a, 0)
        return any(same_value_zero(v, x) for v in _arr(this))

    def _each(this, a):
        items, f, t = _arr(this), arg(a, 0), arg(a, 1)
        if not isinstance(f, JSFunction):
            raise throw_type(f"{to_str(f) if not isinstance(f, JSObject) else 'object'} is not a function")
        n = len(items)
        i = 0
        while i < n and i < len(items):
            yield i, items[i], call(f, t, [items[i], float(i), this])
            i += 1

    @method(AP, "forEach", 1)
    def _(this, a):
        for _ in _each(this, a):
            pass
        return UNDEFINED

    @method(AP, "map", 1)
    def _(this, a):
        return JSArray([r for _, _, r in _each(this, a)])

    @method(AP, "filter", 1)
    def _(this, a):
        return JSArray([v for _, v, r in _each(this, a) if truthy(r)])

    @method(AP, "find", 1)
    def _(this, a):
        for _, v, r in _each(this, a):
            if truthy(r):
                return v
        return UNDEFINED

    @method(AP, "findIndex", 1)
    def _(this, a):
        for i, _, r in _each(this, a):
            if truthy(r):
                return float(i)
        return -1.0

    @method(AP, "some", 1)
    def _(this, a):
        return any(truthy(r) for _, _, r in _each(this, a))

    @method(AP, "every", 1)
    def _(this, a):
        return all(truthy(r) for _, _, r in _each(this, a))

    @method(AP, "reduce", 1)
    def _(this, a):
        items, f = _arr(this), arg(a, 0)
        i = 0
        if len(a) >= 2:
            acc = a[1]
        else:
            if not items:
                raise throw_type("Reduce of empty array with no initial value")
            acc, i = items[0], 1
        while i < len(items):
            acc = call(f, UNDEFINED, [acc, items[i], float(i), this])
            i += 1
        return acc

    @method(AP, "reduceRight", 1)
    def _(this, a):
        items, f = _arr(this), arg(a, 0)
        i = len(items) - 1
        if len(a) >= 2:
            acc = a[1]
        else:
            if not items:
                raise throw_type("Reduce of empty array with no initial value")
            acc, i = items[-1], i - 1
        while i >= 0:
            acc = call(f, UNDEFINED, [acc, items[i], float(i), this])
            i -= 1
        return acc

    @method(AP, "sort", 1)
    def _(this, a):
        items, f = _arr(this), arg(a, 0)
        undef = [x for x in items if x is UNDEFINED]
        rest = [x for x in items if x is not UNDEFINED]
        if isinstance(f, JSFunction):
            def cmp(x, y):
                r = to_num(call(f, UNDEFINED, [x, y]))
                return -1 if r < 0 else (1 if r > 0 else 0)
        else:
            cmp = _default_compare
        rest.sort(key=functools.cmp_to_key(cmp))  # stable, like V8's TimSort
        items[:] = rest + undef
        return this

    @method(AP, "flat")
    def _(this, a):
        depth = 1 if arg(a, 0) is UNDEFINED else to_int(a[0])

        def fl(xs, d):
            out = []
            for x in xs:
                if x.__class__ is JSArray and d > 0:
                    out.extend(fl(x.items, d - 1))
                else:
                    out.append(x)
            return out
        return JSArray(fl(_arr(this), depth))

    @method(AP, "flatMap", 1)
    def _(this, a):
        out = []
        for _, _, r in _each(this, a):
            if r.__class__ is JSArray:
                out.extend(r.items)
            else:
                out.append(r)
        return JSArray(out)

    @method(AP, "fill", 1)
    def _(this, a):
        items = _arr(this)
        n = len(items)
        for i in range(_rel_index(arg(a, 1), n, 0), _rel_index(arg(a, 2), n, n)):
            items[i] = arg(a, 0)
        return this

    @method(AP, "keys")
    def _(this, a):
        return py_iter_object(float(i) for i in range(len(_arr(this))))

    @method(AP, "values")
    def _(this, a):
        return py_iter_object(iterate(this))

    @method(AP, "entries")
    def _(this, a):
        return py_iter_object(JSArray([float(i), v]) for i, v in enumerate(_arr(this)))

    @method(AP, "at", 1)
    def _(this, a):
        items = _arr(this)
        i = to_int(arg(a, 0))
        if i < 0:
            i += len(items)
        return items[i] if 0 <= i < len(items) else UNDEFINED

    AP.define(I.SYM_ITERATOR, AP.props["values"])

    def array_ctor(this, a):
        if len(a) == 1 and a[0].__class__ is float:
            return JSArray([UNDEFINED] * int(a[0]))
        return JSArray(list(a))
    A = native("Array", array_ctor, 1)
    A.define("%construct", lambda a, nt: array_ctor(None, a))
    A.define("prototype", AP)
    AP.define("constructor", A)
    A.define("isArray", native("isArray", lambda t, a: arg(a, 0).__class__ is JSArray, 1))

    def array_from(this, a):
        src, f = arg(a, 0), arg(a, 1)
        if isinstance(src, JSObject) and src.__class__ is not JSArray and src.__class__ is not JSMapObj \
                and src.cls != "%PyIter" and not isinstance(src.get(I.SYM_ITERATOR), JSFunction):
            n = to_int(src.get("length"))
            items = [src.get(str(i)) for i in range(n)]
        else:
            items = list(iterate(src))
        if isinstance(f, JSFunction):
            items = [call(f, UNDEFINED, [v, float(i)]) for i, v in enumerate(items)]
        return JSArray(items)
    A.define("from", native("from", array_from, 1))
    A.define("of", native("of", lambda t, a: JSArray(list(a))))
    g["Array"] = A


# ----------------------------------------------------------------------------- String / Number / Boolean / Symbol
def _this_str(this):
    if this.__class__ is str:
        return this
    if this is UNDEFINED or this is None:
        raise throw_type("String.prototype method called on null or undefined")
    return to_str(this)


def _regexp_of(v):
    return v.props["%re"] if isinstance(v, JSObject) and v.cls == "RegExp" else None


def _match_array(m, s):
    arr = JSArray([m.group(0)] + [UNDEFINED if g is None else g for g in m.groups()])
    arr.props["index"] = float(m.start())
    arr.props["input"] = s
    gd = m.groupdict()
    if gd:
        arr.props["groups"] = from_py({k: (UNDEFINED if v is None else v) for k, v in gd.items()})
    else:
        arr.props["groups"] = UNDEFINED
    return arr


def _setup_string(g):
    SP = I.STRING_PROTO

    @method(SP, "toString")
    def _(this, a):
        return _this_str(this)

    @method(SP, "valueOf")
    def _(this, a):
        return _this_str(this)

    @method(SP, "charAt", 1)
    def _(this, a):
        s, i = _this_str(this), to_int(arg(a, 0))
        return s[i] if 0 <= i < len(s) else ""

    @method(SP, "charCodeAt", 1)
    def _(this, a):
        s, i = _this_str(this), to_int(arg(a, 0))
        return float(ord(s[i])) if 0 <= i < len(s) else math.nan

    @method(SP, "codePointAt", 1)
    def _(this, a):
        s, i = _this_str(this), to_int(arg(a, 0))
        return float(ord(s[i])) if 0 <= i < len(s) else UNDEFINED

    @method(SP, "indexOf", 1)
    def _(this, a):
        s = _this_str(this)
        return float(s.find(to_str(arg(a, 0)), max(0, min(to_int(arg(a, 1)) if len(a) > 1 else 0, len(s)))))

    @method(SP, "lastIndexOf", 1)
    def _(this, a):
        return float(_this_str(this).rfind(to_str(arg(a, 0))))

    @method(SP, "includes", 1)
    def _(this, a):
        return to_str(arg(a, 0)) in _this_str(this)[max(0, to_int(arg(a, 1))) if len(a) > 1 else 0:]

    @method(SP, "startsWith", 1)
    def _(this, a):
        return _this_str(this).startswith(to_str(arg(a, 0)), max(0, to_int(arg(a, 1))) if len(a) > 1 else 0)

    @method(SP, "endsWith", 1)
    def _(this, a):
        s = _this_str(this)
        end = len(s) if arg(a, 1) is UNDEFINED else max(0, min(to_int(a[1]), len(s)))
        return s[:end].endswith(to_str(arg(a, 0)))

    @method(SP, "slice", 2)
    def _(this, a):
        s = _this_str(this)
        n = len(s)
        return s[_rel_index(arg(a, 0), n, 0):_rel_index(arg(a, 1), n, n)]

    @method(SP, "substring", 2)
    def _(this, a):
        s = _this_str(this)
        n = len(s)
        x = max(0, min(to_int(arg(a, 0)), n))
        y = n if arg(a, 1) is UNDEFINED else max(0, min(to_int(a[1]), n))
        return s[min(x, y):max(x, y)]

    @method(SP, "substr", 2)
    def _(this, a):
        s = _this_str(this)
        n = len(s)
        start = _rel_index(arg(a, 0), n, 0)
        ln = n - start if arg(a, 1) is UNDEFINED else max(0, min(to_int(a[1]), n - start))
        return s[start:start + ln]

    @method(SP, "toLowerCase")
    def _(this, a):
        return _this_str(this).lower()

    @method(SP, "toUpperCase")
    def _(this, a):
        return _this_str(this).upper()

    SP.define("toLocaleLowerCase", SP.props["toLowerCase"])
    SP.define("toLocaleUpperCase", SP.props["toUpperCase"])
    from ..jsvalue import _WS as ws

    @method(SP, "trim")
    def _(this, a):
        return _this_str(this).strip(ws)

    @method(SP, "trimStart")
    def _(this, a):
        return _this_str(this).lstrip(ws)

    @method(SP, "trimEnd")
    def _(this, a):
        return _this_str(this).rstrip(ws)

    def _pad(this, a, left):
        s = _this_str(this)
        n = to_int(arg(a, 0))
        fill = " " if arg(a, 1) is UNDEFINED else to_str(a[1])
        if n <= len(s) or not fill:
            return s
        pad = (fill * ((n - len(s)) // len(fill) + 1))[:n - len(s)]
        return pad + s if left else s + pad

    SP.define("padStart", native("padStart", lambda t, a: _pad(t, a, True), 2))
    SP.define("padEnd", native("padEnd", lambda t, a: _pad(t, a, False), 2))

    @method(SP, "repeat", 1)
    def _(this, a):
        n = to_int(arg(a, 0))
        if n < 0:
            raise JSThrow(make_error("RangeError", "Invalid count value"))
        return _this_str(this) * n

    @method(SP, "concat", 1)
    def _(this, a):
        return _this_str(this) + "".join(to_str(x) for x in a)

    @method(SP, "localeCompare", 1)
    def _(this, a):
        s, o = _this_str(this), to_str(arg(a, 0))
        # ICU root collation approximated: case-insensitive first, lower before upper on ties
        ks, ko = (s.lower(), s.swapcase()), (o.lower(), o.swapcase())
        return -1.0 if ks < ko else (1.0 if ks > ko else 0.0)

    @method(SP, "split", 2)
    def _(this, a):
        s, sep = _this_str(this), arg(a, 0)
        limit = None if arg(a, 1) is UNDEFINED else I.to_uint32(a[1])
        rx = _regexp_of(sep)
        if sep is UNDEFINED:
            parts = [s]
        elif rx is not None:
            parts, pos = [], 0
            if s == "":
                parts = [] if rx.match("") else [""]
            else:
                for m in rx.finditer(s):
                    if m.end() == m.start() and (m.start() == 0 or m.start() >= len(s)):
                        continue
                    if m.end() == m.start() and m.start() == pos:
                        continue
                    parts.append(s[pos:m.start()])
                    parts.extend(UNDEFINED if x is None else x for x in m.groups())
                    pos = m.end()
                parts.append(s[pos:])
        else:
            sep = to_str(sep)
            parts = list(s) if sep == "" else s.split(sep)
        if limit is not None:
            parts = parts[:limit]
        return JSArray(parts)

    @method(SP, "match", 1)
    def _(this, a):
        s, r = _this_str(this), arg(a, 0)
        rx = _regexp_of(r)
        if rx is None:
            r = I.make_regexp(I.re.escape(to_str(r)) if r.__class__ is str else to_str(r), "")
            rx = r.props["%re"]
        if truthy(r.get("global")):
            ms = [m.group(0) for m in rx.finditer(s)]
            return JSArray(ms) if ms else None
        m = rx.search(s)
        return _match_array(m, s) if m else None

    @method(SP, "matchAll", 1)
    def _(this, a):
        s = _this_str(this)
        rx = _regexp_of(arg(a, 0))
        return py_iter_object(_match_array(m, s) for m in rx.finditer(s))

    @method(SP, "search", 1)
    def _(this, a):
        rx = _regexp_of(arg(a, 0))
        m = rx.search(_this_str(this)) if rx else None
        return float(m.start()) if m else -1.0

    def _replace(this, a, all_):
        s, pat, rep = _this_str(this), arg(a, 0), arg(a, 1)
        rx = _regexp_of(pat)

        def sub_for(m_text, groups, index, m=None):
            if isinstance(rep, JSFunction):
                return to_str(call(rep, UNDEFINED, [m_text] + [UNDEFINED if x is None else x for x in groups]
                                   + [float(index), s]))
            tpl = to_str(rep)
            if m is not None:
                return I._expand_replacement(tpl, m)
            return tpl.replace("$&", m_text).replace("$$", "$")
        if rx is None:
            pat = to_str(pat)
            if all_:
                if pat == "":
                    return s
                out, pos = [], 0
                while True:
                    i = s.find(pat, pos)
                    if i < 0:
                        break
                    out.append(s[pos:i])
                    out.append(sub_for(pat, (), i))
                    pos = i + len(pat)
                out.append(s[pos:])
                return "".join(out)
            i = s.find(pat)
            if i < 0:
                return s
            return s[:i] + sub_for(pat, (), i) + s[i + len(pat):]
        is_global = truthy(pat.get("global"))
        out, pos = [], 0
        for m in rx.finditer(s):
            out.append(s[pos:m.start()])
            out.append(sub_for(m.group(0), m.groups(), m.start(), m))
            pos = m.end()
            if not is_global:
                break
        out.append(s[pos:])
        return "".join(out)

    SP.define("replace", native("replace", lambda t, a: _replace(t, a, False), 2))
    SP.define("replaceAll", native("replaceAll", lambda t, a: _replace(t, a, True), 2))
    SP.define(I.SYM_ITERATOR, native("[Symbol.iterator]", lambda t, a: py_iter_object(iter(_this_str(t)))))

    def string_ctor(this, a):
        if not a:
            return ""
        v = a[0]
        if v.__class__ is JSSymbol:
            return f"Symbol({v.desc})"
        return to_str(v)
    S = native("String", string_ctor, 1)
    S.define("%construct", lambda a, nt: string_ctor(None, a))  # boxed strings are not modelled
    S.define("prototype", SP)
    SP.define("constructor", S)
    S.define("fromCharCode", native("fromCharCode", lambda t, a: "".join(chr(I.to_uint32(x) & 0xFFFF) for x in a), 1))
    g["String"] = S

    # Number
    NP = I.NUMBER_PROTO

    def _this_num(this):
        if this.__class__ is float:
            return this
        raise throw_type("Number.prototype method called on a non-number")

    @method(NP, "toString", 1)
    def _(this, a):
        x = _this_num(this)
        radix = 10 if arg(a, 0) is UNDEFINED else to_int(a[0])
        if radix == 10:
            return number_to_string(x)
        if x != x:
            return "NaN"
        if x in (math.inf, -math.inf):
            return "Infinity" if x > 0 else "-Infinity"
        digits = "0123456789abcdefghijklmnopqrstuvwxyz"
        neg, x = x < 0, abs(x)
        ip, fp = int(x), x - int(x)
        s = ""
        while True:
            s = digits[ip % radix] + s
            ip //= radix
            if ip == 0:
                break
        if fp > 0:
            s += "."
            for _ in range(52):
                fp *= radix
                d = int(fp)
                s += digits[d]
                fp -= d
                if fp == 0:
                    break
        return ("-" if neg else "") + s

    @method(NP, "toFixed", 1)
    def _(this, a):
        x = _this_num(this)
        d = to_int(arg(a, 0))
        if x != x:
            return "NaN"
        if abs(x) >= 1e21:
            return number_to_string(x)
        from decimal import Decimal, ROUND_HALF_UP
        q = Decimal(x).quantize(Decimal(1).scaleb(-d), rounding=ROUND_HALF_UP)
        s = format(q, "f")
        if q == 0 and x < 0:  # (-0.0001).toFixed(2) is "-0.00"; (-0).toFixed is "0"
            s = "-" + s.lstrip("-") if x != 0 else s.lstrip("-")
        return s

    @method(NP, "toPrecision", 1)
    def _(this, a):
        x = _this_num(this)
        if arg(a, 0) is UNDEFINED:
            return number_to_string(x)
        return format(x, f".{to_int(a[0])}g")

    @method(NP, "valueOf")
    def _(this, a):
        return _this_num(this)

    @method(NP, "toLocaleString")
    def _(this, a):
        x = _this_num(this)
        return f"{x:,.3f}".rstrip("0").rstrip(".") if x == x and abs(x) != math.inf else number_to_string(x)

    def number_ctor(this, a):
        return to_num(a[0]) if a else 0.0
    N = native("Number", number_ctor, 1)
    N.define("%construct", lambda a, nt: number_ctor(None, a))
    N.define("prototype", NP)
    NP.define("constructor", N)
    for k, v in [("MAX_SAFE_INTEGER", 9007199254740991.0), ("MIN_SAFE_INTEGER", -9007199254740991.0),
                 ("MAX_VALUE", 1.7976931348623157e308), ("MIN_VALUE", 5e-324), ("EPSILON", 2.220446049250313e-16),
                 ("POSITIVE_INFINITY", math.inf), ("NEGATIVE_INFINITY", -math.inf), ("NaN", math.nan)]:
        N.define(k, v)
    N.define("isInteger", native("isInteger", lambda t, a: arg(a, 0).__class__ is float and math.isfinite(a[0])
                                  and a[0] == math.floor(a[0]), 1))
    N.define("isSafeInteger", native("isSafeInteger", lambda t, a: arg(a, 0).__class__ is float and math.isfinite(a[0])
                                      and a[0] == math.floor(a[0]) and abs(a[0]) <= 9007199254740991.0, 1))
    N.define("isFinite", native("isFinite", lambda t, a: arg(a, 0).__class__ is float and math.isfinite(a[0]), 1))
    N.define("isNaN", native("isNaN", lambda t, a: arg(a, 0).__class__ is float and a[0] != a[0], 1))
    g["Number"] = N

    def parse_float(this, a):
        s = to_str(arg(a, 0)).lstrip(ws)
        m = I.re.match(r"[+-]?(Infinity|\d+\.?\d*([eE][+-]?\d+)?|\.\d+([eE][+-]?\d+)?)", s)
        if not m:
            return math.nan
        t = m.group(0)
        if "Infinity" in t:
            return -math.inf if t.startswith("-") else math.inf
        return float(t)

    def parse_int(this, a):
        s = to_str(arg(a, 0)).strip(ws)
        radix = 0 if arg(a, 1) is UNDEFINED else I.to_int32(a[1])
        sign = 1
        if s[:1] in ("+", "-"):
            sign = -1 if s[0] == "-" else 1
            s = s[1:]
        if radix == 0:
            radix = 10
            if s[:2].lower() == "0x":
                radix, s = 16, s[2:]
        elif radix == 16 and s[:2].lower() == "0x":
            s = s[2:]
        if not 2 <= radix <= 36:
            return math.nan
        digits = "0123456789abcdefghijklmnopqrstuvwxyz"[:radix]
        i = 0
        while i < len(s) and s[i].lower() in digits:
            i += 1
        if i == 0:
            return math.nan
        return float(sign * int(s[:i], radix))

    N.define("parseFloat", native("parseFloat", parse_float, 1))
    N.define("parseInt", native("parseInt", parse_int, 2))
    g["parseFloat"] = N.props["parseFloat"]
    g["parseInt"] = N.props["parseInt"]
    g["isNaN"] = native("isNaN", lambda t, a: (lambda x: x != x)(to_num(arg(a, 0))), 1)
    g["isFinite"] = native("isFinite", lambda t, a: math.isfinite(to_num(arg(a, 0))), 1)
    g["NaN"] = math.nan
    g["Infinity"] = math.inf

    # Boolean
    BP = I.BOOLEAN_PROTO
    BP.define("toString", native("toString", lambda t, a: "true" if t is True else "false"))
    BP.define("valueOf", native("valueOf", lambda t, a: t))
    B = native("Boolean", lambda t, a: truthy(arg(a, 0)), 1)
    B.define("%construct", lambda a, nt: truthy(arg(a, 0)))
    B.define("prototype", BP)
    BP.define("constructor", B)
    g["Boolean"] = B

    # Symbol
    registry = {}
    Sy = native("Symbol", lambda t, a: JSSymbol(UNDEFINED if arg(a, 0) is UNDEFINED else to_str(a[0])))
    Sy.define("iterator", I.SYM_ITERATOR)
    Sy.define("asyncIterator", JSSymbol("Symbol.asyncIterator"))
    Sy.define("for", native("for", lambda t, a: registry.setdefault(to_str(arg(a, 0)), JSSymbol(to_str(arg(a, 0))))))
    Sy.define("prototype", I.SYMBOL_PROTO)
    I.SYMBOL_PROTO.define("toString", native("toString", lambda t, a: f"Symbol({'' if t.desc is UNDEFINED else t.desc})"))
    g["Symbol"] = Sy


# ----------------------------------------------------------------------------- Math / JSON / Date / console
def _js_round(x):
    if x != x or x in (math.inf, -math.inf):
        return x
    r = math.floor(x + 0.5)
    if r == 0 and (x < 0 or math.copysign(1, x) < 0):
        return -0.0
    return float(r)


def _setup_misc(g, rt):
    M = new_object()
    for k, v in [("PI", math.pi), ("E", math.e), ("LN2", math.log(2)), ("LN10", math.log(10)), ("SQRT2", math.sqrt(2)),
                 ("LOG2E", 1 / math.log(2)), ("LOG10E", 1 / math.log(10)), ("SQRT1_2", math.sqrt(0.5))]:
        M.define(k, v)

    def num1(fn):
        def w(this, a):
            x = to_num(arg(a, 0))
            try:
                return float(fn(x))
            except (ValueError, OverflowError):
                return math.nan
        return w

    def _floor(x):
        return x if (x != x or abs(x) == math.inf) else float(math.floor(x)) if x != 0 else x

    def _ceil(x):
        if x != x or abs(x) == math.inf or x == 0:
            return x
        r = float(math.ceil(x))
        return -0.0 if r == 0 and x < 0 else r

    def _trunc(x):
        if x != x or abs(x) == math.inf or x == 0:
            return x
        r = float(math.trunc(x))
        return -0.0 if r == 0 and x < 0 else r

    def _minmax(is_max):
        def w(this, a):
            if not a:
                return -math.inf if is_max else math.inf
            xs = [to_num(x) for x in a]
            if any(x != x for x in xs):
                return math.nan
            r = xs[0]
            for x in xs[1:]:
                if is_max:
                    if x > r or (x == r == 0 and math.copysign(1, x) > 0):
                        r = x
                else:
                    if x < r or (x == r == 0 and math.copysign(1, x) < 0):
                        r = x
            return r
        return w

    for name, fn in [("floor", num1(_floor)), ("ceil", num1(_ceil)), ("round", num1(_js_round)),
                     ("trunc", num1(_trunc)), ("abs", num1(abs)),
                     ("sqrt", num1(lambda x: math.sqrt(x) if x >= 0 else math.nan)),
                     ("cbrt", num1(lambda x: math.copysign(abs(x) ** (1 / 3), x))),
                     ("log", num1(lambda x: -math.inf if x == 0 else math.log(x))),
                     ("log2", num1(lambda x: -math.inf if x == 0 else math.log2(x))),
                     ("log10", num1(lambda x: -math.inf if x == 0 else math.log10(x))),
                     ("exp", num1(math.exp)), ("sin", num1(math.sin)), ("cos", num1(math.cos)), ("tan", num1(math.tan)),
                     ("atan", num1(math.atan)), ("asin", num1(math.asin)), ("acos", num1(math.acos)),
                     ("sign", num1(lambda x: x if (x != x or x == 0) else math.copysign(1.0, x))),
                     ("max", _minmax(True)), ("min", _minmax(False)),
                     ("pow", lambda t, a: I._pow(arg(a, 0), arg(a, 1))),
                     ("atan2", lambda t, a: math.atan2(to_num(arg(a, 0)), to_num(arg(a, 1)))),
                     ("hypot", lambda t, a: math.hypot(*[to_num(x) for x in a])),
                     ("random", lambda t, a: rt.random())]:
        M.define(name, native(name, fn, 1))
    g["Math"] = M

    J = new_object()
    J.define("stringify", native("stringify", lambda t, a: I.json_stringify(arg(a, 0), arg(a, 1), arg(a, 2)), 3))
    J.define("parse", native("parse", lambda t, a: I.json_parse(to_str(arg(a, 0)), arg(a, 1)), 2))
    g["JSON"] = J

    # Date: epoch milliseconds in "%t"; only UTC rendering
    import datetime
    DP = I.DATE_PROTO

    def _t(this):
        if not (isinstance(this, JSObject) and this.cls == "Date"):
            raise throw_type("this is not a Date object.")
        return this.props["%t"]

    def _dt(this):
        return datetime.datetime(1970, 1, 1, tzinfo=datetime.timezone.utc) + datetime.timedelta(milliseconds=_t(this))

    def _iso(this, a):
        t = _t(this)
        if t != t:
            raise JSThrow(make_error("RangeError", "Invalid time value"))
        d = _dt(this)
        return d.strftime("%Y-%m-%dT%H:%M:%S.") + f"{int(t) % 1000:03d}Z"

    DP.define("getTime", native("getTime", lambda t, a: _t(t)))
    DP.define("valueOf", native("valueOf", lambda t, a: _t(t)))
    DP.define("toISOString", native("toISOString", _iso))
    DP.define("toJSON", native("toJSON", lambda t, a: _iso(t, a) if _t(t) == _t(t) else None))
    DP.define("toString", native("toString", lambda t, a: _dt(t).strftime("%a %b %d %Y %H:%M:%S GMT+0000 (Coordinated Universal Time)") if _t(t) == _t(t) else "Invalid Date"))
    DP.define("toLocaleString", native("toLocaleString", lambda t, a: _dt(t).strftime("%m/%d/%Y, %I:%M:%S %p")))
    DP.define("toLocaleDateString", native("toLocaleDateString", lambda t, a: _dt(t).strftime("%m/%d/%Y")))
    DP.define("toLocaleTimeString", native("toLocaleTimeString", lambda t, a: _dt(t).strftime("%I:%M:%S %p")))
    for nm, fn in [("getFullYear", lambda d: d.year), ("getMonth", lambda d: d.month - 1), ("getDate", lambda d: d.day),
                   ("getDay", lambda d: (d.weekday() + 1) % 7), ("getHours", lambda d: d.hour),
                   ("getMinutes", lambda d: d.minute), ("getSeconds", lambda d: d.second),
                   ("getMilliseconds", lambda d: d.microsecond // 1000), ("getTimezoneOffset", lambda d: 0)]:
        DP.define(nm, native(nm, (lambda fn: lambda t, a: float(fn(_dt(t))) if _t(t) == _t(t) else math.nan)(fn)))
        if nm.startswith("get") and nm != "getTimezoneOffset":
            DP.define("getUTC" + nm[3:], DP.props[nm])

    def _parse_date(s):
        m = I.re.match(r"^(\d{4})-(\d{2})-(\d{2})(?:[T ](\d{2}):(\d{2})(?::(\d{2})(?:\.(\d{1,3}))?)?(Z|[+-]\d{2}:?\d{2})?)?$", s)
        if not m:
            return math.nan
        y, mo, d, h, mi, sec, ms, tz = m.groups()
        try:
            dt = datetime.datetime(int(y), int(mo), int(d), int(h or 0), int(mi or 0), int(sec or 0),
                                   tzinfo=datetime.timezone.utc)
        except ValueError:
            return math.nan
        t = (dt - datetime.datetime(1970, 1, 1, tzinfo=datetime.timezone.utc)).total_seconds() * 1000
        t += int((ms or "0").ljust(3, "0"))
        if tz and tz != "Z":
            sign = 1 if tz[0] == "+" else -1
            hh, mm = int(tz[1:3]), int(tz[-2:])
            t -= sign * (hh * 60 + mm) * 60000
        return float(t)

    def date_construct(a, nt):
        d = JSObject(nt.get("prototype") if isinstance(nt.get("prototype"), JSObject) else DP, "Date")
        if not a:
            t = rt.now()
        elif len(a) == 1:
            v = a[0]
            if isinstance(v, JSObject) and v.cls == "Date":
                t = v.props["%t"]
            elif v.__class__ is str:
                t = _parse_date(v)
            else:
                t = to_num(v)
        else:
            parts = [to_int(x) for x in a] + [0] * 7
            y, mo = parts[0], parts[1]
            day = parts[2] if len(a) > 2 else 1
            dt = datetime.datetime(y + mo // 12, mo % 12 + 1, 1, tzinfo=datetime.timezone.utc) + datetime.timedelta(
                days=day - 1, hours=parts[3], minutes=parts[4], seconds=parts[5], milliseconds=parts[6])
            t = (dt - datetime.datetime(1970, 1, 1, tzinfo=datetime.timezone.utc)).total_seconds() * 1000
        d.define("%t", float(t))
        return d

    D = native("Date", lambda t, a: "Thu Jan 01 1970 00:00:00 GMT+0000 (Coordinated Universal Time)", 7)
    D.define("%construct", date_construct)
    D.define("prototype", DP)
    DP.define("constructor", D)
    D.define("now", native("now", lambda t, a: rt.now()))
    D.define("parse", native("parse", lambda t, a: _parse_date(to_str(arg(a, 0))), 1))
    g["Date"] = D

    con = new_object()

    def log(kind):
        def w(this, a):
            if rt.console is not None:
                rt.console.append((kind, " ".join(v if v.__class__ is str else rt.inspect(v) for v in a)))
            return UNDEFINED
        return w
    for k in ("log", "warn", "error", "info", "debug", "trace", "table", "group", "groupEnd", "time", "timeEnd"):
        con.define(k, native(k, log(k)))
    g["console"] = con


# ----------------------------------------------------------------------------- Error types
def _setup_errors(g):
    def mk(name, parent_proto):
        proto = JSObject(parent_proto, "Error")
        proto.define("name", name)
        proto.define("message", "")
        I.ERROR_PROTOS[name] = proto

        def ctor(a, nt):
            p = nt.get("prototype") if isinstance(nt, JSObject) else proto
            e = JSObject(p if isinstance(p, JSObject) else proto, "Error")
            if arg(a, 0) is not UNDEFINED:
                e.define("message", to_str(a[0]))
            opts = arg(a, 1)
            if isinstance(opts, JSObject) and opts.lookup("cause") is not I._MISSING:
                e.define("cause", opts.get("cause"))
            e.define("stack", f"{name}: {to_str(arg(a, 0)) if a else ''}\n    at <minijs>")
            return e
        f = native(name, lambda t, a: ctor(a, f), 1)
        f.define("%construct", ctor)
        f.define("prototype", proto)
        proto.define("constructor", f)
        f.define("captureStackTrace", native("captureStackTrace", lambda t, a: UNDEFINED))
        g[name] = f
        return proto
    base = mk("Error", I.OBJECT_PROTO)

    @method(base, "toString")
    def _(this, a):
        n, m = to_str(this.get("name")), to_str(this.get("message"))
        return n if not m else (m if not n else f"{n}: {m}")
    for n in ("TypeError", "RangeError", "ReferenceError", "SyntaxError", "EvalError", "URIError"):
        mk(n, base)


# ----------------------------------------------------------------------------- Map / Set / WeakMap / WeakSet
def _setup_collections(g):
    MP, SP = I.MAP_PROTO, I.SET_PROTO

    def _m(this, cls):
        if not (this.__class__ is JSMapObj and this.cls == cls):
            raise throw_type(f"Method {cls}.prototype called on incompatible receiver")
        return this.data

    @method(MP, "get", 1)
    def _(this, a):
        e = _m(this, "Map").get(map_key(arg(a, 0)))
        return e[1] if e is not None else UNDEFINED

    @method(MP, "set", 2)
    def _(this, a):
        k = arg(a, 0)
        if k.__class__ is float and k == 0:
            k = 0.0
        nk = map_key(k)
        old = _m(this, "Map").get(nk)
        this.insert(nk, (old[0] if old is not None else k, arg(a, 1)))
        return this

    @method(MP, "has", 1)
    def _(this, a):
        return map_key(arg(a, 0)) in _m(this, "Map")

    @method(MP, "delete", 1)
    def _(this, a):
        _m(this, "Map")
        return this.remove(map_key(arg(a, 0)))

    @method(MP, "clear")
    def _(this, a):
        _m(this, "Map")
        this.clear()
        return UNDEFINED

    @method(MP, "forEach", 1)
    def _(this, a):
        f, t = arg(a, 0), arg(a, 1)
        for nk in I._live_keys(this):
            k, v = this.data[nk]
            call(f, t, [v, k, this])
        return UNDEFINED

    MP.define("keys", native("keys", lambda t, a: py_iter_object(t.data[nk][0] for nk in I._live_keys(t))))
    MP.define("values", native("values", lambda t, a: py_iter_object(t.data[nk][1] for nk in I._live_keys(t))))
    MP.define("entries", native("entries", lambda t, a: py_iter_object(iterate(t))))
    MP.define(I.SYM_ITERATOR, MP.props["entries"])
    MP.define("size", Accessor(native("size", lambda t, a: float(len(_m(t, "Map"))))))

    @method(SP, "add", 1)
    def _(this, a):
        k = arg(a, 0)
        if k.__class__ is float and k == 0:
            k = 0.0
        d = _m(this, "Set")
        nk = map_key(k)
        if nk not in d:
            this.insert(nk, (k, k))
        return this

    @method(SP, "has", 1)
    def _(this, a):
        return map_key(arg(a, 0)) in _m(this, "Set")

    @method(SP, "delete", 1)
    def _(this, a):
        _m(this, "Set")
        return this.remove(map_key(arg(a, 0)))

    @method(SP, "clear")
    def _(this, a):
        _m(this, "Set")
        this.clear()
        return UNDEFINED

    @method(SP, "forEach", 1)
    def _(this, a):
        f, t = arg(a, 0), arg(a, 1)
        for nk in I._live_keys(this):
            k = this.data[nk][0]
            call(f, t, [k, k, this])
        return UNDEFINED

    SP.define("values", native("values", lambda t, a: py_iter_object(iterate(t))))
    SP.define("keys", SP.props["values"])
    SP.define("entries", native("entries", lambda t, a: py_iter_object(JSArray([v, v]) for v in iterate(t))))
    SP.define(I.SYM_ITERATOR, SP.props["values"])
    SP.define("size", Accessor(native("size", lambda t, a: float(len(_m(t, "Set"))))))

    def mk(name, proto, cls, adder):
        def ctor(a, nt):
            p = nt.get("prototype") if isinstance(nt, JSObject) else proto
            o = JSMapObj(p if isinstance(p, JSObject) else proto, cls)
            src = arg(a, 0)
            if src is not UNDEFINED and src is not None:
                add = proto.props[adder]
                for item in iterate(src):
                    if cls == "Map":
                        call(add, o, [I.get_member(item, "0"), I.get_member(item, "1")])
                    else:
                        call(add, o, [item])
            return o
        f = native(name, lambda t, a: (_ for _ in ()).throw(throw_type(f"Constructor {name} requires 'new'")))
        f.define("%construct", ctor)
        f.define("prototype", proto)
        proto.define("constructor", f)
        g[name] = f
    mk("Map", MP, "Map", "set")
    mk("Set", SP, "Set", "add")
    # WeakMap / WeakSet: identity-keyed, same behaviour minus iteration
    mk("WeakMap", MP, "Map", "set")
    mk("WeakSet", SP, "Set", "add")
    MP.define("constructor", g["Map"])
    SP.define("constructor", g["Set"])

    IP = I.ITER_PROTO

    @method(IP, "next")
    def _(this, a):
        r = new_object()
        try:
            v = next(this.props["%it"])
            r.put_own("value", v)
            r.put_own("done", False)
        except StopIteration:
            r.put_own("value", UNDEFINED)
            r.put_own("done", True)
        return r
    IP.define(I.SYM_ITERATOR, native("[Symbol.iterator]", lambda t, a: t))


# ----------------------------------------------------------------------------- ArrayBuffer / typed arrays
def _setup_typed_arrays(g):
    BP, TP = I.BUFFER_PROTO, I.TYPED_PROTO

    def buffer_ctor(a, nt):
        b = JSObject(BP, "ArrayBuffer")
        b.define("%bytes", bytearray(to_int(arg(a, 0))))
        return b
    B = native("ArrayBuffer", lambda t, a: (_ for _ in ()).throw(throw_type("Constructor ArrayBuffer requires 'new'")), 1)
    B.define("%construct", buffer_ctor)
    B.define("prototype", BP)
    BP.define("constructor", B)
    BP.define("byteLength", Accessor(native("byteLength", lambda t, a: float(len(t.props["%bytes"])))))
    g["ArrayBuffer"] = B

    def _ta(this):
        if this.__class__ is not JSTypedArray:
            raise throw_type("this is not a typed array")
        return this

    @method(TP, "set", 2)
    def _(this, a):
        t, off = _ta(this), to_int(arg(a, 1)) if len(a) > 1 else 0
        for i, v in enumerate(iterate(arg(a, 0))):
            t.put_own(str(off + i), v)
        return UNDEFINED

    @method(TP, "fill", 1)
    def _(this, a):
        t = _ta(this)
        for i in range(_rel_index(arg(a, 1), t.n, 0), _rel_index(arg(a, 2), t.n, t.n)):
            t.put_own(str(i), arg(a, 0))
        return this

    @method(TP, "subarray", 2)
    def _(this, a):
        t = _ta(this)
        lo, hi = _rel_index(arg(a, 0), t.n, 0), _rel_index(arg(a, 1), t.n, t.n)
        out = JSTypedArray(t.proto, t.cls, t.buf, t.fmt, t.size, t.off + lo * t.size, max(hi - lo, 0))
        out.props["%buffer"] = t.props.get("%buffer", UNDEFINED)
        return out

    @method(TP, "slice", 2)
    def _(this, a):
        t = _ta(this)
        lo, hi = _rel_index(arg(a, 0), t.n, 0), _rel_index(arg(a, 1), t.n, t.n)
        n = max(hi - lo, 0)
        buf = bytearray(t.buf[t.off + lo * t.size: t.off + (lo + n) * t.size])
        return JSTypedArray(t.proto, t.cls, buf, t.fmt, t.size, 0, n)

    TP.define("forEach", native("forEach", lambda t, a: [call(arg(a, 0), UNDEFINED, [v, float(i), t])
                                                        for i, v in enumerate(iterate(_ta(t)))] and UNDEFINED, 1))
    TP.define("join", native("join", lambda t, a: ("," if arg(a, 0) is UNDEFINED else to_str(a[0])).join(
        to_str(v) for v in iterate(_ta(t))), 1))
    TP.define(I.SYM_ITERATOR, native("[Symbol.iterator]", lambda t, a: py_iter_object(iterate(_ta(t)))))

    def mk(name, fmt, size):
        proto = JSObject(TP, name)

        def ctor(a, nt):
            x = arg(a, 0)
            if isinstance(x, JSObject) and x.cls == "ArrayBuffer":
                raw = x.props["%bytes"]
                off = to_int(arg(a, 1)) if len(a) > 1 and a[1] is not UNDEFINED else 0
                n = to_int(a[2]) if len(a) > 2 and a[2] is not UNDEFINED else (len(raw) - off) // size
                if off % size or off + n * size > len(raw):
                    raise JSThrow(make_error("RangeError", f"start offset of {name} should be a multiple of {size}"))
                t = JSTypedArray(proto, name, raw, fmt, size, off, n)
                t.props["%buffer"] = x
                return t
            if x.__class__ is float or x is UNDEFINED:
                n = 0 if x is UNDEFINED else to_int(x)
                vals = None
            else:
                vals = list(iterate(x))
                n = len(vals)
            b = buffer_ctor([float(n * size)], None)
            t = JSTypedArray(proto, name, b.props["%bytes"], fmt, size, 0, n)
            t.props["%buffer"] = b
            if vals:
                for i, v in enumerate(vals):
                    t.put_own(str(i), v)
            return t
        f = native(name, lambda t, a: (_ for _ in ()).throw(throw_type(f"Constructor {name} requires 'new'")), 3)
        f.define("%construct", ctor)
        f.define("prototype", proto)
        f.define("BYTES_PER_ELEMENT", float(size))
        proto.define("constructor", f)
        g[name] = f
    for name, fmt, size in (("Uint8Array", "<B", 1), ("Int8Array", "<b", 1), ("Uint16Array", "<H", 2), ("Int16Array", "<h", 2),
                            ("Uint32Array", "<I", 4), ("Int32Array", "<i", 4), ("Float32Array", "<f", 4), ("Float64Array", "<d", 8)):
        mk(name, fmt, size)


# ----------------------------------------------------------------------------- RegExp / Promise
def _setup_regexp_promise(g, rt):
    RP = I.REGEXP_PROTO

    def _exec(this, a):
        s = to_str(arg(a, 0))
        rx = this.props["%re"]
        glob = truthy(this.props.get("global")) or "y" in this.props.get("flags", "")
        start = to_int(this.props.get("lastIndex", 0.0)) if glob else 0
        m = rx.search(s, start) if start <= len(s) else None
        if m is None:
            if glob:
                this.props["lastIndex"] = 0.0
            return None
        if glob:
            this.props["lastIndex"] = float(m.end())
        return I_match(m, s)

    I_match = _match_array
    RP.define("exec", native("exec", _exec, 1))
    RP.define("test", native("test", lambda t, a: _exec(t, a) is not None, 1))
    RP.define("toString", native("toString", lambda t, a: f"/{t.props['source']}/{t.props['flags']}"))

    def ctor(a, nt):
        p, fl = arg(a, 0), arg(a, 1)
        if isinstance(p, JSObject) and p.cls == "RegExp":
            return I.make_regexp(p.props["source"], p.props["flags"] if fl is UNDEFINED else to_str(fl))
        return I.make_regexp(to_str(p), "" if fl is UNDEFINED else to_str(fl))
    R = native("RegExp", lambda t, a: ctor(a, None), 2)
    R.define("%construct", ctor)
    R.define("prototype", RP)
    RP.define("constructor", R)
    g["RegExp"] = R

    PP = I.PROMISE_PROTO
    PP.define("then", native("then", lambda t, a: I.promise_then(t, arg(a, 0), arg(a, 1)), 2))
    PP.define("catch", native("catch", lambda t, a: I.promise_then(t, UNDEFINED, arg(a, 0)), 1))

    def _finally(this, a):
        f = arg(a, 0)

        def ok(t, args):
            call(f, UNDEFINED, [])
            return arg(args, 0)

        def err(t, args):
            call(f, UNDEFINED, [])
            raise JSThrow(arg(args, 0))
        return I.promise_then(this, native("", ok), native("", err))
    PP.define("finally", native("finally", _finally, 1))

    def pctor(a, nt):
        ex = arg(a, 0)
        if not isinstance(ex, JSFunction):
            raise throw_type("Promise resolver is not a function")
        p = I.new_promise()
        res = native("resolve", lambda t, x: I.resolve_promise(p, arg(x, 0)) or UNDEFINED, 1)
        rej = native("reject", lambda t, x: I.reject_promise(p, arg(x, 0)) or UNDEFINED, 1)
        try:
            call(ex, UNDEFINED, [res, rej])
        except JSThrow as e:
            I.reject_promise(p, e.value)
        return p
    P = native("Promise", lambda t, a: (_ for _ in ()).throw(throw_type("Promise constructor requires 'new'")), 1)
    P.define("%construct", pctor)
    P.define("prototype", PP)
    PP.define("constructor", P)

    def p_resolve(this, a):
        v = arg(a, 0)
        if isinstance(v, JSObject) and v.cls == "Promise":
            return v
        p = I.new_promise()
        I.resolve_promise(p, v)
        return p

    def p_reject(this, a):
        p = I.new_promise()
        I.reject_promise(p, arg(a, 0))
        return p

    def p_all(settled):
        def w(this, a):
            items = list(iterate(arg(a, 0)))
            out = I.new_promise()
            results = [UNDEFINED] * len(items)
            left = [len(items)]
            if not items:
                I.resolve_promise(out, JSArray([]))
                return out

            def on(i):
                def cb(state, value):
                    if settled:
                        o = new_object()
                        o.put_own("status", state)
                        o.put_own("value" if state == "fulfilled" else "reason", value)
                        results[i] = o
                    elif state == "rejected":
                        I.reject_promise(out, value)
                        return
                    else:
                        results[i] = value
                    left[0] -= 1
                    if left[0] == 0:
                        I.resolve_promise(out, JSArray(results))
                return cb
            for i, it in enumerate(items):
                I._subscribe(p_resolve(None, [it]), on(i))
            return out
        return w

    def p_race(this, a):
        out = I.new_promise()
        for it in iterate(arg(a, 0)):
            I._subscribe(p_resolve(None, [it]), lambda st, v: I._settle(out, st, v))
        return out
    P.define("resolve", native("resolve", p_resolve, 1))
    P.define("reject", native("reject", p_reject, 1))
    P.define("all", native("all", p_all(False), 1))
    P.define("allSettled", native("allSettled", p_all(True), 1))
    P.define("race", native("race", p_race, 1))
    g["Promise"] = P


# ----------------------------------------------------------------------------- runtime: globals, timers, modules
class Runtime:
    """One JS realm.  `require(path)` runs CommonJS modules from disk; Node built-ins are stubs."""

    def __init__(self, console=None, start_ms=1_700_000_000_000.0, seed=0x9E3779B9, files=None):
        self.console = console  # list collecting (kind, text), or None to drop output
        self.files = files if files is not None else {}  # the in-memory file system behind require("fs")
        self.dirs = set()
        self._now = float(start_ms)
        self._seed = seed & 0xFFFFFFFF
        self.microtasks = []
        self.timers = []  # [due_ms, seq, fn, args, interval_ms | None, id]
        self._timer_seq = 0
        self.modules = {}
        self.node_stubs = {}
        I.Interp.current = self
        g = self.globals = Env(None, fn=True)
        v = g.vars
        _reset_protos()
        _setup_object(v)
        _setup_function(v)
        _setup_array(v)
        _setup_string(v)
        _setup_misc(v, self)
        _setup_errors(v)
        _setup_collections(v)
        _setup_regexp_promise(v, self)
        _setup_typed_arrays(v)
        self._setup_timers(v)
        self._setup_node(v)
        gt = new_object()
        v["globalThis"] = v["global"] = gt
        v["undefined"] = UNDEFINED
        v["this"] = UNDEFINED

    # -- time / randomness
    def now(self):
        self._now += 1.0
        return self._now

    def random(self):
        self._seed = (self._seed * 1664525 + 1013904223) & 0xFFFFFFFF
        return self._seed / 4294967296.0

    def inspect(self, v, depth=0):
        if v.__class__ is str:
            return v if depth == 0 else "'" + v + "'"
        if v.__class__ is JSArray:
            return "[ " + ", ".join(self.inspect(x, depth + 1) for x in v.items) + " ]" if v.items else "[]"
        if isinstance(v, JSFunction):
            return f"[Function: {v.name}]"
        if isinstance(v, JSObject):
            if v.cls == "Error":
                return f"{to_str(v.get('name'))}: {to_str(v.get('message'))}"
            if depth > 2:
                return "[Object]"
            ks = [k for k in v.enumerable_keys() if k.__class__ is str]
            return "{ " + ", ".join(f"{k}: {self.inspect(v.get(k), depth + 1)}" for k in ks) + " }" if ks else "{}"
        return to_str(v) if v.__class__ is not JSSymbol else repr(v)

    # -- queues
    def run_microtasks(self) -> bool:
        ran = False
        while self.microtasks:
            self.microtasks.pop(0)()
            ran = True
        return ran

    def run_next_timer(self) -> bool:
        if not self.timers:
            return False
        self.timers.sort(key=lambda t: (t[0], t[1]))
        due, _, fn, args, interval, tid = self.timers.pop(0)
        self._now = max(self._now, due)
        if interval is not None:
            self._timer_seq += 1
            self.timers.append([due + max(interval, 1.0), self._timer_seq, fn, args, interval, tid])
        call(fn, UNDEFINED, args)
        self.run_microtasks()
        return True

    def run_timers(self, advance_ms=0.0, limit=100000):
        """Advance the clock and fire every timer that is due by then (intervals re-arm)."""
        target = self._now + advance_ms
        n = 0
        while self.timers and n < limit:
            self.timers.sort(key=lambda t: (t[0], t[1]))
            if self.timers[0][0] > target:
                break
            self.run_next_timer()
            n += 1
        self._now = max(self._now, target)
        self.run_microtasks()
        return n

    def _setup_timers(self, v):
        def set_timer(repeat):
            def w(this, a):
                fn = arg(a, 0)
                delay = to_num(arg(a, 1)) if len(a) > 1 else 0.0
                delay = 0.0 if delay != delay else max(delay, 0.0)
                self._timer_seq += 1
                tid = float(self._timer_seq)
                self.timers.append([self._now + delay, self._timer_seq, fn, list(a[2:]), delay if repeat else None, tid])
                h = new_object()
                h.define("%id", tid)
                h.define("unref", native("unref", lambda t, x: t))
                h.define("ref", native("ref", lambda t, x: t))
                h.define("hasRef", native("hasRef", lambda t, x: True))
                return h
            return w

        def clear(this, a):
            h = arg(a, 0)
            tid = h.props.get("%id") if isinstance(h, JSObject) else h
            self.timers = [t for t in self.timers if t[5] != tid]
            return UNDEFINED
        v["setTimeout"] = native("setTimeout", set_timer(False), 2)
        v["setInterval"] = native("setInterval", set_timer(True), 2)
        v["setImmediate"] = native("setImmediate", lambda t, a: set_timer(False)(t, [arg(a, 0), 0.0] + list(a[1:])), 1)
        v["clearTimeout"] = v["clearInterval"] = v["clearImmediate"] = native("clearTimeout", clear, 1)
        v["queueMicrotask"] = native("queueMicrotask", lambda t, a: self.microtasks.append(
            lambda: call(arg(a, 0), UNDEFINED, [])) or UNDEFINED, 1)

    # -- Node
    def _setup_node(self, v):
        proc = new_object()
        proc.put_own("env", new_object())
        proc.put_own("argv", JSArray(["node", "script"]))
        proc.put_own("pid", 4242.0)
        proc.put_own("platform", "linux")
        proc.put_own("version", "v18.0.0-minijs")
        proc.put_own("versions", from_py({"node": "18.0.0"}))
        proc.define("on", native("on", lambda t, a: t))
        proc.define("once", native("once", lambda t, a: t))
        proc.define("exit", native("exit", lambda t, a: (_ for _ in ()).throw(JSThrow(make_error("Error", "process.exit")))))
        proc.define("cwd", native("cwd", lambda t, a: "/"))
        proc.define("nextTick", native("nextTick", lambda t, a: self.microtasks.append(
            lambda: call(arg(a, 0), UNDEFINED, list(a[1:]))) or UNDEFINED))
        proc.define("memoryUsage", native("memoryUsage", lambda t, a: from_py({"heapUsed": 0, "rss": 0})))
        proc.define("uptime", native("uptime", lambda t, a: 1.0))

        def hrtime(this, a):
            t = self.now()
            return JSArray([float(int(t // 1000)), float(int(t % 1000) * 1e6)])
        proc.define("hrtime", native("hrtime", hrtime))
        v["process"] = proc

    def stub_module(self, name, value):
        self.node_stubs[name] = value

    def _builtin_module(self, name):
        if name in self.node_stubs:
            return self.node_stubs[name]
        if name == "events":
            return self._events_module()
        if name == "path":
            p = new_object()
            p.define("join", native("join", lambda t, a: os.path.normpath("/".join(to_str(x) for x in a))))
            p.define("resolve", native("resolve", lambda t, a: os.path.normpath(os.path.join("/", *[to_str(x) for x in a]))))
            p.define("dirname", native("dirname", lambda t, a: os.path.dirname(to_str(arg(a, 0)))))
            p.define("basename", native("basename", lambda t, a: os.path.basename(to_str(arg(a, 0)))))
            p.define("extname", native("extname", lambda t, a: os.path.splitext(to_str(arg(a, 0)))[1]))
            p.put_own("sep", "/")
            return p
        if name == "crypto":
            c = new_object()

            def random_bytes(this, a):
                n = to_int(arg(a, 0))
                data = bytes(int(self.random() * 256) for _ in range(n))
                b = new_object()
                b.define("toString", native("toString", lambda t, x: data.hex()))
                b.put_own("length", float(n))
                return b
            c.define("randomBytes", native("randomBytes", random_bytes, 1))

            def random_uuid(this, a):
                h = "".join(f"{int(self.random() * 16):x}" for _ in range(32))
                return f"{h[:8]}-{h[8:12]}-4{h[13:16]}-a{h[17:20]}-{h[20:32]}"
            c.define("randomUUID", native("randomUUID", random_uuid))

            def create_hash(this, a):
                import hashlib
                hsh = hashlib.new(to_str(arg(a, 0)))
                o = new_object()
                o.define("update", native("update", lambda t, x: hsh.update(to_str(arg(x, 0)).encode()) or t, 1))
                o.define("digest", native("digest", lambda t, x: hsh.hexdigest(), 1))
                return o
            c.define("createHash", native("createHash", create_hash, 1))
            return c
        if name == "fs":
            return self._fs_module()
        if name in ("ws", "http", "https", "net", "os", "util", "zlib", "stream", "url", "child_process"):
            # present but inert: any use on the measured path would surface as "x is not a function"
            return new_object()
        raise JSThrow(make_error("Error", f"Cannot find module '{name}'"))

    def _fs_module(self):
        """An in-memory file system (`self.files`: normalised path -> str), enough for the reference's
        BulletFileStorage (existsSync / mkdirSync / readFileSync / writeFileSync / unlinkSync / readdirSync)."""
        files, dirs = self.files, self.dirs
        norm = lambda v: os.path.normpath(to_str(v))
        fs = new_object()

        def read(this, a):
            p = norm(arg(a, 0))
            if p not in files:
                e = make_error("Error", f"ENOENT: no such file or directory, open '{p}'")
                e.put_own("code", "ENOENT")
                raise JSThrow(e)
            return files[p]

        def write(this, a):
            files[norm(arg(a, 0))] = to_str(arg(a, 1))
            return UNDEFINED

        def exists(this, a):
            p = norm(arg(a, 0))
            return p in files or p in dirs or any(f.startswith(p + "/") for f in files)

        def mkdir(this, a):
            dirs.add(norm(arg(a, 0)))
            return UNDEFINED

        def unlink(this, a):
            files.pop(norm(arg(a, 0)), None)
            return UNDEFINED

        def readdir(this, a):
            p = norm(arg(a, 0))
            return JSArray(sorted({f[len(p) + 1:].split("/")[0] for f in files if f.startswith(p + "/")}))

        def rename(this, a):
            src, dst = norm(arg(a, 0)), norm(arg(a, 1))
            if src in files:
                files[dst] = files.pop(src)
            return UNDEFINED
        for name, fn in (("readFileSync", read), ("writeFileSync", write), ("existsSync", exists), ("mkdirSync", mkdir),
                         ("unlinkSync", unlink), ("readdirSync", readdir), ("renameSync", rename)):
            fs.define(name, native(name, fn, 2))
        return fs

    def _events_module(self):
        src = """
        class EventEmitter {
          constructor() { this._events = {}; }
          on(n, f) { (this._events[n] = this._events[n] || []).push(f); return this; }
          addListener(n, f) { return this.on(n, f); }
          once(n, f) { const w = (...a) => { this.off(n, w); f.apply(this, a); }; w.listener = f; return this.on(n, w); }
          off(n, f) { const l = this._events[n]; if (l) this._events[n] = l.filter((x) => x !== f && x.listener !== f); return this; }
          removeListener(n, f) { return this.off(n, f); }
          removeAllListeners(n) { if (n === undefined) this._events = {}; else delete this._events[n]; return this; }
          emit(n, ...a) { const l = this._events[n]; if (!l || !l.length) return false; for (const f of [...l]) f.apply(this, a); return true; }
          listenerCount(n) { return (this._events[n] || []).length; }
          listeners(n) { return [...(this._events[n] || [])]; }
          setMaxListeners() { return this; }
        }
        EventEmitter.EventEmitter = EventEmitter;
        module.exports = EventEmitter;
        """
        return self.run_module_source(src, "<events>", "/")

    def run_module_source(self, src, filename, dirname):
        module = new_object()
        exports = new_object()
        module.put_own("exports", exports)
        env = Env(self.globals, fn=True)
        env.vars.update({"module": module, "exports": exports, "__filename": filename, "__dirname": dirname,
                         "this": exports,
                         "require": native("require", lambda t, a: self.require(to_str(arg(a, 0)), dirname), 1)})
        prog = I.compile_program(src, filename)
        self.modules[filename] = module
        I.Interp.current = self
        prog(env)
        return module.get("exports")

    def require(self, name, base="/"):
        if name.startswith("node:"):
            name = name[5:]
        if not name.startswith((".", "/")):
            key = "builtin:" + name
            if key not in self.modules:
                m = new_object()
                m.put_own("exports", self._builtin_module(name))
                self.modules[key] = m
            return self.modules[key].get("exports")
        path = os.path.normpath(os.path.join(base, name))
        for cand in (path, path + ".js", os.path.join(path, "index.js")):
            if os.path.isfile(cand):
                path = cand
                break
        else:
            raise JSThrow(make_error("Error", f"Cannot find module '{name}' from {base}"))
        if path in self.modules:
            return self.modules[path].get("exports")
        if path.endswith(".json"):
            with open(path, encoding="utf-8") as f:
                m = new_object()
                m.put_own("exports", I.json_parse(f.read()))
                self.modules[path] = m
                return m.get("exports")
        with open(path, encoding="utf-8") as f:
            src = f.read()
        return self.run_module_source(src, path, os.path.dirname(path))

    def run(self, src, filename="<script>", dirname="/"):
        return self.run_module_source(src, filename, dirname)

    def eval(self, src, **bindings):
        """Evaluate statements with extra global bindings; the value of a trailing `return` is returned."""
        I.Interp.current = self
        env = Env(self.globals, fn=True)
        env.vars["this"] = UNDEFINED
        env.vars["require"] = native("require", lambda t, a: self.require(to_str(arg(a, 0)), "/"), 1)
        for k, v in bindings.items():
            env.vars[k] = v
        r = I.compile_program(src)(env)
        self.run_microtasks()
        if r is not None and r is not I.BREAK and r is not I.CONTINUE:
            return r[1]
        return UNDEFINED

    def call(self, fn, this=UNDEFINED, *args):
        I.Interp.current = self
        r = call(fn, this, list(args))
        self.run_microtasks()
        return r

    def method(self, obj, name, *args):
        I.Interp.current = self
        r = call(I.get_member(obj, name), obj, list(args))
        self.run_microtasks()
        return r

    def new(self, ctor, *args):
        I.Interp.current = self
        return construct(ctor, list(args))


def _reset_protos():
    """Every Runtime is a fresh realm: clear the shared prototype objects before refilling them."""
    for p in (I.OBJECT_PROTO, I.FUNCTION_PROTO, I.ARRAY_PROTO, I.STRING_PROTO, I.NUMBER_PROTO, I.BOOLEAN_PROTO,
              I.SYMBOL_PROTO, I.REGEXP_PROTO, I.PROMISE_PROTO, I.MAP_PROTO, I.SET_PROTO, I.DATE_PROTO, I.ITER_PROTO,
              I.TYPED_PROTO, I.BUFFER_PROTO):
        p.props.clear()
        p.nonenum = None
    I.ERROR_PROTOS.clear()
