"""A small ECMAScript parser: enough of ES2020 to read the reference's own hot-path sources
(src/bullet.js, bullet-crt.js, bullet-query.js, bullet-middleware.js) unmodified.

TEST INFRASTRUCTURE ONLY.  The AST is made of tuples ("kind", ...); see interp.py.
Automatic semicolon insertion is handled the pragmatic way: a statement may end at a `;`,
at a `}`, at end of input, or at a line break (the reference's sources are prettier-formatted,
so no statement continues on the next line after a complete expression except with a leading
`.`, `?`, `:`, operator or `(`-less continuation, all of which the expression parser consumes
before the statement terminator is looked for).
"""
from __future__ import annotations

import re

KEYWORDS = {
    "var", "let", "const", "function", "class", "return", "if", "else", "for", "while", "do", "break",
    "continue", "switch", "case", "default", "try", "catch", "finally", "throw", "new", "delete", "typeof",
    "void", "in", "of", "instanceof", "this", "null", "true", "false", "async", "await", "static", "extends",
    "super", "get", "set",
}
# contextual words that may also be identifiers / property names
SOFT = {"of", "async", "await", "static", "get", "set"}

PUNCT = [
    ">>>=", "...", "===", "!==", "**=", "<<=", ">>=", ">>>", "&&=", "||=", "??=",
    "=>", "==", "!=", "<=", ">=", "&&", "||", "??", "?.", "++", "--", "+=", "-=", "*=", "/=", "%=", "&=", "|=",
    "^=", "<<", ">>", "**",
    "{", "}", "(", ")", "[", "]", ";", ",", "<", ">", "+", "-", "*", "/", "%", "&", "|", "^", "!", "~", "?", ":",
    "=", ".",
]

_ID_START = re.compile(r"[A-Za-z_$]")
_ID = re.compile(r"[A-Za-z_$][A-Za-z0-9_$]*")
_NUM = re.compile(r"0[xX][0-9a-fA-F]+|0[oO][0-7]+|0[bB][01]+|(\d+\.?\d*([eE][+-]?\d+)?|\.\d+([eE][+-]?\d+)?)")


class JSSyntaxError(Exception):
    pass


class Tok:
    __slots__ = ("kind", "val", "pos", "nl")

    def __init__(self, kind, val, pos, nl):
        self.kind, self.val, self.pos, self.nl = kind, val, pos, nl  # nl: a line break precedes the token

    def __repr__(self):
        return f"{self.kind}:{self.val!r}"


_ESC = {"n": "\n", "t": "\t", "r": "\r", "b": "\b", "f": "\f", "v": "\v", "0": "\0"}


def _unescape(s: str, i: int):
    """s[i] is the char after a backslash -> (text, next index)."""
    c = s[i]
    if c == "u":
        if s[i + 1] == "{":
            j = s.index("}", i)
            return chr(int(s[i + 2:j], 16)), j + 1
        return chr(int(s[i + 1:i + 5], 16)), i + 5
    if c == "x":
        return chr(int(s[i + 1:i + 3], 16)), i + 3
    if c == "\n":
        return "", i + 1
    return _ESC.get(c, c), i + 1


def tokenize(src: str):
    toks, i, n, nl = [], 0, len(src), False
    prev_sig = None  # last significant token, to tell `/` (divide) from a regex literal

    def regex_allowed():
        if prev_sig is None:
            return True
        if prev_sig.kind in ("num", "str", "template", "regex"):
            return False
        if prev_sig.kind == "id":
            return False
        if prev_sig.kind == "kw":
            return prev_sig.val not in ("this", "null", "true", "false", "super")
        return prev_sig.val not in (")", "]", "}")

    while i < n:
        c = src[i]
        if c == "\n":
            nl = True
            i += 1
            continue
        if c in " \t\r﻿":
            i += 1
            continue
        if src.startswith("//", i):
            j = src.find("\n", i)
            i = n if j < 0 else j
            continue
        if src.startswith("/*", i):
            j = src.index("*/", i)
            if "\n" in src[i:j]:
                nl = True
            i = j + 2
            continue
        start = i
        if _ID_START.match(c):
            m = _ID.match(src, i)
            word = m.group(0)
            i = m.end()
            t = Tok("kw" if word in KEYWORDS else "id", word, start, nl)
        elif c.isdigit() or (c == "." and i + 1 < n and src[i + 1].isdigit()):
            m = _NUM.match(src, i)
            text = m.group(0)
            i = m.end()
            low = text[:2].lower()
            val = float(int(text[2:], {"0x": 16, "0o": 8, "0b": 2}[low])) if low in ("0x", "0o", "0b") else float(text)
            t = Tok("num", val, start, nl)
        elif c in "\"'":
            i += 1
            out = []
            while src[i] != c:
                if src[i] == "\\":
                    s, i = _unescape(src, i + 1)
                    out.append(s)
                else:
                    out.append(src[i])
                    i += 1
            i += 1
            t = Tok("str", "".join(out), start, nl)
        elif c == "`":
            # template literal -> ("template", [str, expr_src, str, ...]) with expression sources re-tokenized later
            i += 1
            parts, cur = [], []
            while src[i] != "`":
                if src[i] == "\\":
                    s, i = _unescape(src, i + 1)
                    cur.append(s)
                elif src.startswith("${", i):
                    parts.append("".join(cur))
                    cur = []
                    depth, j = 1, i + 2
                    while depth:
                        ch = src[j]
                        if ch == "{":
                            depth += 1
                        elif ch == "}":
                            depth -= 1
                        elif ch in "\"'`":  # skip nested strings
                            q, j = ch, j + 1
                            while src[j] != q:
                                j += 2 if src[j] == "\\" else 1
                        j += 1
                    parts.append(tokenize(src[i + 2:j - 1]))
                    i = j
                else:
                    cur.append(src[i])
                    i += 1
            parts.append("".join(cur))
            i += 1
            t = Tok("template", parts, start, nl)
        elif c == "/" and regex_allowed():
            j, in_class = i + 1, False
            while True:
                ch = src[j]
                if ch == "\\":
                    j += 2
                    continue
                if ch == "[":
                    in_class = True
                elif ch == "]":
                    in_class = False
                elif ch == "/" and not in_class:
                    break
                j += 1
            body = src[i + 1:j]
            m = re.compile(r"[a-z]*").match(src, j + 1)
            i = m.end()
            t = Tok("regex", (body, m.group(0)), start, nl)
        else:
            for p in PUNCT:
                if src.startswith(p, i):
                    i += len(p)
                    t = Tok("p", p, start, nl)
                    break
            else:
                raise JSSyntaxError(f"unexpected character {c!r} at {i}")
        toks.append(t)
        prev_sig = t
        nl = False
    toks.append(Tok("eof", None, n, True))
    return toks


ASSIGN_OPS = {"=", "+=", "-=", "*=", "/=", "%=", "**=", "<<=", ">>=", ">>>=", "&=", "|=", "^=", "&&=", "||=", "??="}
BINARY_PREC = {
    "??": 1, "||": 2, "&&": 3, "|": 4, "^": 5, "&": 6, "==": 7, "!=": 7, "===": 7, "!==": 7,
    "<": 8, ">": 8, "<=": 8, ">=": 8, "instanceof": 8, "in": 8, "<<": 9, ">>": 9, ">>>": 9,
    "+": 10, "-": 10, "*": 11, "/": 11, "%": 11, "**": 12,
}


class Parser:
    def __init__(self, toks):
        self.t, self.i = toks, 0
        self.no_in = False

    # ---- token helpers
    @property
    def cur(self) -> Tok:
        return self.t[self.i]

    def peek(self, k=1) -> Tok:
        return self.t[min(self.i + k, len(self.t) - 1)]

    def at(self, val) -> bool:
        c = self.cur
        return c.kind in ("p", "kw") and c.val == val

    def at_word(self, val) -> bool:
        c = self.cur
        return c.kind in ("kw", "id") and c.val == val

    def eat(self, val) -> bool:
        if self.at(val):
            self.i += 1
            return True
        return False

    def expect(self, val):
        if not self.eat(val):
            raise JSSyntaxError(f"expected {val!r}, got {self.cur!r} at {self.cur.pos}")

    def ident(self) -> str:
        c = self.cur
        if c.kind == "id" or (c.kind == "kw" and c.val in SOFT):
            self.i += 1
            return c.val
        raise JSSyntaxError(f"expected identifier, got {c!r} at {c.pos}")

    def prop_name(self):
        """Property name after `.` or in an object literal / class body: any word, string or number."""
        c = self.cur
        if c.kind in ("id", "kw"):
            self.i += 1
            return c.val
        if c.kind == "str":
            self.i += 1
            return c.val
        if c.kind == "num":
            self.i += 1
            from ..jsvalue import number_to_string
            return number_to_string(c.val)
        raise JSSyntaxError(f"expected property name, got {c!r} at {c.pos}")

    def end_stmt(self):
        if self.eat(";"):
            return
        if self.at("}") or self.cur.kind == "eof" or self.cur.nl:
            return
        raise JSSyntaxError(f"expected ';', got {self.cur!r} at {self.cur.pos}")

    # ---- program / statements
    def program(self):
        body = []
        while self.cur.kind != "eof":
            body.append(self.statement())
        return ("block", body)

    def block(self):
        self.expect("{")
        body = []
        while not self.at("}"):
            body.append(self.statement())
        self.expect("}")
        return ("block", body)

    def statement(self):
        c = self.cur
        if c.kind == "p":
            if c.val == "{":
                return self.block()
            if c.val == ";":
                self.i += 1
                return ("empty",)
        if c.kind == "kw":
            v = c.val
            if v in ("var", "let", "const"):
                d = self.var_decl()
                self.end_stmt()
                return d
            if v == "function" or (v == "async" and self.peek().kind == "kw" and self.peek().val == "function"
                                   and not self.peek().nl):
                f = self.function(is_decl=True)
                return ("funcdecl", f[1], f)
            if v == "class":
                cls = self.class_()
                return ("classdecl", cls[1], cls)
            if v == "return":
                self.i += 1
                arg = None
                if not (self.at(";") or self.at("}") or self.cur.kind == "eof" or self.cur.nl):
                    arg = self.expression()
                self.end_stmt()
                return ("return", arg)
            if v == "if":
                self.i += 1
                self.expect("(")
                test = self.expression()
                self.expect(")")
                cons = self.statement()
                alt = self.statement() if self.eat("else") else None
                return ("if", test, cons, alt)
            if v == "for":
                return self.for_()
            if v == "while":
                self.i += 1
                self.expect("(")
                test = self.expression()
                self.expect(")")
                return ("while", test, self.statement())
            if v == "do":
                self.i += 1
                body = self.statement()
                self.expect("while")
                self.expect("(")
                test = self.expression()
                self.expect(")")
                self.eat(";")
                return ("dowhile", body, test)
            if v in ("break", "continue"):
                self.i += 1
                self.end_stmt()
                return (v,)
            if v == "throw":
                self.i += 1
                arg = self.expression()
                self.end_stmt()
                return ("throw", arg)
            if v == "try":
                self.i += 1
                blk = self.block()
                param, handler, final = None, None, None
                if self.eat("catch"):
                    if self.eat("("):
                        param = self.binding_target()
                        self.expect(")")
                    handler = self.block()
                if self.eat("finally"):
                    final = self.block()
                return ("try", blk, param, handler, final)
            if v == "switch":
                self.i += 1
                self.expect("(")
                disc = self.expression()
                self.expect(")")
                self.expect("{")
                cases = []
                while not self.at("}"):
                    if self.eat("default"):
                        test = None
                    else:
                        self.expect("case")
                        test = self.expression()
                    self.expect(":")
                    body = []
                    while not (self.at("case") or self.at("default") or self.at("}")):
                        body.append(self.statement())
                    cases.append((test, body))
                self.expect("}")
                return ("switch", disc, cases)
        e = self.expression()
        self.end_stmt()
        return ("expr", e)

    def var_decl(self):
        kind = self.cur.val
        self.i += 1
        decls = []
        while True:
            target = self.binding_target()
            init = self.assignment() if self.eat("=") else None
            decls.append((target, init))
            if not self.eat(","):
                break
        return ("var", kind, decls)

    def for_(self):
        self.expect("for")
        self.expect("(")
        init = None
        if self.at("var") or self.at("let") or self.at("const"):
            kind = self.cur.val
            save = self.i
            self.i += 1
            target = self.binding_target()
            if self.at_word("of") or self.at("in"):
                mode = self.cur.val
                self.i += 1
                right = self.assignment() if mode == "of" else self.expression()
                self.expect(")")
                return ("forin" if mode == "in" else "forof", kind, target, right, self.statement())
            self.i = save
            self.no_in = True
            init = self.var_decl()
            self.no_in = False
        elif not self.at(";"):
            self.no_in = True
            init = ("expr", self.expression())
            self.no_in = False
            if self.at_word("of") or self.at("in"):
                raise JSSyntaxError("for-in/of over an existing binding is not supported")
        self.expect(";")
        test = None if self.at(";") else self.expression()
        self.expect(";")
        update = None if self.at(")") else self.expression()
        self.expect(")")
        return ("for", init, test, update, self.statement())

    # ---- binding patterns: ("id", name) | ("objpat", [(key_expr_or_name, target, default)], rest) | ("arrpat", [..], rest)
    def binding_target(self):
        if self.at("{"):
            self.i += 1
            props, rest = [], None
            while not self.at("}"):
                if self.eat("..."):
                    rest = self.binding_target()
                else:
                    if self.at("["):
                        self.i += 1
                        key = ("computed", self.assignment())
                        self.expect("]")
                    else:
                        key = self.prop_name()
                    if self.eat(":"):
                        target = self.binding_target()
                    else:
                        target = ("id", key)
                    default = self.assignment() if self.eat("=") else None
                    props.append((key, target, default))
                if not self.eat(","):
                    break
            self.expect("}")
            return ("objpat", props, rest)
        if self.at("["):
            self.i += 1
            elems, rest = [], None
            while not self.at("]"):
                if self.at(","):
                    self.i += 1
                    elems.append(None)
                    continue
                if self.eat("..."):
                    rest = self.binding_target()
                else:
                    target = self.binding_target()
                    default = self.assignment() if self.eat("=") else None
                    elems.append((target, default))
                if not self.eat(","):
                    break
            self.expect("]")
            return ("arrpat", elems, rest)
        return ("id", self.ident())

    def params(self):
        """( a, b = 1, {c}, ...rest ) -> [(target, default)], rest_target"""
        self.expect("(")
        ps, rest = [], None
        while not self.at(")"):
            if self.eat("..."):
                rest = self.binding_target()
            else:
                target = self.binding_target()
                default = self.assignment() if self.eat("=") else None
                ps.append((target, default))
            if not self.eat(","):
                break
        self.expect(")")
        return ps, rest

    def function(self, is_decl=False, is_method=False, name=None):
        is_async = False
        if not is_method:
            if self.eat("async"):
                is_async = True
            self.expect("function")
            self.eat("*")
            if self.cur.kind == "id" or (self.cur.kind == "kw" and self.cur.val in SOFT and not self.at("(")):
                name = self.ident()
            elif is_decl:
                raise JSSyntaxError("function declaration needs a name")
        ps, rest = self.params()
        body = self.block()
        return ("function", name, ps, rest, body, False, is_async)

    def class_(self):
        self.expect("class")
        name = self.ident() if self.cur.kind == "id" else None
        parent = None
        if self.eat("extends"):
            parent = self.unary_postfix()
        self.expect("{")
        members = []  # (name, function, is_static, kind)
        while not self.at("}"):
            if self.eat(";"):
                continue
            is_static = False
            if self.at("static") and not (self.peek().kind == "p" and self.peek().val == "("):
                self.i += 1
                is_static = True
            is_async, kind = False, "method"
            if self.at("async") and not (self.peek().kind == "p" and self.peek().val == "("):
                self.i += 1
                is_async = True
            if (self.at("get") or self.at("set")) and not (self.peek().kind == "p" and self.peek().val in ("(", "=")):
                kind = self.cur.val
                self.i += 1
            self.eat("*")
            if self.at("["):
                self.i += 1
                key = ("computed", self.assignment())
                self.expect("]")
            else:
                key = self.prop_name()
            if self.at("("):
                f = self.function(is_method=True, name=key if isinstance(key, str) else None)
                f = f[:6] + (is_async,)
                members.append((key, f, is_static, kind))
            else:  # class field
                init = self.assignment() if self.eat("=") else None
                self.end_stmt()
                members.append((key, init, is_static, "field"))
        self.expect("}")
        return ("class", name, parent, members)

    # ---- expressions
    def expression(self):
        e = self.assignment()
        if self.at(","):
            seq = [e]
            while self.eat(","):
                seq.append(self.assignment())
            return ("seq", seq)
        return e

    def is_arrow_ahead(self) -> bool:
        """At `(`: does the matching `)` precede `=>`?  At an identifier: is the next token `=>`?"""
        if self.cur.kind == "id" or (self.cur.kind == "kw" and self.cur.val in SOFT):
            n = self.peek()
            return n.kind == "p" and n.val == "=>" and not n.nl
        if not self.at("("):
            return False
        depth, j = 0, self.i
        while True:
            t = self.t[j]
            if t.kind == "eof":
                return False
            if t.kind == "p":
                if t.val in ("(", "[", "{"):
                    depth += 1
                elif t.val in (")", "]", "}"):
                    depth -= 1
                    if depth == 0:
                        n = self.t[j + 1]
                        return n.kind == "p" and n.val == "=>"
            j += 1

    def arrow(self, is_async=False):
        if self.at("("):
            ps, rest = self.params()
        else:
            ps, rest = [(("id", self.ident()), None)], None
        self.expect("=>")
        if self.at("{"):
            body = self.block()
        else:
            body = ("block", [("return", self.assignment())])
        return ("function", None, ps, rest, body, True, is_async)

    def assignment(self):
        if self.at("async") and not self.peek().nl:
            n = self.peek()
            if n.kind == "kw" and n.val == "function":
                return self.function()
            save = self.i
            self.i += 1
            if self.is_arrow_ahead():
                return self.arrow(is_async=True)
            self.i = save
        if self.is_arrow_ahead():
            return self.arrow()
        left = self.conditional()
        c = self.cur
        if c.kind == "p" and c.val in ASSIGN_OPS:
            self.i += 1
            right = self.assignment()
            if c.val == "=":
                left = self.to_pattern(left)
            return ("assign", c.val, left, right)
        return left

    def to_pattern(self, e):
        """Reinterpret an expression on the left of `=` as a destructuring pattern when it is a literal."""
        k = e[0]
        if k == "object":
            props, rest = [], None
            for p in e[1]:
                if p[0] == "spread":
                    rest = self.to_pattern(p[1])
                else:
                    _, key, val = p
                    default = None
                    if val[0] == "assign" and val[1] == "=":
                        val, default = val[2], val[3]
                    props.append((key, self.to_pattern(val), default))
            return ("objpat", props, rest)
        if k == "array":
            elems, rest = [], None
            for x in e[1]:
                if x is None:
                    elems.append(None)
                elif x[0] == "spread":
                    rest = self.to_pattern(x[1])
                else:
                    default = None
                    if x[0] == "assign" and x[1] == "=":
                        x, default = x[2], x[3]
                    elems.append((self.to_pattern(x), default))
            return ("arrpat", elems, rest)
        if k == "name":
            return ("id", e[1])
        return e  # member expression etc.

    def conditional(self):
        test = self.binary(0)
        if self.eat("?"):
            save, self.no_in = self.no_in, False
            a = self.assignment()
            self.no_in = save
            self.expect(":")
            b = self.assignment()
            return ("cond", test, a, b)
        return test

    def binary(self, min_prec):
        left = self.unary()
        while True:
            c = self.cur
            op = c.val if c.kind in ("p", "kw") else None
            prec = BINARY_PREC.get(op)
            if prec is None or prec < min_prec or (op == "in" and self.no_in):
                return left
            self.i += 1
            right = self.binary(prec if op == "**" else prec + 1)
            left = ("logical" if op in ("&&", "||", "??") else "binary", op, left, right)

    def unary(self):
        c = self.cur
        if c.kind == "p" and c.val in ("!", "-", "+", "~"):
            self.i += 1
            return ("unary", c.val, self.unary())
        if c.kind == "p" and c.val in ("++", "--"):
            self.i += 1
            return ("update", c.val, True, self.unary())
        if c.kind == "kw" and c.val in ("typeof", "void", "delete"):
            self.i += 1
            return ("unary", c.val, self.unary())
        if c.kind == "kw" and c.val == "await" :
            self.i += 1
            return ("await", self.unary())
        e = self.unary_postfix()
        c = self.cur
        if c.kind == "p" and c.val in ("++", "--") and not c.nl:
            self.i += 1
            return ("update", c.val, False, e)
        return e

    def arguments(self):
        self.expect("(")
        args = []
        while not self.at(")"):
            if self.eat("..."):
                args.append(("spread", self.assignment()))
            else:
                args.append(self.assignment())
            if not self.eat(","):
                break
        self.expect(")")
        return args

    def unary_postfix(self):
        if self.at("new"):
            self.i += 1
            if self.at("new"):
                callee = self.unary_postfix()
            else:
                callee = self.primary()
                while True:  # member accesses bind tighter than the `new` arguments
                    if self.eat("."):
                        callee = ("member", callee, ("str", self.prop_name()), False)
                    elif self.at("["):
                        self.i += 1
                        k = self.expression()
                        self.expect("]")
                        callee = ("member", callee, k, False)
                    else:
                        break
            args = self.arguments() if self.at("(") else []
            e = ("new", callee, args)
        else:
            e = self.primary()
        while True:
            c = self.cur
            if c.kind == "p":
                if c.val == ".":
                    self.i += 1
                    e = ("member", e, ("str", self.prop_name()), False)
                    continue
                if c.val == "?.":
                    self.i += 1
                    if self.at("("):
                        e = ("call", e, self.arguments(), True)
                    elif self.at("["):
                        self.i += 1
                        k = self.expression()
                        self.expect("]")
                        e = ("member", e, k, True)
                    else:
                        e = ("member", e, ("str", self.prop_name()), True)
                    continue
                if c.val == "[":
                    self.i += 1
                    save, self.no_in = self.no_in, False
                    k = self.expression()
                    self.no_in = save
                    self.expect("]")
                    e = ("member", e, k, False)
                    continue
                if c.val == "(":
                    e = ("call", e, self.arguments(), False)
                    continue
            if c.kind == "template":
                raise JSSyntaxError("tagged templates are not supported")
            return e

    def primary(self):
        c = self.cur
        k = c.kind
        if k == "num":
            self.i += 1
            return ("num", c.val)
        if k == "str":
            self.i += 1
            return ("str", c.val)
        if k == "template":
            self.i += 1
            parts = []
            for j, part in enumerate(c.val):
                if j % 2 == 0:
                    parts.append(("str", part))
                else:
                    sub = Parser(part)
                    parts.append(sub.expression())
            return ("template", parts)
        if k == "regex":
            self.i += 1
            return ("regex", c.val[0], c.val[1])
        if k == "id":
            self.i += 1
            return ("name", c.val)
        if k == "kw":
            v = c.val
            if v == "this":
                self.i += 1
                return ("this",)
            if v == "null":
                self.i += 1
                return ("null",)
            if v in ("true", "false"):
                self.i += 1
                return ("bool", v == "true")
            if v == "function":
                return self.function()
            if v == "class":
                return self.class_()
            if v == "super":
                self.i += 1
                return ("super",)
            if v in SOFT:
                self.i += 1
                return ("name", v)
        if k == "p":
            if c.val == "(":
                self.i += 1
                save, self.no_in = self.no_in, False
                e = self.expression()
                self.no_in = save
                self.expect(")")
                return e
            if c.val == "[":
                self.i += 1
                elems = []
                while not self.at("]"):
                    if self.at(","):
                        self.i += 1
                        elems.append(None)
                        continue
                    if self.eat("..."):
                        elems.append(("spread", self.assignment()))
                    else:
                        elems.append(self.assignment())
                    if not self.eat(","):
                        break
                self.expect("]")
                return ("array", elems)
            if c.val == "{":
                return self.object_literal()
        raise JSSyntaxError(f"unexpected token {c!r} at {c.pos}")

    def object_literal(self):
        self.expect("{")
        props = []  # ("prop", key, value) | ("spread", expr) | ("accessor", kind, key, fn); key: str | ("computed", expr)
        while not self.at("}"):
            if self.eat("..."):
                props.append(("spread", self.assignment()))
            else:
                is_async = False
                if self.at("async") and self.peek().kind in ("id", "kw", "str") and not self.peek().nl:
                    self.i += 1
                    is_async = True
                if (self.at("get") or self.at("set")) and self.peek().kind in ("id", "kw", "str", "num"):
                    kind = self.cur.val
                    self.i += 1
                    key = self.prop_name()
                    props.append(("accessor", kind, key, self.function(is_method=True, name=key)))
                    if not self.eat(","):
                        break
                    continue
                if self.at("["):
                    self.i += 1
                    key = ("computed", self.assignment())
                    self.expect("]")
                else:
                    key = self.prop_name()
                if self.at("("):
                    f = self.function(is_method=True, name=key if isinstance(key, str) else None)
                    props.append(("prop", key, f[:6] + (is_async,)))
                elif self.eat(":"):
                    props.append(("prop", key, self.assignment()))
                else:  # shorthand, possibly with a default when used as a pattern
                    val = ("name", key)
                    if self.at("="):
                        self.i += 1
                        val = ("assign", "=", ("name", key), self.assignment())
                    props.append(("prop", key, val))
            if not self.eat(","):
                break
        self.expect("}")
        return ("object", props)


def parse(src: str):
    return Parser(tokenize(src)).program()
