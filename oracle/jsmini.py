"""jsmini — a small ECMAScript-subset interpreter (TEST INFRASTRUCTURE, never on the product path).

Why it exists: the reference (triq-org/spectroplot-js) is browser JavaScript and this image has no JS engine
(no node / deno / bun / qjs / V8 / Python-embedded engine), so the reference worker cannot be run as shipped.
jsmini executes the reference's UNMODIFIED source files (lib/worker.js, lib/samples.js, lib/fft_nayuki.js,
lib/polyfill.js, lib/windows.js, the colormap modules, lib/utils.js, lib/parseFreqRate.js) read from
/root/reference at fixture-generation time (tools/make_ref_golden.py), so that tests/golden/ref_*.npz hold outputs
of the reference's own code rather than of a restatement.  Nothing of the reference is copied into this repo.

Semantics implemented (what those files use): ES modules (import / export / export default), let / const / var,
function declarations (hoisted) and expressions, arrow functions, default and rest parameters, `arguments`,
classes (constructor, methods, getters), `new`, `this`, closures, automatic semicolon insertion, template literals,
all arithmetic / bitwise / comparison / logical operators with ToInt32 / ToUint32 / ToNumber conversions,
compound assignment, ++ / --, conditional and comma operators, if / for / for-in / for-of / while / do-while / switch /
break / continue / return / throw / try-catch-finally, object and array literals (spread in calls and arrays),
Array (holes read as undefined, non-index keys such as -3 or NaN become ordinary properties), the typed arrays
(incl. Uint8ClampedArray's clamp + round-half-to-even + NaN -> 0 store conversion), ArrayBuffer(.slice), Math, Object
helpers, String / Number basics, Promise.resolve stub, console.

Numbers are IEEE doubles (Python float) or exact Python ints where the value is integral; Math.log10 / cos / sin come
from the platform libm (V8 uses its own fdlibm port: results may differ by <= 1 ulp, invisible except at exact
quantisation ties).  Not implemented: regular expressions, generators, async, getters on object literals, labels,
destructuring, optional chaining, tagged templates, Symbol, Proxy, Date.
"""
import math
import os
import struct

import numpy as np


# ----------------------------------------------------------------------------------------------- values
class _Undef:
    __slots__ = ()

    def __repr__(self):
        return "undefined"

    def __bool__(self):
        return False


UNDEF = _Undef()


class JSThrow(Exception):
    def __init__(self, value):
        Exception.__init__(self, js_to_string(value) if not isinstance(value, JSObject) else repr(value))
        self.value = value


class JSObject:
    __slots__ = ("props", "proto")

    def __init__(self, proto=None):
        self.props = {}
        self.proto = proto

    def __repr__(self):
        return "JSObject(%s)" % ", ".join(self.props.keys())


class Getter:
    __slots__ = ("fn",)

    def __init__(self, fn):
        self.fn = fn


class JSFunction(JSObject):
    __slots__ = ("params", "body", "env", "arrow", "name", "interp", "uses_args", "hoist", "expr_body", "ctor_kind")

    def __init__(self, interp, name, params, body, env, arrow, uses_args, hoist, expr_body):
        JSObject.__init__(self, interp.function_proto)
        self.interp, self.name, self.params, self.body, self.env = interp, name, params, body, env
        self.arrow, self.uses_args, self.hoist, self.expr_body = arrow, uses_args, hoist, expr_body
        if not arrow:
            po = JSObject(interp.object_proto)
            po.props["constructor"] = self
            self.props["prototype"] = po

    def call(self, this, args):
        env = Scope(self.env)
        v = env.vars
        if not self.arrow:
            v["this"] = this
            if self.uses_args:
                v["arguments"] = JSArray(self.interp, list(args))
        na = len(args)
        for i, (pname, pdef, rest) in enumerate(self.params):
            if rest:
                v[pname] = JSArray(self.interp, list(args[i:]))
            elif i < na and args[i] is not UNDEF:
                v[pname] = args[i]
            elif pdef is not None:
                v[pname] = pdef(env)
            else:
                v[pname] = UNDEF
        if self.expr_body:
            return self.body(env)
        for name in self.hoist[0]:
            v.setdefault(name, UNDEF)
        for name, mk in self.hoist[1]:
            v[name] = mk(env)
        r = self.body(env)
        if r is None or r is BREAK or r is CONTINUE:
            return UNDEF
        return r.value


class NativeFunction(JSObject):
    __slots__ = ("fn", "name", "construct")

    def __init__(self, interp, name, fn, construct=None):
        JSObject.__init__(self, interp.function_proto if interp else None)
        self.fn, self.name, self.construct = fn, name, construct

    def call(self, this, args):
        return self.fn(this, args)


class JSArray(JSObject):
    __slots__ = ("list",)

    def __init__(self, interp, lst):
        JSObject.__init__(self, interp.array_proto)
        self.list = lst


class JSArrayBuffer(JSObject):
    __slots__ = ("data",)

    def __init__(self, interp, data):
        JSObject.__init__(self, interp.object_proto)
        self.data = data            # bytearray


TYPED = {"Uint8Array": np.uint8, "Int8Array": np.int8, "Uint16Array": np.uint16, "Int16Array": np.int16,
         "Uint32Array": np.uint32, "Int32Array": np.int32, "Float32Array": np.float32, "Float64Array": np.float64,
         "Uint8ClampedArray": np.uint8, "BigUint64Array": np.uint64, "BigInt64Array": np.int64}


class JSTypedArray(JSObject):
    __slots__ = ("arr", "kind", "buffer", "is_float")

    def __init__(self, interp, kind, buffer, arr):
        JSObject.__init__(self, interp.object_proto)
        self.kind, self.buffer, self.arr = kind, buffer, arr
        self.is_float = kind.startswith("Float")


class Scope:
    __slots__ = ("vars", "parent")

    def __init__(self, parent):
        self.vars = {}
        self.parent = parent


class _Signal:
    __slots__ = ()


BREAK, CONTINUE = _Signal(), _Signal()


class Return:
    __slots__ = ("value",)

    def __init__(self, value):
        self.value = value


# ----------------------------------------------------------------------------------------------- conversions
def js_typeof(v):
    if v is UNDEF:
        return "undefined"
    if v is None:
        return "object"
    if isinstance(v, bool):
        return "boolean"
    if isinstance(v, (int, float)):
        return "number"
    if isinstance(v, str):
        return "string"
    if isinstance(v, (JSFunction, NativeFunction)):
        return "function"
    return "object"


def to_number(v):
    t = type(v)
    if t is float or t is int:
        return v
    if v is UNDEF:
        return math.nan
    if v is None:
        return 0
    if t is bool:
        return 1 if v else 0
    if t is str:
        s = v.strip()
        if s == "":
            return 0
        try:
            if s[:2].lower() == "0x":
                return int(s, 16)
            f = float(s)
            return int(f) if f.is_integer() and abs(f) < 2 ** 53 and "." not in s and "e" not in s.lower() else f
        except ValueError:
            return math.nan
    if isinstance(v, JSArray):
        if len(v.list) == 0:
            return 0
        if len(v.list) == 1:
            return to_number(v.list[0])
    return math.nan


def to_int32(v):
    if type(v) is int:
        v &= 0xFFFFFFFF
        return v - 0x100000000 if v & 0x80000000 else v
    v = to_number(v)
    if type(v) is float:
        if v != v or v in (math.inf, -math.inf):
            return 0
        v = int(v)              # truncation toward zero
    v &= 0xFFFFFFFF
    return v - 0x100000000 if v & 0x80000000 else v


def to_uint32(v):
    return to_int32(v) & 0xFFFFFFFF


def num_to_str(n):
    if type(n) is int:
        return str(n)
    if n != n:
        return "NaN"
    if n == math.inf:
        return "Infinity"
    if n == -math.inf:
        return "-Infinity"
    if n.is_integer() and abs(n) < 1e21:
        return str(int(n))
    r = repr(n)
    if "e" in r:
        m, e = r.split("e")
        if m.endswith(".0"):
            m = m[:-2]
        e = int(e)
        return "%se%s%d" % (m, "+" if e > 0 else "-", abs(e))
    return r


def js_to_string(v):
    if isinstance(v, str):
        return v
    if v is UNDEF:
        return "undefined"
    if v is None:
        return "null"
    if isinstance(v, bool):
        return "true" if v else "false"
    if isinstance(v, (int, float)):
        return num_to_str(v)
    if isinstance(v, JSArray):
        return ",".join("" if (x is UNDEF or x is None) else js_to_string(x) for x in v.list)
    if isinstance(v, (JSFunction, NativeFunction)):
        return "function %s() { [code] }" % (v.name or "")
    return "[object Object]"


def truthy(v):
    t = type(v)
    if t is bool:
        return v
    if t is float:
        return v == v and v != 0.0
    if t is int:
        return v != 0
    if t is str:
        return v != ""
    return not (v is UNDEF or v is None)


def prop_key(k):
    """ToPropertyKey: ints stay ints (array indices), everything else becomes its JS string."""
    t = type(k)
    if t is int:
        return k if k >= 0 else str(k)
    if t is str:
        if k.isdigit() and (k == "0" or k[0] != "0"):
            return int(k)
        return k
    if t is float:
        if k.is_integer() and 0 <= k < 2 ** 53:
            return int(k)
        return num_to_str(k)
    return js_to_string(k)


def js_div(a, b):
    a, b = to_number(a), to_number(b)
    try:
        return a / b
    except ZeroDivisionError:
        if a != a or a == 0:
            return math.nan
        neg = (a < 0) != (math.copysign(1.0, b) < 0)
        return -math.inf if neg else math.inf
    except OverflowError:
        return float(a) / float(b)


def js_mod(a, b):
    a, b = to_number(a), to_number(b)
    if type(a) is int and type(b) is int and b != 0:
        r = abs(a) % abs(b)
        return -r if a < 0 else r
    try:
        return math.fmod(a, b)
    except (ValueError, ZeroDivisionError):
        return math.nan


def js_add(a, b):
    ta, tb = type(a), type(b)
    if (ta is float or ta is int) and (tb is float or tb is int):
        return a + b
    if isinstance(a, JSObject):
        a = js_to_string(a)
    if isinstance(b, JSObject):
        b = js_to_string(b)
    if isinstance(a, str) or isinstance(b, str):
        return js_to_string(a) + js_to_string(b)
    return to_number(a) + to_number(b)


def js_pow(a, b):
    a, b = to_number(a), to_number(b)
    try:
        r = a ** b
        if isinstance(r, complex):
            return math.nan
        return r
    except ZeroDivisionError:
        return math.inf
    except OverflowError:
        return math.inf


def strict_eq(a, b):
    ta, tb = type(a), type(b)
    if (ta is int or ta is float) and (tb is int or tb is float):
        return (ta is bool) == (tb is bool) and a == b
    if ta is bool or tb is bool:
        return ta is tb and a == b
    if ta is str and tb is str:
        return a == b
    if ta is str or tb is str:
        return False
    return a is b


def loose_eq(a, b):
    if (a is None or a is UNDEF) and (b is None or b is UNDEF):
        return True
    if a is None or a is UNDEF or b is None or b is UNDEF:
        return False
    ta, tb = js_typeof(a), js_typeof(b)
    if ta == tb:
        return strict_eq(a, b)
    if ta in ("object", "function") or tb in ("object", "function"):
        if ta in ("object", "function") and tb in ("object", "function"):
            return a is b
        return loose_eq(js_to_string(a) if ta in ("object", "function") else a, js_to_string(b) if tb in ("object", "function") else b)
    return to_number(a) == to_number(b)


def js_compare(op, a, b):
    if isinstance(a, str) and isinstance(b, str):
        pass
    else:
        a, b = to_number(a), to_number(b)
    if op == "<":
        return a < b
    if op == ">":
        return a > b
    if op == "<=":
        return a <= b
    return a >= b


def clamp_u8(v):
    """Uint8ClampedArray store conversion: NaN -> 0, clamp to [0, 255], round half to even."""
    v = to_number(v)
    if v != v:
        return 0
    if v <= 0:
        return 0
    if v >= 255:
        return 255
    f = math.floor(v)
    d = v - f
    if d < 0.5:
        return int(f)
    if d > 0.5:
        return int(f) + 1
    return int(f) if int(f) % 2 == 0 else int(f) + 1


# ----------------------------------------------------------------------------------------------- tokenizer
KEYWORDS = {"var", "let", "const", "function", "return", "if", "else", "for", "while", "do", "break", "continue", "new",
            "delete", "typeof", "instanceof", "in", "of", "this", "null", "undefined", "true", "false", "class", "extends",
            "import", "export", "default", "from", "throw", "try", "catch", "finally", "switch", "case", "void", "get", "static", "as"}
PUNCT = [">>>=", "...", "===", "!==", "**=", "<<=", ">>=", ">>>", "=>", "==", "!=", "<=", ">=", "&&", "||", "??", "++", "--",
         "+=", "-=", "*=", "/=", "%=", "&=", "|=", "^=", "**", "<<", ">>", "{", "}", "(", ")", "[", "]", ";", ",", "<", ">", "+",
         "-", "*", "/", "%", "&", "|", "^", "!", "~", "?", ":", "=", "."]


class Tok:
    __slots__ = ("t", "v", "nl", "pos")

    def __init__(self, t, v, nl, pos):
        self.t, self.v, self.nl, self.pos = t, v, nl, pos

    def __repr__(self):
        return "%s:%r" % (self.t, self.v)


def tokenize(src):
    toks, i, n, nl = [], 0, len(src), False
    while i < n:
        c = src[i]
        if c == "\n":
            nl = True
            i += 1
        elif c in " \t\r﻿":
            i += 1
        elif src.startswith("//", i):
            j = src.find("\n", i)
            i = n if j < 0 else j
        elif src.startswith("/*", i):
            j = src.find("*/", i + 2)
            if "\n" in src[i:j]:
                nl = True
            i = j + 2
        elif c.isdigit() or (c == "." and i + 1 < n and src[i + 1].isdigit()):
            j = i
            if src[i:i + 2].lower() == "0x":
                j = i + 2
                while j < n and src[j] in "0123456789abcdefABCDEF":
                    j += 1
                val = int(src[i:j], 16)
            else:
                while j < n and (src[j].isdigit() or src[j] == "."):
                    j += 1
                if j < n and src[j] in "eE":
                    k = j + 1
                    if k < n and src[k] in "+-":
                        k += 1
                    if k < n and src[k].isdigit():
                        j = k
                        while j < n and src[j].isdigit():
                            j += 1
                txt = src[i:j]
                val = int(txt) if txt.isdigit() else float(txt)
            toks.append(Tok("num", val, nl, i)); nl = False
            i = j
        elif c.isalpha() or c in "_$":
            j = i + 1
            while j < n and (src[j].isalnum() or src[j] in "_$"):
                j += 1
            toks.append(Tok("id", src[i:j], nl, i)); nl = False
            i = j
        elif c in "'\"":
            j, out = i + 1, []
            while src[j] != c:
                if src[j] == "\\":
                    j += 1
                    e = src[j]
                    if e == "u":
                        out.append(chr(int(src[j + 1:j + 5], 16))); j += 4
                    elif e == "x":
                        out.append(chr(int(src[j + 1:j + 3], 16))); j += 2
                    else:
                        out.append({"n": "\n", "t": "\t", "r": "\r", "0": "\0", "b": "\b", "f": "\f", "v": "\v", "\n": ""}.get(e, e))
                else:
                    out.append(src[j])
                j += 1
            toks.append(Tok("str", "".join(out), nl, i)); nl = False
            i = j + 1
        elif c == "`":
            j, parts, cur = i + 1, [], []
            while src[j] != "`":
                if src[j] == "\\":
                    j += 1
                    cur.append({"n": "\n", "t": "\t"}.get(src[j], src[j])); j += 1
                elif src.startswith("${", j):
                    depth, k = 1, j + 2
                    while depth:
                        if src[k] == "{":
                            depth += 1
                        elif src[k] == "}":
                            depth -= 1
                        k += 1
                    parts.append("".join(cur)); cur = []
                    parts.append(tokenize(src[j + 2:k - 1]))
                    j = k
                else:
                    cur.append(src[j]); j += 1
            parts.append("".join(cur))
            toks.append(Tok("tmpl", parts, nl, i)); nl = False
            i = j + 1
        else:
            for p in PUNCT:
                if src.startswith(p, i):
                    toks.append(Tok("p", p, nl, i)); nl = False
                    i += len(p)
                    break
            else:
                raise SyntaxError("jsmini: unexpected character %r at %d" % (c, i))
    toks.append(Tok("eof", None, True, n))
    return toks


# ----------------------------------------------------------------------------------------------- parser -> AST (tuples)
BINPREC = {"??": 1, "||": 2, "&&": 3, "|": 4, "^": 5, "&": 6, "==": 7, "!=": 7, "===": 7, "!==": 7, "<": 8, ">": 8, "<=": 8,
           ">=": 8, "instanceof": 8, "in": 8, "<<": 9, ">>": 9, ">>>": 9, "+": 10, "-": 10, "*": 11, "/": 11, "%": 11, "**": 12}
ASSIGN_OPS = {"=", "+=", "-=", "*=", "/=", "%=", "&=", "|=", "^=", "<<=", ">>=", ">>>=", "**="}


class Parser:
    def __init__(self, toks, fname="<js>"):
        self.toks, self.i, self.fname = toks, 0, fname
        self.fn_stack = [{"args": False}]
        self.no_in = False

    # -- helpers
    def peek(self, k=0):
        return self.toks[self.i + k]

    def next(self):
        t = self.toks[self.i]
        self.i += 1
        return t

    def is_p(self, v, k=0):
        t = self.toks[self.i + k]
        return t.t == "p" and t.v == v

    def is_id(self, v, k=0):
        t = self.toks[self.i + k]
        return t.t == "id" and t.v == v

    def eat_p(self, v):
        if self.is_p(v):
            self.i += 1
            return True
        return False

    def expect_p(self, v):
        t = self.next()
        if t.t != "p" or t.v != v:
            raise SyntaxError("jsmini %s: expected %r, got %r at %d" % (self.fname, v, t, t.pos))

    def expect_id(self, v=None):
        t = self.next()
        if t.t != "id" or (v is not None and t.v != v):
            raise SyntaxError("jsmini %s: expected identifier %r, got %r at %d" % (self.fname, v, t, t.pos))
        return t.v

    def semicolon(self):
        if self.eat_p(";"):
            return
        t = self.peek()
        if t.t == "eof" or (t.t == "p" and t.v == "}") or t.nl:
            return
        raise SyntaxError("jsmini %s: expected ';', got %r at %d" % (self.fname, t, t.pos))

    # -- program
    def program(self):
        body = []
        while self.peek().t != "eof":
            body.append(self.statement())
        return body

    def statement(self):
        t = self.peek()
        if t.t == "p":
            if t.v == "{":
                return self.block()
            if t.v == ";":
                self.next()
                return ("empty",)
        if t.t == "id":
            v = t.v
            if v in ("var", "let", "const"):
                d = self.var_decl()
                self.semicolon()
                return d
            if v == "function":
                return self.function(decl=True)
            if v == "class":
                return self.klass(decl=True)
            if v == "if":
                self.next(); self.expect_p("(")
                c = self.expression(); self.expect_p(")")
                a = self.statement()
                b = None
                if self.is_id("else"):
                    self.next()
                    b = self.statement()
                return ("if", c, a, b)
            if v == "for":
                return self.for_stmt()
            if v == "while":
                self.next(); self.expect_p("(")
                c = self.expression(); self.expect_p(")")
                return ("while", c, self.statement())
            if v == "do":
                self.next()
                b = self.statement()
                self.expect_id("while"); self.expect_p("(")
                c = self.expression(); self.expect_p(")")
                self.eat_p(";")
                return ("dowhile", c, b)
            if v == "return":
                self.next()
                t2 = self.peek()
                e = None
                if not (t2.nl or t2.t == "eof" or (t2.t == "p" and t2.v in (";", "}"))):
                    e = self.expression()
                self.semicolon()
                return ("return", e)
            if v == "break":
                self.next(); self.semicolon()
                return ("break",)
            if v == "continue":
                self.next(); self.semicolon()
                return ("continue",)
            if v == "throw":
                self.next()
                e = self.expression(); self.semicolon()
                return ("throw", e)
            if v == "try":
                self.next()
                b = self.block()
                param, h, f = None, None, None
                if self.is_id("catch"):
                    self.next()
                    if self.eat_p("("):
                        param = self.expect_id(); self.expect_p(")")
                    h = self.block()
                if self.is_id("finally"):
                    self.next()
                    f = self.block()
                return ("try", b, param, h, f)
            if v == "switch":
                self.next(); self.expect_p("(")
                d = self.expression(); self.expect_p(")"); self.expect_p("{")
                cases = []
                while not self.eat_p("}"):
                    if self.is_id("default"):
                        self.next(); test = None
                    else:
                        self.expect_id("case"); test = self.expression()
                    self.expect_p(":")
                    body = []
                    while not (self.is_id("case") or self.is_id("default") or self.is_p("}")):
                        body.append(self.statement())
                    cases.append((test, body))
                return ("switch", d, cases)
            if v == "import":
                return self.import_stmt()
            if v == "export":
                return self.export_stmt()
        e = self.expression()
        self.semicolon()
        return ("expr", e)

    def block(self):
        self.expect_p("{")
        body = []
        while not self.eat_p("}"):
            body.append(self.statement())
        return ("block", body)

    def var_decl(self):
        kind = self.next().v
        decls = []
        while True:
            if self.is_p("{"):                      # const { a, b: c } = expr
                self.next()
                names = []
                while not self.eat_p("}"):
                    key = self.next().v
                    alias = key
                    if self.eat_p(":"):
                        alias = self.expect_id()
                    names.append((key, alias))
                    if not self.is_p("}"):
                        self.expect_p(",")
                self.expect_p("=")
                decls.append((tuple(names), self.assignment()))
                if not self.eat_p(","):
                    break
                continue
            name = self.expect_id()
            init = None
            if self.eat_p("="):
                init = self.assignment()
            decls.append((name, init))
            if not self.eat_p(","):
                break
        return ("var", kind, decls)

    def for_stmt(self):
        self.next(); self.expect_p("(")
        init = None
        if self.is_p(";"):
            pass
        elif self.peek().t == "id" and self.peek().v in ("var", "let", "const"):
            if self.peek(2).t == "id" and self.peek(2).v in ("in", "of"):
                kind = self.next().v
                name = self.expect_id()
                io = self.next().v
                obj = self.expression(); self.expect_p(")")
                return ("for" + io, kind, name, obj, self.statement())
            self.no_in = True
            init = self.var_decl()
            self.no_in = False
        else:
            self.no_in = True
            init = ("expr", self.expression())
            self.no_in = False
        self.expect_p(";")
        test = None if self.is_p(";") else self.expression()
        self.expect_p(";")
        upd = None if self.is_p(")") else self.expression()
        self.expect_p(")")
        return ("for", init, test, upd, self.statement())

    def params(self):
        ps = []
        self.expect_p("(")
        while not self.eat_p(")"):
            rest = self.eat_p("...")
            name = self.expect_id()
            d = None
            if self.eat_p("="):
                d = self.assignment()
            ps.append((name, d, rest))
            if not self.is_p(")"):
                self.expect_p(",")
        return ps

    def function_body(self):
        self.fn_stack.append({"args": False})
        body = self.block()
        info = self.fn_stack.pop()
        return body, info["args"]

    def function(self, decl=False):
        self.expect_id("function")
        name = None
        if self.peek().t == "id" and not self.is_p("("):
            name = self.next().v
        ps = self.params()
        body, ua = self.function_body()
        node = ("function", name, ps, body, False, ua, False)
        return ("funcdecl", name, node) if decl else node

    def klass(self, decl=False):
        self.expect_id("class")
        name = None
        if self.peek().t == "id" and not self.is_id("extends"):
            name = self.next().v
        if self.is_id("extends"):
            raise SyntaxError("jsmini: class extends not supported")
        self.expect_p("{")
        members = []
        while not self.eat_p("}"):
            if self.eat_p(";"):
                continue
            static = False
            if self.is_id("static") and not self.is_p("(", 1):
                self.next(); static = True
            kind = "method"
            if self.is_id("get") and not self.is_p("(", 1):
                self.next(); kind = "get"
            t = self.next()
            mname = t.v if t.t in ("id", "str") else num_to_str(t.v)
            ps = self.params()
            body, ua = self.function_body()
            members.append((kind, static, mname, ("function", mname, ps, body, False, ua, False)))
        node = ("class", name, members)
        return ("classdecl", name, node) if decl else node

    def import_stmt(self):
        self.expect_id("import")
        if self.peek().t == "str":
            src = self.next().v; self.semicolon()
            return ("import", src, None, [])
        default, names = None, []
        if self.peek().t == "id" and not self.is_p("{"):
            default = self.next().v
            self.eat_p(",")
        if self.eat_p("{"):
            while not self.eat_p("}"):
                a = self.next().v
                b = a
                if self.is_id("as"):
                    self.next(); b = self.expect_id()
                names.append((a, b))
                self.eat_p(",")
        elif self.eat_p("*"):
            self.expect_id("as")
            names.append(("*", self.expect_id()))
        self.expect_id("from")
        src = self.next().v; self.semicolon()
        return ("import", src, default, names)

    def export_stmt(self):
        self.expect_id("export")
        if self.is_id("default"):
            self.next()
            if self.is_id("function"):
                f = self.function(decl=False)
                return ("export_default", f)
            if self.is_id("class"):
                return ("export_default", self.klass(decl=False))
            e = self.assignment(); self.semicolon()
            return ("export_default", e)
        if self.eat_p("{"):
            names = []
            while not self.eat_p("}"):
                a = self.next().v
                b = a
                if self.is_id("as"):
                    self.next(); b = self.next().v
                names.append((a, b))
                self.eat_p(",")
            self.semicolon()
            return ("export_names", names)
        d = self.statement()
        return ("export_decl", d)

    # -- expressions
    def expression(self):
        e = self.assignment()
        if self.is_p(","):
            es = [e]
            while self.eat_p(","):
                es.append(self.assignment())
            return ("seq", es)
        return e

    def arrow_ahead(self):
        """at '(' : is this the parameter list of an arrow function?"""
        depth, k = 0, 0
        while True:
            t = self.peek(k)
            if t.t == "eof":
                return False
            if t.t == "p":
                if t.v in ("(", "[", "{"):
                    depth += 1
                elif t.v in (")", "]", "}"):
                    depth -= 1
                    if depth == 0:
                        return self.is_p("=>", k + 1)
            k += 1

    def arrow_body(self, ps):
        self.expect_p("=>")
        if self.is_p("{"):
            self.fn_stack.append(self.fn_stack[-1])     # arrows share `arguments` with the enclosing function
            body = self.block()
            self.fn_stack.pop()
            return ("function", None, ps, body, True, False, False)
        e = self.assignment()
        return ("function", None, ps, e, True, False, True)

    def assignment(self):
        t = self.peek()
        if t.t == "id" and t.v not in KEYWORDS and self.is_p("=>", 1):
            self.next()
            return self.arrow_body([(t.v, None, False)])
        if t.t == "p" and t.v == "(" and self.arrow_ahead():
            return self.arrow_body(self.params())
        left = self.conditional()
        t = self.peek()
        if t.t == "p" and t.v in ASSIGN_OPS:
            if left[0] not in ("name", "member", "index"):
                raise SyntaxError("jsmini %s: bad assignment target at %d" % (self.fname, t.pos))
            self.next()
            return ("assign", t.v, left, self.assignment())
        return left

    def conditional(self):
        c = self.binary(0)
        if self.eat_p("?"):
            save, self.no_in = self.no_in, False
            a = self.assignment()
            self.no_in = save
            self.expect_p(":")
            b = self.assignment()
            return ("cond", c, a, b)
        return c

    def binary(self, minprec):
        left = self.unary()
        while True:
            t = self.peek()
            op = t.v if (t.t == "p" or (t.t == "id" and t.v in ("instanceof", "in"))) else None
            if op == "in" and self.no_in:
                break
            prec = BINPREC.get(op) if op is not None else None
            if prec is None or prec < minprec:
                break
            self.next()
            right = self.binary(prec if op == "**" else prec + 1)
            left = ("logical", op, left, right) if op in ("&&", "||", "??") else ("bin", op, left, right)
        return left

    def unary(self):
        t = self.peek()
        if t.t == "p" and t.v in ("!", "~", "-", "+"):
            self.next()
            return ("unary", t.v, self.unary())
        if t.t == "p" and t.v in ("++", "--"):
            self.next()
            return ("update", t.v, True, self.unary())
        if t.t == "id" and t.v in ("typeof", "void", "delete"):
            self.next()
            return ("unary", t.v, self.unary())
        e = self.postfix()
        if self.is_p("**"):
            self.next()
            return ("bin", "**", e, self.unary())
        return e

    def postfix(self):
        e = self.call_member()
        t = self.peek()
        if t.t == "p" and t.v in ("++", "--") and not t.nl:
            self.next()
            return ("update", t.v, False, e)
        return e

    def arguments(self):
        args = []
        self.expect_p("(")
        while not self.eat_p(")"):
            if self.eat_p("..."):
                args.append(("spread", self.assignment()))
            else:
                args.append(self.assignment())
            if not self.is_p(")"):
                self.expect_p(",")
        return args

    def call_member(self):
        if self.is_id("new"):
            self.next()
            callee = self.member_only()
            args = self.arguments() if self.is_p("(") else []
            e = ("new", callee, args)
        else:
            e = self.primary()
        while True:
            if self.eat_p("."):
                e = ("member", e, self.next().v)
            elif self.is_p("["):
                self.next()
                save, self.no_in = self.no_in, False
                k = self.expression()
                self.no_in = save
                self.expect_p("]")
                e = ("index", e, k)
            elif self.is_p("("):
                e = ("call", e, self.arguments())
            else:
                return e

    def member_only(self):
        if self.is_id("new"):
            self.next()
            callee = self.member_only()
            args = self.arguments() if self.is_p("(") else []
            e = ("new", callee, args)
        else:
            e = self.primary()
        while True:
            if self.eat_p("."):
                e = ("member", e, self.next().v)
            elif self.is_p("["):
                self.next()
                k = self.expression(); self.expect_p("]")
                e = ("index", e, k)
            else:
                return e

    def primary(self):
        t = self.next()
        if t.t == "num":
            return ("lit", t.v)
        if t.t == "str":
            return ("lit", t.v)
        if t.t == "tmpl":
            parts = []
            for p in t.v:
                parts.append(p if isinstance(p, str) else Parser(p, self.fname).expression())
            return ("tmpl", parts)
        if t.t == "id":
            v = t.v
            if v == "function":
                self.i -= 1
                return self.function()
            if v == "class":
                self.i -= 1
                return self.klass()
            if v == "this":
                return ("name", "this")
            if v == "null":
                return ("lit", None)
            if v == "undefined":
                return ("name", "undefined")
            if v == "true":
                return ("lit", True)
            if v == "false":
                return ("lit", False)
            if v == "arguments":
                self.fn_stack[-1]["args"] = True
            return ("name", v)
        if t.t == "p":
            if t.v == "(":
                save, self.no_in = self.no_in, False
                e = self.expression()
                self.no_in = save
                self.expect_p(")")
                return e
            if t.v == "[":
                items = []
                while not self.eat_p("]"):
                    if self.is_p(","):
                        self.next(); items.append(("lit", UNDEF)); continue
                    if self.eat_p("..."):
                        items.append(("spread", self.assignment()))
                    else:
                        items.append(self.assignment())
                    if not self.is_p("]"):
                        self.expect_p(",")
                return ("array", items)
            if t.v == "{":
                props = []
                while not self.eat_p("}"):
                    kt = self.next()
                    if kt.t == "p" and kt.v == "[":
                        key = self.assignment(); self.expect_p("]")
                    elif kt.t == "p" and kt.v == "...":
                        props.append(("spread", None, self.assignment()))
                        self.eat_p(",")
                        continue
                    else:
                        key = ("lit", kt.v if kt.t in ("id", "str") else num_to_str(kt.v))
                    if self.is_p("("):
                        ps = self.params()
                        body, ua = self.function_body()
                        props.append(("kv", key, ("function", key[1], ps, body, False, ua, False)))
                    elif self.eat_p(":"):
                        props.append(("kv", key, self.assignment()))
                    else:
                        props.append(("kv", key, ("name", kt.v)))
                    if not self.is_p("}"):
                        self.expect_p(",")
                return ("object", props)
        raise SyntaxError("jsmini %s: unexpected token %r at %d" % (self.fname, t, t.pos))


# ----------------------------------------------------------------------------------------------- compiler: AST -> closures
def _lookup(env, name):
    e = env
    while e is not None:
        v = e.vars
        if name in v:
            return v[name]
        e = e.parent
    raise JSThrow("ReferenceError: %s is not defined" % name)


def _assign_name(env, name, val):
    e = env
    while e is not None:
        if name in e.vars:
            e.vars[name] = val
            return
        last = e
        e = e.parent
    last.vars[name] = val           # sloppy-mode global


def collect_hoists(stmts, top=True):
    """(var names, [(function name, node)]) declared in a function body (not crossing nested functions)."""
    names, funcs = [], []

    def walk(s, top_level):
        if not isinstance(s, tuple) or not s:
            return
        k = s[0]
        if k == "var" and s[1] == "var":
            for n, _ in s[2]:
                names.extend([a for _k, a in n] if isinstance(n, tuple) else [n])
        elif k == "funcdecl" and top_level:
            funcs.append((s[1], s[2]))
        elif k == "block":
            for x in s[1]:
                walk(x, False)
        elif k == "if":
            walk(s[2], False); walk(s[3], False)
        elif k == "for":
            walk(s[1], False); walk(s[4], False)
        elif k in ("forin", "forof"):
            if s[1] == "var":
                names.append(s[2])
            walk(s[4], False)
        elif k in ("while", "dowhile"):
            walk(s[2], False)
        elif k == "try":
            walk(s[1], False); walk(s[3], False); walk(s[4], False)
        elif k == "switch":
            for _, body in s[2]:
                for x in body:
                    walk(x, False)
        elif k == "export_decl":
            walk(s[1], top_level)
    for s in stmts:
        walk(s, True)
    return names, funcs


def has_lexical(stmts):
    for s in stmts:
        if s[0] in ("classdecl", "funcdecl") or (s[0] == "var" and s[1] != "var"):
            return True
    return False


class Interp:
    def __init__(self, root=None):
        self.root = root
        self.modules = {}
        self.object_proto = JSObject(None)
        self.function_proto = JSObject(self.object_proto)
        self.array_proto = JSObject(self.object_proto)
        self.globals = Scope(None)
        self.log = []
        _install_builtins(self)

    # ---- property access
    def get_prop(self, obj, key):
        if type(obj) is JSArray:
            if type(key) is int:
                lst = obj.list
                return lst[key] if 0 <= key < len(lst) else obj.props.get(str(key), UNDEF) if key < 0 else UNDEF
            key = prop_key(key)
            if type(key) is int:
                lst = obj.list
                return lst[key] if key < len(lst) else UNDEF
            if key == "length":
                return len(obj.list)
        elif type(obj) is JSTypedArray:
            if type(key) is not int:
                key = prop_key(key)
            if type(key) is int:
                a = obj.arr
                if 0 <= key < len(a):
                    v = a[key]
                    return float(v) if obj.is_float else int(v)
                return UNDEF
            if key == "length":
                return len(obj.arr)
            if key == "buffer":
                return obj.buffer
            if key == "byteLength":
                return obj.arr.nbytes
            if key in self.typed_methods:
                return self.typed_methods[key]
        elif isinstance(obj, JSObject):
            if type(obj) is JSArrayBuffer:
                if key == "byteLength":
                    return len(obj.data)
                if key == "slice":
                    return self.arraybuffer_slice
            key = prop_key(key)
            if type(key) is int:
                key = str(key)
        elif isinstance(obj, str):
            key = prop_key(key)
            if type(key) is int:
                return obj[key] if key < len(obj) else UNDEF
            if key == "length":
                return len(obj)
            return self.string_methods.get(key, UNDEF)
        elif isinstance(obj, (int, float)) and not isinstance(obj, bool):
            return self.number_methods.get(key, UNDEF)
        elif obj is UNDEF or obj is None:
            raise JSThrow("TypeError: cannot read property %s of %s" % (js_to_string(key), js_to_string(obj)))
        else:
            return UNDEF
        o = obj
        while o is not None:
            p = o.props
            if key in p:
                v = p[key]
                if type(v) is Getter:
                    return v.fn.call(obj, [])
                return v
            o = o.proto
        return UNDEF

    def set_prop(self, obj, key, val):
        if type(obj) is JSArray:
            if type(key) is not int:
                key = prop_key(key)
            if type(key) is int and key >= 0:
                lst = obj.list
                if key < len(lst):
                    lst[key] = val
                else:
                    lst.extend([UNDEF] * (key - len(lst)))
                    lst.append(val)
                return
            if key == "length":
                del obj.list[int(val):]
                return
            obj.props[str(key)] = val
            return
        if type(obj) is JSTypedArray:
            if type(key) is not int:
                key = prop_key(key)
            if type(key) is int:
                a = obj.arr
                if 0 <= key < len(a):
                    if obj.kind == "Uint8ClampedArray":
                        a[key] = clamp_u8(val)
                    elif obj.is_float:
                        a[key] = to_number(val)
                    else:
                        bits = a.itemsize * 8
                        v = to_number(val)
                        if type(v) is float:
                            v = 0 if (v != v or v in (math.inf, -math.inf)) else int(v)
                        v &= (1 << bits) - 1
                        if a.dtype.kind == "i" and v >> (bits - 1):
                            v -= 1 << bits
                        a[key] = v
                return
            obj.props[str(key)] = val
            return
        if isinstance(obj, JSObject):
            key = prop_key(key)
            obj.props[str(key) if type(key) is int else key] = val
            return
        if obj is UNDEF or obj is None:
            raise JSThrow("TypeError: cannot set property %s of %s" % (js_to_string(key), js_to_string(obj)))

    def call(self, fn, this, args):
        if not isinstance(fn, (JSFunction, NativeFunction)):
            raise JSThrow("TypeError: %s is not a function" % js_to_string(fn))
        return fn.call(this, args)

    def construct(self, fn, args):
        if isinstance(fn, NativeFunction):
            if fn.construct is None:
                raise JSThrow("TypeError: not a constructor")
            return fn.construct(args)
        if not isinstance(fn, JSFunction):
            raise JSThrow("TypeError: %s is not a constructor" % js_to_string(fn))
        proto = fn.props.get("prototype")
        obj = JSObject(proto if isinstance(proto, JSObject) else self.object_proto)
        r = fn.call(obj, args)
        return r if isinstance(r, JSObject) else obj

    def iterate(self, v):
        if isinstance(v, JSArray):
            return list(v.list)
        if isinstance(v, JSTypedArray):
            return [float(x) if v.is_float else int(x) for x in v.arr]
        if isinstance(v, str):
            return list(v)
        raise JSThrow("TypeError: not iterable")

    def keys_of(self, v):
        if isinstance(v, JSArray):
            return [str(i) for i, x in enumerate(v.list)] + list(v.props.keys())
        if isinstance(v, JSTypedArray):
            return [str(i) for i in range(len(v.arr))]
        if isinstance(v, str):
            return [str(i) for i in range(len(v))]
        if isinstance(v, JSObject):
            return list(v.props.keys())
        return []

    # ---- compile
    def c_args(self, args):
        fs = [(a[0] == "spread", self.c_expr(a[1] if a[0] == "spread" else a)) for a in args]
        if not any(s for s, _ in fs):
            fl = [f for _, f in fs]
            n = len(fl)
            if n == 0:
                return lambda env: []
            if n == 1:
                f0 = fl[0]
                return lambda env: [f0(env)]
            if n == 2:
                f0, f1 = fl
                return lambda env: [f0(env), f1(env)]
            return lambda env: [f(env) for f in fl]

        def ev(env):
            out = []
            for s, f in fs:
                if s:
                    out.extend(self.iterate(f(env)))
                else:
                    out.append(f(env))
            return out
        return ev

    def c_function(self, node):
        _, name, params, body, arrow, uses_args, expr_body = node
        cparams = [(p, self.c_expr(d) if d is not None else None, rest) for p, d, rest in params]
        if expr_body:
            cbody, hoist = self.c_expr(body), ([], [])
        else:
            stmts = body[1]
            names, funcs = collect_hoists(stmts)
            hoist = (names, [(n, self.c_function(f)) for n, f in funcs])
            cbody = self.c_block_body(stmts)
        interp = self

        def make(env):
            return JSFunction(interp, name, cparams, cbody, env, arrow, uses_args, hoist, expr_body)
        return make

    def c_class(self, node):
        _, name, members = node
        ctor = None
        ms = []
        for kind, static, mname, f in members:
            if mname == "constructor" and kind == "method":
                ctor = self.c_function(f)
            else:
                ms.append((kind, static, mname, self.c_function(f)))
        interp = self

        def make(env):
            if ctor is not None:
                cf = ctor(env)
            else:
                cf = JSFunction(interp, name, [], lambda e: None, env, False, False, ([], []), False)
            cf.name = name
            proto = cf.props["prototype"]
            for kind, static, mname, mk in ms:
                fn = mk(env)
                target = cf if static else proto
                target.props[mname] = Getter(fn) if kind == "get" else fn
            return cf
        return make

    def c_block_body(self, stmts):
        cs = [self.c_stmt(s) for s in stmts]
        if len(cs) == 1:
            return cs[0]

        def run(env):
            for s in cs:
                r = s(env)
                if r is not None:
                    return r
            return None
        return run

    def c_stmt(self, s):
        k = s[0]
        if k == "expr":
            e = self.c_expr(s[1])

            def run_expr(env):
                e(env)
            return run_expr
        if k == "var":
            kind = s[1]
            decls = [(n, self.c_expr(i) if i is not None else None) for n, i in s[2]]
            if any(isinstance(n, tuple) for n, _ in decls):
                get_prop = self.get_prop

                def run_destruct(env):
                    for n, i in decls:
                        if isinstance(n, tuple):
                            o = i(env)
                            for key, alias in n:
                                env.vars[alias] = get_prop(o, key)
                        else:
                            env.vars[n] = i(env) if i is not None else UNDEF
                return run_destruct
            if kind == "var":
                def run_var(env):
                    for n, i in decls:
                        if i is not None:
                            _assign_name(env, n, i(env))
                return run_var
            if len(decls) == 1:
                n0, i0 = decls[0]
                if i0 is None:
                    def run_let0(env):
                        env.vars[n0] = UNDEF
                    return run_let0

                def run_let1(env):
                    env.vars[n0] = i0(env)
                return run_let1

            def run_let(env):
                v = env.vars
                for n, i in decls:
                    v[n] = i(env) if i is not None else UNDEF
            return run_let
        if k == "block":
            body = self.c_block_body(s[1]) if s[1] else (lambda env: None)
            if has_lexical(s[1]):
                funcs = [(x[1], self.c_function(x[2])) for x in s[1] if x[0] == "funcdecl"]

                def run_block(env):
                    e2 = Scope(env)
                    for n, mk in funcs:
                        e2.vars[n] = mk(e2)
                    return body(e2)
                return run_block
            return body
        if k == "if":
            c, a = self.c_expr(s[1]), self.c_stmt(s[2])
            b = self.c_stmt(s[3]) if s[3] is not None else None
            if b is None:
                def run_if(env):
                    if truthy(c(env)):
                        return a(env)
                return run_if

            def run_ifelse(env):
                if truthy(c(env)):
                    return a(env)
                return b(env)
            return run_ifelse
        if k == "for":
            init = self.c_stmt(s[1]) if s[1] is not None else None
            test = self.c_expr(s[2]) if s[2] is not None else None
            upd = self.c_expr(s[3]) if s[3] is not None else None
            body = self.c_stmt(s[4])
            scoped = s[1] is not None and s[1][0] == "var" and s[1][1] != "var"

            def run_for(env):
                e2 = Scope(env) if scoped else env
                if init is not None:
                    init(e2)
                while test is None or truthy(test(e2)):
                    r = body(e2)
                    if r is not None:
                        if r is BREAK:
                            break
                        if r is not CONTINUE:
                            return r
                    if upd is not None:
                        upd(e2)
                return None
            return run_for
        if k in ("forin", "forof"):
            kind, name, obj, body = s[1], s[2], self.c_expr(s[3]), self.c_stmt(s[4])
            interp = self

            def run_forio(env):
                o = obj(env)
                items = interp.keys_of(o) if k == "forin" else interp.iterate(o)
                for it in items:
                    e2 = Scope(env) if kind != "var" else env
                    if kind != "var":
                        e2.vars[name] = it
                    else:
                        _assign_name(env, name, it)
                    r = body(e2)
                    if r is not None:
                        if r is BREAK:
                            break
                        if r is not CONTINUE:
                            return r
                return None
            return run_forio
        if k == "while":
            c, body = self.c_expr(s[1]), self.c_stmt(s[2])

            def run_while(env):
                while truthy(c(env)):
                    r = body(env)
                    if r is not None:
                        if r is BREAK:
                            break
                        if r is not CONTINUE:
                            return r
                return None
            return run_while
        if k == "dowhile":
            c, body = self.c_expr(s[1]), self.c_stmt(s[2])

            def run_dowhile(env):
                while True:
                    r = body(env)
                    if r is not None:
                        if r is BREAK:
                            break
                        if r is not CONTINUE:
                            return r
                    if not truthy(c(env)):
                        break
                return None
            return run_dowhile
        if k == "return":
            e = self.c_expr(s[1]) if s[1] is not None else None
            if e is None:
                return lambda env: Return(UNDEF)
            return lambda env: Return(e(env))
        if k == "break":
            return lambda env: BREAK
        if k == "continue":
            return lambda env: CONTINUE
        if k == "empty":
            return lambda env: None
        if k == "throw":
            e = self.c_expr(s[1])

            def run_throw(env):
                raise JSThrow(e(env))
            return run_throw
        if k == "try":
            b = self.c_stmt(s[1])
            param = s[2]
            h = self.c_stmt(s[3]) if s[3] is not None else None
            f = self.c_stmt(s[4]) if s[4] is not None else None

            def run_try(env):
                try:
                    try:
                        return b(env)
                    except JSThrow as ex:
                        if h is None:
                            raise
                        e2 = Scope(env)
                        if param:
                            e2.vars[param] = ex.value
                        return h(e2)
                finally:
                    if f is not None:
                        f(env)
            return run_try
        if k == "switch":
            d = self.c_expr(s[1])
            cases = [(self.c_expr(t) if t is not None else None, [self.c_stmt(x) for x in body]) for t, body in s[2]]

            def run_switch(env):
                v = d(env)
                e2 = Scope(env)
                start = None
                for i, (t, _) in enumerate(cases):
                    if t is not None and strict_eq(v, t(e2)):
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
                    for st in body:
                        r = st(e2)
                        if r is not None:
                            if r is BREAK:
                                return None
                            return r
                return None
            return run_switch
        if k == "funcdecl":
            return lambda env: None         # hoisted by the enclosing function / block / module
        if k == "classdecl":
            mk = self.c_class(s[2])
            name = s[1]

            def run_class(env):
                env.vars[name] = mk(env)
            return run_class
        if k == "import":
            src, default, names = s[1], s[2], s[3]
            interp = self

            def run_import(env):
                ex = interp.load_module(src, env.vars.get("__dir__", interp.root))
                if default:
                    env.vars[default] = ex.get("default", UNDEF)
                for a, b in names:
                    if a == "*":
                        o = JSObject(interp.object_proto)
                        o.props.update(ex)
                        env.vars[b] = o
                    else:
                        env.vars[b] = ex.get(a, UNDEF)
            return run_import
        if k == "export_default":
            node = s[1]
            e = self.c_class(node) if node[0] == "class" else self.c_expr(node)
            nm = node[1] if node[0] in ("class", "function") else None

            def run_exd(env):
                v = e(env)
                if nm:
                    env.vars[nm] = v
                _lookup(env, "__exports__")["default"] = v
            return run_exd
        if k == "export_names":
            names = s[1]

            def run_exn(env):
                ex = _lookup(env, "__exports__")
                for a, b in names:
                    ex[b] = _lookup(env, a)
            return run_exn
        if k == "export_decl":
            d = s[1]
            run = self.c_stmt(d)
            if d[0] == "var":
                names = [x for n, _ in d[2] for x in ([a for _k, a in n] if isinstance(n, tuple) else [n])]
            else:
                names = [d[1]]

            def run_exdecl(env):
                run(env)
                ex = _lookup(env, "__exports__")
                for n in names:
                    ex[n] = _lookup(env, n)
            return run_exdecl
        raise SyntaxError("jsmini: cannot compile statement %r" % (k,))

    def c_expr(self, e):
        k = e[0]
        interp = self
        get_prop, set_prop = self.get_prop, self.set_prop
        if k == "lit":
            v = e[1]
            return lambda env: v
        if k == "name":
            name = e[1]
            if name == "undefined":
                return lambda env: UNDEF

            def run_name(env):
                en = env
                while en is not None:
                    vs = en.vars
                    if name in vs:
                        return vs[name]
                    en = en.parent
                raise JSThrow("ReferenceError: %s is not defined" % name)
            return run_name
        if k == "member":
            o, key = self.c_expr(e[1]), e[2]

            def run_member(env):
                return get_prop(o(env), key)
            return run_member
        if k == "index":
            o, kx = self.c_expr(e[1]), self.c_expr(e[2])

            def run_index(env):
                return get_prop(o(env), kx(env))
            return run_index
        if k == "call":
            callee, args = e[1], self.c_args(e[2])
            if callee[0] in ("member", "index"):
                o = self.c_expr(callee[1])
                kx = (lambda env, kk=callee[2]: kk) if callee[0] == "member" else self.c_expr(callee[2])

                def run_mcall(env):
                    this = o(env)
                    fn = get_prop(this, kx(env))
                    a = args(env)
                    try:
                        return fn.call(this, a)
                    except AttributeError:
                        raise JSThrow("TypeError: %s is not a function" % js_to_string(kx(env)))
                return run_mcall
            f = self.c_expr(callee)

            def run_call(env):
                fn = f(env)
                a = args(env)
                try:
                    return fn.call(UNDEF, a)
                except AttributeError:
                    raise JSThrow("TypeError: %s is not a function" % js_to_string(fn))
            return run_call
        if k == "new":
            f, args = self.c_expr(e[1]), self.c_args(e[2])
            return lambda env: interp.construct(f(env), args(env))
        if k == "function":
            return self.c_function(e)
        if k == "class":
            return self.c_class(e)
        if k == "array":
            items = [(x[0] == "spread", self.c_expr(x[1] if x[0] == "spread" else x)) for x in e[1]]

            def run_array(env):
                out = []
                for sp, f in items:
                    if sp:
                        out.extend(interp.iterate(f(env)))
                    else:
                        out.append(f(env))
                return JSArray(interp, out)
            return run_array
        if k == "object":
            props = [(p[0], self.c_expr(p[1]) if p[1] is not None else None, self.c_expr(p[2])) for p in e[1]]

            def run_object(env):
                o = JSObject(interp.object_proto)
                for kind, kf, vf in props:
                    if kind == "spread":
                        src = vf(env)
                        if isinstance(src, JSObject):
                            o.props.update(src.props)
                    else:
                        key = prop_key(kf(env))
                        o.props[str(key) if type(key) is int else key] = vf(env)
                return o
            return run_object
        if k == "tmpl":
            parts = [p if isinstance(p, str) else self.c_expr(p) for p in e[1]]
            return lambda env: "".join(p if isinstance(p, str) else js_to_string(p(env)) for p in parts)
        if k == "seq":
            fs = [self.c_expr(x) for x in e[1]]

            def run_seq(env):
                r = UNDEF
                for f in fs:
                    r = f(env)
                return r
            return run_seq
        if k == "cond":
            c, a, b = self.c_expr(e[1]), self.c_expr(e[2]), self.c_expr(e[3])
            return lambda env: a(env) if truthy(c(env)) else b(env)
        if k == "logical":
            op, a, b = e[1], self.c_expr(e[2]), self.c_expr(e[3])
            if op == "&&":
                def run_and(env):
                    v = a(env)
                    return b(env) if truthy(v) else v
                return run_and
            if op == "||":
                def run_or(env):
                    v = a(env)
                    return v if truthy(v) else b(env)
                return run_or

            def run_nullish(env):
                v = a(env)
                return b(env) if (v is None or v is UNDEF) else v
            return run_nullish
        if k == "unary":
            op, a = e[1], None
            if op == "typeof" and e[2][0] == "name":
                nm = e[2][1]

                def run_typeof_name(env):
                    try:
                        return js_typeof(_lookup(env, nm)) if nm != "undefined" else "undefined"
                    except JSThrow:
                        return "undefined"
                return run_typeof_name
            if op == "delete":
                t = e[2]
                if t[0] in ("member", "index"):
                    o = self.c_expr(t[1])
                    kx = (lambda env, kk=t[2]: kk) if t[0] == "member" else self.c_expr(t[2])

                    def run_delete(env):
                        ob = o(env)
                        key = prop_key(kx(env))
                        if isinstance(ob, JSObject):
                            ob.props.pop(str(key) if type(key) is int else key, None)
                        return True
                    return run_delete
                return lambda env: True
            a = self.c_expr(e[2])
            if op == "!":
                return lambda env: not truthy(a(env))
            if op == "-":
                def run_neg(env):
                    v = a(env)
                    if type(v) is float:
                        return -v
                    v = to_number(v)
                    return -0.0 if (type(v) is int and v == 0) else -v
                return run_neg
            if op == "+":
                return lambda env: to_number(a(env))
            if op == "~":
                if e[2][0] == "unary" and e[2][1] == "~":        # ~~x : ToInt32(x)
                    inner = self.c_expr(e[2][2])
                    return lambda env: to_int32(inner(env))
                return lambda env: ~to_int32(a(env))
            if op == "typeof":
                return lambda env: js_typeof(a(env))
            if op == "void":
                def run_void(env):
                    a(env)
                    return UNDEF
                return run_void
        if k == "bin":
            return self.c_binary(e[1], self.c_expr(e[2]), self.c_expr(e[3]))
        if k == "update":
            op, prefix, target = e[1], e[2], e[3]
            d = 1 if op == "++" else -1
            if target[0] == "name":
                name = target[1]

                def run_upd_name(env):
                    en = env
                    while en is not None:
                        vs = en.vars
                        if name in vs:
                            old = vs[name]
                            if type(old) is not int and type(old) is not float:
                                old = to_number(old)
                            new = old + d
                            vs[name] = new
                            return new if prefix else old
                        en = en.parent
                    raise JSThrow("ReferenceError: %s is not defined" % name)
                return run_upd_name
            o = self.c_expr(target[1])
            kx = (lambda env, kk=target[2]: kk) if target[0] == "member" else self.c_expr(target[2])

            def run_upd_member(env):
                ob, key = o(env), kx(env)
                old = to_number(get_prop(ob, key))
                set_prop(ob, key, old + d)
                return old + d if prefix else old
            return run_upd_member
        if k == "assign":
            op, target, val = e[1], e[2], self.c_expr(e[3])
            binop = None if op == "=" else self.c_binary_fn(op[:-1])
            if target[0] == "name":
                name = target[1]
                if binop is None:
                    def run_assign_name(env):
                        v = val(env)
                        en = env
                        while en is not None:
                            vs = en.vars
                            if name in vs:
                                vs[name] = v
                                return v
                            last = en
                            en = en.parent
                        last.vars[name] = v
                        return v
                    return run_assign_name

                def run_cassign_name(env):
                    old = _lookup(env, name)
                    v = binop(old, val(env))
                    _assign_name(env, name, v)
                    return v
                return run_cassign_name
            o = self.c_expr(target[1])
            kx = (lambda env, kk=target[2]: kk) if target[0] == "member" else self.c_expr(target[2])
            if binop is None:
                def run_assign_member(env):
                    ob, key = o(env), kx(env)
                    v = val(env)
                    set_prop(ob, key, v)
                    return v
                return run_assign_member

            def run_cassign_member(env):
                ob, key = o(env), kx(env)
                v = binop(get_prop(ob, key), val(env))
                set_prop(ob, key, v)
                return v
            return run_cassign_member
        raise SyntaxError("jsmini: cannot compile expression %r" % (k,))

    def c_binary_fn(self, op):
        interp = self
        if op == "+":
            return js_add
        if op == "-":
            return lambda a, b: to_number(a) - to_number(b)
        if op == "*":
            return lambda a, b: to_number(a) * to_number(b)
        if op == "/":
            return js_div
        if op == "%":
            return js_mod
        if op == "**":
            return js_pow
        if op == "&":
            return lambda a, b: to_int32(to_int32(a) & to_int32(b))
        if op == "|":
            return lambda a, b: to_int32(to_int32(a) | to_int32(b))
        if op == "^":
            return lambda a, b: to_int32(to_int32(a) ^ to_int32(b))
        if op == "<<":
            return lambda a, b: to_int32(to_int32(a) << (to_uint32(b) & 31))
        if op == ">>":
            return lambda a, b: to_int32(a) >> (to_uint32(b) & 31)
        if op == ">>>":
            return lambda a, b: to_uint32(a) >> (to_uint32(b) & 31)
        if op == "===":
            return strict_eq
        if op == "!==":
            return lambda a, b: not strict_eq(a, b)
        if op == "==":
            return loose_eq
        if op == "!=":
            return lambda a, b: not loose_eq(a, b)
        if op in ("<", ">", "<=", ">="):
            return lambda a, b: js_compare(op, a, b)
        if op == "instanceof":
            def inst(a, b):
                if isinstance(b, NativeFunction):
                    return interp.native_instanceof(a, b)
                if not isinstance(a, JSObject) or not isinstance(b, JSObject):
                    return False
                p = b.props.get("prototype")
                o = a.proto
                while o is not None:
                    if o is p:
                        return True
                    o = o.proto
                return False
            return inst
        if op == "in":
            def has(a, b):
                key = prop_key(a)
                if isinstance(b, JSArray) and type(key) is int:
                    return key < len(b.list)
                o = b
                key = str(key) if type(key) is int else key
                while isinstance(o, JSObject):
                    if key in o.props:
                        return True
                    o = o.proto
                return False
            return has
        raise SyntaxError("jsmini: operator %s" % op)

    def c_binary(self, op, a, b):
        # fast paths for plain numbers, slow paths through the generic operator
        slow = self.c_binary_fn(op)
        if op == "+":
            def run_add(env):
                x, y = a(env), b(env)
                tx, ty = type(x), type(y)
                if (tx is float or tx is int) and (ty is float or ty is int):
                    return x + y
                return slow(x, y)
            return run_add
        if op == "-":
            def run_sub(env):
                x, y = a(env), b(env)
                tx, ty = type(x), type(y)
                if (tx is float or tx is int) and (ty is float or ty is int):
                    return x - y
                return slow(x, y)
            return run_sub
        if op == "*":
            def run_mul(env):
                x, y = a(env), b(env)
                tx, ty = type(x), type(y)
                if (tx is float or tx is int) and (ty is float or ty is int):
                    return x * y
                return slow(x, y)
            return run_mul
        if op in ("<", ">", "<=", ">="):
            import operator
            pyop = {"<": operator.lt, ">": operator.gt, "<=": operator.le, ">=": operator.ge}[op]

            def run_cmp(env):
                x, y = a(env), b(env)
                tx, ty = type(x), type(y)
                if (tx is float or tx is int) and (ty is float or ty is int):
                    return pyop(x, y)
                return slow(x, y)
            return run_cmp
        return lambda env: slow(a(env), b(env))

    # ---- modules
    def run_source(self, src, fname="<js>", dirname=None, scope=None):
        ast = Parser(tokenize(src), fname).program()
        env = scope or Scope(self.globals)
        env.vars.setdefault("__exports__", {})
        env.vars["__dir__"] = dirname or self.root
        names, funcs = collect_hoists(ast)
        for n in names:
            env.vars.setdefault(n, UNDEF)
        for n, f in funcs:
            env.vars[n] = self.c_function(f)(env)
        body = self.c_block_body(ast) if ast else (lambda e: None)
        body(env)
        return env

    def load_module(self, spec, base):
        path = os.path.normpath(os.path.join(base, spec))
        if not os.path.exists(path) and os.path.exists(path + ".js"):
            path += ".js"
        if path in self.modules:
            return self.modules[path]
        ex = {}
        self.modules[path] = ex
        env = Scope(self.globals)
        env.vars["__exports__"] = ex
        with open(path) as f:
            self.run_source(f.read(), path, os.path.dirname(path), env)
        return ex

    def drain(self):
        """Run queued promise reactions (microtasks) until none is left."""
        n = 0
        while self.jobs:
            fn, v = self.jobs.pop(0)
            fn(v)
            n += 1
        return n

    def promise_state(self, p):
        return p.state, p.value

    def require_file(self, path, require=None):
        """CommonJS module: `require`, `module`, `exports` in scope; returns module.exports."""
        path = os.path.normpath(path)
        if path in self.modules:
            return self.modules[path]["exports"]
        mod = JSObject(self.object_proto)
        mod.props["exports"] = JSObject(self.object_proto)
        self.modules[path] = mod.props
        env = Scope(self.globals)
        env.vars["module"] = mod
        env.vars["exports"] = mod.props["exports"]
        base = os.path.dirname(path)

        def req(this, a):
            spec = js_to_string(a[0])
            if require is not None:
                r = require(spec)
                if r is not None:
                    return r
            p2 = os.path.join(base, spec)
            return self.require_file(p2 if os.path.exists(p2) else p2 + ".js", require)
        env.vars["require"] = NativeFunction(self, "require", req)
        with open(path) as f:
            self.run_source(f.read(), path, base, env)
        return mod.props["exports"]

    def native_instanceof(self, a, b):
        nm = b.name
        if nm == "Array":
            return isinstance(a, JSArray)
        if nm == "ArrayBuffer":
            return isinstance(a, JSArrayBuffer)
        if nm in TYPED:
            return isinstance(a, JSTypedArray) and a.kind == nm
        if nm == "Object":
            return isinstance(a, JSObject)
        if nm == "Function":
            return isinstance(a, (JSFunction, NativeFunction))
        if nm == "Promise":
            return isinstance(a, self.JSPromise)
        return False

    # ---- host helpers
    def to_py(self, v):
        """JS value -> plain Python (lists / dicts / numpy arrays) for the fixture writer."""
        if isinstance(v, JSArray):
            return [self.to_py(x) for x in v.list]
        if isinstance(v, JSTypedArray):
            return v.arr.copy()
        if isinstance(v, JSArrayBuffer):
            return bytes(v.data)
        if isinstance(v, (JSFunction, NativeFunction)):
            return v
        if isinstance(v, JSObject):
            return {k: self.to_py(x) for k, x in v.props.items()}
        if v is UNDEF:
            return None
        return v

    def from_py(self, v):
        if isinstance(v, (bytes, bytearray)):
            return JSArrayBuffer(self, bytearray(v))
        if isinstance(v, np.ndarray):
            return JSArray(self, [self.from_py(x) for x in v.tolist()])
        if isinstance(v, (list, tuple)):
            return JSArray(self, [self.from_py(x) for x in v])
        if isinstance(v, dict):
            o = JSObject(self.object_proto)
            for k, x in v.items():
                o.props[k] = self.from_py(x)
            return o
        if isinstance(v, (np.integer,)):
            return int(v)
        if isinstance(v, (np.floating,)):
            return float(v)
        return v


# ----------------------------------------------------------------------------------------------- builtins
def _install_builtins(I):
    G = I.globals.vars

    def native(name, fn, construct=None):
        return NativeFunction(I, name, fn, construct)

    def arg(a, i, d=UNDEF):
        return a[i] if i < len(a) else d

    # ---- Math
    M = JSObject(I.object_proto)

    def m1(f):
        def g(this, a):
            x = to_number(arg(a, 0))
            try:
                return f(x)
            except (ValueError, OverflowError):
                return math.nan
        return g

    def log_like(f):
        def g(x):
            if x != x:
                return math.nan
            if x == 0:
                return -math.inf
            if x < 0:
                return math.nan
            if x == math.inf:
                return math.inf
            return f(x)
        return g

    def js_round(x):
        if x != x or x in (math.inf, -math.inf):
            return x
        r = math.floor(x)                 # nearest integer, ties toward +Infinity; x - floor(x) is exact
        return r + 1 if x - r >= 0.5 else r

    def js_minmax(is_max):
        def g(this, a):
            r = -math.inf if is_max else math.inf
            for x in a:
                x = to_number(x)
                if x != x:
                    return math.nan
                if (x > r) if is_max else (x < r):
                    r = x
            return r
        return g

    def trig(f):
        def g(x):
            if x != x or x in (math.inf, -math.inf):
                return math.nan
            return f(x)
        return g
    M.props.update({
        "PI": math.pi, "E": math.e, "LN2": math.log(2), "LN10": math.log(10), "LOG2E": 1 / math.log(2), "LOG10E": 1 / math.log(10),
        "SQRT2": math.sqrt(2), "SQRT1_2": math.sqrt(0.5),
        "cos": native("cos", m1(trig(math.cos))), "sin": native("sin", m1(trig(math.sin))), "tan": native("tan", m1(trig(math.tan))),
        "atan": native("atan", m1(math.atan)), "exp": native("exp", m1(lambda x: math.exp(x) if x < 709.78 else math.inf)),
        "atan2": native("atan2", lambda t, a: math.atan2(to_number(arg(a, 0)), to_number(arg(a, 1)))),
        "log": native("log", m1(log_like(math.log))), "log10": native("log10", m1(log_like(math.log10))),
        "log2": native("log2", m1(log_like(math.log2))), "sqrt": native("sqrt", m1(lambda x: math.sqrt(x) if x >= 0 else math.nan)),
        "abs": native("abs", m1(abs)), "floor": native("floor", m1(lambda x: x if (x != x or x in (math.inf, -math.inf)) else math.floor(x))),
        "ceil": native("ceil", m1(lambda x: x if (x != x or x in (math.inf, -math.inf)) else math.ceil(x))),
        "trunc": native("trunc", m1(lambda x: x if (x != x or x in (math.inf, -math.inf)) else math.trunc(x))),
        "round": native("round", m1(js_round)), "sign": native("sign", m1(lambda x: x if x != x else (x > 0) - (x < 0))),
        "max": native("max", js_minmax(True)), "min": native("min", js_minmax(False)),
        "pow": native("pow", lambda t, a: js_pow(arg(a, 0), arg(a, 1))),
        "hypot": native("hypot", lambda t, a: math.hypot(*[to_number(x) for x in a])),
        "random": native("random", lambda t, a: 0.5),
    })
    G["Math"] = M
    G["NaN"], G["Infinity"] = math.nan, math.inf
    G["isNaN"] = native("isNaN", lambda t, a: to_number(arg(a, 0)) != to_number(arg(a, 0)))
    G["isFinite"] = native("isFinite", lambda t, a: math.isfinite(to_number(arg(a, 0))))

    def parse_int(t, a):
        s = js_to_string(arg(a, 0)).strip()
        radix = to_int32(arg(a, 1, 10)) or 10
        sign = 1
        if s[:1] in "+-":
            sign = -1 if s[0] == "-" else 1
            s = s[1:]
        if radix == 16 and s[:2].lower() == "0x":
            s = s[2:]
        elif s[:2].lower() == "0x" and len(a) < 2:
            radix, s = 16, s[2:]
        digs = "0123456789abcdefghijklmnopqrstuvwxyz"[:radix]
        j = 0
        while j < len(s) and s[j].lower() in digs:
            j += 1
        if j == 0:
            return math.nan
        return sign * int(s[:j], radix)

    def parse_float(t, a):
        s = js_to_string(arg(a, 0)).strip()
        j, seen_e, seen_dot = 0, False, False
        while j < len(s):
            ch = s[j]
            if ch.isdigit():
                pass
            elif ch in "+-" and (j == 0 or s[j - 1] in "eE"):
                pass
            elif ch == "." and not seen_dot and not seen_e:
                seen_dot = True
            elif ch in "eE" and not seen_e and j > 0:
                seen_e = True
            else:
                break
            j += 1
        while j > 0:
            try:
                f = float(s[:j])
                return int(f) if f.is_integer() and abs(f) < 2 ** 53 else f
            except ValueError:
                j -= 1
        return math.nan
    G["parseInt"], G["parseFloat"] = native("parseInt", parse_int), native("parseFloat", parse_float)

    # ---- Object
    def obj_ctor(t, a):
        v = arg(a, 0)
        return v if isinstance(v, JSObject) else JSObject(I.object_proto)
    O = native("Object", obj_ctor, lambda a: obj_ctor(None, a))
    O.props["prototype"] = I.object_proto
    O.props["keys"] = native("keys", lambda t, a: JSArray(I, list(I.keys_of(arg(a, 0)))))
    O.props["values"] = native("values", lambda t, a: JSArray(I, [I.get_prop(arg(a, 0), k) for k in I.keys_of(arg(a, 0))]))
    O.props["entries"] = native("entries", lambda t, a: JSArray(I, [JSArray(I, [k, I.get_prop(arg(a, 0), k)]) for k in I.keys_of(arg(a, 0))]))

    def obj_assign(t, a):
        tgt = a[0]
        for s in a[1:]:
            if isinstance(s, JSObject):
                for k in I.keys_of(s):
                    I.set_prop(tgt, k, I.get_prop(s, k))
        return tgt
    O.props["assign"] = native("assign", obj_assign)

    def define_property(t, a):
        o, k, d = a[0], a[1], a[2]
        v = I.get_prop(d, "value")
        g = I.get_prop(d, "get")
        o.props[js_to_string(k)] = Getter(g) if g is not UNDEF else v
        return o
    O.props["defineProperty"] = native("defineProperty", define_property)
    O.props["create"] = native("create", lambda t, a: JSObject(arg(a, 0) if isinstance(arg(a, 0), JSObject) else None))
    O.props["freeze"] = native("freeze", lambda t, a: arg(a, 0))
    G["Object"] = O
    I.object_proto.props["hasOwnProperty"] = native("hasOwnProperty", lambda t, a: (
        (type(prop_key(arg(a, 0))) is int and isinstance(t, JSArray) and prop_key(arg(a, 0)) < len(t.list))
        or (str(prop_key(arg(a, 0))) in t.props)))
    I.object_proto.props["toString"] = native("toString", lambda t, a: js_to_string(t))

    # ---- Function.prototype
    FP = I.function_proto
    FP.props["call"] = native("call", lambda t, a: I.call(t, arg(a, 0), list(a[1:])))
    FP.props["apply"] = native("apply", lambda t, a: I.call(t, arg(a, 0), I.iterate(arg(a, 1)) if len(a) > 1 and arg(a, 1) not in (None, UNDEF) else []))

    def fn_bind(t, a):
        this, pre = arg(a, 0), list(a[1:])
        return native("bound", lambda t2, a2: I.call(t, this, pre + list(a2)))
    FP.props["bind"] = native("bind", fn_bind)

    # ---- Array
    def array_ctor(a):
        if len(a) == 1 and isinstance(a[0], (int, float)) and not isinstance(a[0], bool):
            n = a[0]
            if n < 0 or n != int(n):
                raise JSThrow("RangeError: Invalid array length")
            return JSArray(I, [UNDEF] * int(n))
        return JSArray(I, list(a))
    A = native("Array", lambda t, a: array_ctor(a), array_ctor)
    A.props["prototype"] = I.array_proto
    A.props["isArray"] = native("isArray", lambda t, a: isinstance(arg(a, 0), JSArray))

    def array_from(t, a):
        src, fn = arg(a, 0), arg(a, 1)
        if isinstance(src, JSObject) and not isinstance(src, (JSArray, JSTypedArray)):
            n = int(to_number(I.get_prop(src, "length")) or 0)
            items = [I.get_prop(src, i) for i in range(n)]
        else:
            items = I.iterate(src)
        if fn is not UNDEF:
            items = [I.call(fn, UNDEF, [x, i]) for i, x in enumerate(items)]
        return JSArray(I, items)
    A.props["from"] = native("from", array_from)
    G["Array"] = A
    AP = I.array_proto.props

    def seq_of(t):
        return t.list if isinstance(t, JSArray) else I.iterate(t)

    def a_fill(t, a):
        v = arg(a, 0)
        n = len(t.list) if isinstance(t, JSArray) else len(t.arr)
        s = to_int32(arg(a, 1, 0))
        e = n if arg(a, 2) is UNDEF else to_int32(arg(a, 2))
        s = max(n + s, 0) if s < 0 else min(s, n)
        e = max(n + e, 0) if e < 0 else min(e, n)
        for i in range(s, e):
            I.set_prop(t, i, v)
        return t

    def a_push(t, a):
        t.list.extend(a)
        return len(t.list)

    def a_slice(t, a):
        lst = seq_of(t)
        n = len(lst)
        s = to_int32(arg(a, 0, 0))
        e = n if arg(a, 1) is UNDEF else to_int32(arg(a, 1))
        s = max(n + s, 0) if s < 0 else min(s, n)
        e = max(n + e, 0) if e < 0 else min(e, n)
        return JSArray(I, list(lst[s:e]))

    def a_reduce(t, a):
        fn = a[0]
        lst = seq_of(t)
        i = 0
        if len(a) > 1:
            acc = a[1]
        else:
            acc, i = lst[0], 1
        while i < len(lst):
            acc = I.call(fn, UNDEF, [acc, lst[i], i, t])
            i += 1
        return acc

    def a_index_of(t, a):
        for i, x in enumerate(seq_of(t)):
            if strict_eq(x, arg(a, 0)):
                return i
        return -1

    def a_sort(t, a):
        import functools
        fn = arg(a, 0)
        if fn is UNDEF:
            t.list.sort(key=js_to_string)
        else:
            t.list.sort(key=functools.cmp_to_key(lambda x, y: (lambda r: -1 if r < 0 else (1 if r > 0 else 0))(to_number(I.call(fn, UNDEF, [x, y])))))
        return t
    AP.update({
        "fill": native("fill", a_fill), "push": native("push", a_push),
        "pop": native("pop", lambda t, a: t.list.pop() if t.list else UNDEF),
        "shift": native("shift", lambda t, a: t.list.pop(0) if t.list else UNDEF),
        "unshift": native("unshift", lambda t, a: (t.list.__setitem__(slice(0, 0), list(a)), len(t.list))[1]),
        "slice": native("slice", a_slice),
        "map": native("map", lambda t, a: JSArray(I, [I.call(a[0], arg(a, 1), [x, i, t]) for i, x in enumerate(list(seq_of(t)))])),
        "forEach": native("forEach", lambda t, a: ([I.call(a[0], arg(a, 1), [x, i, t]) for i, x in enumerate(list(seq_of(t)))], UNDEF)[1]),
        "filter": native("filter", lambda t, a: JSArray(I, [x for i, x in enumerate(list(seq_of(t))) if truthy(I.call(a[0], arg(a, 1), [x, i, t]))])),
        "some": native("some", lambda t, a: any(truthy(I.call(a[0], arg(a, 1), [x, i, t])) for i, x in enumerate(list(seq_of(t))))),
        "every": native("every", lambda t, a: all(truthy(I.call(a[0], arg(a, 1), [x, i, t])) for i, x in enumerate(list(seq_of(t))))),
        "find": native("find", lambda t, a: next((x for i, x in enumerate(list(seq_of(t))) if truthy(I.call(a[0], arg(a, 1), [x, i, t]))), UNDEF)),
        "reduce": native("reduce", a_reduce), "indexOf": native("indexOf", a_index_of),
        "includes": native("includes", lambda t, a: a_index_of(t, a) >= 0),
        "join": native("join", lambda t, a: (js_to_string(arg(a, 0)) if arg(a, 0) is not UNDEF else ",").join(
            "" if (x is UNDEF or x is None) else js_to_string(x) for x in seq_of(t))),
        "concat": native("concat", lambda t, a: JSArray(I, list(t.list) + [y for x in a for y in (x.list if isinstance(x, JSArray) else [x])])),
        "reverse": native("reverse", lambda t, a: (t.list.reverse(), t)[1]),
        "sort": native("sort", a_sort),
    })

    # ---- ArrayBuffer and typed arrays
    def ab_ctor(a):
        return JSArrayBuffer(I, bytearray(int(to_number(arg(a, 0, 0)))))
    AB = native("ArrayBuffer", lambda t, a: ab_ctor(a), ab_ctor)
    G["ArrayBuffer"] = AB

    def ab_slice(t, a):
        n = len(t.data)
        s = to_number(arg(a, 0, 0))
        e = n if arg(a, 1) is UNDEF else to_number(arg(a, 1))
        s = 0 if s != s else int(s)
        e = 0 if e != e else int(e)
        s = max(n + s, 0) if s < 0 else min(s, n)
        e = max(n + e, 0) if e < 0 else min(e, n)
        return JSArrayBuffer(I, bytearray(t.data[s:max(s, e)]))
    I.arraybuffer_slice = native("slice", ab_slice)

    def typed_ctor(kind):
        dt = np.dtype(TYPED[kind]).newbyteorder("<")

        def make(a):
            x = arg(a, 0, 0)
            if isinstance(x, JSArrayBuffer):
                off = int(to_number(arg(a, 1, 0)))
                avail = len(x.data) - off
                if arg(a, 2) is UNDEF:
                    if avail % dt.itemsize:
                        raise JSThrow("RangeError: byte length of %s should be a multiple of %d" % (kind, dt.itemsize))
                    cnt = avail // dt.itemsize
                else:
                    cnt = int(to_number(a[2]))
                arr = np.frombuffer(x.data, dtype=dt, count=cnt, offset=off)
                return JSTypedArray(I, kind, x, arr)
            if isinstance(x, (JSArray, JSTypedArray)):
                items = I.iterate(x)
                buf = JSArrayBuffer(I, bytearray(len(items) * dt.itemsize))
                ta = JSTypedArray(I, kind, buf, np.frombuffer(buf.data, dtype=dt))
                for i, v in enumerate(items):
                    I.set_prop(ta, i, v)
                return ta
            n = int(to_number(x))
            buf = JSArrayBuffer(I, bytearray(n * dt.itemsize))
            return JSTypedArray(I, kind, buf, np.frombuffer(buf.data, dtype=dt))
        f = native(kind, lambda t, a: make(a), make)
        f.props["BYTES_PER_ELEMENT"] = dt.itemsize
        f.props["from"] = native("from", lambda t, a: make([array_from(None, a)]))
        return f
    for kind in TYPED:
        G[kind] = typed_ctor(kind)

    def ta_subarray(t, a):
        n = len(t.arr)
        s = to_int32(arg(a, 0, 0))
        e = n if arg(a, 1) is UNDEF else to_int32(arg(a, 1))
        s = max(n + s, 0) if s < 0 else min(s, n)
        e = max(n + e, 0) if e < 0 else min(e, n)
        return JSTypedArray(I, t.kind, t.buffer, t.arr[s:e])

    def ta_set(t, a):
        if isinstance(a[0], JSTypedArray) and a[0].kind == t.kind:
            off = int(to_number(arg(a, 1, 0)))
            if off + len(a[0].arr) > len(t.arr):
                raise JSThrow("RangeError: offset is out of bounds")
            t.arr[off:off + len(a[0].arr)] = a[0].arr
            return UNDEF
        src, off = I.iterate(a[0]), int(to_number(arg(a, 1, 0)))
        for i, v in enumerate(src):
            I.set_prop(t, off + i, v)
        return UNDEF
    I.typed_methods = {"fill": AP["fill"], "subarray": native("subarray", ta_subarray), "set": native("set", ta_set),
                       "slice": native("slice", lambda t, a: G[t.kind].construct([a_slice(t, a)])),
                       "forEach": AP["forEach"], "map": AP["map"], "reduce": AP["reduce"], "join": AP["join"], "indexOf": AP["indexOf"]}

    # ---- String / Number methods
    def s_method(f):
        return native(f.__name__, lambda t, a: f(js_to_string(t), a))

    def s_slice(s, a):
        n = len(s)
        b = to_int32(arg(a, 0, 0))
        e = n if arg(a, 1) is UNDEF else to_int32(arg(a, 1))
        b = max(n + b, 0) if b < 0 else min(b, n)
        e = max(n + e, 0) if e < 0 else min(e, n)
        return s[b:e]

    def s_substring(s, a):
        n = len(s)
        b = min(max(to_int32(arg(a, 0, 0)), 0), n)
        e = n if arg(a, 1) is UNDEF else min(max(to_int32(arg(a, 1)), 0), n)
        if b > e:
            b, e = e, b
        return s[b:e]

    def s_split(s, a):
        sep = arg(a, 0)
        if sep is UNDEF:
            return JSArray(I, [s])
        sep = js_to_string(sep)
        return JSArray(I, list(s) if sep == "" else s.split(sep))

    def s_char_code(s, a):
        i = to_int32(arg(a, 0, 0))
        return ord(s[i]) if 0 <= i < len(s) else math.nan
    I.string_methods = {
        "toUpperCase": s_method(lambda s, a: s.upper()), "toLowerCase": s_method(lambda s, a: s.lower()),
        "trim": s_method(lambda s, a: s.strip()), "slice": s_method(s_slice), "substring": s_method(s_substring),
        "substr": s_method(lambda s, a: s_slice(s, [arg(a, 0, 0), UNDEF if arg(a, 1) is UNDEF else to_int32(arg(a, 0, 0)) + to_int32(arg(a, 1))])),
        "indexOf": s_method(lambda s, a: s.find(js_to_string(arg(a, 0)), max(to_int32(arg(a, 1, 0)), 0))),
        "lastIndexOf": s_method(lambda s, a: s.rfind(js_to_string(arg(a, 0)))),
        "startsWith": s_method(lambda s, a: s.startswith(js_to_string(arg(a, 0)), to_int32(arg(a, 1, 0)))),
        "endsWith": s_method(lambda s, a: s.endswith(js_to_string(arg(a, 0)))),
        "includes": s_method(lambda s, a: js_to_string(arg(a, 0)) in s),
        "charAt": s_method(lambda s, a: s[to_int32(arg(a, 0, 0))] if 0 <= to_int32(arg(a, 0, 0)) < len(s) else ""),
        "charCodeAt": s_method(s_char_code), "split": s_method(s_split),
        "toString": s_method(lambda s, a: s), "concat": s_method(lambda s, a: s + "".join(js_to_string(x) for x in a)),
        "repeat": s_method(lambda s, a: s * max(to_int32(arg(a, 0, 0)), 0)),
        "padStart": s_method(lambda s, a: s.rjust(to_int32(arg(a, 0, 0)), (js_to_string(arg(a, 1, " ")) or " ")[0])),
    }

    def n_to_fixed(t, a):
        d = to_int32(arg(a, 0, 0))
        x = float(t)
        if x != x:
            return "NaN"
        from decimal import Decimal, ROUND_HALF_UP
        q = Decimal(1).scaleb(-d)
        return str(Decimal(x).quantize(q, rounding=ROUND_HALF_UP))

    def n_to_string(t, a):
        r = arg(a, 0)
        if r is UNDEF or to_int32(r) == 10:
            return num_to_str(t)
        r, n = to_int32(r), int(t)
        digs, out, neg = "0123456789abcdefghijklmnopqrstuvwxyz", "", n < 0
        n = abs(n)
        while True:
            out = digs[n % r] + out
            n //= r
            if not n:
                break
        return ("-" if neg else "") + out
    I.number_methods = {"toFixed": native("toFixed", n_to_fixed), "toString": native("toString", n_to_string),
                        "toPrecision": native("toPrecision", lambda t, a: "%.*g" % (to_int32(arg(a, 0, 6)), float(t)))}

    def number_fn(t, a):
        return to_number(arg(a, 0, 0))
    N = native("Number", number_fn)
    N.props.update({"MAX_SAFE_INTEGER": 2 ** 53 - 1, "EPSILON": 2.0 ** -52, "MAX_VALUE": 1.7976931348623157e308,
                    "isInteger": native("isInteger", lambda t, a: isinstance(arg(a, 0), (int, float)) and not isinstance(arg(a, 0), bool) and float(arg(a, 0)).is_integer()),
                    "isFinite": native("isFinite", lambda t, a: isinstance(arg(a, 0), (int, float)) and math.isfinite(arg(a, 0))),
                    "isNaN": native("isNaN", lambda t, a: isinstance(arg(a, 0), float) and arg(a, 0) != arg(a, 0)),
                    "parseFloat": G["parseFloat"], "parseInt": G["parseInt"]})
    G["Number"] = N
    G["String"] = native("String", lambda t, a: js_to_string(arg(a, 0, "")))
    G["Boolean"] = native("Boolean", lambda t, a: truthy(arg(a, 0)))

    # ---- errors, Promise stub, console
    def err_ctor(name):
        def make(a):
            o = JSObject(I.object_proto)
            o.props["name"], o.props["message"] = name, js_to_string(arg(a, 0, ""))
            return o
        return native(name, lambda t, a: make(a), make)
    for nm in ("Error", "TypeError", "RangeError", "ReferenceError", "SyntaxError"):
        G[nm] = err_ctor(nm)

    # Promise: settle + reaction jobs on a microtask queue that the host drains (Interp.drain) after each entry into JS
    class JSPromise(JSObject):
        __slots__ = ("state", "value", "reactions")

        def __init__(self):
            JSObject.__init__(self, promise_proto)
            self.state, self.value, self.reactions = "pending", UNDEF, []
    promise_proto = JSObject(I.object_proto)
    I.jobs = []

    def settle(p, state, v):
        if p.state != "pending":
            return
        if state == "fulfilled" and isinstance(v, JSPromise):
            subscribe(v, lambda x: settle(p, "fulfilled", x), lambda x: settle(p, "rejected", x))
            return
        p.state, p.value = state, v
        for on_f, on_r in p.reactions:
            I.jobs.append((on_f if state == "fulfilled" else on_r, v))
        p.reactions = []

    def subscribe(p, on_f, on_r):
        if p.state == "pending":
            p.reactions.append((on_f, on_r))
        else:
            I.jobs.append((on_f if p.state == "fulfilled" else on_r, p.value))

    def promise_then(t, a):
        on_f, on_r = arg(a, 0), arg(a, 1)
        p2 = JSPromise()

        def run(handler, fallback_state):
            def job(v):
                if isinstance(handler, (JSFunction, NativeFunction)):
                    try:
                        settle(p2, "fulfilled", handler.call(UNDEF, [v]))
                    except JSThrow as ex:
                        settle(p2, "rejected", ex.value)
                else:
                    settle(p2, fallback_state, v)
            return job
        subscribe(t, run(on_f, "fulfilled"), run(on_r, "rejected"))
        return p2
    promise_proto.props["then"] = native("then", promise_then)
    promise_proto.props["catch"] = native("catch", lambda t, a: promise_then(t, [UNDEF, arg(a, 0)]))

    def promise_construct(a):
        p = JSPromise()
        res = native("resolve", lambda t, x: (settle(p, "fulfilled", arg(x, 0)), UNDEF)[1])
        rej = native("reject", lambda t, x: (settle(p, "rejected", arg(x, 0)), UNDEF)[1])
        try:
            I.call(a[0], UNDEF, [res, rej])
        except JSThrow as ex:
            settle(p, "rejected", ex.value)
        return p

    def promise_resolve(t, a):
        v = arg(a, 0)
        if isinstance(v, JSPromise):
            return v
        p = JSPromise()
        settle(p, "fulfilled", v)
        return p

    def promise_reject(t, a):
        p = JSPromise()
        settle(p, "rejected", arg(a, 0))
        return p

    def promise_all(t, a):
        items = I.iterate(a[0])
        p = JSPromise()
        out, left = [UNDEF] * len(items), [len(items)]
        if not items:
            settle(p, "fulfilled", JSArray(I, []))
        for i, it in enumerate(items):
            def on_f(v, i=i):
                out[i] = v
                left[0] -= 1
                if left[0] == 0:
                    settle(p, "fulfilled", JSArray(I, out))
            subscribe(promise_resolve(None, [it]), on_f, lambda v: settle(p, "rejected", v))
        return p
    P = native("Promise", lambda t, a: UNDEF, promise_construct)
    P.props["prototype"] = promise_proto
    P.props["resolve"] = native("resolve", promise_resolve)
    P.props["reject"] = native("reject", promise_reject)
    P.props["all"] = native("all", promise_all)
    G["Promise"] = P
    I.JSPromise = JSPromise
    con = JSObject(I.object_proto)
    for nm in ("log", "warn", "error", "info", "debug", "time", "timeEnd"):
        con.props[nm] = native(nm, lambda t, a: (I.log.append(" ".join(js_to_string(x) for x in a)), UNDEF)[1])
    G["console"] = con
    G["undefined"] = UNDEF
    gobj = JSObject(I.object_proto)
    G["globalThis"] = gobj
