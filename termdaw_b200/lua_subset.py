"""A reader for the subset of Lua that termdaw project scripts use.

termdaw executes the project file with mlua / Lua 5.3 (`state.rs:83-159`); every registered global
only appends a tuple to a seed vector, so a project script is, in practice, a list of calls with
literal arguments plus a few global variables (`project.lua`, `examples/*.lua`).  No Lua interpreter
exists in this image, so this module evaluates exactly that subset:

    statements   name = expr | local name = expr | f(args) | for i = a, b [, step] do ... end
    expressions  numbers, "strings" / 'strings', true / false / nil, { e, e, ... } (array tables),
                 names, unary -, binary + - * / % .. with Lua precedence, parentheses
    comments     -- line   and   --[[ block ]]         optional ';' after a statement

Anything else raises `LuaSubsetError` (the reference would print the mlua error and leave the
project unloaded, `state.rs:160-164`).  Calls are dispatched to a `Project`
(termdaw_b200.state.Project), whose methods are the registered globals under their real names.
"""
from __future__ import annotations

import re
from typing import Any, Dict, List

from .state import Project


class LuaSubsetError(ValueError):
    pass


_TOKEN = re.compile(r"""
    (?P<ws>\s+)
  | (?P<blockcomment>--\[\[.*?\]\])
  | (?P<comment>--[^\n]*)
  | (?P<number>0[xX][0-9a-fA-F]+|(?:\d+\.?\d*|\.\d+)(?:[eE][+-]?\d+)?)
  | (?P<string>"(?:\\.|[^"\\])*"|'(?:\\.|[^'\\])*')
  | (?P<name>[A-Za-z_][A-Za-z_0-9]*)
  | (?P<op>\.\.|[=(){},;+\-*/%])
""", re.X | re.S)

KEYWORDS = {"local", "for", "do", "end", "true", "false", "nil"}
_ESC = {"n": "\n", "t": "\t", "\\": "\\", '"': '"', "'": "'", "r": "\r", "0": "\0"}


def tokenize(text: str):
    pos, out = 0, []
    while pos < len(text):
        m = _TOKEN.match(text, pos)
        if not m:
            line = text.count("\n", 0, pos) + 1
            raise LuaSubsetError(f"line {line}: unexpected character {text[pos]!r}")
        pos = m.end()
        kind = m.lastgroup
        if kind in ("ws", "comment", "blockcomment"):
            continue
        val = m.group(kind)
        line = text.count("\n", 0, m.start()) + 1
        if kind == "number":
            out.append(("num", int(val, 16) if val[:2] in ("0x", "0X") else
                        (int(val) if re.fullmatch(r"\d+", val) else float(val)), line))
        elif kind == "string":
            body = re.sub(r"\\(.)", lambda e: _ESC.get(e.group(1), e.group(1)), val[1:-1])
            out.append(("str", body, line))
        elif kind == "name":
            out.append(("kw" if val in KEYWORDS else "name", val, line))
        else:
            out.append(("op", val, line))
    out.append(("eof", None, text.count("\n") + 1))
    return out


class _Interp:
    def __init__(self, tokens, functions: Dict[str, Any]):
        self.t, self.i = tokens, 0
        self.functions = functions
        self.globals: Dict[str, Any] = {}
        self.scopes: List[Dict[str, Any]] = []

    # -- token helpers
    def peek(self):
        return self.t[self.i]

    def next(self):
        tok = self.t[self.i]
        self.i += 1
        return tok

    def accept(self, kind, val=None):
        k, v, _ = self.peek()
        if k == kind and (val is None or v == val):
            self.i += 1
            return True
        return False

    def expect(self, kind, val=None):
        k, v, line = self.peek()
        if k != kind or (val is not None and v != val):
            raise LuaSubsetError(f"line {line}: expected {val or kind}, found {v!r}")
        return self.next()

    # -- names
    def lookup(self, name, line):
        for sc in reversed(self.scopes):
            if name in sc:
                return sc[name]
        if name in self.globals:
            return self.globals[name]
        if name in self.functions:
            raise LuaSubsetError(f"line {line}: function {name!r} used as a value (outside the supported subset)")
        return None                                      # undefined global is nil in Lua

    # -- expressions: concat < additive < multiplicative < unary < primary
    def expr(self):
        left = self.additive()
        if self.accept("op", ".."):
            right = self.expr()                          # right associative
            return _tostr(left) + _tostr(right)
        return left

    def additive(self):
        v = self.multiplicative()
        while True:
            k, op, line = self.peek()
            if k == "op" and op in "+-" and len(op) == 1:
                self.next()
                r = self.multiplicative()
                v = _arith(op, v, r, line)
            else:
                return v

    def multiplicative(self):
        v = self.unary()
        while True:
            k, op, line = self.peek()
            if k == "op" and op in ("*", "/", "%"):
                self.next()
                r = self.unary()
                v = _arith(op, v, r, line)
            else:
                return v

    def unary(self):
        k, v, line = self.peek()
        if k == "op" and v == "-":
            self.next()
            x = self.unary()
            if not isinstance(x, (int, float)) or isinstance(x, bool):
                raise LuaSubsetError(f"line {line}: unary minus on a non-number")
            return -x
        return self.primary()

    def primary(self):
        k, v, line = self.next()
        if k in ("num", "str"):
            return v
        if k == "kw" and v in ("true", "false", "nil"):
            return {"true": True, "false": False, "nil": None}[v]
        if k == "name":
            if self.peek()[:2] == ("op", "("):
                raise LuaSubsetError(f"line {line}: call of {v!r} inside an expression is outside the supported subset")
            return self.lookup(v, line)
        if k == "op" and v == "(":
            x = self.expr()
            self.expect("op", ")")
            return x
        if k == "op" and v == "{":
            items = []
            while not self.accept("op", "}"):
                items.append(self.expr())
                if not (self.accept("op", ",") or self.accept("op", ";")):
                    self.expect("op", "}")
                    break
            return items
        raise LuaSubsetError(f"line {line}: unexpected token {v!r}")

    # -- statements
    def block(self, until_end: bool):
        while True:
            k, v, line = self.peek()
            if k == "eof":
                if until_end:
                    raise LuaSubsetError(f"line {line}: 'end' expected")
                return
            if k == "kw" and v == "end":
                if not until_end:
                    raise LuaSubsetError(f"line {line}: unexpected 'end'")
                self.next()
                return
            self.statement()

    def skip_block(self):
        depth = 1
        while depth:
            k, v, line = self.next()
            if k == "eof":
                raise LuaSubsetError(f"line {line}: 'end' expected")
            if k == "kw" and v in ("do",):
                depth += 1
            elif k == "kw" and v == "end":
                depth -= 1

    def statement(self):
        k, v, line = self.next()
        if k == "op" and v == ";":
            return
        if k == "kw" and v == "local":
            name = self.expect("name")[1]
            val = None
            if self.accept("op", "="):
                val = self.expr()
            (self.scopes[-1] if self.scopes else self.globals)[name] = val
            self.accept("op", ";")
            return
        if k == "kw" and v == "for":
            var = self.expect("name")[1]
            self.expect("op", "=")
            a = self.expr()
            self.expect("op", ",")
            b = self.expr()
            step = 1
            if self.accept("op", ","):
                step = self.expr()
            self.expect("kw", "do")
            if step == 0:
                raise LuaSubsetError(f"line {line}: 'for' step is zero")
            body_start = self.i
            x, ran = a, False
            while (step > 0 and x <= b) or (step < 0 and x >= b):
                self.i = body_start
                self.scopes.append({var: x})
                self.block(until_end=True)
                self.scopes.pop()
                ran = True
                x = x + step
            if not ran:
                self.i = body_start
                self.skip_block()
            self.accept("op", ";")
            return
        if k == "name":
            if self.accept("op", "("):
                args = []
                if not self.accept("op", ")"):
                    while True:
                        args.append(self.expr())
                        if self.accept("op", ")"):
                            break
                        self.expect("op", ",")
                fn = self.functions.get(v)
                if fn is None:
                    raise LuaSubsetError(f"line {line}: attempt to call a nil value (global '{v}')")
                try:
                    fn(*args)
                except TypeError as e:
                    raise LuaSubsetError(f"line {line}: bad arguments to '{v}': {e}") from None
                self.accept("op", ";")
                return
            if self.accept("op", "="):
                val = self.expr()
                for sc in reversed(self.scopes):
                    if v in sc:
                        sc[v] = val
                        break
                else:
                    self.globals[v] = val
                self.accept("op", ";")
                return
        raise LuaSubsetError(f"line {line}: unsupported statement starting with {v!r}")


def _tostr(x) -> str:
    if isinstance(x, bool) or x is None or isinstance(x, list):
        raise LuaSubsetError("attempt to concatenate a non-string, non-number value")
    if isinstance(x, float):
        return repr(x) if x != int(x) else f"{x:.1f}"
    return str(x)


def _arith(op, a, b, line):
    for x in (a, b):
        if isinstance(x, bool) or not isinstance(x, (int, float)):
            raise LuaSubsetError(f"line {line}: attempt to perform arithmetic on a non-number")
    if op == "+":
        return a + b
    if op == "-":
        return a - b
    if op == "*":
        return a * b
    if op == "/":
        if b == 0:
            return float("inf") if a > 0 else float("-inf") if a < 0 else float("nan")
        return a / b                                     # Lua 5.3: '/' is always float division
    if b == 0:
        raise LuaSubsetError(f"line {line}: attempt to perform 'n%%0'")
    return a - (a // b) * b if isinstance(a, int) and isinstance(b, int) else a - __import__("math").floor(a / b) * b


# the registered globals (state.rs:103-157) and the argument types mlua converts them to
_GLOBALS = ["set_length", "set_render_samplerate", "set_render_bitdepth", "set_output_file", "load_sample",
            "load_resource", "load_midi_floww", "declare_stream", "load_lv2", "parameter", "add_sum",
            "add_normalize", "add_sampleloop", "add_sample_multi", "add_sample_lerp", "add_debug_sine",
            "add_synth", "add_sampsyn", "add_lv2fx", "add_adsr", "add_bandpass", "connect", "set_output"]
_ALIASES = ["add_samplemulti", "add_samplelerp", "add_wavetable", "add_envelope"]      # BASELINE.json paraphrases


def run_script(text: str, project: Project | None = None) -> Project:
    """Execute a project script; returns the `Project` holding the seed vectors."""
    p = project or Project()
    fns = {n: getattr(p, n) for n in _GLOBALS + _ALIASES}
    it = _Interp(tokenize(text), fns)
    it.block(until_end=False)
    return p


def run_file(path: str, project: Project | None = None) -> Project:
    with open(path, "r") as f:
        return run_script(f.read(), project)
