"""A small interpreter for the reference's PATUS stencil definitions (<test>/<test>.stc).

TEST INFRASTRUCTURE.  The suite carries, next to every C / Fortran kernel, a second, declarative definition of the same
stencil for the PATUS code generator (e.g. /root/reference/jacobi/jacobi.stc:1-13): grids, scalar parameters, the
iteration domain and the update expression with explicit neighbour offsets.  For the three Fortran tests (no gfortran in
the build container) this is the one machine-readable definition the reference holds that does not need a Fortran
compiler, so the oracle's restatement of jacobi and sincos is pinned against it (tests/test_stc_pin.py), and the committed
fixtures tests/golden/{jacobi,sincos}_*.npz are generated from it (tests/golden/make_golden_stc.py).

The file is parsed with a hand-written tokenizer and recursive-descent parser (no eval of reference text); expressions
are evaluated with numpy over the whole domain box at once, in the precision of the arrays passed in.
"""
from __future__ import annotations

import re

import numpy as np

_TOKEN = re.compile(r"\s*(?:(\d+\.\d*(?:[eE][-+]?\d+)?|\.\d+(?:[eE][-+]?\d+)?|\d+(?:[eE][-+]?\d+)?)|([A-Za-z_]\w*)|(\.\.|[-+*/()\[\],;={}]))")


def _strip_comments(text: str) -> str:
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    return re.sub(r"//[^\n]*", " ", text)


def _tokenize(text: str):
    pos, out = 0, []
    text = text.rstrip()
    while pos < len(text):
        m = _TOKEN.match(text, pos)
        if not m:
            raise ValueError(f"stc: cannot tokenize at {text[pos:pos + 30]!r}")
        num, ident, punct = m.groups()
        out.append(("num", num) if num is not None else ("id", ident) if ident is not None else ("p", punct))
        pos = m.end()
    return out


class _Parser:
    def __init__(self, toks):
        self.t, self.i = toks, 0

    def peek(self, k=0):
        return self.t[self.i + k] if self.i + k < len(self.t) else ("eof", "")

    def next(self):
        tok = self.peek()
        self.i += 1
        return tok

    def expect(self, val):
        tok = self.next()
        if tok[1] != val:
            raise ValueError(f"stc: expected {val!r}, got {tok[1]!r}")

    # ---- expressions -> AST tuples -------------------------------------------------------------
    def expr(self):
        node = self.term()
        while self.peek()[1] in ("+", "-"):
            op = self.next()[1]
            node = (op, node, self.term())
        return node

    def term(self):
        node = self.factor()
        while self.peek()[1] in ("*", "/"):
            op = self.next()[1]
            node = (op, node, self.factor())
        return node

    def factor(self):
        kind, val = self.next()
        if kind == "num":
            return ("num", float(val))
        if val == "-":
            return ("neg", self.factor())
        if val == "+":
            return self.factor()
        if val == "(":
            node = self.expr()
            self.expect(")")
            return node
        if kind == "id":
            if self.peek()[1] == "(":                      # function call
                self.next()
                arg = self.expr()
                self.expect(")")
                return ("call", val, arg)
            if self.peek()[1] == "[":                      # grid reference
                return self.gridref(val)
            return ("var", val)
        raise ValueError(f"stc: unexpected token {val!r}")

    def gridref(self, name):
        self.expect("[")
        offs = []
        while True:
            offs.append(self.index())
            if self.peek()[1] == ",":
                self.next()
                continue
            break
        time = None
        if self.peek()[1] == ";":
            self.next()
            time = self.index(var="t")[1]
        self.expect("]")
        return ("grid", name, tuple(o[1] for o in offs), tuple(o[0] for o in offs), time)

    def index(self, var=None):
        kind, v = self.next()
        if kind != "id" or (var and v != var):
            raise ValueError(f"stc: bad index {v!r}")
        off = 0
        if self.peek()[1] in ("+", "-"):
            sign = 1 if self.next()[1] == "+" else -1
            kind, n = self.next()
            if kind != "num":
                raise ValueError("stc: index offset must be an integer")
            off = sign * int(float(n))
        return v, off


class Stencil:
    """Parsed <test>.stc: .grids (name -> ndims), .params [names], .domain [(lo_ast, hi_ast)] inclusive bounds,
    .stmts [("local", name, ast) | ("store", gridref_ast, ast)]."""

    def __init__(self, text: str):
        p = _Parser(_tokenize(_strip_comments(text)))
        p.expect("stencil")
        self.name = p.next()[1]
        p.expect("(")
        self.grids, self.params = {}, []
        while p.peek()[1] != ")":
            words = []
            while p.peek()[0] == "id":
                words.append(p.next()[1])
            if "grid" in words:
                nd = 0
                p.expect("(")
                while p.peek()[1] != ")":
                    p.expr()
                    p.expect("..")
                    p.expr()
                    nd += 1
                    if p.peek()[1] == ",":
                        p.next()
                p.expect(")")
                self.grids[words[-1]] = nd
            elif "param" in words:
                self.params.append(words[-1])
            else:
                raise ValueError(f"stc: unknown declaration {words}")
            if p.peek()[1] == ",":
                p.next()
        p.expect(")")
        p.expect("{")
        p.expect("domainsize")
        p.expect("=")
        p.expect("(")
        self.domain = []
        while p.peek()[1] != ")":
            lo = p.expr()
            p.expect("..")
            hi = p.expr()
            self.domain.append((lo, hi))
            if p.peek()[1] == ",":
                p.next()
        p.expect(")")
        p.expect(";")
        p.expect("operation")
        p.expect("{")
        self.stmts = []
        while p.peek()[1] != "}":
            if p.peek()[1] == "float":
                p.next()
                name = p.next()[1]
                p.expect("=")
                self.stmts.append(("local", name, p.expr()))
            else:
                name = p.next()[1]
                ref = p.gridref(name)
                p.expect("=")
                self.stmts.append(("store", ref, p.expr()))
            p.expect(";")

    # ---- evaluation ------------------------------------------------------------------------------
    def offsets(self):
        """Every (grid, offsets, time) the update reads -- the stencil's support."""
        found = set()

        def walk(n):
            if n[0] == "grid":
                found.add((n[1], n[2], n[4]))
            elif n[0] in ("+", "-", "*", "/"):
                walk(n[1]); walk(n[2])
            elif n[0] == "neg":
                walk(n[1])
            elif n[0] == "call":
                walk(n[2])
        for s in self.stmts:
            walk(s[2])
        return found

    def domain_box(self, dims: dict):
        def ev(n):
            if n[0] == "num":
                return int(n[1])
            if n[0] == "var":
                return int(dims[n[1]])
            if n[0] in "+-":
                return ev(n[1]) + ev(n[2]) if n[0] == "+" else ev(n[1]) - ev(n[2])
            raise ValueError("stc: domain bound too complex")
        return [(ev(lo), ev(hi) + 1) for lo, hi in self.domain]        # half-open

    def apply(self, dims: dict, params: dict, read, write):
        """One sweep.  read(grid, time) -> ndarray shaped (ns, ny, nx) / (ny, nx); write(grid, time) -> ndarray to store
        into (only the domain box is written)."""
        box = self.domain_box(dims)
        nd = len(box)
        axes = {"x": nd - 1, "y": nd - 2, "z": nd - 3}
        local = {}

        def ev(n):
            k = n[0]
            if k == "num":
                return n[1]
            if k == "var":
                return local[n[1]] if n[1] in local else params[n[1]]
            if k == "neg":
                return -ev(n[1])
            if k == "call":
                f = {"sin": np.sin, "cos": np.cos, "sqrt": np.sqrt, "exp": np.exp}[n[1]]
                return f(ev(n[2]))
            if k == "grid":
                arr = read(n[1], n[4])
                sl = [None] * nd
                for off, var in zip(n[2], n[3]):
                    lo, hi = box[{"x": 0, "y": 1, "z": 2}[var]]          # the domain is listed (x, y, z)
                    sl[axes[var]] = slice(lo + off, hi + off)
                return arr[tuple(sl)]
            a, b = ev(n[1]), ev(n[2])
            return a + b if k == "+" else a - b if k == "-" else a * b if k == "*" else a / b
        for s in self.stmts:
            if s[0] == "local":
                local[s[1]] = ev(s[2])
            else:
                ref = s[1]
                out = write(ref[1], ref[4])
                sl = [None] * nd
                for off, var in zip(ref[2], ref[3]):
                    lo, hi = box[{"x": 0, "y": 1, "z": 2}[var]]
                    sl[axes[var]] = slice(lo + off, hi + off)
                out[tuple(sl)] = ev(s[2])


# How each .stc names what the drivers' arrays hold: (grid, time) -> array slot (driver init order, include/b200_stencil.h),
# parameter -> index into the scalars the driver draws.  time None = the definition gives no time index.
BINDINGS = {
    "laplacian": {"read": {("U", 0): 0}, "write": {("U", 1): 1}, "params": ["alpha", "beta"]},
    "wave13pt": {"read": {("U", -1): 0, ("U", 0): 1}, "write": {("U", 1): 2}, "params": ["c0", "c1", "c2"]},
    "divergence": {"read": {("Ux", None): 1, ("Uy", None): 2, ("Uz", None): 3}, "write": {("V", 0): 0}, "params": ["alpha", "beta", "gamma"]},
    "gradient": {"read": {("V", None): 0}, "write": {("Ux", 0): 1, ("Uy", 0): 2, ("Uz", 0): 3}, "params": ["alpha", "beta", "gamma"]},
    "lapgsrb": {"read": {("U", 0): 0}, "write": {("U", 1): 1}, "params": ["c0", "c1", "c2", "c3"]},
    "jacobi": {"read": {("U", 0): 0}, "write": {("U", 1): 1}, "params": ["c0", "c1", "c2"]},
    "gaussblur": {"read": {("U", 0): 0}, "write": {("U", 1): 1}, "params": ["s0", "s1", "s2", "s4", "s5", "s8"]},
    "gameoflife": {"read": {("U", 0): 0}, "write": {("U", 1): 1}, "params": []},
    "sincos": {"read": {("U", None): 0, ("V", None): 1}, "write": {("UV", 0): 2}, "params": []},
    "tricubic": {"read": {("U", 0): 0, ("a", None): 2, ("b", None): 3, ("c", None): 4}, "write": {("U", 1): 1}, "params": []},
}


def stc_sweep(stencil: Stencil, test: str, nx: int, ny: int, ns: int, scalars, arrays):
    """One sweep of `test` as its .stc defines it, on the driver's arrays (slot order, flat, x fastest); in place."""
    b = BINDINGS[test]
    nd = len(stencil.domain)
    shape = (ns, ny, nx) if nd == 3 else (ny, nx)
    views = [a.reshape(shape) for a in arrays]
    params = {name: arrays[0].dtype.type(scalars[i]) for i, name in enumerate(b["params"])}
    stencil.apply({"nx": nx, "ny": ny, "ns": ns}, params,
                  read=lambda g, t: views[b["read"][(g, t)]], write=lambda g, t: views[b["write"][(g, t)]])
