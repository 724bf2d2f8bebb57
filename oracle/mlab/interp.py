"""ORACLE TOOLING (test infrastructure only) -- a minimal MATLAB-subset interpreter.

Purpose: execute the reference's own `.m` sources (read in place from /root/reference, never
copied) in this MATLAB-less container, so that the numpy oracle can be pinned against outputs
of the reference code itself for every stage except the qpOASES MEX call.  It implements
exactly the language subset those files use: function files with sub-functions, multiple
return values, `~` placeholders, anonymous functions, if/for/while, N-d arrays with
column-major linear indexing, `end` arithmetic in subscripts, matrix literals with MATLAB's
whitespace rules, and ~45 builtins.  It is NOT a general MATLAB.

    from oracle.mlab.interp import Matlab
    ml = Matlab(["/root/reference/spline", ...])
    A, B, d = ml.call("rk2_kinematic_curvilinear", x, u, kappa, dt, nargout=3)
"""
import math
import os
import re

import numpy as np
from scipy import integrate as _integrate


class MatlabError(Exception):
    pass


# ----------------------------------------------------------------------------- tokenizer
TOKEN_RE = re.compile(r"""
    (?P<num>(\d+(\.(?![*/\\^'])\d*)?([eE][+-]?\d+)?|\.\d+([eE][+-]?\d+)?))
  | (?P<id>[A-Za-z_]\w*)
  | (?P<op>\.\*|\./|\.\^|\.'|==|~=|<=|>=|&&|\|\||\.(?=[A-Za-z_])|[-+*/\\^'<>=&|~:,;()\[\]{}@])
""", re.X)

KEYWORDS = {"function", "end", "if", "elseif", "else", "for", "while", "break", "continue", "return"}


class Tok:
    __slots__ = ("kind", "val", "sp", "line")

    def __init__(self, kind, val, sp, line):
        self.kind, self.val, self.sp, self.line = kind, val, sp, line

    def __repr__(self):
        return f"{self.kind}:{self.val!r}"


def tokenize(src):
    toks = []
    i, n, line = 0, len(src), 1
    depth = 0          # bracket depth [] {} (newlines inside are row separators)
    pdepth = 0
    sp = False
    while i < n:
        c = src[i]
        if c in " \t\r":
            sp = True
            i += 1
            continue
        if c == "%":
            while i < n and src[i] != "\n":
                i += 1
            continue
        if src.startswith("...", i):
            while i < n and src[i] != "\n":
                i += 1
            i += 1
            line += 1
            sp = True
            continue
        if c == "\n":
            toks.append(Tok("nl", "\n", sp, line))
            line += 1
            i += 1
            sp = False
            continue
        if c == '"':
            j = src.index('"', i + 1)
            toks.append(Tok("str", src[i + 1:j], sp, line))
            i = j + 1
            sp = False
            continue
        if c == "'":
            prev = toks[-1] if toks else None
            is_transpose = (prev is not None and not (sp and depth > 0)
                            and (prev.kind in ("num", "id") or prev.val in (")", "]", "}", "'", ".'"))
                            and not (prev.kind == "id" and prev.val in KEYWORDS and prev.val != "end"))
            if not is_transpose:
                j = i + 1
                out = []
                while True:
                    if src[j] == "'":
                        if j + 1 < n and src[j + 1] == "'":
                            out.append("'")
                            j += 2
                            continue
                        break
                    out.append(src[j])
                    j += 1
                toks.append(Tok("str", "".join(out), sp, line))
                i = j + 1
                sp = False
                continue
        m = TOKEN_RE.match(src, i)
        if not m:
            raise MatlabError(f"line {line}: cannot tokenize {src[i:i+20]!r}")
        if m.group("num"):
            toks.append(Tok("num", float(m.group("num")), sp, line))
        elif m.group("id"):
            v = m.group("id")
            toks.append(Tok("kw" if v in KEYWORDS else "id", v, sp, line))
        else:
            v = m.group("op")
            if v in "[{":
                depth += 1
            elif v in "]}":
                depth -= 1
            toks.append(Tok("op", v, sp, line))
        i = m.end()
        sp = False
    toks.append(Tok("nl", "\n", False, line))
    toks.append(Tok("eof", None, False, line))
    return toks


# ----------------------------------------------------------------------------- parser
class Parser:
    def __init__(self, toks):
        self.t = toks
        self.p = 0

    def peek(self, k=0):
        return self.t[self.p + k]

    def next(self):
        tok = self.t[self.p]
        self.p += 1
        return tok

    def accept(self, val):
        if self.peek().val == val and self.peek().kind in ("op", "kw"):
            return self.next()
        return None

    def expect(self, val):
        tok = self.next()
        if tok.val != val:
            raise MatlabError(f"line {tok.line}: expected {val!r}, got {tok.val!r}")
        return tok

    def skip_nl(self):
        while self.peek().kind == "nl" or (self.peek().kind == "op" and self.peek().val in ";,"):
            self.next()

    # --- file level
    def parse_file(self):
        funcs = []
        self.skip_nl()
        while self.peek().kind != "eof":
            funcs.append(self.parse_function())
            self.skip_nl()
        return funcs

    def parse_function(self):
        self.expect("function")
        outs = []
        # forms: function name(...), function out = name(...), function [o1,o2] = name(...)
        if self.peek().val == "[":
            self.next()
            while self.peek().val != "]":
                tok = self.next()
                if tok.val == ",":
                    continue
                outs.append(tok.val)
            self.next()
            self.expect("=")
            name = self.next().val
        else:
            name = self.next().val
            if self.peek().val == "=":
                self.next()
                outs = [name]
                name = self.next().val
        args = []
        if self.peek().val == "(":
            self.next()
            while self.peek().val != ")":
                tok = self.next()
                if tok.val == ",":
                    continue
                args.append(tok.val)
            self.next()
        body = self.parse_block(("function", "eof"), file_level=True)
        return ("func", name, args, outs, body)

    def parse_block(self, stops, file_level=False):
        stmts = []
        while True:
            self.skip_nl()
            tok = self.peek()
            if tok.kind == "eof" or (tok.kind == "kw" and tok.val in stops):
                if file_level and tok.kind == "kw" and tok.val == "end":
                    self.next()          # `end` closing a function
                    continue
                return stmts
            if file_level and tok.kind == "kw" and tok.val == "end":
                self.next()
                return stmts
            stmts.append(self.parse_statement())

    def parse_statement(self):
        tok = self.peek()
        if tok.kind == "kw":
            if tok.val == "if":
                return self.parse_if()
            if tok.val == "for":
                self.next()
                paren = self.accept("(")
                var = self.next().val
                self.expect("=")
                rng = self.parse_expr()
                if paren:
                    self.expect(")")
                body = self.parse_block(("end",))
                self.expect("end")
                return ("for", var, rng, body)
            if tok.val == "while":
                self.next()
                cond = self.parse_expr()
                body = self.parse_block(("end",))
                self.expect("end")
                return ("while", cond, body)
            if tok.val in ("break", "continue", "return"):
                self.next()
                return (tok.val,)
            raise MatlabError(f"line {tok.line}: unexpected keyword {tok.val}")
        # multi-assignment [a, b] = f(...)
        if tok.val == "[" and tok.kind == "op":
            save = self.p
            try:
                targets = self.parse_lhs_list()
                if self.peek().val == "=" and self.peek(1).val != "=":
                    self.next()
                    rhs = self.parse_expr()
                    return ("massign", targets, rhs, self.end_stmt())
            except MatlabError:
                pass
            self.p = save
        expr = self.parse_expr()
        if self.peek().val == "=" and self.peek().kind == "op":
            self.next()
            rhs = self.parse_expr()
            return ("assign", expr, rhs, self.end_stmt())
        return ("expr", expr, self.end_stmt())

    def end_stmt(self):
        """consume the statement terminator; returns True if output is suppressed"""
        tok = self.peek()
        if tok.kind == "op" and tok.val == ";":
            self.next()
            return True
        if tok.kind == "op" and tok.val == ",":
            self.next()
        return False

    def parse_lhs_list(self):
        self.expect("[")
        targets = []
        while self.peek().val != "]":
            if self.peek().val == ",":
                self.next()
                continue
            if self.peek().val == "~":
                self.next()
                targets.append(None)
                continue
            targets.append(self.parse_postfix(in_matrix=True))
        self.next()
        return targets

    def parse_if(self):
        self.expect("if")
        cond = self.parse_expr()
        body = self.parse_block(("elseif", "else", "end"))
        clauses = [(cond, body)]
        orelse = None
        while True:
            tok = self.next()
            if tok.val == "elseif":
                c = self.parse_expr()
                b = self.parse_block(("elseif", "else", "end"))
                clauses.append((c, b))
            elif tok.val == "else":
                orelse = self.parse_block(("end",))
            elif tok.val == "end":
                break
        return ("if", clauses, orelse)

    # --- expressions (precedence climbing, MATLAB order)
    def parse_expr(self, in_matrix=False):
        return self.parse_oror(in_matrix)

    def parse_oror(self, im):
        left = self.parse_andand(im)
        while self.peek().val == "||":
            self.next()
            left = ("oror", left, self.parse_andand(im))
        return left

    def parse_andand(self, im):
        left = self.parse_or(im)
        while self.peek().val == "&&":
            self.next()
            left = ("andand", left, self.parse_or(im))
        return left

    def parse_or(self, im):
        left = self.parse_and(im)
        while self.peek().val == "|" and self.peek().kind == "op":
            self.next()
            left = ("bin", "|", left, self.parse_and(im))
        return left

    def parse_and(self, im):
        left = self.parse_cmp(im)
        while self.peek().val == "&" and self.peek().kind == "op":
            self.next()
            left = ("bin", "&", left, self.parse_cmp(im))
        return left

    def parse_cmp(self, im):
        left = self.parse_range(im)
        while self.peek().kind == "op" and self.peek().val in ("==", "~=", "<", "<=", ">", ">="):
            op = self.next().val
            left = ("bin", op, left, self.parse_range(im))
        return left

    def parse_range(self, im):
        first = self.parse_add(im)
        if self.peek().val == ":" and self.peek().kind == "op" and self.peek(1).val not in (")", ","):
            self.next()
            second = self.parse_add(im)
            if self.peek().val == ":" and self.peek().kind == "op":
                self.next()
                third = self.parse_add(im)
                return ("range", first, second, third)
            return ("range", first, None, second)
        return first

    def _is_elem_sep(self, im):
        """inside [], `a -b` / `a +b` starts a new element; `a - b` and `a-b` are binary"""
        tok, nxt = self.peek(), self.peek(1)
        return im and tok.sp and not nxt.sp

    def parse_add(self, im):
        left = self.parse_mul(im)
        while self.peek().kind == "op" and self.peek().val in ("+", "-"):
            if self._is_elem_sep(im):
                break
            op = self.next().val
            left = ("bin", op, left, self.parse_mul(im))
        return left

    def parse_mul(self, im):
        left = self.parse_unary(im)
        while self.peek().kind == "op" and self.peek().val in ("*", "/", "\\", ".*", "./"):
            op = self.next().val
            left = ("bin", op, left, self.parse_unary(im))
        return left

    def parse_unary(self, im):
        tok = self.peek()
        if tok.kind == "op" and tok.val in ("-", "+", "~"):
            self.next()
            operand = self.parse_unary(im)
            return ("un", tok.val, operand)
        return self.parse_power(im)

    def parse_power(self, im):
        base = self.parse_postfix(im)
        while self.peek().kind == "op" and self.peek().val in ("^", ".^"):
            op = self.next().val
            # exponent binds tighter than unary minus on its right: 2^-1
            if self.peek().val in ("-", "+") and self.peek().kind == "op":
                sgn = self.next().val
                expo = ("un", sgn, self.parse_postfix(im))
            else:
                expo = self.parse_postfix(im)
            base = ("bin", op, base, expo)
        return base

    def parse_postfix(self, in_matrix=False):
        node = self.parse_primary()
        while True:
            tok = self.peek()
            if tok.kind == "op" and tok.val == "(" and not (in_matrix and tok.sp):
                self.next()
                args = self.parse_args(")")
                node = ("index", node, args)
            elif tok.kind == "op" and tok.val == "{" and not (in_matrix and tok.sp):
                self.next()
                args = self.parse_args("}")
                node = ("cellindex", node, args)
            elif tok.kind == "op" and tok.val in ("'", ".'") and not (in_matrix and tok.sp):
                self.next()
                node = ("transpose", node)
            elif tok.kind == "op" and tok.val == ".":
                self.next()                      # struct field s.name (parsed; only dict values can be read)
                node = ("field", node, self.next().val)
            else:
                return node

    def parse_args(self, close):
        args = []
        while self.peek().kind == "nl":
            self.next()
        if self.peek().val == close:
            self.next()
            return args
        while True:
            if self.peek().val == ":" and self.peek(1).val in (",", close):
                self.next()
                args.append(("colon",))
            else:
                args.append(self.parse_expr())
            tok = self.next()
            if tok.val == close:
                return args
            if tok.val != ",":
                raise MatlabError(f"line {tok.line}: expected , or {close} got {tok.val!r}")

    def parse_primary(self):
        tok = self.next()
        if tok.kind == "num":
            return ("num", tok.val)
        if tok.kind == "str":
            return ("str", tok.val)
        if tok.kind == "id":
            return ("name", tok.val)
        if tok.kind == "kw" and tok.val == "end":
            return ("endval",)
        if tok.kind == "op":
            if tok.val == "(":
                e = self.parse_expr()
                self.expect(")")
                return ("paren", e)
            if tok.val == "[":
                return self.parse_matrix()
            if tok.val == "{":
                rows = self.parse_matrix_rows("}")
                return ("cell", rows)
            if tok.val == "@":
                if self.peek().val == "(":
                    self.next()
                    params = []
                    while self.peek().val != ")":
                        t2 = self.next()
                        if t2.val != ",":
                            params.append(t2.val)
                    self.next()
                    body = self.parse_expr()
                    return ("anon", params, body)
                return ("fhandle", self.next().val)
        raise MatlabError(f"line {tok.line}: unexpected token {tok.val!r}")

    def parse_matrix(self):
        return ("matrix", self.parse_matrix_rows("]"))

    def parse_matrix_rows(self, close):
        rows, row = [], []
        while True:
            tok = self.peek()
            if tok.kind == "op" and tok.val == close:
                self.next()
                if row:
                    rows.append(row)
                return rows
            if tok.kind == "nl" or (tok.kind == "op" and tok.val == ";"):
                self.next()
                if row:
                    rows.append(row)
                row = []
                continue
            if tok.kind == "op" and tok.val == ",":
                self.next()
                continue
            row.append(self.parse_expr(in_matrix=True))


# ----------------------------------------------------------------------------- values
def M(a):
    """to MATLAB numeric array (>= 2-D float64 or bool)"""
    if isinstance(a, np.ndarray):
        if a.ndim == 0:
            return a.reshape(1, 1)
        if a.ndim == 1:
            return a.reshape(1, -1)
        return a
    if isinstance(a, (bool, np.bool_)):
        return np.array([[bool(a)]])
    if isinstance(a, (int, float, np.floating, np.integer)):
        return np.array([[float(a)]])
    return a


def scalar(a):
    if isinstance(a, np.ndarray):
        if a.size != 1:
            raise MatlabError("scalar expected")
        return float(a.reshape(-1)[0])
    return float(a)


def is_true(a):
    a = M(a)
    return a.size > 0 and bool(np.all(a != 0))


def colmajor(a):
    return np.reshape(a, -1, order="F")


class Break(Exception):
    pass


class Continue(Exception):
    pass


class Return(Exception):
    pass


class Colon:
    pass


class MFunction:
    def __init__(self, interp, node, siblings):
        _, self.name, self.args, self.outs, self.body = node
        self.interp = interp
        self.siblings = siblings

    def __call__(self, *args, nargout=1):
        env = {}
        if len(args) > len(self.args):
            raise MatlabError(f"{self.name}: too many arguments")
        for name, val in zip(self.args, args):
            if name != "~":
                env[name] = val
        frame = Frame(self.interp, env, self.siblings, nargin=len(args), nargout=nargout)
        try:
            frame.run(self.body)
        except Return:
            pass
        outs = []
        for o in self.outs[:max(nargout, 1)]:
            if o not in env:
                raise MatlabError(f"{self.name}: output {o} not assigned")
            outs.append(env[o])
        return outs


# ----------------------------------------------------------------------------- interpreter
class Matlab:
    def __init__(self, paths, overrides=None):
        self.paths = list(paths)
        self.cache = {}
        self.overrides = dict(overrides or {})     # name -> python callable(*args, nargout=) -> list
        self.calls = {}

    def load(self, name):
        if name in self.cache:
            return self.cache[name]
        for d in self.paths:
            f = os.path.join(d, name + ".m")
            if os.path.exists(f):
                with open(f) as fh:
                    src = fh.read()
                nodes = Parser(tokenize(src)).parse_file()
                sib = {}
                funcs = [MFunction(self, nd, sib) for nd in nodes]
                for fn in funcs:
                    sib[fn.name] = fn
                self.cache[name] = funcs[0]
                return funcs[0]
        self.cache[name] = None
        return None

    def run_script_lines(self, path, first, last, env):
        """Execute lines first..last (1-based, inclusive) of a SCRIPT file in the workspace `env`
        (a dict that is updated in place), e.g. a block of main.m.  Returns env."""
        with open(path) as fh:
            lines = fh.read().split("\n")
        src = "\n".join(lines[first - 1:last]) + "\n"
        stmts = Parser(tokenize(src)).parse_block(("eof",))
        frame = Frame(self, env, {})
        try:
            frame.run(stmts)
        except (Break, Return):
            pass
        return env

    def call(self, name, *args, nargout=1):
        fn = self.overrides.get(name) or self.load(name)
        if fn is None:
            raise MatlabError(f"unknown function {name}")
        self.calls[name] = self.calls.get(name, 0) + 1
        outs = fn(*[M(a) if not callable(a) and not isinstance(a, str) else a for a in args], nargout=nargout)
        return outs[0] if nargout <= 1 else outs[:nargout]


class Frame:
    def __init__(self, interp, env, siblings, nargin=0, nargout=1):
        self.I = interp
        self.env = env
        self.sib = siblings
        self.nargin, self.nargout = nargin, nargout
        self.end_stack = []

    # --- statements
    def run(self, stmts):
        for s in stmts:
            self.exec(s)

    def exec(self, s):
        k = s[0]
        if k == "assign":
            self.assign(s[1], self.eval(s[2]))
        elif k == "massign":
            targets = s[1]
            vals = self.eval_multi(s[2], len(targets))
            if len(vals) < len([t for t in targets]):
                raise MatlabError("not enough outputs")
            for t, v in zip(targets, vals):
                if t is not None:
                    self.assign(t, v)
        elif k == "expr":
            v = self.eval_multi(s[1], 0)
            if v:
                self.env["ans"] = v[0]
        elif k == "if":
            for cond, body in s[1]:
                if is_true(self.eval(cond)):
                    self.run(body)
                    return
            if s[2] is not None:
                self.run(s[2])
        elif k == "for":
            rng = M(self.eval(s[2]))
            for c in range(rng.shape[1] if rng.size else 0):
                self.env[s[1]] = rng[:, c:c + 1] if rng.shape[0] > 1 else M(rng[0, c])
                try:
                    self.run(s[3])
                except Break:
                    break
                except Continue:
                    continue
        elif k == "while":
            while is_true(self.eval(s[1])):
                try:
                    self.run(s[2])
                except Break:
                    break
                except Continue:
                    continue
        elif k == "break":
            raise Break()
        elif k == "continue":
            raise Continue()
        elif k == "return":
            raise Return()
        else:
            raise MatlabError(f"unknown statement {k}")

    # --- assignment
    def assign(self, target, val):
        if target[0] == "name":
            self.env[target[1]] = val
            return
        if target[0] == "index" and target[1][0] == "name":
            name = target[1][1]
            cur = self.env.get(name)
            if cur is None:
                cur = np.zeros((0, 0))
            cur = M(cur)
            self.env[name] = self.index_assign(cur, target[2], M(val))
            return
        if target[0] == "cellindex" and target[1][0] == "name":
            # c{k} = v on a cell array held as a python list (value semantics: copy on write)
            name = target[1][1]
            cur = list(self.env.get(name) or [])
            idx = int(scalar(self.eval(target[2][0]))) - 1
            while len(cur) <= idx:
                cur.append(np.zeros((0, 0)))
            cur[idx] = val
            self.env[name] = cur
            return
        raise MatlabError(f"unsupported assignment target {target[0]}")

    def subs(self, arr, args):
        """evaluate subscripts against arr -> list of index arrays (zero-based) or Colon"""
        out = []
        n = len(args)
        for pos, a in enumerate(args):
            if a[0] == "colon":
                out.append(Colon())
                continue
            self.end_stack.append((arr, pos, n))
            try:
                v = self.eval(a)
            finally:
                self.end_stack.pop()
            v = M(v)
            if v.dtype == bool:
                idx = np.flatnonzero(colmajor(v))
            else:
                idx = colmajor(v)
                if np.any(idx != np.round(idx)) or np.any(idx < 1):
                    raise MatlabError(f"bad subscript {idx}")
                idx = idx.astype(np.int64) - 1
            out.append((idx, v.shape))
        return out

    def dim_size(self, arr, pos, n):
        shp = arr.shape
        if n == 1:
            return arr.size
        if pos < n - 1:
            return shp[pos] if pos < len(shp) else 1
        return int(np.prod(shp[pos:])) if pos < len(shp) else 1

    def index_read(self, arr, args):
        arr = M(arr)
        subs = self.subs(arr, args)
        n = len(subs)
        if n == 0:
            return arr
        if n == 1:
            s = subs[0]
            flat = colmajor(arr)
            if isinstance(s, Colon):
                return flat.reshape(-1, 1).copy()
            idx, shp = s
            if idx.size and idx.max() >= flat.size:
                raise MatlabError("index exceeds array bounds")
            res = flat[idx]
            if arr.shape[0] == 1 and arr.ndim == 2 and (len(shp) == 2 and (shp[0] == 1 or shp[1] == 1)):
                return res.reshape(1, -1)      # row vector source keeps row orientation for vector index
            if len(shp) == 2 and (shp[0] == 1 or shp[1] == 1) and (arr.ndim > 2 or (arr.shape[0] != 1 and arr.shape[1] != 1)):
                return res.reshape(shp, order="F")
            if len(shp) == 2 and (shp[0] == 1 or shp[1] == 1):
                # vector source, vector index: orientation of the source
                return res.reshape(-1, 1) if arr.shape[1] == 1 else res.reshape(1, -1)
            return res.reshape(shp, order="F")
        # N-d: collapse trailing dims into the last subscript
        shp = list(arr.shape) + [1] * max(0, n - arr.ndim)
        if n < len(shp):
            shp = shp[:n - 1] + [int(np.prod(shp[n - 1:]))]
        a = arr.reshape(shp, order="F")
        ix = []
        for d, s in enumerate(subs):
            if isinstance(s, Colon):
                ix.append(np.arange(shp[d]))
            else:
                if s[0].size and s[0].max() >= shp[d]:
                    raise MatlabError("index exceeds array bounds")
                ix.append(s[0])
        res = a[np.ix_(*ix)]
        # drop trailing singleton dims beyond 2
        while res.ndim > 2 and res.shape[-1] == 1:
            res = res.reshape(res.shape[:-1])
        return res

    def index_assign(self, arr, args, val):
        subs = self.subs(arr, args)
        n = len(subs)
        if n == 1:
            s = subs[0]
            if isinstance(s, Colon):
                flat = np.broadcast_to(colmajor(val), (arr.size,)) if val.size in (1, arr.size) else None
                if flat is None:
                    raise MatlabError("A(:) = B size mismatch")
                return np.array(flat, dtype=np.float64).reshape(arr.shape, order="F")
            idx, _ = s
            need = int(idx.max()) + 1 if idx.size else 0
            if arr.size == 0:
                arr = np.zeros((1, need))
            elif need > arr.size:
                if arr.shape[0] == 1:
                    arr = np.hstack([arr, np.zeros((1, need - arr.size))])
                elif arr.shape[1] == 1:
                    arr = np.vstack([arr, np.zeros((need - arr.size, 1))])
                else:
                    raise MatlabError("linear index growth of a matrix")
            flat = colmajor(arr).astype(np.float64).copy()
            v = colmajor(val)
            if v.size not in (1, idx.size):
                raise MatlabError(f"assignment size mismatch: {idx.size} <- {v.size}")
            flat[idx] = v
            return flat.reshape(arr.shape, order="F")
        shp = list(arr.shape) + [1] * max(0, n - arr.ndim)
        if n < len(shp):
            raise MatlabError("assignment with fewer subscripts than dims")
        ix = []
        new_shape = list(shp)
        for d, s in enumerate(subs):
            if isinstance(s, Colon):
                if shp[d] == 0 and val.ndim > d:
                    new_shape[d] = val.shape[d] if d < val.ndim else 1
                ix.append(None)
            else:
                need = int(s[0].max()) + 1 if s[0].size else 0
                new_shape[d] = max(new_shape[d], need)
                ix.append(s[0])
        if new_shape != shp:
            grown = np.zeros(new_shape)
            if arr.size:
                grown[tuple(slice(0, k) for k in shp)] = arr.reshape(shp, order="F")
            a = grown
        else:
            a = arr.reshape(shp, order="F").astype(np.float64).copy()
        ix = [np.arange(a.shape[d]) if i is None else i for d, i in enumerate(ix)]
        tgt_shape = tuple(len(i) for i in ix)
        if val.size == 1:
            a[np.ix_(*ix)] = scalar(val)
        else:
            v = val
            vs = [k for k in v.shape if k != 1]
            ts = [k for k in tgt_shape if k != 1]
            if vs != ts:
                raise MatlabError(f"assignment shape mismatch {tgt_shape} <- {v.shape}")
            a[np.ix_(*ix)] = v.reshape(tgt_shape, order="F")
        while a.ndim > 2 and a.shape[-1] == 1:
            a = a.reshape(a.shape[:-1])
        return a

    # --- expressions
    def eval_multi(self, node, nargout):
        """evaluate allowing multiple return values (function calls)"""
        if node[0] == "index" and node[1][0] == "name" and node[1][1] not in self.env:
            args = [self.eval_arg(a) for a in node[2]]
            return self.call_function(node[1][1], args, nargout)
        if node[0] == "name" and node[1] not in self.env:
            return self.call_function(node[1], [], nargout)
        return [self.eval(node)]

    def eval_arg(self, a):
        if a[0] == "colon":
            return ":"
        return self.eval(a)

    def call_function(self, name, args, nargout):
        if name in self.sib:
            return self.sib[name](*args, nargout=max(nargout, 1))
        if name in self.I.overrides:
            self.I.calls[name] = self.I.calls.get(name, 0) + 1
            return self.I.overrides[name](*args, nargout=max(nargout, 1))
        fn = self.I.load(name)
        if fn is not None:
            self.I.calls[name] = self.I.calls.get(name, 0) + 1
            return fn(*args, nargout=max(nargout, 1))
        b = BUILTINS.get(name)
        if b is None:
            raise MatlabError(f"undefined function or variable '{name}'")
        res = b(self, args, max(nargout, 1))
        return res if isinstance(res, list) else [res]

    def eval(self, node):
        k = node[0]
        if k == "num":
            return np.array([[node[1]]])
        if k == "str":
            return node[1]
        if k == "paren":
            return self.eval(node[1])
        if k == "name":
            if node[1] in self.env:
                return self.env[node[1]]
            return self.call_function(node[1], [], 1)[0]
        if k == "endval":
            if not self.end_stack:
                raise MatlabError("`end` outside of a subscript")
            arr, pos, n = self.end_stack[-1]
            return np.array([[float(self.dim_size(M(arr), pos, n))]])
        if k == "index":
            base = node[1]
            if base[0] == "name" and base[1] not in self.env:
                args = [self.eval_arg(a) for a in node[2]]
                return self.call_function(base[1], args, 1)[0]
            target = self.eval(base)
            if callable(target):
                args = [self.eval_arg(a) for a in node[2]]
                r = target(*args)
                return r[0] if isinstance(r, list) else r
            return self.index_read(target, node[2])
        if k == "cellindex":
            target = self.eval(node[1])
            idx = int(scalar(self.eval(node[2][0]))) - 1
            return target[idx]
        if k == "cell":
            return [self.eval(e) for row in node[1] for e in row]
        if k == "field":
            target = self.eval(node[1])
            if not isinstance(target, dict) or node[2] not in target:
                raise MatlabError(f"no field {node[2]!r}")
            return target[node[2]]
        if k == "transpose":
            v = M(self.eval(node[1]))
            return v.T.copy()
        if k == "un":
            v = M(self.eval(node[2]))
            if node[1] == "-":
                return -v.astype(np.float64)
            if node[1] == "+":
                return v
            return ~(v != 0)
        if k == "bin":
            lhs, rhs = self.eval(node[2]), self.eval(node[3])
            if isinstance(lhs, str) or isinstance(rhs, str):
                return self.string_binop(node[1], lhs, rhs)
            return self.binop(node[1], M(lhs), M(rhs))
        if k == "andand":
            return np.array([[is_true(self.eval(node[1])) and is_true(self.eval(node[2]))]])
        if k == "oror":
            return np.array([[is_true(self.eval(node[1])) or is_true(self.eval(node[2]))]])
        if k == "range":
            a = scalar(self.eval(node[1]))
            b = scalar(self.eval(node[3]))
            st = 1.0 if node[2] is None else scalar(self.eval(node[2]))
            if st == 0 or (st > 0 and a > b) or (st < 0 and a < b):
                return np.zeros((1, 0))
            nel = int(math.floor((b - a) / st * (1 + 1e-15) + 1e-10)) + 1
            return (a + st * np.arange(nel)).reshape(1, -1)
        if k == "matrix":
            rows = []
            for row in node[1]:
                elems = [M(self.eval(e)) for e in row]
                elems = [e.astype(np.float64) if e.dtype == bool else e for e in elems]
                elems = [e for e in elems if e.size or len(elems) == 1]
                if elems:
                    rows.append(np.concatenate(elems, axis=1) if len(elems) > 1 else elems[0])
            rows = [r for r in rows if r.size]
            if not rows:
                return np.zeros((0, 0))
            return np.concatenate(rows, axis=0) if len(rows) > 1 else rows[0]
        if k == "anon":
            params, body = node[1], node[2]
            captured = dict(self.env)
            frame_sib, interp = self.sib, self.I

            def fn(*args, nargout=1, _p=params, _b=body, _c=captured):
                env = dict(_c)
                for nme, v in zip(_p, args):
                    env[nme] = M(v) if not callable(v) and not isinstance(v, str) else v
                return [Frame(interp, env, frame_sib).eval(_b)]
            return fn
        if k == "fhandle":
            name = node[1]
            return lambda *args, nargout=1: self.call_function(name, list(args), nargout)
        raise MatlabError(f"cannot evaluate {k}")

    @staticmethod
    def string_binop(op, a, b):
        """string scalars ("..."): comparison gives a logical scalar, + concatenates (numbers are
        converted the way MATLAB's string() prints integers)"""
        def txt(v):
            if isinstance(v, str):
                return v
            f = float(scalar(v))
            return str(int(f)) if f == int(f) else repr(f)
        if op == "==":
            return np.array([[isinstance(a, str) and isinstance(b, str) and a == b]])
        if op == "~=":
            return np.array([[not (isinstance(a, str) and isinstance(b, str) and a == b)]])
        if op == "+":
            return txt(a) + txt(b)
        raise MatlabError(f"unsupported string operator {op}")

    def binop(self, op, a, b):
        a = a.astype(np.float64) if a.dtype == bool and op not in ("&", "|") else a
        b = b.astype(np.float64) if b.dtype == bool and op not in ("&", "|") else b
        if op == "+":
            return a + b
        if op == "-":
            return a - b
        if op == ".*":
            return a * b
        if op == "./":
            with np.errstate(divide="ignore", invalid="ignore"):
                return a / b
        if op == ".^":
            return np.power(a, b)
        if op == "*":
            if a.size == 1 or b.size == 1:
                return a * b
            return a @ b
        if op == "/":
            if b.size == 1:
                with np.errstate(divide="ignore", invalid="ignore"):
                    return a / b
            return np.linalg.solve(b.T, a.T).T
        if op == "\\":
            if a.size == 1:
                return b / a
            return np.linalg.solve(a, b)
        if op == "^":
            if a.size == 1 and b.size == 1:
                return np.power(a, b)
            if b.size == 1 and a.shape[0] == a.shape[1]:
                return np.linalg.matrix_power(a, int(scalar(b)))
            raise MatlabError("unsupported ^")
        if op in ("==", "~=", "<", "<=", ">", ">="):
            f = {"==": np.equal, "~=": np.not_equal, "<": np.less, "<=": np.less_equal,
                 ">": np.greater, ">=": np.greater_equal}[op]
            return f(a, b)
        if op == "&":
            return (a != 0) & (b != 0)
        if op == "|":
            return (a != 0) | (b != 0)
        raise MatlabError(f"unknown operator {op}")


# ----------------------------------------------------------------------------- builtins
def _dims(args):
    if len(args) == 1:
        a = M(args[0])
        if a.size == 1:
            n = int(scalar(a))
            return (n, n)
        return tuple(int(v) for v in colmajor(a))
    return tuple(int(scalar(a)) for a in args)


def _elementwise(f):
    return lambda fr, args, no: f(M(args[0]).astype(np.float64))


def b_size(fr, args, nargout):
    a = args[0]
    shp = (1, len(a)) if isinstance(a, (str, list)) else M(a).shape
    if len(args) > 1:
        d = int(scalar(args[1])) - 1
        return np.array([[float(shp[d] if d < len(shp) else 1)]])
    if nargout <= 1:
        return np.array([list(map(float, shp))])
    out = [float(s) for s in shp[:nargout]]
    if nargout < len(shp):
        out[-1] = float(np.prod(shp[nargout - 1:]))
    while len(out) < nargout:
        out.append(1.0)
    return [np.array([[v]]) for v in out]


def b_length(fr, args, no):
    a = args[0]
    if isinstance(a, (str, list)):
        return np.array([[float(len(a))]])
    a = M(a)
    return np.array([[float(0 if a.size == 0 else max(a.shape))]])


def b_repmat(fr, args, no):
    a = M(args[0])
    reps = _dims(args[1:])
    if len(reps) == 1:
        reps = (reps[0], reps[0])
    while a.ndim < len(reps):
        a = a.reshape(a.shape + (1,))
    return np.tile(a, reps)


def b_spdiags(fr, args, no):
    """spdiags(B, d, m, n): column k of B on diagonal d(k) (MATLAB placement rule)."""
    Bm, d, m, n = M(args[0]).astype(np.float64), colmajor(M(args[1])).astype(int), int(scalar(args[2])), int(scalar(args[3]))
    if Bm.shape[0] == 1 and len(d) == 1 and Bm.shape[1] > 1:
        Bm = Bm.T
    A = np.zeros((m, n))
    for k, dk in enumerate(d):
        col = Bm[:, k]
        for j in range(n):
            i = j - dk
            if 0 <= i < m:
                # m >= n: element taken from B(j, k); m < n: from B(i, k)
                src = j if m >= n else i
                if src < col.size:
                    A[i, j] = col[src]
    return A


def b_minmax(which):
    def f(fr, args, nargout):
        a = M(args[0]).astype(np.float64)
        if len(args) >= 2 and not (isinstance(args[1], np.ndarray) and M(args[1]).size == 0):
            b = M(args[1]).astype(np.float64)
            return np.minimum(a, b) if which == "min" else np.maximum(a, b)
        red = np.min if which == "min" else np.max
        arg = np.argmin if which == "min" else np.argmax
        if a.shape[0] == 1 or a.shape[1] == 1:
            v = colmajor(a)
            if nargout > 1:
                return [np.array([[red(v)]]), np.array([[float(arg(v) + 1)]])]
            return np.array([[red(v)]])
        return red(a, axis=0, keepdims=True)
    return f


def b_sum(fr, args, no):
    a = M(args[0]).astype(np.float64)
    if len(args) > 1:
        return np.sum(a, axis=int(scalar(args[1])) - 1, keepdims=True)
    if a.shape[0] == 1 or a.shape[1] == 1:
        return np.array([[a.sum()]])
    return a.sum(axis=0, keepdims=True)


def b_cumsum(fr, args, no):
    a = M(args[0]).astype(np.float64)
    if a.shape[0] == 1:
        return np.cumsum(a, axis=1)
    return np.cumsum(a, axis=0)


def b_find(fr, args, no):
    a = colmajor(M(args[0]))
    idx = np.flatnonzero(a != 0) + 1.0
    if len(args) > 1:
        idx = idx[:int(scalar(args[1]))]
    shape_row = M(args[0]).shape[0] == 1
    return idx.reshape(1, -1) if shape_row else idx.reshape(-1, 1)


def b_integral(fr, args, no):
    """integral(f, a, b): MATLAB's adaptive Gauss-Kronrod (AbsTol 1e-10, RelTol 1e-6); restated
    with QUADPACK at tighter tolerance (smooth integrands: both agree to ~1e-12)."""
    f, a, b = args[0], scalar(args[1]), scalar(args[2])

    def g(t):
        r = f(np.array([[t]]))
        r = r[0] if isinstance(r, list) else r
        return scalar(r)
    return np.array([[_integrate.quad(g, a, b, epsabs=1e-13, epsrel=1e-13, limit=200)[0]]])


def b_exist(fr, args, no):
    name = args[0]
    return np.array([[1.0 if name in fr.env else 0.0]])


def b_zeros(fill):
    def f(fr, args, no):
        shp = _dims(args) if args else (1, 1)
        return np.full(shp, fill, dtype=np.float64)
    return f


def b_eye(fr, args, no):
    shp = _dims(args)
    return np.eye(shp[0], shp[1] if len(shp) > 1 else shp[0])


def b_linspace(fr, args, no):
    n = int(scalar(args[2])) if len(args) > 2 else 100
    return np.linspace(scalar(args[0]), scalar(args[1]), n).reshape(1, -1)


def b_reshape(fr, args, no):
    a = M(args[0])
    shp = _dims(args[1:])
    return a.reshape(shp, order="F")


def b_norm(fr, args, no):
    return np.array([[np.linalg.norm(colmajor(M(args[0])))]])


def b_mod(fr, args, no):
    x, y = M(args[0]).astype(np.float64), M(args[1]).astype(np.float64)
    return x - np.floor(x / y) * y


def b_display(fr, args, no):
    print("display:", args[0])
    return []


def b_inf(fr, args, no):
    if not args:
        return np.array([[np.inf]])
    return np.full(_dims(args), np.inf)


def b_numel(fr, args, no):
    return np.array([[float(M(args[0]).size)]])


def b_angdiff(fr, args, no):
    """Robotics System Toolbox angdiff(alpha, beta): beta - alpha wrapped to [-pi, pi]
    (wrapToPi: values already inside the interval are returned unchanged; outside it
    mod(theta + pi, 2 pi) - pi, with positive multiples of 2 pi mapped to +pi)."""
    if len(args) != 2:
        raise MatlabError("angdiff: two-argument form only")
    th = M(args[1]).astype(np.float64) - M(args[0]).astype(np.float64)
    out = th.copy()
    pos = (th < -math.pi) | (th > math.pi)
    t2 = th[pos] + math.pi
    w = t2 - np.floor(t2 / (2 * math.pi)) * (2 * math.pi)
    w[(w == 0) & (t2 > 0)] = 2 * math.pi
    out[pos] = w - math.pi
    return out


BUILTINS = {
    "angdiff": b_angdiff,
    "tic": lambda fr, a, no: [], "toc": lambda fr, a, no: np.array([[0.0]]),
    "zeros": b_zeros(0.0), "ones": b_zeros(1.0), "eye": b_eye, "inf": b_inf, "Inf": b_inf,
    "pi": lambda fr, a, no: np.array([[math.pi]]),
    "size": b_size, "length": b_length, "numel": b_numel, "repmat": b_repmat, "spdiags": b_spdiags,
    "sin": _elementwise(np.sin), "cos": _elementwise(np.cos), "tan": _elementwise(np.tan),
    "atan": _elementwise(np.arctan), "exp": _elementwise(np.exp), "sqrt": _elementwise(np.sqrt),
    "abs": _elementwise(np.abs), "floor": _elementwise(np.floor), "ceil": _elementwise(np.ceil),
    "sec": _elementwise(lambda v: 1.0 / np.cos(v)),
    "atan2": lambda fr, a, no: np.arctan2(M(a[0]), M(a[1])),
    "mod": b_mod, "min": b_minmax("min"), "max": b_minmax("max"), "sum": b_sum, "cumsum": b_cumsum,
    "find": b_find, "integral": b_integral, "exist": b_exist, "linspace": b_linspace,
    "reshape": b_reshape, "norm": b_norm, "display": b_display, "isempty": lambda fr, a, no: np.array([[M(a[0]).size == 0]]),
    "full": lambda fr, a, no: M(a[0]), "sparse": lambda fr, a, no: M(a[0]),
    "double": lambda fr, a, no: M(a[0]).astype(np.float64),
    "true": lambda fr, a, no: np.array([[True]]), "false": lambda fr, a, no: np.array([[False]]),
    "dot": lambda fr, a, no: np.array([[float(colmajor(M(a[0])) @ colmajor(M(a[1])))]]),
}
