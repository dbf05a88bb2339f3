"""mini_r -- a restricted evaluator for the R language, TEST INFRASTRUCTURE ONLY.

Why it exists.  The reference (DongyueXie/ebpmf.alpha) is 100 % R, ships no golden vectors, and
GNU R is installed neither in this image nor on the GPU box.  The oracle in this directory (vga_numpy.py,
vga_oracle.c) is a hand transcription of the reference, so on its own it cannot rule out a transcription error.
This module closes that particular gap: it parses the reference's OWN source files (R/vga_solver_mat.R,
R/ebpmf_log.R, R/ebpmf_log_controls.R -- read from /root/reference, never copied) and executes their functions
statement by statement with base-R semantics implemented here on NumPy: column-major matrices, vector
recycling, R's operator precedence (`%*%` above `*`, unary minus below `^`), lazy arguments (promises), 1-based /
negative / logical indexing, replacement assignment through `$`, `[[` and `[`, long-double `sum()`/`colSums()`.
tests/golden/make_golden_minir.py uses it to generate tests/golden/minir_reference/, and tests/test_mini_r.py
re-runs the reference through it whenever /root/reference is present.

What it is NOT.  It is not GNU R: libm (`exp`, `log`, `lgamma`) and BLAS summation order differ from an R build
at the 1e-16 level, packages (Matrix, Rfast, flashier, vebpm) are stood in for by the few builtins the hot path
touches (`Diagonal`, `Pmin`, `rowMaxs`, `fitted` ...), and only the syntax and the builtins these three files use
are implemented -- anything else raises `RError("unsupported ...")` instead of guessing.  Numbers produced with it
are "the reference's source text under a re-implemented evaluator"; tests/golden/make_golden.R remains the route
to outputs of GNU R itself.

Only tests/ and tests/golden/ scripts import this module; nothing under ebpmf.alpha_b200/ does (tests/test_abi.py).
"""
from __future__ import annotations

import math
import re

import numpy as np

try:
    from scipy.special import gammaln as _gammaln
except Exception:  # pragma: no cover
    _gammaln = None

LD = np.longdouble


class RError(Exception):
    """An R-level error (stop(), a missing value in `if`, an unsupported construct)."""


class _Break(Exception):
    pass


class _Next(Exception):
    pass


class _Return(Exception):
    def __init__(self, value):
        self.value = value


# ---------------------------------------------------------------------------------------------------------------
# lexer
# ---------------------------------------------------------------------------------------------------------------
_NUM = re.compile(r"(0[xX][0-9a-fA-F]+|(\d+\.?\d*|\.\d+)([eE][+-]?\d+)?)[Li]?")
_ID = re.compile(r"[A-Za-z.][A-Za-z0-9._]*")
_OPS = [":::", "<<-", "%%", "<-", "->", "<=", ">=", "==", "!=", "&&", "||", "|>", "::", "+", "-", "*", "/", "^", "<", ">", "=", "!", "&",
        "|", "~", "?", ":", "$", "@", ",", ";"]


def tokenize(src):
    """-> list of (kind, value, line).  Newlines are tokens only where R treats them as terminators: at top level
    and directly inside `{`; inside `(` and `[` they are white space."""
    toks, stack, i, line, n = [], [], 0, 1, len(src)
    while i < n:
        ch = src[i]
        if ch == "\n":
            if not stack or stack[-1] == "{":
                toks.append(("nl", "\n", line))
            line += 1
            i += 1
        elif ch in " \t\r\f":
            i += 1
        elif ch == "#":
            while i < n and src[i] != "\n":
                i += 1
        elif ch in "\"'":
            j, buf = i + 1, []
            while j < n and src[j] != ch:
                if src[j] == "\\":
                    j += 1
                    buf.append({"n": "\n", "t": "\t", "\\": "\\", "'": "'", '"': '"', "0": "\0"}.get(src[j], "\\" + src[j]))
                else:
                    if src[j] == "\n":
                        line += 1
                    buf.append(src[j])
                j += 1
            if j >= n:
                raise RError("unterminated string at line %d" % line)
            toks.append(("str", "".join(buf), line))
            i = j + 1
        elif ch == "`":
            j = src.index("`", i + 1)
            toks.append(("id", src[i + 1:j], line))
            i = j + 1
        elif ch.isdigit() or (ch == "." and i + 1 < n and src[i + 1].isdigit()):
            m = _NUM.match(src, i)
            t = m.group(0)
            body = t.rstrip("Li")
            toks.append(("num", float(int(body, 16)) if body[:2].lower() == "0x" else float(body), line))
            i = m.end()
        elif ch.isalpha() or ch == ".":
            m = _ID.match(src, i)
            toks.append(("id", m.group(0), line))
            i = m.end()
        elif ch == "%":
            j = src.index("%", i + 1)
            toks.append(("op", src[i:j + 1], line))
            i = j + 1
        elif ch in "({":
            stack.append(ch)
            toks.append(("op", ch, line))
            i += 1
        elif ch == "[":
            if src.startswith("[[", i):
                stack.append("[[")
                toks.append(("op", "[[", line))
                i += 2
            else:
                stack.append("[")
                toks.append(("op", "[", line))
                i += 1
        elif ch in ")}":
            if not stack or stack[-1] != {")": "(", "}": "{"}[ch]:
                raise RError("unbalanced '%s' at line %d" % (ch, line))
            stack.pop()
            toks.append(("op", ch, line))
            i += 1
        elif ch == "]":
            if stack and stack[-1] == "[[" and src.startswith("]]", i):
                stack.pop()
                toks.append(("op", "]]", line))
                i += 2
            elif stack and stack[-1] == "[":
                stack.pop()
                toks.append(("op", "]", line))
                i += 1
            else:
                raise RError("unbalanced ']' at line %d" % line)
        else:
            for op in _OPS:
                if src.startswith(op, i):
                    toks.append(("op", op, line))
                    i += len(op)
                    break
            else:
                raise RError("unexpected character %r at line %d" % (ch, line))
    if stack:
        raise RError("unexpected end of input: open %r" % stack[-1])
    toks.append(("eof", None, line))
    return toks


# ---------------------------------------------------------------------------------------------------------------
# parser (precedence climbing; binding powers follow ?Syntax)
# ---------------------------------------------------------------------------------------------------------------
#           op: (left binding power, right-associative)
_BINARY = {"?": (1, False), "=": (5, True), "<-": (10, True), "<<-": (10, True), "->": (11, False), "~": (15, False),
           "||": (20, False), "|": (20, False), "&&": (25, False), "&": (25, False),
           "==": (35, False), "!=": (35, False), "<": (35, False), ">": (35, False), "<=": (35, False), ">=": (35, False),
           "+": (40, False), "-": (40, False), "*": (45, False), "/": (45, False), "|>": (50, False), ":": (55, False),
           "^": (70, True)}
_SPECIAL_BP = 50      # %any%
_UNARY_MINUS_BP = 60  # below `^` and above `:`  (-2^2 == -4, -1:3 == (-1):3)
_NOT_BP = 30
_POSTFIX_BP = 80


class Parser:
    def __init__(self, src):
        self.t = tokenize(src)
        self.i = 0

    def peek(self):
        return self.t[self.i]

    def next(self):
        tok = self.t[self.i]
        self.i += 1
        return tok

    def is_op(self, v):
        k, val, _ = self.t[self.i]
        return k == "op" and val == v

    def expect(self, v):
        k, val, line = self.next()
        if k != "op" or val != v:
            raise RError("parse error at line %d: expected %r, found %r" % (line, v, val))

    def skip_nl(self):
        while self.t[self.i][0] == "nl":
            self.i += 1

    def program(self):
        out = []
        while True:
            while self.t[self.i][0] == "nl" or self.is_op(";"):
                self.i += 1
            if self.t[self.i][0] == "eof":
                return out
            out.append(self.expr(0))

    def expr(self, rbp):
        left = self.prefix()
        while True:
            k, v, _ = self.peek()
            if k != "op":
                break
            if v in ("(", "[", "[[", "$", "@"):
                if _POSTFIX_BP <= rbp:
                    break
                left = self.postfix(left)
                continue
            if v in ("::", ":::"):
                self.next()
                left = ("name", self.next()[1])      # pkg::name -> name (packages are stood in for by builtins)
                continue
            if v.startswith("%"):
                lbp, right = _SPECIAL_BP, False
            elif v in _BINARY:
                lbp, right = _BINARY[v]
            else:
                break
            if lbp <= rbp:
                break
            self.next()
            self.skip_nl()
            rhs = self.expr(lbp - 1 if right else lbp)
            if v in ("=", "<-", "<<-"):
                left = ("assign", left, rhs, v == "<<-")
            elif v == "->":
                left = ("assign", rhs, left, False)
            else:
                left = ("binop", v, left, rhs)
        return left

    def prefix(self):
        k, v, line = self.next()
        if k == "num":
            return ("num", v)
        if k == "str":
            return ("str", v)
        if k == "id":
            if v == "function":
                return self.function()
            if v == "if":
                self.expect("(")
                c = self.expr(0)
                self.expect(")")
                self.skip_nl()
                a = self.expr(5 - 1)          # `if (c) x = 1` takes the whole assignment
                b = None
                j = self.i
                while self.t[j][0] == "nl":   # `}` newline `else` inside braces
                    j += 1
                if self.t[j][0] == "id" and self.t[j][1] == "else":
                    self.i = j + 1
                    self.skip_nl()
                    b = self.expr(5 - 1)
                return ("if", c, a, b)
            if v == "for":
                self.expect("(")
                var = self.next()[1]
                kk, vv, ll = self.next()
                if vv != "in":
                    raise RError("parse error at line %d: expected 'in'" % ll)
                seq = self.expr(0)
                self.expect(")")
                self.skip_nl()
                return ("for", var, seq, self.expr(5 - 1))
            if v == "while":
                self.expect("(")
                c = self.expr(0)
                self.expect(")")
                self.skip_nl()
                return ("while", c, self.expr(5 - 1))
            if v == "repeat":
                self.skip_nl()
                return ("while", ("name", "TRUE"), self.expr(5 - 1))
            if v == "break":
                return ("break",)
            if v == "next":
                return ("next",)
            if v == "NULL":
                return ("null",)
            return ("name", v)
        if k == "op":
            if v == "(":
                e = self.expr(0)
                self.expect(")")
                return ("paren", e)
            if v == "{":
                body = []
                while True:
                    while self.t[self.i][0] == "nl" or self.is_op(";"):
                        self.i += 1
                    if self.is_op("}"):
                        self.next()
                        return ("block", body)
                    body.append(self.expr(0))
            if v in ("-", "+"):
                return ("unop", v, self.expr(_UNARY_MINUS_BP))
            if v == "!":
                return ("unop", "!", self.expr(_NOT_BP))
            if v == "~":
                return ("unop", "~", self.expr(15))
        raise RError("parse error at line %d: unexpected %r" % (line, v))

    def function(self):
        self.expect("(")
        params = []
        while not self.is_op(")"):
            name = self.next()[1]
            default = None
            if self.is_op("="):
                self.next()
                default = self.expr(5)
            params.append((name, default))
            if self.is_op(","):
                self.next()
        self.expect(")")
        self.skip_nl()
        return ("function", params, self.expr(5 - 1))

    def args(self, close):
        """Arguments of a call or subscript up to `close`; an empty position (x[i, ], x[, j]) is (None, None)."""
        out = []
        if self.is_op(close):
            self.next()
            return out
        while True:
            if self.is_op(",") or self.is_op(close):
                out.append((None, None))
            else:
                k, v, _ = self.peek()
                name = None
                if k in ("id", "str") and self.t[self.i + 1][0] == "op" and self.t[self.i + 1][1] == "=":
                    name = v
                    self.i += 2
                if name is not None and (self.is_op(",") or self.is_op(close)):
                    out.append((name, None))
                else:
                    out.append((name, self.expr(5)))
            if self.is_op(","):
                self.next()
            elif self.is_op(close):
                self.next()
                return out
            else:
                raise RError("parse error at line %d: expected ',' or %r" % (self.peek()[2], close))

    def postfix(self, left):
        v = self.next()[1]
        if v == "(":
            return ("call", left, self.args(")"))
        if v == "[":
            return ("index", left, self.args("]"), False)
        if v == "[[":
            return ("index", left, self.args("]]"), True)
        k, name, line = self.next()
        if k not in ("id", "str"):
            raise RError("parse error at line %d after %s" % (line, v))
        return ("dollar" if v == "$" else "slot", left, name)


def parse(src):
    return Parser(src).program()


# ---------------------------------------------------------------------------------------------------------------
# values
# ---------------------------------------------------------------------------------------------------------------
class RList:
    """An R list (generic vector) with optional names."""

    def __init__(self, vals=None, names=None, cls=None):
        self.vals = list(vals or [])
        self.names = list(names) if names is not None else [None] * len(self.vals)
        self.cls = cls

    def copy(self):
        return RList(self.vals, self.names, self.cls)

    def find(self, name, partial=False):
        for i, nm in enumerate(self.names):
            if nm == name:
                return i
        if partial:                                   # `$` matches partially when unambiguous
            hits = [i for i, nm in enumerate(self.names) if nm and nm.startswith(name)]
            if len(hits) == 1:
                return hits[0]
        return -1

    def get(self, name, partial=False):
        i = self.find(name, partial)
        return self.vals[i] if i >= 0 else None

    def set(self, name, value):
        i = self.find(name)
        if value is None:                             # x$a = NULL removes the element
            if i >= 0:
                del self.vals[i]
                del self.names[i]
            return
        if i >= 0:
            self.vals[i] = value
        else:
            self.vals.append(value)
            self.names.append(name)

    def __len__(self):
        return len(self.vals)


class REnvObj:
    """An environment as a value (`new.env()`): `$` reads, `$<-` writes IN PLACE (reference semantics)."""

    def __init__(self):
        self.vars = {}


class RDiag:
    """Matrix::Diagonal(n, x): only ever an operand of %*% here (exact column / row scaling, no summation)."""

    def __init__(self, n, x):
        self.n = int(n)
        self.x = np.resize(np.asarray(x, dtype=np.float64).ravel(), self.n) if np.size(x) else np.ones(self.n)


class RSparse:
    """A dgCMatrix stand-in: dense storage (arithmetic on it is elementwise, as for the Matrix classes) plus the
    `@x` slot = the non-zeros in column-major order, which is all sum_lfactorial_sparseMat reads."""

    def __init__(self, dense):
        self.dense = np.asfortranarray(np.asarray(dense, dtype=np.float64))

    def slot(self, name):
        d = self.dense
        nz = d.ravel(order="F")
        if name == "x":
            return nz[nz != 0]
        if name == "Dim":
            return np.array(d.shape, dtype=np.float64)
        raise RError("unsupported slot @%s" % name)


class Closure:
    def __init__(self, params, body, env, name=None):
        self.params, self.body, self.env, self.name = params, body, env, name


class Builtin:
    def __init__(self, fn, name, rawnames=False):
        self.fn, self.name, self.rawnames = fn, name, rawnames      # rawnames: argument names are data (list(), c())


class Promise:
    __slots__ = ("expr", "env", "val", "done")

    def __init__(self, expr, env):
        self.expr, self.env, self.val, self.done = expr, env, None, False


class Env:
    def __init__(self, parent=None):
        self.vars = {}
        self.parent = parent

    def lookup(self, name):
        e = self
        while e is not None:
            if name in e.vars:
                return e, e.vars[name]
            e = e.parent
        return None, None

    def set(self, name, value):
        self.vars[name] = value

    def set_super(self, name, value):
        e = self.parent
        while e is not None:
            if name in e.vars:
                e.vars[name] = value
                return
            if e.parent is None:
                e.vars[name] = value
                return
            e = e.parent


_MISSING = object()


def num(x):
    return np.atleast_1d(np.asarray(x, dtype=np.float64))


def is_str(v):
    return isinstance(v, np.ndarray) and v.dtype == object


def strs(*s):
    a = np.empty(len(s), dtype=object)
    for i, x in enumerate(s):
        a[i] = x
    return a


def dense(v):
    """The numeric array behind a value (RSparse -> its dense image)."""
    if isinstance(v, RSparse):
        return v.dense
    if isinstance(v, np.ndarray):
        return v if type(v) is np.ndarray else v.view(np.ndarray)     # a flagged subclass (integer storage) is plain data
    if isinstance(v, (int, float, bool, np.floating, np.integer, np.bool_)):
        return np.atleast_1d(np.asarray(v))
    raise RError("non-numeric argument (%s)" % type(v).__name__)


def flat(a):
    return a.ravel(order="F")


def rlen(v):
    if v is None:
        return 0
    if isinstance(v, RList):
        return len(v)
    if isinstance(v, (Closure, Builtin)):
        return 1
    return int(dense(v).size)


def scalar(v, what="argument"):
    a = dense(v)
    if a.size < 1:
        raise RError("%s of length zero" % what)
    return a.ravel(order="F")[0]


def as_logical_scalar(v, what="argument"):
    x = scalar(v, what)
    if isinstance(x, (float, np.floating)) and math.isnan(x):
        raise RError("missing value where TRUE/FALSE needed")
    if isinstance(x, str):
        raise RError("argument is not interpretable as logical")
    return bool(x)


def r_string(v):
    x = scalar(v)
    return x if isinstance(x, str) else fmt_num(x)


def fmt_num(x):
    if isinstance(x, (bool, np.bool_)):
        return "TRUE" if x else "FALSE"
    x = float(x)
    if math.isnan(x):
        return "NA"
    if math.isinf(x):
        return "Inf" if x > 0 else "-Inf"
    if x == int(x) and abs(x) < 1e15:
        return str(int(x))
    return "%.15g" % x


# ---------------------------------------------------------------------------------------------------------------
# arithmetic with R's recycling rule
# ---------------------------------------------------------------------------------------------------------------
def _result_dim(a, b):
    da = a.shape if a.ndim == 2 else None
    db = b.shape if b.ndim == 2 else None
    if da and db and da != db:
        raise RError("non-conformable arrays")
    return da or db


def recycle2(a, b):
    """Two arrays brought to a common shape by R's rule: the shorter one is repeated along the column-major
    order of the longer; a matrix operand gives the result its dim."""
    if a.shape == b.shape:
        return a, b, None
    dim = _result_dim(a, b)
    if a.size == 1 or b.size == 1:
        if dim is None:
            return a.ravel(), b.ravel(), None
        return (a if a.ndim == 2 else a.ravel()[0]), (b if b.ndim == 2 else b.ravel()[0]), None
    if a.size == 0 or b.size == 0:
        return a.ravel()[:0], b.ravel()[:0], None
    n = max(a.size, b.size)
    if dim is not None and n != dim[0] * dim[1]:
        raise RError("dims do not match the length of object")
    fa = flat(a) if a.size == n else np.resize(flat(a), n)
    fb = flat(b) if b.size == n else np.resize(flat(b), n)
    return fa, fb, dim


def _finish(r, dim):
    r = np.asarray(r)
    if dim is not None:
        return np.reshape(r, dim, order="F")
    return np.atleast_1d(r)


def r_pow(x, y):
    """R_pow (arithmetic.c): x^2 is x*x, everything else is C pow()."""
    if np.size(y) == 1 and float(np.ravel(y)[0]) == 2.0:
        return x * x
    return np.power(x, y)


def arith(op, a, b):
    a, b = dense(a), dense(b)
    if a.dtype == object or b.dtype == object:
        raise RError("non-numeric argument to binary operator")
    a = a.astype(np.float64, copy=False)
    b = b.astype(np.float64, copy=False)
    x, y, dim = recycle2(a, b)
    with np.errstate(all="ignore"):
        if op == "+":
            r = x + y
        elif op == "-":
            r = x - y
        elif op == "*":
            r = x * y
        elif op == "/":
            r = x / y
        elif op == "^":
            r = r_pow(x, y)
        elif op == "%%":
            r = np.where(np.isinf(y) & np.isfinite(x), np.where((x >= 0) == (y > 0), x, y), x - np.floor(x / y) * y)
        elif op == "%/%":
            r = np.floor(x / y)
        else:
            raise RError("unsupported operator %s" % op)
    return _finish(r, dim)


def compare(op, a, b):
    a, b = dense(a), dense(b)
    if a.dtype == object or b.dtype == object:
        x, y, dim = recycle2(np.asarray(a, dtype=object), np.asarray(b, dtype=object))
        if op not in ("==", "!="):
            raise RError("unsupported comparison of strings")
        r = np.asarray(x == y) if op == "==" else np.asarray(x != y)
        return _finish(r.astype(bool), dim)
    x, y, dim = recycle2(a.astype(np.float64, copy=False), b.astype(np.float64, copy=False))
    with np.errstate(invalid="ignore"):
        r = {"<": np.less, ">": np.greater, "<=": np.less_equal, ">=": np.greater_equal, "==": np.equal,
             "!=": np.not_equal}[op](x, y)
    na = np.isnan(x) | np.isnan(y)
    if np.any(na):                                   # a logical NA is carried as NaN in a float array
        r = np.where(na, np.nan, np.asarray(r, dtype=np.float64))
    return _finish(r, dim)


def logic(op, a, b):
    x, y, dim = recycle2(dense(a).astype(np.float64), dense(b).astype(np.float64))
    nx, ny = np.isnan(x), np.isnan(y)
    tx, ty = (x != 0) & ~nx, (y != 0) & ~ny
    if op == "&":
        r = tx & ty
        na = (nx | ny) & ~((~tx & ~nx) | (~ty & ~ny))
    else:
        r = tx | ty
        na = (nx | ny) & ~r
    if np.any(na):
        r = np.where(na, np.nan, np.asarray(r, dtype=np.float64))
    return _finish(r, dim)


def matmul(a, b):
    if isinstance(b, RDiag):
        A = dense(a)
        if A.ndim != 2 or A.shape[1] != b.n:
            raise RError("non-conformable arguments")
        return np.asfortranarray(A * b.x[None, :])
    if isinstance(a, RDiag):
        B = dense(b)
        if B.ndim != 2 or B.shape[0] != a.n:
            raise RError("non-conformable arguments")
        return np.asfortranarray(a.x[:, None] * B)
    A, B = dense(a).astype(np.float64), dense(b).astype(np.float64)
    if A.ndim == 1 and B.ndim == 1:
        A, B = (A[None, :], B[:, None]) if A.size == B.size else (A[:, None], B[None, :])
    elif A.ndim == 1:
        A = A[None, :] if B.shape[0] == A.size else A[:, None]
    elif B.ndim == 1:
        B = B[:, None] if A.shape[1] == B.size else B[None, :]
    if A.shape[1] != B.shape[0]:
        raise RError("non-conformable arguments")
    return np.asfortranarray(A @ B)


# ---------------------------------------------------------------------------------------------------------------
# indexing
# ---------------------------------------------------------------------------------------------------------------
def resolve_index(idx, n, names=None, extend=False):
    """An R subscript -> 0-based positions.  Positions past the end are kept (reads give NA, writes extend)."""
    if idx is _MISSING:
        return np.arange(n)
    a = dense(idx)
    if a.dtype == object:
        if names is None:
            raise RError("subscript out of bounds")
        out = []
        for s in a.ravel():
            out.append(names.index(s) if s in names else (n + len([o for o in out if o >= n]) if extend else -1))
            if out[-1] == -1:
                raise RError("subscript out of bounds: %s" % s)
        return np.asarray(out, dtype=np.int64)
    if a.dtype == bool:
        m = np.resize(flat(a), max(n, a.size)) if a.size else np.zeros(0, bool)
        return np.nonzero(m)[0]
    f = flat(a).astype(np.float64)
    if f.size and np.all(f[~np.isnan(f)] <= 0) and np.any(f < 0):
        keep = np.ones(n, dtype=bool)
        drop = (-f[f < 0]).astype(np.int64) - 1
        keep[drop[drop < n]] = False
        return np.nonzero(keep)[0]
    if np.any(f < 0):
        raise RError("can't mix positive and negative subscripts")
    f = f[f != 0]
    if np.any(np.isnan(f)):
        raise RError("NA subscripts are not supported")
    return f.astype(np.int64) - 1


def take(vec, pos):
    """vec[pos] with NA for positions past the end."""
    n = vec.size
    if pos.size and pos.max(initial=-1) >= n:
        out = np.full(pos.size, np.nan) if vec.dtype != object else np.array([None] * pos.size, dtype=object)
        ok = pos < n
        out[ok] = vec[pos[ok]]
        return out
    return vec[pos]


def index_get(obj, idx, drop=True):
    if obj is None:
        return None
    if hasattr(obj, "r_index_get"):                    # an object that brings its own `[` (oracle/minir_io.py's data frame)
        return obj.r_index_get(idx, drop)
    if isinstance(obj, RList):
        pos = resolve_index(idx[0], len(obj), obj.names)
        return RList([obj.vals[i] if i < len(obj) else None for i in pos], [obj.names[i] if i < len(obj) else None for i in pos],
                     obj.cls)
    a = dense(obj)
    if len(idx) == 1:
        i0 = idx[0]
        if i0 is not _MISSING and isinstance(i0, np.ndarray) and i0.ndim == 2 and i0.dtype == bool and i0.shape == a.shape:
            return flat(a)[flat(i0)]
        return take(flat(a), resolve_index(i0, a.size))
    if len(idx) == 2:
        if a.ndim != 2:
            raise RError("incorrect number of dimensions")
        r = resolve_index(idx[0], a.shape[0])
        c = resolve_index(idx[1], a.shape[1])
        if (r.size and r.max() >= a.shape[0]) or (c.size and c.max() >= a.shape[1]):
            raise RError("subscript out of bounds")
        sub = np.asfortranarray(a[np.ix_(r, c)])
        if drop and (sub.shape[0] == 1 or sub.shape[1] == 1):
            return flat(sub).copy()
        return sub
    raise RError("unsupported number of subscripts")


def index_get2(obj, i):
    """obj[[i]]"""
    if isinstance(obj, RList):
        x = scalar(i)
        if isinstance(x, str):
            k = obj.find(x)
            return obj.vals[k] if k >= 0 else None
        k = int(x) - 1
        if k < 0 or k >= len(obj):
            raise RError("subscript out of bounds")
        return obj.vals[k]
    a = dense(obj)
    k = int(scalar(i)) - 1
    if k < 0 or k >= a.size:
        raise RError("subscript out of bounds")
    return flat(a)[k:k + 1].copy()


def index_set(obj, idx, value):
    """`obj[idx] <- value` as a pure function of obj."""
    if hasattr(obj, "r_index_set"):
        return obj.r_index_set(idx, value)
    if isinstance(obj, RList):
        new = obj.copy()
        pos = resolve_index(idx[0], len(new), new.names, extend=True)
        vals = value.vals if isinstance(value, RList) else [value]
        for k, i in enumerate(pos):
            while i >= len(new.vals):
                new.vals.append(None)
                new.names.append(None)
            new.vals[i] = vals[k % len(vals)]
        return new
    v = dense(value)
    if obj is None:
        a = np.zeros(0, dtype=v.dtype if v.dtype == object else np.float64)
    else:
        a = dense(obj)
    if v.dtype != object and a.dtype != object:
        a = a.astype(np.float64)          # a copy: R's value semantics
        v = v.astype(np.float64, copy=False)
    else:
        a = a.astype(object)
    if len(idx) == 1:
        i0 = idx[0]
        if i0 is not _MISSING and isinstance(i0, np.ndarray) and i0.ndim == 2 and i0.dtype == bool and i0.shape == a.shape:
            pos = np.nonzero(flat(i0))[0]
        else:
            pos = resolve_index(i0, a.size)
        if pos.size == 0:
            return a
        if v.size == 0:
            raise RError("replacement has length zero")
        shape = a.shape
        f = flat(a).copy()
        if pos.max() >= f.size:
            ext = np.full(pos.max() + 1, np.nan) if f.dtype != object else np.array([None] * (pos.max() + 1), dtype=object)
            ext[:f.size] = f
            f, shape = ext, None
        f[pos] = np.resize(flat(v), pos.size)
        return f if shape is None or len(shape) == 1 else np.reshape(f, shape, order="F")
    if len(idx) == 2:
        if a.ndim != 2:
            raise RError("incorrect number of subscripts on matrix")
        r = resolve_index(idx[0], a.shape[0])
        c = resolve_index(idx[1], a.shape[1])
        if (r.size and r.max() >= a.shape[0]) or (c.size and c.max() >= a.shape[1]):
            raise RError("subscript out of bounds")
        a = np.array(a, order="F")
        if r.size * c.size:
            if v.size == 0 or (r.size * c.size) % v.size:
                raise RError("number of items to replace is not a multiple of replacement length")
            a[np.ix_(r, c)] = np.reshape(np.resize(flat(v), r.size * c.size), (r.size, c.size), order="F")
        return a
    raise RError("unsupported number of subscripts")


def index_set2(obj, i, value):
    """`obj[[i]] <- value`"""
    if obj is None:
        obj = RList()
    if isinstance(obj, RList):
        new = obj.copy()
        x = scalar(i)
        if isinstance(x, str):
            new.set(x, value)
            return new
        k = int(x) - 1
        if value is None:
            if k < len(new):
                del new.vals[k]
                del new.names[k]
            return new
        while k >= len(new.vals):
            new.vals.append(None)
            new.names.append(None)
        new.vals[k] = value
        return new
    return index_set(obj, [i], value)


# ---------------------------------------------------------------------------------------------------------------
# the evaluator
# ---------------------------------------------------------------------------------------------------------------
class Interp:
    def __init__(self, echo=None):
        self.globalenv = Env()
        self.echo = echo if echo is not None else (lambda s: None)     # print()/cat() sink
        self.warnings = []
        g = self.globalenv
        g.set("TRUE", np.array([True]))
        g.set("FALSE", np.array([False]))
        g.set("T", np.array([True]))
        g.set("F", np.array([False]))
        g.set("Inf", num(np.inf))
        g.set("NaN", num(np.nan))
        g.set("NA", num(np.nan))
        g.set("NA_real_", num(np.nan))
        g.set("pi", num(math.pi))
        _install_builtins(self)

    # -- public -------------------------------------------------------------------------------------------------
    def source(self, path):
        with open(path) as f:
            return self.run(f.read())

    def run(self, src, env=None):
        env = env or self.globalenv
        val = None
        for e in parse(src):
            val = self.eval(e, env)
        return val

    def call(self, name, *args, **kwargs):
        fn = self.globalenv.lookup(name)[1]
        if fn is None:
            raise RError("could not find function \"%s\"" % name)
        return self.apply(fn, [(None, to_r(a)) for a in args] + [(k, to_r(v)) for k, v in kwargs.items()])

    def def_builtin(self, name, fn, rawnames=False):
        self.globalenv.set(name, Builtin(fn, name, rawnames))

    def set(self, name, value):
        self.globalenv.set(name, to_r(value))

    def get(self, name):
        return self.force(self.globalenv.lookup(name)[1])

    # -- evaluation ---------------------------------------------------------------------------------------------
    def force(self, v):
        if isinstance(v, Promise):
            if not v.done:
                v.val = self.eval(v.expr, v.env)
                v.done = True
                v.env = None
            return v.val
        return v

    def lookup(self, name, env):
        e, v = env.lookup(name)
        if e is None:
            raise RError("object '%s' not found" % name)
        if v is _MISSING:
            raise RError("argument \"%s\" is missing, with no default" % name)
        return self.force(v)

    def lookup_fn(self, name, env):
        e = env
        while e is not None:
            if name in e.vars:
                v = self.force(e.vars[name])
                if isinstance(v, (Closure, Builtin)):
                    return v
            e = e.parent
        raise RError("could not find function \"%s\"" % name)

    def eval(self, e, env):
        k = e[0]
        if k == "num":
            return num(e[1])
        if k == "str":
            return strs(e[1])
        if k == "name":
            return self.lookup(e[1], env)
        if k == "null":
            return None
        if k == "paren":
            return self.eval(e[1], env)
        if k == "block":
            val = None
            for s in e[1]:
                val = self.eval(s, env)
            return val
        if k == "binop":
            return self.binop(e, env)
        if k == "unop":
            v = self.eval(e[2], env)
            if e[1] == "-":
                return np.negative(dense(v).astype(np.float64))
            if e[1] == "+":
                return v
            if e[1] == "!":
                a = dense(v).astype(np.float64)
                r = (a == 0)
                return np.where(np.isnan(a), np.nan, r.astype(np.float64)) if np.any(np.isnan(a)) else r
            raise RError("unsupported unary operator %s" % e[1])
        if k == "assign":
            val = self.eval(e[2], env)
            if isinstance(val, Closure) and val.name is None and e[1][0] == "name":
                val.name = e[1][1]
            self.assign(e[1], val, env, e[3])
            return val
        if k == "call":
            return self.call_expr(e, env)
        if k == "index":
            obj = self.eval(e[1], env)
            drop, idx = True, []
            for nm, ex in e[2]:
                if nm == "drop":
                    drop = as_logical_scalar(self.eval(ex, env))
                elif nm == "exact":
                    pass
                else:
                    idx.append(_MISSING if ex is None else self.eval(ex, env))
            if e[3]:
                if len(idx) != 1:
                    raise RError("unsupported [[ with %d subscripts" % len(idx))
                return index_get2(obj, idx[0])
            return index_get(obj, idx, drop)
        if k == "dollar":
            obj = self.eval(e[1], env)
            if obj is None:
                return None
            if isinstance(obj, REnvObj):
                return obj.vars.get(e[2])
            if hasattr(obj, "r_dollar"):
                return obj.r_dollar(e[2])
            if not isinstance(obj, RList):
                raise RError("$ operator is invalid for atomic vectors")
            return obj.get(e[2], partial=True)
        if k == "slot":
            obj = self.eval(e[1], env)
            if isinstance(obj, RSparse):
                return obj.slot(e[2])
            raise RError("no slot of name \"%s\" for this object" % e[2])
        if k == "if":
            if as_logical_scalar(self.eval(e[1], env), "argument"):
                return self.eval(e[2], env)
            return self.eval(e[3], env) if e[3] is not None else None
        if k == "for":
            seq = self.eval(e[2], env)
            items = seq.vals if isinstance(seq, RList) else ([] if seq is None else [flat(dense(seq))[i:i + 1] for i in range(rlen(seq))])
            for it in items:
                env.set(e[1], it)
                try:
                    self.eval(e[3], env)
                except _Break:
                    break
                except _Next:
                    continue
            return None
        if k == "while":
            while as_logical_scalar(self.eval(e[1], env)):
                try:
                    self.eval(e[2], env)
                except _Break:
                    break
                except _Next:
                    continue
            return None
        if k == "break":
            raise _Break()
        if k == "next":
            raise _Next()
        if k == "function":
            return Closure(e[1], e[2], env)
        raise RError("unsupported syntax node %s" % k)

    def binop(self, e, env):
        op = e[1]
        if op in ("&&", "||"):
            a = as_logical_scalar(self.eval(e[2], env), "invalid 'x'")
            if op == "&&" and not a:
                return np.array([False])
            if op == "||" and a:
                return np.array([True])
            return np.array([as_logical_scalar(self.eval(e[3], env), "invalid 'y'")])
        a = self.eval(e[2], env)
        b = self.eval(e[3], env)
        if op in ("+", "-", "*", "/", "^", "%%", "%/%"):
            return arith(op, a, b)
        if op in ("<", ">", "<=", ">=", "==", "!="):
            return compare(op, a, b)
        if op in ("&", "|"):
            return logic(op, a, b)
        if op == ":":
            lo, hi = float(scalar(a)), float(scalar(b))
            if math.isnan(lo) or math.isnan(hi):
                raise RError("NA/NaN argument")
            n = int(math.floor(abs(hi - lo) + 1e-10)) + 1
            return lo + np.arange(n, dtype=np.float64) * (1.0 if hi >= lo else -1.0)
        if op == "%*%":
            return matmul(a, b)
        if op == "%in%":
            x, t = dense(a), dense(b) if b is not None else np.zeros(0)
            return np.isin(flat(x), flat(t))
        if op == "%o%":
            return np.asfortranarray(np.outer(flat(dense(a)), flat(dense(b))))
        raise RError("unsupported operator %s" % op)

    # -- assignment ---------------------------------------------------------------------------------------------
    def assign(self, target, value, env, superassign=False):
        k = target[0]
        if k in ("name", "str"):
            if superassign:
                env.set_super(target[1], value)
            else:
                env.set(target[1], value)
            return
        if k == "paren":
            return self.assign(target[1], value, env, superassign)
        if k in ("index", "dollar", "slot", "call"):
            inner = target[1] if k != "call" else (target[2][0][1] if target[2] else None)
            if inner is None:
                raise RError("invalid assignment target")
            cur = self.current(inner, env)
            if k == "dollar" and isinstance(cur, REnvObj):
                cur.vars[target[2]] = value
                return
            if k == "dollar":
                new = RList() if cur is None else cur
                if not isinstance(new, RList):
                    raise RError("$<- on an atomic vector is not supported")
                new = new.copy()
                new.set(target[2], value)
            elif k == "index":
                idx = [_MISSING if ex is None else self.eval(ex, env) for nm, ex in target[2] if nm not in ("drop", "exact")]
                new = index_set2(cur, idx[0], value) if target[3] else index_set(cur, idx, value)
            elif k == "call":
                fname = target[1][1] if target[1][0] == "name" else None
                fn = self.lookup_fn("%s<-" % fname, env)
                extra = [(nm, self.eval(ex, env)) for nm, ex in target[2][1:]]
                new = self.apply(fn, [(None, cur)] + extra + [("value", value)])
            else:
                raise RError("@<- is not supported")
            return self.assign(inner, new, env, superassign)
        raise RError("invalid assignment target (%s)" % k)

    def current(self, expr, env):
        """The value an inner replacement target has now (NULL if the name does not exist yet)."""
        if expr[0] == "name":
            e, v = env.lookup(expr[1])
            return None if e is None else self.force(v)
        return self.eval(expr, env)

    # -- calls --------------------------------------------------------------------------------------------------
    def call_expr(self, e, env):
        fexpr = e[1]
        if fexpr[0] == "name":
            name = fexpr[1]
            if name == "return":
                raise _Return(self.eval(e[2][0][1], env) if e[2] and e[2][0][1] is not None else None)
            if name == "missing":
                ee, v = env.lookup(e[2][0][1][1])
                return np.array([v is _MISSING])
            if name in ("quote", "substitute"):
                raise RError("unsupported: %s" % name)
            if name in ("suppressWarnings", "suppressMessages", "invisible", "suppressPackageStartupMessages"):
                return self.eval(e[2][0][1], env) if e[2] else None
            if name in ("library", "require", "requireNamespace"):
                return None
            if name == "rm":
                for nm, ex in e[2]:
                    if ex is not None and ex[0] == "name":
                        env.vars.pop(ex[1], None)
                return None
            if name == "on.exit":
                return None
            fn = self.lookup_fn(name, env)
        else:
            fn = self.eval(fexpr, env)
            if not isinstance(fn, (Closure, Builtin)):
                raise RError("attempt to apply non-function")
        args = []
        for nm, ex in e[2]:
            if ex is not None and ex == ("name", "..."):          # forward the caller's dots
                args += env.lookup("...")[1] or []
            elif ex is None:
                args.append((nm, _MISSING))
            else:
                args.append((nm, self.eval(ex, env) if isinstance(fn, Builtin) else Promise(ex, env)))
        return self.apply(fn, args)

    def apply(self, fn, args):
        """args: list of (name or None, value / Promise / _MISSING)."""
        if isinstance(fn, Builtin):
            pos = [self.force(v) for nm, v in args if nm is None]
            kw = {nm: self.force(v) for nm, v in args if nm is not None}
            return fn.fn(self, *pos, **(kw if fn.rawnames else {k.replace(".", "_"): v for k, v in kw.items()}))
        if not isinstance(fn, Closure):
            raise RError("attempt to apply non-function")
        fenv = Env(fn.env)
        names = [p for p, _ in fn.params]
        bound = {}
        rest = []
        for nm, v in args:                       # 1. exact names
            if nm is not None and nm in names and nm not in bound:
                bound[nm] = v
            else:
                rest.append((nm, v))
        rest2 = []
        for nm, v in rest:                       # 2. partial names
            if nm is not None:
                hits = [p for p in names if p.startswith(nm) and p not in bound and p != "..."]
                if len(hits) == 1:
                    bound[hits[0]] = v
                    continue
                if "..." not in names:
                    raise RError("unused argument (%s) in call to %s" % (nm, fn.name or "function"))
            rest2.append((nm, v))
        free = [p for p in names if p not in bound and p != "..."]
        dots = []
        for nm, v in rest2:                      # 3. positions
            if nm is None and free:
                bound[free.pop(0)] = v
            elif "..." in names:
                dots.append((nm, v))
            else:
                raise RError("unused argument in call to %s" % (fn.name or "function"))
        for p, default in fn.params:
            if p == "...":
                fenv.set("...", dots)
            elif p in bound and bound[p] is not _MISSING:
                fenv.set(p, bound[p])
            elif default is not None:
                fenv.set(p, Promise(default, fenv))
            else:
                fenv.set(p, _MISSING)
        try:
            val = self.eval(fn.body, fenv)
        except _Return as r:
            val = r.value
        return val


def to_r(v):
    """Python / NumPy value -> interpreter value."""
    if v is None or isinstance(v, (RList, RDiag, RSparse, Closure, Builtin, REnvObj)) or hasattr(v, "sexp"):
        return v                                           # (an object with .sexp: an external pointer of tests/r_mock.py)
    if isinstance(v, str):
        return strs(v)
    if hasattr(v, "toarray"):                          # scipy.sparse -> the dgCMatrix stand-in
        return RSparse(v.toarray())
    if isinstance(v, dict):
        return RList([to_r(x) for x in v.values()], list(v.keys()))
    if isinstance(v, (list, tuple)) and any(isinstance(x, (np.ndarray, dict, list, RList)) or x is None for x in v):
        return RList([to_r(x) for x in v])
    a = np.asarray(v)
    if a.dtype == bool:
        return np.atleast_1d(a)
    if a.dtype.kind in "US":
        return strs(*[str(x) for x in a.ravel()])
    a = np.atleast_1d(np.asarray(a, dtype=np.float64))
    if a.ndim == 2:
        a = np.asfortranarray(a)
    return a


def to_py(v):
    """Interpreter value -> NumPy / dict."""
    if isinstance(v, RList):
        if all(nm for nm in v.names) and v.names:
            return {nm: to_py(x) for nm, x in zip(v.names, v.vals)}
        return [to_py(x) for x in v.vals]
    if isinstance(v, RSparse):
        return v.dense
    return v


# ---------------------------------------------------------------------------------------------------------------
# builtins: the part of base R (and of Matrix / Rfast / matrixStats) the three reference files call
# ---------------------------------------------------------------------------------------------------------------
def _ld_sum(a):
    return float(np.sum(np.asarray(a, dtype=LD)))


def _math1(f):
    def g(it, x):
        a = dense(x).astype(np.float64)
        with np.errstate(all="ignore"):
            return f(a)
    return g


def _matrix(it, data=np.nan, nrow=None, ncol=None, byrow=False, dimnames=None):
    d = flat(dense(data)) if data is not None else np.zeros(0)
    byrow = as_logical_scalar(byrow)
    if nrow is None and ncol is None:
        nr, nc = d.size, 1
    elif nrow is None:
        nc = int(scalar(ncol))
        nr = int(math.ceil(d.size / nc)) if nc else 0
    elif ncol is None:
        nr = int(scalar(nrow))
        nc = int(math.ceil(d.size / nr)) if nr else 0
    else:
        nr, nc = int(scalar(nrow)), int(scalar(ncol))
    if d.size == 0:
        d = np.full(1, np.nan)
    full = np.resize(d, nr * nc)
    if byrow:
        return np.asfortranarray(np.reshape(full, (nr, nc), order="C"))
    return np.reshape(full, (nr, nc), order="F")


def _c(it, *args, **kw):
    items = list(args) + list(kw.values())
    names = [None] * len(args) + list(kw.keys())
    if any(isinstance(x, (RList, Closure, Builtin)) for x in items):
        out = RList()
        for nm, x in zip(names, items):
            if isinstance(x, RList):
                out.vals += x.vals
                out.names += x.names
            elif x is not None:
                out.vals.append(x)
                out.names.append(nm)
        return out
    parts = [flat(dense(x)) for x in items if x is not None]
    if not parts:
        return None
    if any(p.dtype == object for p in parts):
        return strs(*[s if isinstance(s, str) else fmt_num(s) for p in parts for s in p])
    if all(p.dtype == bool for p in parts):
        return np.concatenate(parts)
    return np.concatenate([p.astype(np.float64) for p in parts])


def _list(it, *args, **kw):
    return RList(list(args) + list(kw.values()), [None] * len(args) + list(kw.keys()))


def _bind(axis):
    def g(it, *args, **kw):
        mats = []
        for x in args:
            if x is None:
                continue
            a = dense(x).astype(np.float64)
            mats.append(a if a.ndim == 2 else (a[:, None] if axis == 1 else a[None, :]))
        if not mats:
            return None
        other = 1 - axis
        m = max(a.shape[other] for a in mats)
        out = []
        for a in mats:
            if a.shape[other] != m:                                  # a vector argument is recycled
                v = np.resize(flat(a), m)
                a = v[:, None] if axis == 1 else v[None, :]
            out.append(a)
        return np.asfortranarray(np.concatenate(out, axis=axis))
    return g


def _rep(it, x, times=None, each=None, length_out=None):
    a = flat(dense(x))
    if each is not None:
        a = np.repeat(a, int(scalar(each)))
    if times is not None:
        t = dense(times)
        a = np.tile(a, int(scalar(t))) if t.size == 1 else np.repeat(a, flat(t).astype(int))
    if length_out is not None:
        a = np.resize(a, int(scalar(length_out)))
    return a


def _seq(it, from_=None, to=None, by=None, length_out=None, **kw):
    if "from" in kw:
        from_ = kw["from"]
    if from_ is not None and to is None and by is None and length_out is None:
        a = dense(from_)
        n = a.size if a.size != 1 else int(scalar(a))
        return np.arange(1, n + 1, dtype=np.float64)
    lo = float(scalar(from_)) if from_ is not None else 1.0
    if length_out is not None and to is not None:
        n = int(scalar(length_out))
        hi = float(scalar(to))
        if n == 1:
            return num(lo)
        out = lo + np.arange(n, dtype=np.float64) * ((hi - lo) / (n - 1))   # seq.int: from + i*by ...
        out[-1] = hi                                                       # ... with the last value set to `to`
        return out
    hi = float(scalar(to))
    step = float(scalar(by)) if by is not None else (1.0 if hi >= lo else -1.0)
    n = int(math.floor((hi - lo) / step + 1e-10)) + 1
    return lo + np.arange(max(n, 0), dtype=np.float64) * step


def _cut(it, x, breaks, labels=None, right=True, include_lowest=False, **kw):
    """base::cut.default for a break COUNT or a vector of breaks, labels = FALSE (integer codes)."""
    if labels is None or as_logical_scalar(labels):
        raise RError("unsupported: cut() with labels")
    xv = flat(dense(x)).astype(np.float64)
    b = flat(dense(breaks)).astype(np.float64)
    if b.size == 1:
        if math.isnan(b[0]) or b[0] < 2:
            raise RError("invalid number of intervals")
        nb = int(b[0] + 1)
        rx = (np.nanmin(xv), np.nanmax(xv))
        dx = rx[1] - rx[0]
        if dx == 0:
            dx = abs(rx[0]) if rx[0] != 0 else 1.0
            b = _seq(it, num(rx[0] - dx / 1000), num(rx[1] + dx / 1000), length_out=num(nb))
        else:
            b = _seq(it, num(rx[0]), num(rx[1]), length_out=num(nb))
            b[0] = rx[0] - dx / 1000
            b[-1] = rx[1] + dx / 1000
    else:
        b = np.sort(b)
    if as_logical_scalar(right):                                            # (b_k, b_{k+1}]
        code = np.searchsorted(b, xv, side="left").astype(np.float64)
    else:
        code = np.searchsorted(b, xv, side="right").astype(np.float64)
    code[(code < 1) | (code > b.size - 1)] = np.nan
    return code


def _split(it, x, f, **kw):
    xv, fv = flat(dense(x)), flat(dense(f))
    if isinstance(x, np.ndarray) and x.ndim == 2:
        raise RError("unsupported: split() of a matrix")
    levels = sorted(set(v for v in fv.tolist() if not (isinstance(v, float) and math.isnan(v))))
    return RList([xv[fv == lv] for lv in levels], [fmt_num(lv) if not isinstance(lv, str) else lv for lv in levels])


def _lapply(it, X, FUN, *extra, **kw):
    items = X.vals if isinstance(X, RList) else [flat(dense(X))[i:i + 1] for i in range(rlen(X))]
    names = X.names if isinstance(X, RList) else [None] * len(items)
    return RList([it.apply(FUN, [(None, v)] + [(None, a) for a in extra] + list(kw.items())) for v in items], names)


def _sapply(it, X, FUN, *extra, **kw):
    res = _lapply(it, X, FUN, *extra, **kw)
    if all(isinstance(v, np.ndarray) and v.size == 1 for v in res.vals):
        return _c(it, *res.vals)
    return res


def _apply(it, X, MARGIN, FUN, *extra):
    a = dense(X)
    m = int(scalar(MARGIN))
    cols = [it.apply(FUN, [(None, (a[i, :] if m == 1 else a[:, i]).copy())] + [(None, e) for e in extra]) for i in range(a.shape[m - 1])]
    if all(np.size(c) == 1 for c in cols):
        return np.concatenate([flat(dense(c)) for c in cols])
    return np.asfortranarray(np.stack([flat(dense(c)) for c in cols], axis=1))


def _modify_list(it, x, val, keep_null=False):
    """utils::modifyList for flat lists (nested lists are replaced recursively like R does)."""
    out = x.copy()
    keep = as_logical_scalar(keep_null)
    for nm, v in zip(val.names, val.vals):
        if nm is None:
            continue
        i = out.find(nm)
        if isinstance(v, RList) and i >= 0 and isinstance(out.vals[i], RList):
            out.vals[i] = _modify_list(it, out.vals[i], v, keep_null)
        elif v is None and not keep:
            out.set(nm, None)
        elif i >= 0:
            out.vals[i] = v
        else:
            out.vals.append(v)
            out.names.append(nm)
    return out


def _paste(sep_default):
    def g(it, *args, sep=None, collapse=None):
        s = sep_default if sep is None else r_string(sep)
        parts = [[x if isinstance(x, str) else fmt_num(x) for x in flat(dense(a))] for a in args if a is not None and rlen(a)]
        if not parts:
            return strs("")
        n = max(len(p) for p in parts)
        out = [s.join(p[i % len(p)] for p in parts) for i in range(n)]
        if collapse is not None:
            out = [r_string(collapse).join(out)]
        return strs(*out)
    return g


def _print(it, x=None, *a, **kw):
    it.echo(_show(x))
    return x


def _show(x):
    if x is None:
        return "NULL"
    if isinstance(x, RList):
        return "list(%s)" % ", ".join("%s=%s" % (nm, _show(v)) for nm, v in zip(x.names, x.vals))
    if isinstance(x, (Closure, Builtin)):
        return "<function>"
    a = dense(x)
    return " ".join(v if isinstance(v, str) else fmt_num(v) for v in flat(a)[:20])


def _stop(it, *args, **kw):
    raise RError("".join(r_string(a) for a in args))


def _warning(it, *args, **kw):
    it.warnings.append("".join(r_string(a) for a in args))
    return None


def _extreme(f, nanf):
    def g(it, *args, na_rm=False):
        parts = [flat(dense(a)).astype(np.float64) for a in args if a is not None]
        a = np.concatenate(parts) if parts else np.zeros(0)
        if a.size == 0:
            return num(-np.inf if f is np.max else np.inf)
        if as_logical_scalar(na_rm):
            a = a[~np.isnan(a)]
        return num(f(a))                     # NaN propagates, as NA/NaN does through R's max()/min()
    return g


def _pminmax(f):
    def g(it, *args, **kw):
        out = dense(args[0]).astype(np.float64)
        for b in args[1:]:
            x, y, dim = recycle2(out, dense(b).astype(np.float64))
            out = _finish(f(x, y), dim)
        return out
    return g


def _rfast_pmin(it, x, y, na_rm=False):
    """Rfast::Pmin: elementwise minimum of two equally sized numeric matrices / vectors."""
    a, b = dense(x).astype(np.float64), dense(y).astype(np.float64)
    if a.size != b.size:
        raise RError("Pmin: arguments of different size")
    r = np.where(flat(b) < flat(a), flat(b), flat(a))
    return np.reshape(r, a.shape, order="F") if a.ndim == 2 else r


def _which(it, x, **kw):
    a = flat(dense(x))
    if a.dtype != bool:
        a = np.where(np.isnan(a.astype(np.float64)), False, a != 0)
    return (np.nonzero(a)[0] + 1).astype(np.float64)


def _dim(it, x):
    if hasattr(x, "r_dim"):
        return x.r_dim()
    if isinstance(x, (np.ndarray, RSparse)) and dense(x).ndim == 2:
        return np.array(dense(x).shape, dtype=np.float64)
    return None


def _nrow(it, x):
    d = _dim(it, x)
    return None if d is None else d[:1]


def _ncol(it, x):
    d = _dim(it, x)
    return None if d is None else d[1:]


def _as_matrix(it, x, *a, **kw):
    a = dense(x)
    return np.asfortranarray(a) if a.ndim == 2 else np.reshape(a, (a.size, 1), order="F")


def _as_vector(it, x, *a, **kw):
    if isinstance(x, RList) or x is None:
        return x
    return flat(dense(x)).copy()


def _t(it, x):
    a = dense(x)
    return np.asfortranarray(a.T) if a.ndim == 2 else a[None, :]


def _as_matprod_operands(x, y, op):
    """The shapes R's do_matprod (array.c) gives vector operands of crossprod (op 1) / tcrossprod (op 2)."""
    a = dense(x).astype(np.float64)
    b = a if y is None else dense(y).astype(np.float64)
    if a.ndim != 2 and b.ndim != 2:                      # both vectors: column vectors
        return a[:, None], b[:, None]
    if a.ndim != 2:                                      # x a vector, y a matrix
        nry, ncy = b.shape
        if op == 1:
            a = a[:, None] if a.size == nry else None    # crossprod: x is a column vector
        else:
            a = a[None, :] if a.size == ncy else (a[:, None] if ncy == 1 else None)      # tcrossprod: a ROW vector if it fits
    elif b.ndim != 2:                                    # y a vector, x a matrix
        nrx, ncx = a.shape
        if op == 1:
            b = b[:, None] if b.size == nrx else (b[None, :] if nrx == 1 else None)
        else:
            b = b[None, :] if b.size == ncx else (b[:, None] if ncx == 1 else None)
    if a is None or b is None:
        raise RError("non-conformable arguments")
    return a, b


def _tcrossprod(it, x, y=None):
    a, b = _as_matprod_operands(x, y, 2)
    if a.shape[1] != b.shape[1]:
        raise RError("non-conformable arguments")
    return np.asfortranarray(a @ b.T)


def _crossprod(it, x, y=None):
    a, b = _as_matprod_operands(x, y, 1)
    if a.shape[0] != b.shape[0]:
        raise RError("non-conformable arguments")
    return np.asfortranarray(a.T @ b)


def _fitted(it, object, **kw):
    """fitted.flash (flashier): L_pm %*% t(F_pm).  The stand-in fit carries L_pm / F_pm."""
    if not isinstance(object, RList) or object.get("L_pm") is None or object.get("F_pm") is None:
        raise RError("fitted(): the stand-in flash fit needs L_pm and F_pm")
    return _tcrossprod(it, object.get("L_pm"), object.get("F_pm"))


def _is(it, obj, class2=None):
    c = r_string(class2)
    if c in ("sparseMatrix", "dgCMatrix", "Matrix"):
        return np.array([isinstance(obj, RSparse)])
    if c == "matrix":
        return np.array([isinstance(obj, np.ndarray) and obj.ndim == 2])
    raise RError("unsupported: is(, '%s')" % c)


def _round(it, x, digits=0):
    a = dense(x).astype(np.float64)
    d = scalar(digits)
    if math.isinf(d) or math.isnan(d):
        return a
    return np.round(a, int(d))


def _names_set(it, x, value=None):
    if isinstance(x, RList):
        new = x.copy()
        new.names = [None] * len(new) if value is None else list(flat(dense(value)))
        return new
    return x


def _class_set(it, x, value=None):
    if isinstance(x, RList):
        new = x.copy()
        new.cls = value
        return new
    return x


def _install_builtins(it):
    d = it.def_builtin
    for name, f in [("exp", np.exp), ("log", None), ("sqrt", np.sqrt), ("abs", np.abs), ("floor", np.floor), ("ceiling", np.ceil),
                    ("log10", np.log10), ("log2", np.log2), ("log1p", np.log1p), ("expm1", np.expm1), ("is.na", np.isnan),
                    ("is.nan", np.isnan), ("is.finite", np.isfinite), ("is.infinite", np.isinf), ("sign", np.sign)]:
        if f is not None:
            d(name, _math1(f))

    def _log(it_, x, base=None):
        a = dense(x).astype(np.float64)
        with np.errstate(all="ignore"):
            r = np.log(a)
            return r if base is None else r / math.log(float(scalar(base)))
    d("log", _log)
    d("lgamma", lambda it_, x: _gammaln(dense(x).astype(np.float64)))
    d("lfactorial", lambda it_, x: _gammaln(dense(x).astype(np.float64) + 1.0))      # lfactorial(x) = lgamma(x + 1)
    d("sum", lambda it_, *a, na_rm=False: num(_ld_sum(np.concatenate([flat(dense(x)).astype(LD) for x in a if x is not None] or [np.zeros(0)]))))
    d("mean", lambda it_, x, **kw: num(float(np.mean(flat(dense(x)).astype(LD)))))
    d("prod", lambda it_, *a: num(float(np.prod(np.concatenate([flat(dense(x)).astype(LD) for x in a])))))
    d("max", _extreme(np.max, np.nanmax))
    d("min", _extreme(np.min, np.nanmin))
    d("range", lambda it_, *a, **kw: np.concatenate([_extreme(np.min, None)(it_, *a), _extreme(np.max, None)(it_, *a)]))
    d("pmin", _pminmax(np.minimum))
    d("pmax", _pminmax(np.maximum))
    d("Pmin", _rfast_pmin)
    d("colSums", lambda it_, x, **kw: np.asarray(np.sum(dense(x).astype(LD), axis=0), dtype=np.float64))
    d("rowSums", lambda it_, x, **kw: np.asarray(np.sum(dense(x).astype(LD), axis=1), dtype=np.float64))
    d("colMeans", lambda it_, x, **kw: np.asarray(np.sum(dense(x).astype(LD), axis=0) / dense(x).shape[0], dtype=np.float64))
    d("rowMeans", lambda it_, x, **kw: np.asarray(np.sum(dense(x).astype(LD), axis=1) / dense(x).shape[1], dtype=np.float64))
    # the package takes rowMaxs from Rfast and colMaxs from matrixStats (NAMESPACE).  Rfast::rowMaxs(x, value = FALSE)
    # returns the POSITIONS of the row maxima (1-based, first on ties) unless value = TRUE; matrixStats::colMaxs the values.
    d("rowMaxs", lambda it_, x, value=False, **kw: np.max(dense(x), axis=1) if as_logical_scalar(value) else np.argmax(dense(x), axis=1) + 1.0)
    d("colMaxs", lambda it_, x, **kw: np.max(dense(x), axis=0))
    d("cumsum", lambda it_, x: np.cumsum(flat(dense(x)).astype(np.float64)))
    d("diff", lambda it_, x, **kw: np.diff(flat(dense(x)).astype(np.float64)))
    d("all", lambda it_, *a, **kw: np.array([all(bool(np.all(dense(x))) for x in a)]))
    d("any", lambda it_, *a, **kw: np.array([any(bool(np.any(dense(x))) for x in a)]))
    d("which", _which)
    d("which.max", lambda it_, x: num(int(np.argmax(flat(dense(x)))) + 1))
    d("which.min", lambda it_, x: num(int(np.argmin(flat(dense(x)))) + 1))
    d("length", lambda it_, x: num(rlen(x)))
    d("nrow", _nrow)
    d("ncol", _ncol)
    d("NROW", lambda it_, x: num(dense(x).shape[0]))
    d("NCOL", lambda it_, x: num(dense(x).shape[1] if dense(x).ndim == 2 else 1))
    d("dim", _dim)
    d("matrix", _matrix)
    d("as.matrix", _as_matrix)
    d("as.vector", _as_vector)
    d("as.numeric", lambda it_, x, **kw: flat(dense(x)).astype(np.float64))
    d("as.double", lambda it_, x, **kw: flat(dense(x)).astype(np.float64))
    d("as.integer", lambda it_, x, **kw: np.trunc(flat(dense(x)).astype(np.float64)))
    d("as.logical", lambda it_, x, **kw: flat(dense(x)) != 0)
    d("as.character", lambda it_, x, **kw: strs(*[v if isinstance(v, str) else fmt_num(v) for v in flat(dense(x))]))
    d("drop", lambda it_, x: flat(dense(x)).copy() if 1 in dense(x).shape else x)
    d("numeric", lambda it_, length=0: np.zeros(int(scalar(length)) if not isinstance(length, int) else length))
    d("vector", lambda it_, mode="list", length=0: RList([None] * int(scalar(length))) if r_string(mode) == "list" else np.zeros(int(scalar(length))))
    d("t", _t)
    d("tcrossprod", _tcrossprod)
    d("crossprod", _crossprod)
    d("diag", lambda it_, x, **kw: np.asfortranarray(np.diag(flat(dense(x)))) if dense(x).ndim == 1 and dense(x).size > 1 else (np.diag(dense(x)).copy() if dense(x).ndim == 2 else np.asfortranarray(np.eye(int(scalar(x))))))
    d("Diagonal", lambda it_, n, x=None: RDiag(scalar(n), dense(x) if x is not None else np.ones(0)))
    d("Matrix", lambda it_, data, sparse=False, **kw: RSparse(dense(data)) if as_logical_scalar(sparse) else np.asfortranarray(dense(data)))
    d("is", _is)
    d("is.null", lambda it_, x: np.array([x is None]))
    d("is.list", lambda it_, x: np.array([isinstance(x, RList)]))
    d("is.matrix", lambda it_, x: np.array([isinstance(x, np.ndarray) and x.ndim == 2]))
    d("is.function", lambda it_, x: np.array([isinstance(x, (Closure, Builtin))]))
    d("is.numeric", lambda it_, x: np.array([isinstance(x, np.ndarray) and x.dtype != object]))
    d("c", _c, rawnames=True)
    d("list", _list, rawnames=True)
    d("unlist", lambda it_, x, **kw: _c(it_, *x.vals) if isinstance(x, RList) else x)
    d("cbind", _bind(1))
    d("rbind", _bind(0))
    d("rep", _rep)
    d("rep_len", lambda it_, x, length_out: np.resize(flat(dense(x)), int(scalar(length_out))))
    d("seq", _seq)
    d("seq.int", _seq)
    d("seq_len", lambda it_, n: np.arange(1, int(scalar(n)) + 1, dtype=np.float64))
    d("seq_along", lambda it_, x: np.arange(1, rlen(x) + 1, dtype=np.float64))
    d("rev", lambda it_, x: flat(dense(x))[::-1].copy())
    d("cut", _cut)
    d("split", _split)
    d("lapply", _lapply)
    d("sapply", _sapply)
    d("apply", _apply)
    d("modifyList", _modify_list)
    d("names", lambda it_, x: strs(*[nm or "" for nm in x.names]) if isinstance(x, RList) and any(x.names) else None)
    d("names<-", _names_set)
    d("storage.mode<-", lambda it_, x, value=None: np.array(dense(x), dtype=np.float64, order="F"))      # only ever 'double' here
    d("character", lambda it_, length=0: strs(*[""] * int(scalar(length)) if not isinstance(length, int) else [""] * length))
    d("integer", lambda it_, length=0: np.zeros(int(scalar(length)) if not isinstance(length, int) else length))
    d("source", lambda it_, file, **kw: it_.source(r_string(file)))
    d("class<-", _class_set)
    d("class", lambda it_, x: x.cls if isinstance(x, RList) and x.cls is not None else strs("list" if isinstance(x, RList) else "numeric"))
    d("paste", _paste(" "))
    d("paste0", _paste(""))
    d("print", _print)
    d("cat", lambda it_, *a, **kw: it_.echo(" ".join(_show(x) if not is_str(x) else " ".join(flat(x)) for x in a)))
    d("message", lambda it_, *a, **kw: it_.echo("".join(r_string(x) for x in a)))
    d("stop", _stop)
    d("warning", _warning)
    d("gc", lambda it_, *a, **kw: None)
    d("Sys.time", lambda it_: num(0.0))
    d("difftime", lambda it_, a, b, units=None, **kw: arith("-", a, b))
    d("round", _round)
    d("ifelse", lambda it_, test, yes, no: _finish(np.where(flat(dense(test)) != 0, np.resize(flat(dense(yes)), dense(test).size), np.resize(flat(dense(no)), dense(test).size)), dense(test).shape if dense(test).ndim == 2 else None))
    d("identical", lambda it_, a, b, **kw: np.array([_identical(a, b)]))
    d("isTRUE", lambda it_, x: np.array([isinstance(x, np.ndarray) and x.size == 1 and x.dtype == bool and bool(x[0])]))
    d("nchar", lambda it_, x, **kw: num([len(s) for s in flat(dense(x))]))
    d("exists", lambda it_, x, **kw: np.array([it_.globalenv.lookup(r_string(x))[0] is not None]))
    d("do.call", lambda it_, what, args, **kw: it_.apply(what if not is_str(what) else it_.lookup_fn(r_string(what), it_.globalenv), list(zip(args.names, args.vals))))
    d("dnorm", _dnorm)
    d("saveRDS", lambda it_, *a, **kw: None)
    d("fitted", _fitted)


def _dnorm(it, x, mean=0.0, sd=1.0, log=False):
    x, m, dim = recycle2(dense(x).astype(np.float64), dense(mean).astype(np.float64))
    z, s, dim2 = recycle2(np.asarray(x - m), dense(sd).astype(np.float64))
    z = z / s
    r = -(0.918938533204672741780329736406 + 0.5 * z * z + np.log(s))
    return _finish(r if as_logical_scalar(log) else np.exp(r), dim or dim2)


def _identical(a, b):
    if a is None or b is None:
        return a is None and b is None
    if isinstance(a, RList) or isinstance(b, RList):
        return (isinstance(a, RList) and isinstance(b, RList) and a.names == b.names and len(a) == len(b)
                and all(_identical(x, y) for x, y in zip(a.vals, b.vals)))
    x, y = dense(a), dense(b)
    return x.shape == y.shape and bool(np.all((x == y) | ((x != x) & (y != y))))
