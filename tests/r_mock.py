"""Runs the R side of the boundary without R -- TEST INFRASTRUCTURE.

* `ebpmf.alpha_b200/r/src/r_glue.c` (the `.Call` shim, unmodified) is compiled together with
  tests/r_api_stub/mock_r.c, a functional mock of the R C-API subset the glue uses, and linked to the shipped
  libebpmf_b200.so: tests/_build/libr_glue_mock.so.
* `ebpmf.alpha_b200/r/R/vga_cuda.R` and `ebpmf_log_vga_block.R` (the R side, unmodified) are executed by oracle/mini_r.py; its `.Call(C_xxx,
  ...)` goes through `RSession.dot_call`: interpreter values -> SEXPs -> the glue's registered routine (arity
  checked against the glue's own R_registerRoutines table, Rf_error caught, PROTECT balance checked) -> SEXPs ->
  interpreter values.

So `vga_pois_solver_mat_newton(M, X, S, Beta, Sigma2, ...)` evaluated in an RSession travels the whole product path a
user of the R package would: R wrapper -> C glue -> C ABI -> CUDA.  Neither GNU R nor its headers are involved, and
nothing here is shipped."""
import ctypes as C
import os
import subprocess

import numpy as np
import scipy.sparse as sp

from oracle import mini_r as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GLUE = os.path.join(ROOT, "ebpmf.alpha_b200", "r", "src", "r_glue.c")
RWRAP = os.path.join(ROOT, "ebpmf.alpha_b200", "r", "R", "vga_cuda.R")
RBLOCK = os.path.join(ROOT, "ebpmf.alpha_b200", "r", "R", "ebpmf_log_vga_block.R")
STUB = os.path.join(ROOT, "tests", "r_api_stub")
LIBDIR = os.path.join(ROOT, "ebpmf.alpha_b200", "lib")
BUILD = os.path.join(ROOT, "tests", "_build")
SO = os.path.join(BUILD, "libr_glue_mock.so")

NILSXP, LGLSXP, INTSXP, REALSXP, STRSXP, VECSXP, EXTPTRSXP, S4SXP = 0, 10, 13, 14, 16, 19, 22, 25


class RCallError(R.RError):
    """An R-level error raised by the glue through Rf_error()."""


class RInt(np.ndarray):
    """An integer vector (INTSXP): a float64 array flagged as integer -- all the interpreter needs to hand the glue
    the storage mode its INTEGER() accessors expect.  Arithmetic drops the flag (results are plain doubles)."""


def as_int(a):
    a = R.dense(a) if isinstance(a, (np.ndarray, R.RSparse)) else np.atleast_1d(np.asarray(a))
    return np.trunc(np.asarray(a, dtype=np.float64)).view(RInt)


class RExt:
    """An external pointer (reference semantics, may carry attributes)."""

    def __init__(self, sexp):
        self.sexp = sexp
        self.attrs = {}


_lib = None


def build():
    os.makedirs(BUILD, exist_ok=True)
    srcs = [GLUE, os.path.join(STUB, "mock_r.c")]
    if os.path.exists(SO) and all(os.path.getmtime(SO) >= os.path.getmtime(s) for s in srcs + [os.path.join(STUB, "Rinternals.h")]):
        return SO
    cmd = ["/usr/bin/gcc", "-std=gnu99", "-O1", "-g", "-shared", "-fPIC", "-Wall", "-Wextra", "-Werror", "-Wno-cast-function-type",
           "-Wno-unused-parameter", "-I", STUB, "-I", os.path.join(ROOT, "include"), *srcs, "-o", SO,
           "-L", LIBDIR, "-lebpmf_b200", "-Wl,-rpath," + LIBDIR]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building the glue against the mock R API failed:\n" + r.stderr)
    return SO


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        vp, ci, cl, cs = C.c_void_p, C.c_int, C.c_long, C.c_char_p
        for name, res, args in [("mockr_init", None, []), ("mockr_n_routines", ci, []), ("mockr_routine_name", cs, [ci]),
                                ("mockr_routine_nargs", ci, [ci]), ("mockr_dot_call", vp, [cs, ci, C.POINTER(vp)]),
                                ("mockr_last_error", cs, []), ("mockr_max_protect", ci, []), ("mockr_live_objects", cl, []),
                                ("mockr_nil", vp, []), ("mockr_alloc", vp, [ci, cl]), ("mockr_data", vp, [vp]), ("mockr_length", cl, [vp]),
                                ("mockr_type", ci, [vp]), ("mockr_set_attr", None, [vp, cs, vp]), ("mockr_get_attr", vp, [vp, cs]),
                                ("mockr_new_s4", vp, []), ("mockr_vector_elt", vp, [vp, cl]), ("mockr_set_vector_elt", None, [vp, cl, vp]),
                                ("mockr_mkchar", vp, [cs]), ("mockr_char", cs, [vp]), ("mockr_used_allocator", ci, [vp]),
                                ("mockr_free", None, [vp])]:
            f = getattr(L, name)
            f.restype, f.argtypes = res, args
        L.mockr_init()
        _lib = L
    return _lib


def routines():
    L = lib()
    return {L.mockr_routine_name(i).decode(): L.mockr_routine_nargs(i) for i in range(L.mockr_n_routines())}


# ---------------------------------------------------------------------------------------------------------------
# interpreter values <-> SEXPs
# ---------------------------------------------------------------------------------------------------------------
def _vector(sxptype, arr, ctype):
    L = lib()
    a = np.ascontiguousarray(arr.ravel(order="F"), dtype=ctype)
    s = L.mockr_alloc(sxptype, a.size)
    if a.size:
        C.memmove(L.mockr_data(s), a.ctypes.data, a.nbytes)
    if arr.ndim == 2:
        L.mockr_set_attr(s, b"dim", _vector(INTSXP, np.array(arr.shape), np.int32))
    return s


def _strings(items):
    L = lib()
    s = L.mockr_alloc(STRSXP, len(items))
    for i, x in enumerate(items):
        L.mockr_set_vector_elt(s, i, L.mockr_mkchar(str(x).encode()))
    return s


def to_sexp(v):
    L = lib()
    if v is None:
        return L.mockr_nil()
    if isinstance(v, RExt):
        return v.sexp
    if isinstance(v, R.RSparse):                       # a dgCMatrix: S4 object whose slots are attributes
        m = sp.csc_matrix(v.dense)
        m.sort_indices()
        s = L.mockr_new_s4()
        L.mockr_set_attr(s, b"class", _strings(["dgCMatrix"]))
        L.mockr_set_attr(s, b"i", _vector(INTSXP, m.indices, np.int32))
        L.mockr_set_attr(s, b"p", _vector(INTSXP, m.indptr, np.int32))
        L.mockr_set_attr(s, b"x", _vector(REALSXP, m.data, np.float64))
        L.mockr_set_attr(s, b"Dim", _vector(INTSXP, np.array(m.shape), np.int32))
        return s
    if isinstance(v, R.RList):
        s = L.mockr_alloc(VECSXP, len(v))
        for i, x in enumerate(v.vals):
            L.mockr_set_vector_elt(s, i, to_sexp(x))
        if any(v.names):
            L.mockr_set_attr(s, b"names", _strings([nm or "" for nm in v.names]))
        return s
    if isinstance(v, RInt):
        return _vector(INTSXP, np.asarray(v), np.int32)
    if isinstance(v, np.ndarray):
        if v.dtype == object:
            return _strings(list(v.ravel()))
        if v.dtype == bool:
            return _vector(LGLSXP, v, np.int32)
        return _vector(REALSXP, v, np.float64)
    raise R.RError("cannot pass a %s through .Call" % type(v).__name__)


def from_sexp(s):
    L = lib()
    t = L.mockr_type(s)
    if t == NILSXP:
        return None
    if t == EXTPTRSXP:
        return RExt(s)
    n = L.mockr_length(s)
    if t in (REALSXP, INTSXP, LGLSXP):
        ctype = np.float64 if t == REALSXP else np.int32
        a = np.ctypeslib.as_array(C.cast(L.mockr_data(s), C.POINTER(C.c_double if t == REALSXP else C.c_int)), shape=(max(n, 1),))[:n]
        a = np.array(a, dtype=ctype)
        out = a.astype(np.float64) if t != LGLSXP else a != 0
        dim = L.mockr_get_attr(s, b"dim")
        if L.mockr_type(dim) == INTSXP:
            d = from_sexp(dim)
            out = np.reshape(out, (int(d[0]), int(d[1])), order="F")
        return out.view(RInt) if t == INTSXP else out
    if t == STRSXP:
        return R.strs(*[L.mockr_char(L.mockr_vector_elt(s, i)).decode() for i in range(n)])
    if t == VECSXP:
        names = L.mockr_get_attr(s, b"names")
        nm = list(from_sexp(names)) if L.mockr_type(names) == STRSXP else [None] * n
        out = R.RList([from_sexp(L.mockr_vector_elt(s, i)) for i in range(n)], nm)
        extra = L.mockr_get_attr(s, b"n_updates_per_batch")
        out.attrs = {"n_updates_per_batch": from_sexp(extra)} if L.mockr_type(extra) != NILSXP else {}
        return out
    if t == S4SXP:                                      # a dgCMatrix
        g = lambda nm: from_sexp(L.mockr_get_attr(s, nm))                      # noqa: E731
        d = g(b"Dim")
        m = sp.csc_matrix((np.asarray(g(b"x")), np.asarray(g(b"i")).astype(np.int32), np.asarray(g(b"p")).astype(np.int32)),
                          shape=(int(d[0]), int(d[1])))
        return R.RSparse(m.toarray())
    raise R.RError("cannot read a SEXP of type %d back" % t)


# ---------------------------------------------------------------------------------------------------------------
# an "R session" with the package's R wrappers loaded over the compiled glue
# ---------------------------------------------------------------------------------------------------------------
class RSession:
    def __init__(self, reference_root=None, echo=None):
        """With `reference_root`, the reference's own R files are sourced FIRST and the wrappers of vga_cuda.R
        then replace the functions of the hot path, which is how the drop-in is installed."""
        self.L = lib()
        self.calls = []                                 # (routine, max PROTECT depth) of every .Call made
        if reference_root is not None:
            from oracle import minir_reference as mr
            self.it = mr.load_reference(reference_root, echo=echo)
        else:
            self.it = R.Interp(echo=echo)
        it = self.it
        for name in routines():                         # what useDynLib(..., .registration = TRUE) gives the namespace
            it.set(name, name)
        it.def_builtin(".Call", self._dot_call_builtin)
        it.def_builtin("new.env", lambda it_, **kw: R.REnvObj())
        it.def_builtin("emptyenv", lambda it_: None)
        it.def_builtin("Sys.getenv", lambda it_, x, unset="": R.strs(os.environ.get(R.r_string(x), R.r_string(unset))))
        it.def_builtin("as.integer", lambda it_, x, **kw: as_int(np.array([float(s) for s in R.dense(x).ravel()]) if R.is_str(x) else x))
        it.def_builtin("storage.mode<-", lambda it_, x, value=None: np.array(R.dense(x), dtype=np.float64, order="F"))
        it.def_builtin("as", self._as)
        it.def_builtin("is", self._is)
        it.def_builtin("new", self._new)
        it.def_builtin("vapply", lambda it_, X, FUN, FUN_VALUE=None, **kw: it_.get("sapply").fn(it_, X, FUN))
        it.def_builtin("attr", lambda it_, x, which: getattr(x, "attrs", {}).get(R.r_string(which)))
        it.def_builtin("attr<-", self._attr_set)
        it.source(RWRAP)
        it.source(RBLOCK)

    @staticmethod
    def _as(it_, obj, cls):
        c = R.r_string(cls)
        if c == "dgCMatrix":
            return obj if isinstance(obj, R.RSparse) else R.RSparse(R.dense(obj))
        raise R.RError("unsupported: as(, '%s')" % c)

    @staticmethod
    def _is(it_, obj, class2=None):
        c = R.r_string(class2)
        if c in ("dgCMatrix", "sparseMatrix", "Matrix"):
            return np.array([isinstance(obj, R.RSparse)])
        raise R.RError("unsupported: is(, '%s')" % c)

    @staticmethod
    def _new(it_, cls, **slots):
        if R.r_string(cls) != "dgCMatrix":
            raise R.RError("unsupported: new('%s')" % R.r_string(cls))
        dim = np.asarray(slots["Dim"]).astype(int)
        m = sp.csc_matrix((np.asarray(slots["x"], dtype=np.float64), np.asarray(slots["i"]).astype(np.int32),
                           np.asarray(slots["p"]).astype(np.int32)), shape=(dim[0], dim[1]))
        return R.RSparse(m.toarray())

    @staticmethod
    def _attr_set(it_, x, which, value=None):
        if not hasattr(x, "attrs"):
            raise R.RError("unsupported: attr<- on this object")
        x.attrs[R.r_string(which)] = value
        return x

    def _dot_call_builtin(self, it_, symbol, *args):
        return self.dot_call(R.r_string(symbol), *args)

    def dot_call(self, name, *args):
        sx = [to_sexp(a) for a in args]
        before = [self._fingerprint(s) for s in sx]
        arr = (C.c_void_p * max(len(sx), 1))(*sx)
        try:
            out = self.L.mockr_dot_call(name.encode(), len(sx), arr)
            if not out:
                raise RCallError(self.L.mockr_last_error().decode())
            if before != [self._fingerprint(s) for s in sx]:
                raise AssertionError("%s wrote to one of its arguments (R shares them: copy-on-modify)" % name)
            self.calls.append((name, self.L.mockr_max_protect()))
            res = from_sexp(out)
        finally:
            for s in sx:                                # arguments were copies made for this call
                if self.L.mockr_type(s) not in (NILSXP, EXTPTRSXP):
                    self.L.mockr_free(s)
        if self.L.mockr_type(out) not in (NILSXP, EXTPTRSXP):
            self.last_used_allocator = self._any_allocator(out)
            self._free_tree(out)
        return res

    def _fingerprint(self, s):
        L = self.L
        t = L.mockr_type(s)
        if t in (REALSXP, INTSXP, LGLSXP) and L.mockr_length(s):
            nbytes = L.mockr_length(s) * (8 if t == REALSXP else 4)
            return hash(C.string_at(L.mockr_data(s), min(nbytes, 1 << 24)))
        return 0

    def _any_allocator(self, s):
        L = self.L
        if L.mockr_type(s) == VECSXP:
            return any(self._any_allocator(L.mockr_vector_elt(s, i)) for i in range(L.mockr_length(s)))
        return bool(L.mockr_used_allocator(s))

    def _free_tree(self, s):
        L = self.L
        if L.mockr_type(s) == VECSXP:
            for i in range(L.mockr_length(s)):
                e = L.mockr_vector_elt(s, i)
                if L.mockr_type(e) not in (NILSXP, EXTPTRSXP):
                    self._free_tree(e)
        L.mockr_free(s)

    def call(self, fname, *args, **kw):
        return self.it.call(fname, *args, **kw)

    def close(self):
        """What R's garbage collector does at exit: run the external pointers' finalizers."""
        env = self.it.globalenv.vars.get(".ebpmf_b200_env")
        if isinstance(env, R.REnvObj):
            for v in list(env.vars.values()):
                if isinstance(v, RExt):
                    self.L.mockr_free(v.sexp)
            env.vars.clear()
