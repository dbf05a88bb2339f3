"""ctypes binding for oracle/liboracle_vga.so (the C transcription of the oracle).

TEST INFRASTRUCTURE ONLY -- see vga_oracle.c.  Build with ``make -C oracle``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

VAR_TYPE = {"constant": 0, "by_row": 1, "by_col": 2}   # R/ebpmf_log.R:178-192


class OracleNaN(FloatingPointError):
    pass


def build():
    subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle_vga.so")
        if not os.path.exists(path):
            build()
        _LIB = C.CDLL(path)
        _LIB.oracle_sum_lfactorial.restype = C.c_double
    return _LIB


def _d(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def _f(a):
    """n x p -> column-major flat double buffer."""
    return np.asfortranarray(np.asarray(a, dtype=np.float64))


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _operand(a, n, p):
    """(buffer, row stride, col stride) for an R-recycled operand."""
    a = np.asarray(a, dtype=np.float64)
    if a.ndim == 0 or a.size == 1:
        return _d(a.reshape(1)), 0, 0
    if a.ndim == 1:
        if a.size == n:
            return _d(a), 1, 0
        raise ValueError("vector operand must have length n")
    if a.shape == (n, p):
        return _f(a), 1, n
    raise ValueError("bad operand shape %r" % (a.shape,))


def _check(rc):
    if rc == 1:
        raise OracleNaN("missing value where TRUE/FALSE needed")
    if rc != 0:
        raise RuntimeError("oracle error %d" % rc)


def max_threads():
    return int(lib().oracle_max_threads())


def vga_newton(M, X, S, Beta, Sigma2, maxiter=1000, tol=1e-8, nthreads=0, by_col_sigma2=False):
    M = _f(M); n, p = M.shape
    X = _f(X)
    Sb, Srs, Scs = _operand(S, n, p)
    Bb, Brs, Bcs = _operand(Beta, n, p)
    if by_col_sigma2:
        Vb, Vrs, Vcs = _d(Sigma2), 0, 1
    else:
        Vb, Vrs, Vcs = _operand(Sigma2, n, p)
    Mo = np.empty((n, p), order="F"); Vo = np.empty((n, p), order="F")
    u = C.c_int(); cv = C.c_int()
    tr = np.full(maxiter, np.nan)
    rc = lib().oracle_vga_newton(C.c_longlong(n), C.c_longlong(p), _p(M), _p(X),
                                 _p(Sb), C.c_longlong(Srs), C.c_longlong(Scs),
                                 _p(Bb), C.c_longlong(Brs), C.c_longlong(Bcs),
                                 _p(Vb), C.c_longlong(Vrs), C.c_longlong(Vcs),
                                 C.c_int(maxiter), C.c_double(tol), _p(Mo), _p(Vo),
                                 C.byref(u), C.byref(cv), _p(tr), C.c_int(nthreads))
    _check(rc)
    return dict(M=Mo, V=Vo, n_updates=u.value, converged=bool(cv.value), maxabs_f=tr[:u.value + 1])


def vga_newton_rstyle(M, X, S, Beta, Sigma2, maxiter=1000, tol=1e-8):
    M = _f(M); n, p = M.shape
    X = _f(X); Beta = _f(Beta); Sigma2 = _f(Sigma2)
    Mo = np.empty((n, p), order="F"); Vo = np.empty((n, p), order="F")
    u = C.c_int(); cv = C.c_int()
    rc = lib().oracle_vga_newton_rstyle(C.c_longlong(n), C.c_longlong(p), _p(M), _p(X), C.c_double(float(S)),
                                        _p(Beta), _p(Sigma2), C.c_int(maxiter), C.c_double(tol),
                                        _p(Mo), _p(Vo), C.byref(u), C.byref(cv))
    _check(rc)
    return dict(M=Mo, V=Vo, n_updates=u.value, converged=bool(cv.value))


def vga_fixed_iter(M, V, Y, S, Beta, Sigma2, maxiter=1000, nthreads=0):
    M = _f(M); n, p = M.shape
    Y = _f(Y)
    Sb, Srs, Scs = _operand(S, n, p)
    Bb, Brs, Bcs = _operand(Beta, n, p)
    Vb, Vrs, Vcs = _operand(Sigma2, n, p)
    V0 = None if V is None else _f(np.broadcast_to(np.asarray(V, dtype=np.float64), (n, p)))
    Mo = np.empty((n, p), order="F"); Vo = np.empty((n, p), order="F")
    rc = lib().oracle_vga_fixed_iter(C.c_longlong(n), C.c_longlong(p), _p(M),
                                     None if V0 is None else _p(V0), _p(Y),
                                     _p(Sb), C.c_longlong(Srs), C.c_longlong(Scs),
                                     _p(Bb), C.c_longlong(Brs), C.c_longlong(Bcs),
                                     _p(Vb), C.c_longlong(Vrs), C.c_longlong(Vcs),
                                     C.c_int(maxiter), _p(Mo), _p(Vo), C.c_int(nthreads))
    _check(rc)
    return dict(M=Mo, V=Vo)


def vga_low_memory(M, Y, l0, f0, EL, EF, sigma2, var_type, maxiter=1000, tol=1e-8, nthreads=0):
    M = _f(M); n, p = M.shape
    Y = _f(Y); EL = _f(np.asarray(EL).reshape(n, -1)); EF = _f(np.asarray(EF).reshape(p, -1))
    K = EL.shape[1]
    l0 = _d(np.broadcast_to(np.asarray(l0, dtype=np.float64), (n,)))
    f0 = _d(np.broadcast_to(np.asarray(f0, dtype=np.float64), (p,)))
    s2 = _d(np.asarray(sigma2).reshape(-1))
    Mo = np.empty((n, p), order="F")
    u = C.c_int(); cv = C.c_int()
    rc = lib().oracle_vga_low_memory(C.c_longlong(n), C.c_longlong(p), C.c_longlong(K), _p(M), _p(Y),
                                     _p(l0), _p(f0), _p(EL), _p(EF), _p(s2),
                                     C.c_int(VAR_TYPE.get(var_type, -1)), C.c_int(maxiter), C.c_double(tol),
                                     _p(Mo), C.byref(u), C.byref(cv), C.c_int(nthreads))
    _check(rc)
    return dict(M=Mo, n_updates=u.value, converged=bool(cv.value))


def partials(Y, M, V, nthreads=0):
    Y = _f(Y); M = _f(M); V = _f(V); n, p = M.shape
    out = (C.c_double * 4)()
    col = np.empty(p); row = np.empty(n)
    rc = lib().oracle_partials(C.c_longlong(n), C.c_longlong(p), _p(Y), _p(M), _p(V),
                               C.byref(out, 0), C.byref(out, 8), C.byref(out, 16), C.byref(out, 24),
                               _p(col), _p(row), C.c_int(nthreads))
    _check(rc)
    return dict(sym=out[0], ssexp=out[1], slogv=out[2], sum_v=out[3], col_v=col, row_v=row)


def sum_lfactorial(x):
    x = _d(np.asarray(x).reshape(-1))
    return float(lib().oracle_sum_lfactorial(C.c_longlong(x.size), _p(x)))


def vga_step(M, Y, EL, EF, sigma2, var_type, maxiter=10, tol=1e-5, nthreads=0):
    """R/ebpmf_log.R:302-327 (un-batched VGA step), C transcription."""
    M = _f(M); n, p = M.shape
    Y = _f(Y); EL = _f(np.asarray(EL).reshape(n, -1)); EF = _f(np.asarray(EF).reshape(p, -1))
    K = EL.shape[1]
    s2 = _d(np.asarray(sigma2).reshape(-1))
    vt = VAR_TYPE[var_type]
    Mo = np.empty((n, p), order="F"); Vo = np.empty((n, p), order="F")
    sums = np.empty(3); vs = np.empty({0: 1, 1: n, 2: p}[vt])
    u = C.c_int(); cv = C.c_int()
    rc = lib().oracle_vga_step(C.c_longlong(n), C.c_longlong(p), C.c_longlong(K), _p(M), _p(Y),
                               _p(EL), _p(EF), _p(s2), C.c_int(vt), C.c_int(maxiter), C.c_double(tol),
                               _p(Mo), _p(Vo), _p(sums), _p(vs), C.byref(u), C.byref(cv), C.c_int(nthreads))
    _check(rc)
    return dict(M=Mo, V=Vo, sym=sums[0], ssexp=sums[1], slogv=sums[2],
                v_sum=(float(vs[0]) if vt == 0 else vs), n_updates=u.value, converged=bool(cv.value))
