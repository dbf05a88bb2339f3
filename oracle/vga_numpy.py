"""CPU ORACLE (NumPy spec) for the ebpmf_log VGA Newton hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it, and only as the checker.  The product path
(``ebpmf.alpha_b200``) never imports anything under ``oracle/``.

PARITY: the reference (DongyueXie/ebpmf.alpha, 100 % R) ships no golden vectors, no
assertions and no known-answer tests for this path, and GNU R is not installed here or
on the GPU box.  What pins this file: (i) the reference's own source, executed statement
by statement by the restricted R evaluator oracle/mini_r.py (tests/golden/
minir_reference.npz, tests/test_minir_reference.py) -- this transcription reproduces it
bit for bit; (ii) it follows the R source operator by operator (citations below, all
relative to /root/reference); (iii) an independent C transcription
(``oracle/vga_oracle.c``) must agree with it to <= 1e-13; (iv) mathematical invariants
and an mpmath high-precision root (tests/).  NOT pinned against GNU R output: that needs
tests/golden/make_golden.R to be run by someone who has R ("parity unpinned" in that sense).

R evaluation rules that are reproduced on purpose:
  * matrices are column-major; a length-n vector recycles DOWN each column;
  * ``for(i in 1:maxiter)``: test, then update; break at i* => i*-1 updates;
  * the stop test is GLOBAL: ``max(abs(f)) < tol`` over every entry passed in;
  * NaN in ``f`` makes ``if()`` raise an error in R -> FloatingPointError here;
  * on exhaustion ``temp`` (hence V) is one iteration stale;
  * ``x^2`` is ``x*x`` and ``x^(-2)`` is ``pow(x,-2)`` (R_pow);
  * ``%*%`` binds tighter than ``*``;
  * ``sum``/``colSums``/``rowSums`` accumulate in long double.
"""
from __future__ import annotations

import numpy as np

try:  # scipy is only needed for sparse Y and lgamma
    import scipy.sparse as _sp
    from scipy.special import gammaln as _gammaln
except Exception:  # pragma: no cover
    _sp = None
    _gammaln = None

SLOGV_CONST = 0.9189385  # literal in R/ebpmf_log.R:276,318 (NOT 0.5*log(2*pi))


class RError(FloatingPointError):
    """Raised where R would stop with 'missing value where TRUE/FALSE needed'."""


# --------------------------------------------------------------------------
# helpers that restate base-R semantics
# --------------------------------------------------------------------------
def _as_dense(Y):
    if _sp is not None and _sp.issparse(Y):
        return np.asarray(Y.toarray(), dtype=np.float64)
    return np.asarray(Y, dtype=np.float64)


def _recycle(a, n, p):
    """R recycling of a scalar / length-n vector / n x p matrix against an
    n x p column-major matrix (a length-n vector maps row i -> a[i])."""
    a = np.asarray(a, dtype=np.float64)
    if a.ndim == 0 or a.size == 1:
        return a.reshape(())
    if a.ndim == 1:
        if a.size == n:
            return a.reshape(n, 1)
        if a.size == n * p:
            return a.reshape((n, p), order="F")
        raise ValueError("vector length %d does not recycle against %dx%d" % (a.size, n, p))
    if a.shape != (n, p):
        raise ValueError("non-conformable arrays")
    return a


def _rsum(a):
    """base R sum(): long double accumulator."""
    return float(np.sum(np.asarray(a), dtype=np.longdouble))


def _rcolsums(a):
    return np.asarray(np.sum(np.asarray(a), axis=0, dtype=np.longdouble), dtype=np.float64)


def _rrowsums(a):
    return np.asarray(np.sum(np.asarray(a), axis=1, dtype=np.longdouble), dtype=np.float64)


def _global_stop(f, tol):
    """``if(max(abs(f))<tol)`` -- R/vga_solver_mat.R:26,85,100."""
    mx = np.max(np.abs(f))
    if np.isnan(mx):
        raise RError("missing value where TRUE/FALSE needed")
    return bool(mx < tol)


# --------------------------------------------------------------------------
# a1: vga_pois_solver_mat_newton -- R/vga_solver_mat.R:6-41
# --------------------------------------------------------------------------
def vga_pois_solver_mat_newton(M, X, S, Beta, Sigma2, maxiter=1000, tol=1e-8,
                               return_V=True, info=None):
    """Line-by-line restatement of R/vga_solver_mat.R:6-41.

    Returns ``{"M":…, "V":…}`` (or ``M`` when ``return_V`` is false), exactly
    the R shapes.  ``info`` (optional dict) receives ``n_updates`` (the global
    update count u), ``converged`` and ``maxabs_f`` per tested iteration.
    """
    if maxiter < 1:
        raise ValueError("maxiter must be >= 1")
    M = np.array(M, dtype=np.float64, order="F")
    n, p = M.shape
    X = _as_dense(X)
    S = _recycle(S, n, p)
    Beta = _recycle(Beta, n, p)
    Sigma2 = _recycle(Sigma2, n, p)

    const0 = (Sigma2 * X + Beta + 1)            # :9   ((Sigma2*X)+Beta)+1
    const1 = 1 / Sigma2                          # :10
    const2 = Sigma2 / 2                          # :11
    const3 = Beta / Sigma2                       # :12
    M = np.where(const0 - 1 < M, const0 - 1, M)  # :16  Rfast::Pmin = std::min(M, const0-1): (y < x) ? y : x, a NaN in M stays
    n_updates = 0
    converged = False
    trace = []
    for _i in range(1, maxiter + 1):             # :21
        temp = (const0 - M)                      # :22
        sexp = S * np.exp(M + const2 / temp)     # :23
        f = X - sexp - M * const1 + const3       # :25
        trace.append(float(np.max(np.abs(f))))
        if _global_stop(f, tol):                 # :26
            converged = True
            break
        M = M - f / (-sexp * (1 + const2 / (temp * temp)) - const1)   # :31 (temp^2 == temp*temp)
        n_updates += 1
    if info is not None:
        info.update(n_updates=n_updates, converged=converged, maxabs_f=trace)
    if return_V:
        return {"M": M, "V": Sigma2 / temp}      # :36  (temp may be stale)
    return M                                     # :38


# --------------------------------------------------------------------------
# a9: vga_pois_solver_mat_newton_fixed_iter -- R/vga_solver_mat.R:44-65
# --------------------------------------------------------------------------
def vga_pois_solver_mat_newton_fixed_iter(M, V, Y, S, Beta, Sigma2, maxiter=1000):
    """Restatement of R/vga_solver_mat.R:44-65 (no stop test, per entry)."""
    if maxiter < 1:
        raise ValueError("maxiter must be >= 1")
    M = np.array(M, dtype=np.float64, order="F")
    n, p = M.shape
    Y = _as_dense(Y)
    S = _recycle(S, n, p)
    Beta = _recycle(Beta, n, p)
    Sigma2 = _recycle(Sigma2, n, p)
    if V is None:                                # :47-49
        V = Sigma2 / 2
    V = np.broadcast_to(_recycle(V, n, p), (n, p)).astype(np.float64)
    for _i in range(maxiter):                    # :52
        temp = S * np.exp(M + V / 2)             # :54
        M = M - (Y / temp - 1 - (M - Beta) / Sigma2 / temp) / (-1 - 1 / Sigma2 / temp)   # :56
        temp = S * np.exp(M + V / 2) * V / 2     # :58
        V = V / np.exp((-1 - V / 2 / Sigma2 / temp + 0.5 / temp)
                       / (-(1 + V / 2) - V / 2 / Sigma2 / temp))                         # :62
    return {"M": M, "V": V}                      # :64


# --------------------------------------------------------------------------
# a10: vga_pois_solver_mat_newton_low_memory -- R/vga_solver_mat.R:71-109
# --------------------------------------------------------------------------
def vga_pois_solver_mat_newton_low_memory(M, Y, l0, f0, EL, EF, sigma2, var_type,
                                          maxiter=1000, tol=1e-8, info=None):
    """Restatement of R/vga_solver_mat.R:71-109, quirks included: clamp to
    Beta (:78), multiplicative l0_i*f0_j scale (:83,97), sigma2-less derivative
    0.5*(const0-M)^(-2) (:90,105), returns M only (:108)."""
    if maxiter < 1:
        raise ValueError("maxiter must be >= 1")
    M = np.array(M, dtype=np.float64, order="F")
    n, p = M.shape                               # :75-76
    Y = _as_dense(Y)
    l0 = _recycle(l0, n, p)                      # length-n vector: row i -> l0[i]
    f0 = np.asarray(f0, dtype=np.float64).reshape(1, -1) if np.size(f0) > 1 \
        else np.asarray(f0, dtype=np.float64).reshape(())
    EL = np.asarray(EL, dtype=np.float64).reshape(n, -1)
    EF = np.asarray(EF, dtype=np.float64).reshape(p, -1)
    Beta = EL @ EF.T                             # :77 tcrossprod
    M = np.where(Beta < M, Beta, M)              # :78  std::min(M, Beta)
    n_updates = 0
    converged = False
    franges = []
    if var_type == "by_col":
        s2 = np.asarray(sigma2, dtype=np.float64).reshape(1, p)
        const0 = Y * s2 + Beta + 1               # :81  Y %*% Diagonal(p,sigma2) + Beta + 1
        for _i in range(1, maxiter + 1):         # :82
            sexp = l0 * (np.exp(M + (1 / (const0 - M)) * (s2 / 2)) * f0)    # :83
            f = Y - sexp + (Beta - M) * (1 / s2)                            # :84
            if _global_stop(f, tol):             # :85
                converged = True
                break
            M = M - f / (-sexp * (1 + 0.5 * np.power(const0 - M, -2.0)) - (1 / s2))   # :90
            n_updates += 1
    if var_type == "by_row" or var_type == "constant":                       # :94
        s2 = _recycle(sigma2, n, p)              # length n (row i) or scalar
        const0 = s2 * Y + Beta + 1               # :95
        for _i in range(1, maxiter + 1):         # :96
            sexp = l0 * (np.exp(M + (1 / (const0 - M)) * s2 / 2) * f0)      # :97
            f = Y - sexp + (Beta - M) / s2                                   # :98
            franges.append((float(np.min(f)), float(np.max(f))))             # :99 print(range(f))
            if _global_stop(f, tol):             # :100
                converged = True
                break
            M = M - f / (-sexp * (1 + 0.5 * np.power(const0 - M, -2.0)) - 1 / s2)   # :105
            n_updates += 1
    if info is not None:
        info.update(n_updates=n_updates, converged=converged, franges=franges)
    return np.asarray(M)                         # :108


# --------------------------------------------------------------------------
# a3: adjust_var_shape / my_ifelse -- R/ebpmf_log.R:475-494
# --------------------------------------------------------------------------
def adjust_var_shape(sigma2, var_type, n, p):
    sigma2 = np.asarray(sigma2, dtype=np.float64)
    if var_type == "constant":
        return np.full((n, p), float(sigma2.reshape(-1)[0]))
    if var_type == "by_row":
        return np.repeat(sigma2.reshape(n, 1), p, axis=1)
    if var_type == "by_col":
        return np.repeat(sigma2.reshape(1, p), n, axis=0)
    raise ValueError("Non-supported var type")


# --------------------------------------------------------------------------
# a4/a5: ELBO partial sums and v_sum -- R/ebpmf_log.R:274-287,316-327
# --------------------------------------------------------------------------
def elbo_partials(Y, M, V):
    Y = _as_dense(Y)
    sym = _rsum(Y * M)                                   # :316
    ssexp = _rsum(np.exp(M + V / 2))                     # :317
    slogv = _rsum(np.log(V) / 2 + SLOGV_CONST)           # :318
    return sym, ssexp, slogv


def v_sum_of(V, var_type):
    if var_type == "constant":
        return _rsum(V)                                  # :322
    if var_type == "by_col":
        return _rcolsums(V)                              # :324
    if var_type == "by_row":
        return _rrowsums(V)                              # :326
    raise ValueError("Non-supported var type")


# --------------------------------------------------------------------------
# a6: ebpmf_log_update_sigma2 -- R/ebpmf_log.R:413-447
# --------------------------------------------------------------------------
def ebpmf_log_update_sigma2(fit_flash, sigma2, v_sum, var_type, cap_var_mean_ratio, a0, b0, n, p):
    """``fit_flash`` is a mapping with ``R2`` (flash_fit$R2) and ``EL``/``EF``
    (flash_fit$EF[[1]], [[2]]; fitted(fit_flash) == EL %*% t(EF))."""
    R2 = np.asarray(fit_flash["R2"], dtype=np.float64)
    sigma2 = np.array(sigma2, dtype=np.float64)
    v_sum = np.asarray(v_sum, dtype=np.float64)
    cap = cap_var_mean_ratio

    def fitted():
        return np.asarray(fit_flash["EL"]) @ np.asarray(fit_flash["EF"]).T

    def gate(mx):
        with np.errstate(divide="ignore", invalid="ignore"):
            return mx > (np.log(cap) - np.log(np.exp(sigma2) - 1) - sigma2 / 2)

    if var_type == "constant":
        new = ((v_sum + _rsum(R2)) / 2 + b0) / (n * p / 2 + a0 + 1)          # :417,420
        if cap > 0:
            if bool(np.all(gate(np.max(fitted())))):
                sigma2 = np.asarray(new, dtype=np.float64)
        else:
            sigma2 = np.asarray(new, dtype=np.float64)
    elif var_type == "by_row":
        new = ((v_sum + R2) / 2 + b0) / (p / 2 + a0 + 1)                     # :428,431
        if cap > 0:
            # :425 rowMaxs is Rfast's (NAMESPACE: importFrom(Rfast,rowMaxs)) and Rfast::rowMaxs(x, value = FALSE)
            # returns the POSITION of each row's maximum (1-based, first on ties), not the value: the reference
            # compares a column index with the threshold.  Kept as is.
            idx = gate(np.argmax(fitted(), axis=1) + 1.0)
            sigma2[idx] = new[idx]
        else:
            sigma2 = new
    elif var_type == "by_col":
        new = ((v_sum + R2) / 2 + b0) / (n / 2 + a0 + 1)                     # :438,441
        if cap > 0:
            idx = gate(np.max(fitted(), axis=0))                             # :435 colMaxs
            sigma2[idx] = new[idx]
        else:
            sigma2 = new
    else:
        raise ValueError("Non-supported var type")
    return sigma2


# --------------------------------------------------------------------------
# a7: calc_ebpmf_log_obj -- R/ebpmf_log.R:406-410
# --------------------------------------------------------------------------
def calc_ebpmf_log_obj(n, p, sym, ssexp, slogv, sv, sigma2, R2, KL_LF, const, ss, a0, b0):
    sigma2 = np.asarray(sigma2, dtype=np.float64)
    sv = np.asarray(sv, dtype=np.float64)
    R2 = np.asarray(R2, dtype=np.float64)
    val = (sym - ssexp + slogv + 0.5 * n * p
           - _rsum(ss * np.log(2 * np.pi * sigma2) / 2)
           - _rsum(sv / 2 / sigma2)
           - _rsum(R2 / 2 / sigma2)
           + const + KL_LF
           - _rsum((a0 + 1) * np.log(sigma2))
           - _rsum(b0 / sigma2))
    return float(val)


# --------------------------------------------------------------------------
# a8: sum_lfactorial_sparseMat -- R/ebpmf_log.R:470-472 (and dense :175)
# --------------------------------------------------------------------------
def sum_lfactorial_sparseMat(Y):
    """sum(lfactorial(Y@x)); for a dense Y this is sum(lfactorial(Y))."""
    if _sp is not None and _sp.issparse(Y):
        x = np.asarray(Y.tocsc().data, dtype=np.float64)
    else:
        x = np.asarray(Y, dtype=np.float64).reshape(-1)
    return _rsum(_gammaln(x + 1.0))


# --------------------------------------------------------------------------
# a11: row batching -- R/ebpmf_log.R:195-199
# --------------------------------------------------------------------------
def r_cut_batches(n, n_batch):
    """split(1:n, cut(seq_along(1:n), n_batch, labels=FALSE)) -> list of 0-based
    index arrays.  base::cut.default with a break COUNT: n_batch+1 equally spaced
    breaks on range(x), outer breaks widened by diff(range)/1000, right-closed."""
    x = np.arange(1, n + 1, dtype=np.float64)
    if n_batch < 2:
        raise ValueError("cut: invalid number of intervals")
    rx = (1.0, float(n))
    dx = rx[1] - rx[0]
    if dx == 0:
        dx = abs(rx[0])
    breaks = np.linspace(rx[0], rx[1], n_batch + 1)
    breaks[0] = rx[0] - dx / 1000
    breaks[-1] = rx[1] + dx / 1000
    code = np.searchsorted(breaks, x, side="left")      # (b_k, b_{k+1}]
    return [np.nonzero(code == k)[0] for k in range(1, n_batch + 1) if np.any(code == k)]


# --------------------------------------------------------------------------
# the fused VGA step of ebpmf_log's outer loop -- R/ebpmf_log.R:254-334
# --------------------------------------------------------------------------
def ebpmf_log_vga_step(M, Y, EL, EF, sigma2, var_type, maxiter_vga=10, vga_tol=1e-5,
                       batches=None):
    """One VGA step exactly as ebpmf_log runs it.

    ``batches is None``  -> the un-batched branch R/ebpmf_log.R:302-327.
    ``batches`` (list of 0-based row index arrays) -> the batched branch :254-293
    (stop test scoped to each batch; by_row v_sum concatenated).
    Returns dict(M, V, sym, ssexp, slogv, v_sum, n_updates, converged).
    """
    M = np.array(M, dtype=np.float64, order="F")
    n, p = M.shape
    Yd = _as_dense(Y)
    EL = np.asarray(EL, dtype=np.float64).reshape(n, -1)
    EF = np.asarray(EF, dtype=np.float64).reshape(p, -1)
    sigma2 = np.asarray(sigma2, dtype=np.float64).reshape(-1)
    if batches is None:
        info = {}
        res = vga_pois_solver_mat_newton(M, Yd, 1, EL @ EF.T,
                                         adjust_var_shape(sigma2, var_type, n, p),
                                         maxiter=maxiter_vga, tol=vga_tol, return_V=True, info=info)
        sym, ssexp, slogv = elbo_partials(Yd, res["M"], res["V"])
        return dict(M=res["M"], V=res["V"], sym=sym, ssexp=ssexp, slogv=slogv,
                    v_sum=v_sum_of(res["V"], var_type),
                    n_updates=[info["n_updates"]], converged=[info["converged"]])
    Mout = M.copy(order="F")
    Vout = np.empty_like(Mout)
    sym = ssexp = slogv = 0.0
    v_sum = 0.0
    v_rows = []
    n_updates, converged = [], []
    for b in batches:                                                       # :259
        y_b = Yd[b, :]
        sig = sigma2[b] if var_type == "by_row" else adjust_var_shape(sigma2, var_type, len(b), p)   # :267
        info = {}
        res = vga_pois_solver_mat_newton(M[b, :], y_b, 1, EL[b, :] @ EF.T, sig,
                                         maxiter=maxiter_vga, tol=vga_tol, return_V=True, info=info)
        Mout[b, :] = res["M"]                                               # :270
        Vout[b, :] = res["V"]
        s1, s2_, s3 = elbo_partials(y_b, res["M"], res["V"])
        sym += s1; ssexp += s2_; slogv += s3                                # :274-276
        if var_type == "by_row":
            v_rows.append(_rrowsums(res["V"]))                              # :282
        elif var_type == "by_col":
            v_sum = v_sum + _rcolsums(res["V"])                             # :284
        else:
            v_sum = v_sum + _rsum(res["V"])                                 # :286
        n_updates.append(info["n_updates"]); converged.append(info["converged"])
    if var_type == "by_row":
        v_sum = np.concatenate(v_rows)                                      # :292-293
    return dict(M=Mout, V=Vout, sym=sym, ssexp=ssexp, slogv=slogv, v_sum=v_sum,
                n_updates=n_updates, converged=converged)


# --------------------------------------------------------------------------
# N4: calc_split_PMF_obj_flashier -- inst/code/Poisson_MF_splitting.R:373-391 (archived driver; un-fused ELBO)
# --------------------------------------------------------------------------
def calc_split_PMF_obj_flashier(Y, S, sigma2, M, V, R2, KL_LF, const, var_type):
    """``R2`` is fit_flash$flash.fit$R2 (:374).  sv / ss per var_type (:377-388); the value (:389), every sum()
    at its own length."""
    Y = np.asarray(Y, dtype=np.float64); M = np.asarray(M, dtype=np.float64); V = np.asarray(V, dtype=np.float64)
    S = np.asarray(S, dtype=np.float64)
    sigma2 = np.asarray(sigma2, dtype=np.float64); R2 = np.asarray(R2, dtype=np.float64)
    n, p = Y.shape
    if var_type == "by_row":
        sv, ss = V.sum(axis=1), p
    elif var_type == "by_col":
        sv, ss = V.sum(axis=0), n
    elif var_type == "constant":
        sv, ss = V.sum(), n * p
    else:
        raise ValueError("Non-supported var type")
    return (_rsum(Y * M - S * np.exp(M + V / 2) + np.log(2 * np.pi * V) / 2 + 0.5) - _rsum(ss * np.log(2 * np.pi * sigma2) / 2)
            - _rsum(sv / 2 / sigma2) - _rsum(R2 / 2 / sigma2) + const + KL_LF)
