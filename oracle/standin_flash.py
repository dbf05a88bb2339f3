"""A deterministic stand-in for the flashier side of `ebpmf_log` -- TEST INFRASTRUCTURE ONLY.

`ebpmf_log` (R/ebpmf_log.R:116-403) alternates the VGA step (the hot path rebuilt in this repository) with an
EBMF fit by flashier (out of scope, SURVEY.md section 8: `EL`, `EF`, `R2`, `KL` cross the boundary as inputs).  To
run the reference's own driver function end to end -- under oracle/mini_r.py -- and the CUDA path through the same
outer loop, both sides need *something* that turns the current M into factors.  This is that something: a
rank-K truncated SVD of M minus the two fixed offset factors, shrunk by a constant, with residual sums of squares
as R2 and a made-up KL term.  It is NOT an EBMF fit and makes no statistical claim; it only has to be a
deterministic, smooth function of M so that both sides see the same inputs.

The same function is installed into the interpreter as `ebpmf_log_flash_init` / `ebpmf_log_flash_update` /
`flash_fit_get_KL` (tests/golden/make_golden_minir.py) and called directly by the product-side loop in
tests/test_mini_r.py."""
import numpy as np


def standin_flash(M, l0, f0, K, var_type, shrink=0.9):
    """-> dict(EL n x (K+2), EF p x (K+2), R2, KL1, KL2).  Factor 1 = (l0, 1), factor 2 = (1, f0), as ebpmf_log sets
    them up (R/ebpmf_log_flash_init.R, R/ebpmf_log_flash_update.R:17-18)."""
    M = np.asarray(M, dtype=np.float64)
    n, p = M.shape
    l0 = np.asarray(l0, dtype=np.float64).reshape(n)
    f0 = np.asarray(f0, dtype=np.float64).reshape(p)
    U, s, Vt = np.linalg.svd(M - l0[:, None] - f0[None, :], full_matrices=False)
    EL = np.column_stack([l0, np.ones(n), U[:, :K] * (shrink * s[:K])])
    EF = np.column_stack([np.ones(p), f0, Vt[:K].T])
    resid2 = (M - EL @ EF.T) ** 2
    if var_type == "by_col":
        R2 = resid2.sum(axis=0) + 0.25
    elif var_type == "by_row":
        R2 = resid2.sum(axis=1) + 0.25
    elif var_type == "constant":
        R2 = np.array([resid2.sum() + 0.25])
    else:
        raise ValueError("Non-supported var type")
    KL1 = -1e-3 * (EL ** 2).sum(axis=0)
    KL2 = -1e-3 * (EF ** 2).sum(axis=0)
    return dict(EL=np.asfortranarray(EL), EF=np.asfortranarray(EF), R2=R2, KL1=KL1, KL2=KL2)


def outer_loop(Y, l0, f0, M_init, sigma2_init, var_type, K, outer_iters, step=None, update_sigma2=None, calc_obj=None, const=0.0,
               a0=1.0, b0=1.0, block=None):
    """`ebpmf_log`'s outer loop (R/ebpmf_log.R:228-352) around a backend's VGA step, with `standin_flash` where the
    reference calls flashier -- the product-side / oracle-side counterpart of minir_reference.run_driver.

    step(M, EL, EF, sigma2) -> dict(M, V, v_sum, sym, ssexp, slogv); update_sigma2(fit, sigma2, v_sum) -> sigma2;
    calc_obj(n, p, sym, ssexp, slogv, v_sum, sigma2, R2, KL_LF, const, ss, a0, b0) -> float.
    Instead of step + update_sigma2, `block(M, fit, sigma2)` may do both and return the new sigma2 in its dict (the
    shape of ebpmf_b200_vga_block, ebpmf.alpha_b200/r/R/ebpmf_log_vga_block.R).
    Returns dict(elbo_trace, sigma2, M, V)."""
    Y = np.asarray(Y, dtype=np.float64)
    n, p = Y.shape
    ss = {"by_col": n, "by_row": p, "constant": n * p}[var_type]            # var_offset_for_obj, R/ebpmf_log.R:178-192
    M = np.asfortranarray(M_init)
    fit = standin_flash(M, l0, f0, K, var_type)                             # ebpmf_log_flash_init
    sigma2 = np.sqrt(np.asarray(sigma2_init, dtype=np.float64)) ** 2        # sigma2 = fit_flash$residuals_sd^2, :229
    obj = [-np.inf]
    res = None
    for _ in range(outer_iters):
        if block is not None:
            res = block(M, fit, sigma2)                                     # :254-334 in one call
            M, sigma2 = res["M"], res["sigma2"]
        else:
            res = step(M, fit["EL"], fit["EF"], sigma2)                     # :254-327
            M = res["M"]
            sigma2 = update_sigma2(dict(R2=fit["R2"], EL=fit["EL"], EF=fit["EF"]), sigma2, res["v_sum"])     # :329-332
        fit = standin_flash(M, l0, f0, K, var_type)                         # ebpmf_log_flash_update, :343-349
        KL_LF = float(np.sum(fit["KL1"])) + float(np.sum(fit["KL2"]))       # :352
        obj.append(float(calc_obj(n, p, res["sym"], res["ssexp"], res["slogv"], res["v_sum"], sigma2, fit["R2"], KL_LF, const, ss, a0, b0)))
    return dict(elbo_trace=np.array(obj), sigma2=np.atleast_1d(np.asarray(sigma2, dtype=np.float64)), M=np.asarray(M), V=res.get("V"))
