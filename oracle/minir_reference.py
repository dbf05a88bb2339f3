"""Runs the REFERENCE'S OWN R SOURCE for the hot path under oracle/mini_r.py -- TEST INFRASTRUCTURE ONLY.

The reference's files are read from a checkout (`/root/reference` in the build container; it does not exist on the
GPU box, so nothing here is called by `-m gpu` tests, smoke() or bench.py -- they use the committed outputs,
tests/golden/minir_reference.npz, made by tests/golden/make_golden_minir.py).  Three levels:

* `run_functions`  -- the exported functions of the path called directly, the same calls as
  tests/golden/make_golden.R makes under GNU R: `vga_pois_solver_mat_newton` (R/vga_solver_mat.R:6-41),
  `_fixed_iter` (:44-65), `_low_memory` (:71-109), `adjust_var_shape` (R/ebpmf_log.R:475-486),
  `ebpmf_log_update_sigma2` (:413-447), `calc_ebpmf_log_obj` (:406-410), `sum_lfactorial_sparseMat` (:470-472);
* `run_loop_blocks` -- the statements of `ebpmf_log`'s outer loop that are NOT a function of their own (the
  inline `sym`/`ssexp`/`slogv`/`v_sum` sums, the `cut()` row batches and the batched branch with its per-batch stop
  scope, R/ebpmf_log.R:195-203 and :254-334), taken out of the parsed body of `ebpmf_log` and executed in place;
* `run_driver` -- `ebpmf_log()` itself, with the out-of-scope callees (vebpm's init, flashier's fit) replaced by
  oracle/standin_flash.py at exactly the boundary SURVEY.md section 8b draws.
"""
import hashlib
import os

import numpy as np

from . import mini_r as R
from .standin_flash import standin_flash

FILES = ("R/vga_solver_mat.R", "R/ebpmf_log.R", "R/ebpmf_log_controls.R")
CASES = {"def": (10, 1e-5), "exh": (2, 1e-5), "tight": (1000, 1e-8)}


def provenance(ref_root):
    out = {}
    for f in FILES:
        with open(os.path.join(ref_root, f), "rb") as fh:
            out[f] = hashlib.sha256(fh.read()).hexdigest()
    return out


def load_reference(ref_root, echo=None):
    """An interpreter with the three reference files sourced and the out-of-scope callees stood in for."""
    it = R.Interp(echo=echo)
    for f in FILES:
        it.source(os.path.join(ref_root, f))

    # ---- stand-ins for what SURVEY.md section 8 puts outside the boundary ----------------------------------------
    def ebnm_stub(it_, x, s, *a, **kw):                       # posterior mean of mixed sign -> sign constraint 0
        return R.to_r({"posterior": {"mean": x}})

    def init_stub(it_, Y, l0, f0, sigma2_init, var_type, M_init, *a, **kw):      # vebpm's initialisation
        return R.RList([sigma2_init, M_init], ["sigma2_init", "M_init"])

    def fit_of(M, l0, f0, K, var_type, sigma2):
        s = standin_flash(M, l0, f0, K, var_type)
        flash_fit = R.RList([np.asfortranarray(M), R.RList([s["EL"], s["EF"]]), s["R2"], R.RList([s["KL1"], s["KL2"]]),
                             R.RList([np.asarray(l0, dtype=np.float64).ravel(), np.asarray(f0, dtype=np.float64).ravel(), R.num(K), var_type])],
                            ["Y", "EF", "R2", "KL", "standin"])
        return R.RList([flash_fit, s["EL"], s["EF"], np.sqrt(R.dense(sigma2)), R.num(K + 2)],
                       ["flash_fit", "L_pm", "F_pm", "residuals_sd", "n_factors"])

    def flash_init_stub(it_, M, sigma2, l0, f0, ones_n, ones_p, *a, **kw):
        return fit_of(R.dense(M), R.dense(l0), R.dense(f0), int(it_.get("standin_K")[0]), R.r_string(it_.get("standin_var_type")), sigma2)

    def flash_update_stub(it_, fit_flash, sigma2, *a, **kw):
        ff = fit_flash.get("flash_fit")
        l0, f0, K, vt = ff.get("standin").vals
        return fit_of(R.dense(ff.get("Y")), l0, f0, int(K[0]), vt if isinstance(vt, str) else R.r_string(vt), sigma2)

    it.def_builtin("ebnm_point_normal", ebnm_stub)
    it.def_builtin("ebnm_normal", ebnm_stub)
    it.def_builtin("ebpmf_log_init", init_stub)
    it.def_builtin("ebpmf_log_flash_init", flash_init_stub)
    it.def_builtin("ebpmf_log_flash_update", flash_update_stub)
    it.def_builtin("flash_fit_get_KL", lambda it_, ff, n: ff.get("KL").vals[int(R.scalar(n)) - 1])
    return it


def _fit_flash(M, EL, EF, R2):
    """The fields of a flash fit the hot path reads: flash_fit$Y, flash_fit$EF, flash_fit$R2, fitted()."""
    return R.RList([R.RList([np.asfortranarray(M), R.RList([np.asfortranarray(EL), np.asfortranarray(EF)]), np.asarray(R2, dtype=np.float64)],
                            ["Y", "EF", "R2"]), np.asfortranarray(EL), np.asfortranarray(EF)], ["flash_fit", "L_pm", "F_pm"])


def update_count(it, M0, Y, Beta, S2, maxiter, tol, M_final):
    """The reference does not return its update count; it is the smallest k whose maxiter = k run ends at the same
    M (the same device tests/golden/make_golden.R uses); 0 when the first stop test already passes."""
    M1 = it.call("vga_pois_solver_mat_newton", M0, Y, 1, Beta, S2, maxiter=1, tol=np.inf, return_V=False)
    if np.array_equal(M1, M_final):
        return 0
    for k in range(1, maxiter + 1):
        Mk = it.call("vga_pois_solver_mat_newton", M0, Y, 1, Beta, S2, maxiter=k, tol=tol, return_V=False)
        if np.array_equal(Mk, M_final):
            return k
    return maxiter


def run_functions(it, G):
    """G: the committed inputs (tests/golden/vga_golden.npz).  Returns the arrays make_golden.R would write."""
    out = {}
    Y = np.asarray(G["Y"], dtype=np.float64)
    n, p = Y.shape
    for vt in ("by_col", "by_row", "constant"):
        M0, EL, EF, sigma2, R2 = (np.asarray(G["%s_%s" % (vt, k)], dtype=np.float64) for k in ("M0", "EL", "EF", "sigma2", "R2"))
        Beta = it.call("tcrossprod", EL, EF)                                        # fitted(fit_flash), R/ebpmf_log.R:305
        S2 = it.call("adjust_var_shape", sigma2, vt, n, p)                          # :306
        for tag, (maxiter, tol) in CASES.items():
            res = it.call("vga_pois_solver_mat_newton", M0, Y, 1, Beta, S2, maxiter=maxiter, tol=tol, return_V=True)
            M, V = res.get("M"), res.get("V")
            u = update_count(it, M0, Y, Beta, S2, maxiter, tol, M)
            # the inline sums of the loop body (R/ebpmf_log.R:316-327) are not a function: they are taken from the
            # parsed body of ebpmf_log and executed in place (run_loop_blocks), on the same call
            blk = run_loop_blocks(it, Y, M0, EL, EF, sigma2, R2, vt, maxiter_vga=maxiter, vga_tol=tol, est_sigma2=False)
            assert np.array_equal(blk["M"], M) and np.array_equal(blk["V"], V)
            sums = {k: np.atleast_1d(np.asarray(blk[k], dtype=np.float64)) for k in ("sym", "ssexp", "slogv", "v_sum")}
            k0 = "%s_%s_" % (vt, tag)
            out[k0 + "M"], out[k0 + "V"], out[k0 + "v_sum"] = M, V, sums["v_sum"]
            out[k0 + "scal"] = np.array([sums["sym"][0], sums["ssexp"][0], sums["slogv"][0], float(u)])
            if tag == "def":
                s2new = it.call("ebpmf_log_update_sigma2", _fit_flash(M, EL, EF, R2), sigma2, sums["v_sum"], vt, 0, 1, 1, n, p)
                ss = {"by_col": n, "by_row": p, "constant": n * p}[vt]              # var_offset_for_obj, :178-192
                obj = it.call("calc_ebpmf_log_obj", n, p, sums["sym"], sums["ssexp"], sums["slogv"], sums["v_sum"],
                              s2new, R2, -123.4, -5.5, ss, 1, 1)
                out[k0 + "sigma2_new"], out[k0 + "obj"] = s2new, obj
                # the cap gate of ebpmf_log_update_sigma2 (cap_var_mean_ratio > 0), R/ebpmf_log.R:415-419,424-429,435-440
                out[k0 + "sigma2_new_cap"] = it.call("ebpmf_log_update_sigma2", _fit_flash(M, EL, EF, R2), sigma2, sums["v_sum"],
                                                     vt, 2.0, 1, 1, n, p)
                # ... and with a ratio so large (threshold ~ 20) that the gate SELECTS: for by_row what it compares with the
                # threshold is Rfast::rowMaxs' column position, for by_col matrixStats::colMaxs' value
                out[k0 + "sigma2_new_cap_big"] = it.call("ebpmf_log_update_sigma2", _fit_flash(M, EL, EF, R2), sigma2, sums["v_sum"],
                                                         vt, 1e9, 1, 1, n, p)
    M0, EL, EF, sigma2 = (np.asarray(G["by_col_" + k], dtype=np.float64) for k in ("M0", "EL", "EF", "sigma2"))
    Beta = it.call("tcrossprod", EL, EF)
    S2 = it.call("adjust_var_shape", sigma2, "by_col", n, p)
    fi = it.call("vga_pois_solver_mat_newton_fixed_iter", M0, None, Y, 1, Beta, S2, maxiter=1)
    out["fixed1_M"], out["fixed1_V"] = fi.get("M"), fi.get("V")
    fi = it.call("vga_pois_solver_mat_newton_fixed_iter", M0, None, Y, 1, Beta, S2, maxiter=10)
    out["fixed10_M"], out["fixed10_V"] = fi.get("M"), fi.get("V")
    for vt in ("by_col", "by_row", "constant"):                                     # _low_memory prints range(f) for by_row/constant
        s2 = np.asarray(G[vt + "_sigma2"], dtype=np.float64)
        lm = it.call("vga_pois_solver_mat_newton_low_memory", np.log(Y + 0.5), R.RSparse(Y), G["L0"], G["F0"],
                     np.asarray(G[vt + "_EL"])[:, 2:], np.asarray(G[vt + "_EF"])[:, 2:], s2, vt, maxiter=1000, tol=1e-8)
        out["lowmem_M" if vt == "by_col" else "lowmem_%s_M" % vt] = lm
    out["lfact"] = it.call("sum_lfactorial_sparseMat", R.RSparse(Y))
    return {k: np.asarray(R.dense(v), dtype=np.float64) for k, v in out.items()}


def _find(stmts, pred):
    for i, s in enumerate(stmts):
        if pred(s):
            return i
    raise R.RError("statement not found in the parsed body of ebpmf_log")


def loop_statements(it):
    """(n_batch assignment, the `## split data` if, the VGA if/else of the outer loop) out of ebpmf_log's body."""
    body = it.get("ebpmf_log").body[1]
    i_nb = _find(body, lambda s: s[0] == "assign" and s[1] == ("name", "n_batch"))
    split_if = body[i_nb + 1]
    assert split_if[0] == "if" and split_if[1][0] == "binop" and split_if[1][2] == ("name", "n_batch")
    loop = body[_find(body, lambda s: s[0] == "for" and s[1] == "iter")]
    lbody = loop[3][1]
    vga_if = lbody[_find(lbody, lambda s: s[0] == "if" and s[1][0] == "binop" and s[1][2] == ("name", "n_batch"))]
    return body[i_nb], split_if, vga_if


def run_loop_blocks(it, Y, M0, EL, EF, sigma2, R2, var_type, batch_size=np.inf, maxiter_vga=10, vga_tol=1e-5, est_sigma2=True,
                    cap_var_mean_ratio=0.0):
    """One pass of ebpmf_log's outer-loop VGA block (R/ebpmf_log.R:254-334), executed in place on a prepared state."""
    Y = np.asarray(Y, dtype=np.float64)
    n, p = Y.shape
    env = R.Env(it.globalenv)
    for k, v in dict(Y=np.asfortranarray(Y), n=n, p=p, var_type=var_type, sigma2=sigma2, fit_flash=_fit_flash(M0, EL, EF, R2)).items():
        env.set(k, R.to_r(v))
    it.run("general_control = modifyList(ebpmf_log_general_control_default(), list(batch_size=%s, save_latent_M=TRUE), keep.null=TRUE)"
           % ("Inf" if np.isinf(batch_size) else repr(int(batch_size))), env)
    it.run("vga_control = modifyList(ebpmf_log_vga_control_default(), list(maxiter_vga=%r, vga_tol=%r))" % (maxiter_vga, vga_tol), env)
    it.run("sigma2_control = modifyList(ebpmf_log_sigma2_control_default(), list(est_sigma2=%s, cap_var_mean_ratio=%r))"
           % ("TRUE" if est_sigma2 else "FALSE", cap_var_mean_ratio), env)
    nb_stmt, split_if, vga_if = loop_statements(it)
    it.eval(nb_stmt, env)
    it.eval(split_if, env)
    it.eval(vga_if, env)
    g = lambda k: np.asarray(R.dense(it.lookup(k, env)), dtype=np.float64)          # noqa: E731
    out = dict(M=R.dense(env.vars["fit_flash"].get("flash_fit").get("Y")), v_sum=g("v_sum"), sym=g("sym")[0], ssexp=g("ssexp")[0],
               slogv=g("slogv")[0], sigma2=g("sigma2"), n_batch=int(g("n_batch")[0]))
    if out["n_batch"] > 1:
        out["batch_row0"] = np.array([int(b[0]) - 1 for b in env.vars["batches"].vals] + [n], dtype=np.int64)
    else:
        out["V"] = g("V")
    return out


def run_driver(it, Y, l0, f0, M_init, sigma2_init, var_type, K=3, outer_iters=4, batch_size=np.inf, sparse=False):
    """ebpmf_log() itself for `outer_iters` outer iterations (conv_tol = -Inf so that it never stops early).
    With row batches the reference cannot return M: `save_latent_M = TRUE` reads a `V` the batched branch never
    assigns (R/ebpmf_log.R:389 vs :254-300, "object 'V' not found" -- reproduced by the evaluator), so batched runs
    are made with save_latent_M = FALSE and compared through the ELBO trace and sigma2."""
    batched = bool(np.isfinite(batch_size) and batch_size < np.asarray(Y).shape[0])
    it.set("standin_K", K)
    it.set("standin_var_type", var_type)
    Yr = R.RSparse(Y) if sparse else np.asfortranarray(np.asarray(Y, dtype=np.float64))
    gc = R.to_r(dict(maxiter=outer_iters, conv_tol=-np.inf, save_latent_M=np.array([not batched]), batch_size=batch_size, verbose=np.array([False])))
    ic = R.to_r(dict(M_init=np.asfortranarray(M_init), sigma2_init=np.asarray(sigma2_init, dtype=np.float64)))
    fit = it.call("ebpmf_log", Yr, np.asarray(l0, dtype=np.float64).reshape(-1, 1), np.asarray(f0, dtype=np.float64).reshape(-1, 1),
                  var_type=var_type, general_control=gc, init_control=ic, verbose=np.array([False]))
    ff = fit.get("fit_flash").get("flash_fit")
    out = dict(elbo_trace=R.dense(fit.get("elbo_trace")), sigma2=R.dense(fit.get("sigma2")), K_trace=R.dense(fit.get("K_trace")),
               elbo=R.dense(fit.get("elbo")))
    if not batched:
        out["M"], out["V"] = R.dense(ff.get("Y")), R.dense(ff.get("V"))
    return out
