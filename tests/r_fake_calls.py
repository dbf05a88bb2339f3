"""A TEST DOUBLE for `.Call`: answers the glue's routines from the CPU oracle instead of the compiled glue + GPU.

Only for the build container, where the reference checkout exists but no GPU does: it lets tests/test_r_boundary.py
run the reference's own `ebpmf_log()` with the drop-in installed and check the R-level plumbing (what is passed where,
with which storage mode) against the reference's traces.  The arithmetic behind it is the oracle's, so it says
nothing about the CUDA path -- the GPU tests of the same file do that, through the real glue."""
import numpy as np

import r_mock
from oracle import mini_r as R, vga_numpy as vn


def _list(d, attrs=None):
    out = R.RList(list(d.values()), list(d.keys()))
    out.attrs = attrs or {}
    return out


def _step_list(r, want_v):
    return _list(dict(M=np.asfortranarray(r["M"]), V=np.asfortranarray(r["V"]) if want_v else None,
                      v_sum=np.atleast_1d(np.asarray(r["v_sum"], dtype=np.float64)), sym=np.array([r["sym"]]), ssexp=np.array([r["ssexp"]]),
                      slogv=np.array([r["slogv"]]), n_updates=r_mock.as_int([max(np.atleast_1d(r["n_updates"]))]),
                      converged=np.array([bool(np.all(r["converged"]))])))


class OracleBackedCalls:
    """Install with `session.dot_call = OracleBackedCalls(session)`; keeps the resident state a context would keep."""

    def __init__(self, session):
        self.s, self.Y, self.M, self.row0, self.log = session, None, None, None, []

    def __call__(self, name, *args):
        a = [x if isinstance(x, r_mock.RExt) else r_mock.from_sexp(r_mock.to_sexp(x)) for x in args]      # values must survive the SEXP trip
        self.log.append(name)
        is_int = lambda i: isinstance(a[i], r_mock.RInt)                                                 # noqa: E731
        is_real = lambda i: isinstance(a[i], np.ndarray) and not is_int(i) and a[i].dtype == np.float64   # noqa: E731
        vt = lambda x: R.r_string(x)                                                                     # noqa: E731
        f = lambda x: float(x[0])                                                                        # noqa: E731
        if name == "C_ebpmf_ctx_new":
            assert is_int(0)
            return r_mock.RExt(None)
        if name == "C_ebpmf_ctx_set_y":
            self.Y = a[1].dense if isinstance(a[1], R.RSparse) else a[1]
            assert self.Y.ndim == 2
            return None
        if name == "C_ebpmf_ctx_set_m":
            assert is_real(1) and a[1].ndim == 2
            self.M = a[1]
            return None
        if name == "C_ebpmf_sum_lfactorial":
            return np.array([vn.sum_lfactorial_sparseMat(self.Y if a[1] is None else a[1])])
        if name == "C_ebpmf_vga_step_resident":
            dims, EL, EF, s2, v, mi, tol, wv = a[1:]
            assert is_int(1) and is_int(6) and wv.dtype == bool and list(dims) == list(self.M.shape)
            r = vn.ebpmf_log_vga_step(self.M, self.Y, EL, EF, s2, vt(v), int(mi[0]), f(tol))
            self.M = r["M"]
            return _step_list(r, bool(wv[0]))
        if name == "C_ebpmf_batched_new":
            assert isinstance(a[0], R.RSparse) and is_real(1) and is_int(2) and is_int(3)
            self.Y, self.row0 = a[0].dense, a[1].astype(int)
            return r_mock.RExt(None)
        if name == "C_ebpmf_batched_vga_step":
            M, EL, EF, s2, v, mi, tol, wv, nb = a[1:]
            assert is_int(9) and int(nb[0]) == len(self.row0) - 1
            batches = [np.arange(self.row0[i], self.row0[i + 1]) for i in range(len(self.row0) - 1)]
            r = vn.ebpmf_log_vga_step(M, self.Y, EL, EF, s2, vt(v), int(mi[0]), f(tol), batches=batches)
            out = _step_list(r, bool(wv[0]))
            out.attrs = {"n_updates_per_batch": r_mock.as_int(r["n_updates"])}
            return out
        if name == "C_ebpmf_group_new":
            assert is_int(0) and a[0].size >= 1
            self.group_devices = [int(v) for v in a[0]]
            return r_mock.RExt(None)
        if name == "C_ebpmf_group_set_y":
            assert isinstance(a[1], R.RSparse)            # the group wants the dgCMatrix (rows are dealt by 256-row blocks)
            self.Y = a[1].dense
            return None
        if name == "C_ebpmf_group_set_m":
            assert is_real(1) and a[1].ndim == 2
            self.M = a[1]
            return None
        if name == "C_ebpmf_group_sum_lfactorial":
            return np.array([vn.sum_lfactorial_sparseMat(self.Y)])
        if name == "C_ebpmf_group_vga_step":
            M, dims, EL, EF, s2, v, mi, tol, wv = a[1:]
            assert is_int(2) and is_int(7) and wv.dtype == bool and list(dims) == list(self.Y.shape)
            r = vn.ebpmf_log_vga_step(self.M if M is None else M, self.Y, EL, EF, s2, vt(v), int(mi[0]), f(tol))
            self.M = r["M"]
            return _step_list(r, bool(wv[0]))
        if name == "C_vga_pois_solver_mat_newton":
            M, X, S, Beta, S2, mi, tol, rv = a[1:]
            n, p = M.shape
            X = X.dense if isinstance(X, R.RSparse) else X
            shape = lambda o: o.reshape((n, p), order="F") if o.size == n * p else (o if o.size != 1 else float(o[0]))     # noqa: E731
            r = vn.vga_pois_solver_mat_newton(M, X, shape(S), shape(Beta), shape(S2), int(mi[0]), f(tol))
            return _list(dict(M=np.asfortranarray(r["M"]), V=np.asfortranarray(r["V"]))) if rv[0] else np.asfortranarray(r["M"])
        if name == "C_ebpmf_log_update_sigma2":
            s2, vs, R2, v, cap, a0, b0, n, p, EL, EF = a[1:]
            r = vn.ebpmf_log_update_sigma2(dict(R2=R2, EL=EL, EF=EF), s2, vs if vt(v) != "constant" else f(vs), vt(v), f(cap), f(a0), f(b0),
                                           int(n[0]), int(p[0]))
            return np.atleast_1d(np.asarray(r, dtype=np.float64))
        if name == "C_calc_ebpmf_log_obj":
            n, p, sym, ssexp, slogv, sv, s2, R2, KL, const, ss, a0, b0 = a[1:]
            return np.array([vn.calc_ebpmf_log_obj(int(n[0]), int(p[0]), f(sym), f(ssexp), f(slogv), sv, s2, R2, f(KL), f(const), f(ss), f(a0), f(b0))])
        raise AssertionError("the test double does not model " + name)
