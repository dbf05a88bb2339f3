"""Host-side mirror of the reference's R operator interface for the VGA hot path.

Same names, argument order / meaning, defaults, return shapes and error
behaviour as the R functions (paths relative to the reference root):

  vga_pois_solver_mat_newton             R/vga_solver_mat.R:6-41
  vga_pois_solver_mat_newton_fixed_iter  R/vga_solver_mat.R:44-65
  vga_pois_solver_mat_newton_low_memory  R/vga_solver_mat.R:71-109
  ebpmf_log_update_sigma2                R/ebpmf_log.R:413-447
  calc_ebpmf_log_obj                     R/ebpmf_log.R:406-410
  sum_lfactorial_sparseMat               R/ebpmf_log.R:470-472
  adjust_var_shape                       R/ebpmf_log.R:475-486

R itself is not installed here, so this Python layer plays the part of the
`.Call` wrappers in ebpmf.alpha_b200/r/R/vga_cuda.R: it only marshals arrays
(column-major fp64, int32 CSC slots) into the C ABI of include/ebpmf_b200.h.
All arithmetic runs in the CUDA library; nothing here computes on the CPU and
nothing imports ``oracle/``.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import EbpmfError, SolveInfo, VAR_TYPE

try:
    import scipy.sparse as _sp
except Exception:  # pragma: no cover
    _sp = None


# ---------------------------------------------------------------------------
# marshalling helpers
# ---------------------------------------------------------------------------
def _f(a):
    return np.asfortranarray(np.asarray(a, dtype=np.float64))


def _vec(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(-1))


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _is_sparse(Y):
    return _sp is not None and _sp.issparse(Y)


def _csc_slots(Y):
    """dgCMatrix slots (@p, @i, @x) of a scipy sparse matrix."""
    Yc = Y.tocsc()
    if not Yc.has_sorted_indices:
        Yc = Yc.sorted_indices()
    return (np.ascontiguousarray(Yc.indptr, dtype=np.int32), np.ascontiguousarray(Yc.indices, dtype=np.int32),
            np.ascontiguousarray(Yc.data, dtype=np.float64))


def _operand(a, n, p, what):
    """An operand R would recycle against an n x p matrix -> (buffer, EBPMF_OP_* kind)."""
    a = np.asarray(a, dtype=np.float64)
    if a.ndim == 0 or a.size == 1:
        return _vec(a), _lib.OP_SCALAR
    if a.ndim == 1:
        if a.size == n:
            return _vec(a), _lib.OP_ROWVEC      # R recycles a length-n vector down each column
        raise ValueError("%s: a vector operand must have length nrow(M) = %d" % (what, n))
    if a.shape == (n, p):
        return _f(a), _lib.OP_MATRIX
    raise ValueError("%s: non-conformable arrays" % what)


class _PoolBlock:
    """Owner of one block of the library's host pool; gives it back when the last view dies (what R's garbage
    collector does through the R_allocator_t of the .Call glue)."""

    def __init__(self, nbytes):
        lib = _lib.load()
        self._lib = lib
        self.ptr = lib.ebpmf_host_alloc(C.c_size_t(max(1, int(nbytes))))
        if not self.ptr:
            raise MemoryError("ebpmf_host_alloc(%d) failed" % nbytes)

    def __del__(self):  # pragma: no cover
        try:
            if self.ptr:
                self._lib.ebpmf_host_free(C.c_void_p(self.ptr))
                self.ptr = None
        except Exception:
            pass


def host_matrix(n, p):
    """A column-major n x p float64 result matrix in the library's warm host pool (ebpmf_host_alloc): the Python
    stand-in for the R glue's ``allocVector3(REALSXP, n*p, &ebpmf_allocator)``.  The block returns to the pool when
    the array (and every view of it) is garbage collected."""
    blk = _PoolBlock(8 * int(n) * int(p))
    buf = (C.c_double * (int(n) * int(p))).from_address(blk.ptr)
    buf._ebpmf_owner = blk                       # the ctypes array keeps the block alive; NumPy keeps the ctypes array
    a = np.frombuffer(buf, dtype=np.float64, count=int(n) * int(p)).reshape((int(p), int(n))).T
    return a


def host_pool_stats():
    out = np.zeros(4)
    _lib.load().ebpmf_host_pool_stats(_ptr(out))
    return dict(fresh=int(out[0]), reused=int(out[1]), kept_bytes=float(out[2]), live_blocks=int(out[3]))


def host_pool_trim():
    _lib.load().ebpmf_host_pool_trim()


_DEFAULT_CTX = {}
_WHICH = {"M": 0, "V": 1, "M_in": 2}     # get_rows / get_block: result M, V, the M the last step started from


def default_context(device=0):
    """One lazily created context per device (the R glue keeps one in an external pointer)."""
    ctx = _DEFAULT_CTX.get(device)
    if ctx is None:
        ctx = VgaContext(device)
        _DEFAULT_CTX[device] = ctx
    return ctx


# ---------------------------------------------------------------------------
# the context: resident Y / M / factors / sigma2 of the fused ebpmf_log step
# ---------------------------------------------------------------------------
class VgaContext:
    """Owns one ``ebpmf_ctx`` (device buffers, streams, optional NCCL communicator)."""

    def __init__(self, device=0):
        self._lib = _lib.load()
        h = C.c_void_p()
        _lib.check(self._lib.ebpmf_ctx_create(int(device), C.byref(h)))
        self._h = h
        self.device = device
        self.n = self.p = None
        self.var_type = None

    def close(self):
        if getattr(self, "_h", None):
            self._lib.ebpmf_ctx_destroy(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        _lib.check(rc, self._h)

    # -- multi-GPU --------------------------------------------------------
    @staticmethod
    def nccl_unique_id():
        buf = C.create_string_buffer(128)
        _lib.check(_lib.load().ebpmf_nccl_unique_id(buf))
        return buf.raw

    def comm_init(self, rank, world, unique_id):
        buf = C.create_string_buffer(bytes(unique_id), 128)
        self._ck(self._lib.ebpmf_ctx_comm_init(self._h, int(rank), int(world), buf))

    def peer_count(self):
        out = C.c_int()
        self._ck(self._lib.ebpmf_ctx_peer_count(self._h, C.byref(out)))
        return out.value

    # -- resident state ---------------------------------------------------
    def set_y(self, Y):
        """Y: scipy sparse (-> dgCMatrix slots) or dense n x p."""
        if _is_sparse(Y):
            n, p = Y.shape
            cp, ri, x = _csc_slots(Y)
            self._ck(self._lib.ebpmf_ctx_set_y_csc(self._h, n, p, _ptr(cp), _ptr(ri), _ptr(x)))
        else:
            Yd = _f(Y)
            n, p = Yd.shape
            self._ck(self._lib.ebpmf_ctx_set_y_dense(self._h, n, p, _ptr(Yd)))
        self.n, self.p = n, p

    def set_y_csc(self, n, p, colptr, rowidx, x):
        cp = np.ascontiguousarray(colptr, dtype=np.int32)
        ri = np.ascontiguousarray(rowidx, dtype=np.int32)
        xx = np.ascontiguousarray(x, dtype=np.float64)
        self._ck(self._lib.ebpmf_ctx_set_y_csc(self._h, n, p, _ptr(cp), _ptr(ri), _ptr(xx)))
        self.n, self.p = n, p

    def set_m(self, M):
        M = _f(M)
        if M.shape != (self.n, self.p):
            raise ValueError("non-conformable arrays")
        self._ck(self._lib.ebpmf_ctx_set_m(self._h, _ptr(M)))

    def set_factors(self, EL, EF):
        EL = _f(np.asarray(EL, dtype=np.float64).reshape(self.n, -1))
        EF = _f(np.asarray(EF, dtype=np.float64).reshape(self.p, -1))
        if EL.shape[1] != EF.shape[1]:
            raise ValueError("non-conformable arguments")
        self._ck(self._lib.ebpmf_ctx_set_factors(self._h, EL.shape[1], _ptr(EL), _ptr(EF)))

    def set_sigma2(self, sigma2, var_type):
        if var_type not in VAR_TYPE:
            raise EbpmfError(4, "Non-supported var type")
        s2 = _vec(sigma2)
        want = {"by_col": self.p, "by_row": self.n, "constant": 1}[var_type]
        if s2.size != want:
            raise ValueError("sigma2 must have length %d for var_type %r" % (want, var_type))
        self._ck(self._lib.ebpmf_ctx_set_sigma2(self._h, VAR_TYPE[var_type], _ptr(s2)))
        self.var_type = var_type

    def vga_step(self, maxiter=10, tol=1e-5, want_v=False):
        info = SolveInfo()
        self._ck(self._lib.ebpmf_ctx_vga_step(self._h, int(maxiter), float(tol), int(bool(want_v)), C.byref(info)))
        return info

    def rewind_m(self):
        self._ck(self._lib.ebpmf_ctx_rewind_m(self._h))

    def get_m(self, out=None):
        out = np.empty((self.n, self.p), order="F") if out is None else out
        if out.shape != (self.n, self.p) or not out.flags.f_contiguous or out.dtype != np.float64:
            raise ValueError("out must be an n x p column-major float64 array")
        self._ck(self._lib.ebpmf_ctx_get_m(self._h, _ptr(out)))
        return out

    def get_v(self):
        out = np.empty((self.n, self.p), order="F")
        self._ck(self._lib.ebpmf_ctx_get_v(self._h, _ptr(out)))
        return out

    def get_v_sum(self):
        ln = {"by_col": self.p, "by_row": self.n, "constant": 1}[self.var_type]
        out = np.empty(ln)
        self._ck(self._lib.ebpmf_ctx_get_v_sum(self._h, _ptr(out)))
        return out if self.var_type != "constant" else float(out[0])

    def get_rows(self, row0, nrows, which="M"):
        out = np.empty((nrows, self.p), order="F")
        self._ck(self._lib.ebpmf_ctx_get_rows(self._h, _WHICH[which], int(row0), int(nrows), _ptr(out)))
        return out

    def get_block(self, row0, nrows, col0, ncols, which="M"):
        out = np.empty((nrows, ncols), order="F")
        self._ck(self._lib.ebpmf_ctx_get_block(self._h, _WHICH[which], int(row0), int(nrows), int(col0), int(ncols),
                                               _ptr(out)))
        return out

    def shape(self):
        n, p, K, nnz = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
        self._ck(self._lib.ebpmf_ctx_get_shape(self._h, C.byref(n), C.byref(p), C.byref(K), C.byref(nnz)))
        return n.value, p.value, K.value, nnz.value

    def get_y_csc(self):
        n, p, _, nnz = self.shape()
        cp = np.empty(p + 1, dtype=np.int32); ri = np.empty(max(nnz, 1), dtype=np.int32); x = np.empty(max(nnz, 1))
        self._ck(self._lib.ebpmf_ctx_get_y_csc(self._h, _ptr(cp), _ptr(ri), _ptr(x)))
        return cp, ri[:nnz], x[:nnz]

    def get_factors(self):
        n, p, K, _ = self.shape()
        EL = np.empty((n, K), order="F"); EF = np.empty((p, K), order="F")
        self._ck(self._lib.ebpmf_ctx_get_factors(self._h, _ptr(EL), _ptr(EF)))
        return EL, EF

    def get_sigma2(self):
        ln = {"by_col": self.p, "by_row": self.n, "constant": 1}[self.var_type]
        out = np.empty(ln)
        self._ck(self._lib.ebpmf_ctx_get_sigma2(self._h, _ptr(out)))
        return out

    def sum_lfactorial(self):
        out = C.c_double()
        self._ck(self._lib.ebpmf_ctx_sum_lfactorial(self._h, C.byref(out)))
        return out.value

    def update_sigma2(self, R2, a0=1.0, b0=1.0, cap_var_mean_ratio=0.0, n_total=None, p_total=None):
        R2 = _vec(R2)
        ln = {"by_col": self.p, "by_row": self.n, "constant": 1}[self.var_type]
        if self.var_type == "constant":
            R2 = _vec(np.sum(R2))                       # sum(fit_flash$flash_fit$R2), R/ebpmf_log.R:417
        if R2.size != ln:
            raise ValueError("R2 must have length %d" % ln)
        out = np.empty(ln)
        self._ck(self._lib.ebpmf_ctx_update_sigma2(self._h, _ptr(R2), float(a0), float(b0), float(cap_var_mean_ratio),
                                                   int(n_total or self.n), int(p_total or self.p), _ptr(out)))
        return out

    def vga_step_host(self, M, EL, EF, sigma2, var_type, maxiter=10, tol=1e-5, return_V=False, M_out=None):
        """The whole step with host buffers (what the .Call shim does each outer iteration)."""
        if var_type not in VAR_TYPE:
            raise EbpmfError(4, "Non-supported var type")
        M = _f(M)
        n, p = M.shape                                  # the library checks them against the resident Y
        EL = _f(np.asarray(EL, dtype=np.float64).reshape(n, -1))
        EF = _f(np.asarray(EF, dtype=np.float64).reshape(p, -1))
        s2 = _vec(sigma2)
        ln = {"by_col": p, "by_row": n, "constant": 1}[var_type]
        Mo = host_matrix(n, p) if M_out is None else M_out     # any host memory works; the default is the warm pool
        if Mo.shape != (n, p) or not Mo.flags.f_contiguous or Mo.dtype != np.float64:
            raise ValueError("M_out must be an n x p column-major float64 array")
        Vo = np.empty((n, p), order="F") if return_V else None
        vs = np.empty(ln)
        info = SolveInfo()
        self._ck(self._lib.ebpmf_vga_step_host(self._h, n, p, _ptr(M), EL.shape[1], _ptr(EL), _ptr(EF), VAR_TYPE[var_type],
                                               _ptr(s2), int(maxiter), float(tol), _ptr(Mo), _ptr(Vo), _ptr(vs),
                                               C.byref(info)))
        self.var_type = var_type
        return dict(M=Mo, V=Vo, v_sum=(vs if var_type != "constant" else float(vs[0])), sym=info.sym,
                    ssexp=info.ssexp, slogv=info.slogv, n_updates=info.n_updates,
                    converged=bool(info.converged), passes=info.passes, kernel_ms=info.kernel_ms)

    def vga_step_resident_host(self, EL, EF, sigma2, var_type, maxiter=10, tol=1e-5, return_V=False, M_out=None):
        """The per-iteration call of ebpmf_log's loop once M is resident (set_m once): factors and
        sigma2 in, M [, V], v_sum out.  R feeds res$M back unchanged, so nothing else is needed."""
        if var_type not in VAR_TYPE:
            raise EbpmfError(4, "Non-supported var type")
        EL = _f(np.asarray(EL, dtype=np.float64).reshape(self.n, -1))
        EF = _f(np.asarray(EF, dtype=np.float64).reshape(self.p, -1))
        s2 = _vec(sigma2)
        ln = {"by_col": self.p, "by_row": self.n, "constant": 1}[var_type]
        Mo = host_matrix(self.n, self.p) if M_out is None else M_out
        Vo = host_matrix(self.n, self.p) if return_V else None
        vs = np.empty(ln)
        info = SolveInfo()
        if Mo.shape != (self.n, self.p) or not Mo.flags.f_contiguous or Mo.dtype != np.float64:
            raise ValueError("M_out must be an n x p column-major float64 array")
        self._ck(self._lib.ebpmf_vga_step_resident_host(self._h, Mo.shape[0], Mo.shape[1], EL.shape[1], _ptr(EL), _ptr(EF),
                                                        VAR_TYPE[var_type],
                                                        _ptr(s2), int(maxiter), float(tol), _ptr(Mo), _ptr(Vo),
                                                        _ptr(vs), C.byref(info)))
        self.var_type = var_type
        return dict(M=Mo, V=Vo, v_sum=(vs if var_type != "constant" else float(vs[0])), sym=info.sym,
                    ssexp=info.ssexp, slogv=info.slogv, n_updates=info.n_updates,
                    converged=bool(info.converged), passes=info.passes, kernel_ms=info.kernel_ms)

    def sim_data_log(self, n, p, K, row0=0, seed=12345, log_offset=0.0, background_unif_range=(1.0, 2.0),
                     pi0=0.8, noise=0.2, warm_start=False):
        nnz = C.c_int64()
        self._ck(self._lib.ebpmf_ctx_sim_data_log(self._h, n, p, K, row0, seed, float(log_offset),
                                                  float(background_unif_range[0]), float(background_unif_range[1]),
                                                  float(pi0), float(noise), int(bool(warm_start)), C.byref(nnz)))
        self.n, self.p, self.var_type = n, p, "by_col"
        return nnz.value

    def set_option(self, name, value):
        """'delta' (0 = exact update count, the default), 'freeze', 'history', 'speculate', 'pipe_chunks', ...: see
        include/ebpmf_b200.h."""
        self._ck(self._lib.ebpmf_ctx_set_option(self._h, name.encode(), float(value)))

    def counters(self):
        out = np.zeros(9)
        self._ck(self._lib.ebpmf_ctx_counters(self._h, _ptr(out), 9))
        keys = ("topup_units", "unit_evals", "topup_tiles", "launches", "spec_bytes", "recopy_bytes", "tiles", "have_history",
                "last_kernel_ms")
        return dict(zip(keys, (float(v) for v in out)))

    def m_products(self, EL_w, EF_w, want_r2_col=True, want_r2_row=True):
        """What flashier needs from the resident M, computed on the device (N1): M %*% EF_w, t(M) %*% EL_w and the
        column / row sums of (M - EL_w %*% t(EF_w))^2 (R/ebpmf_log_flash_update.R:35,55-64)."""
        EL_w = _f(np.asarray(EL_w, dtype=np.float64).reshape(self.n, -1))
        EF_w = _f(np.asarray(EF_w, dtype=np.float64).reshape(self.p, -1))
        if EL_w.shape[1] != EF_w.shape[1]:
            raise ValueError("non-conformable arguments")
        Kw = EL_w.shape[1]
        MF = np.empty((self.n, Kw), order="F"); MtL = np.empty((self.p, Kw), order="F")
        rc = np.empty(self.p) if want_r2_col else None
        rr = np.empty(self.n) if want_r2_row else None
        ms = (C.c_double * 3)()
        self._ck(self._lib.ebpmf_ctx_m_products(self._h, Kw, _ptr(EL_w), _ptr(EF_w), _ptr(MF), _ptr(MtL), _ptr(rc), _ptr(rr), ms))
        return dict(MF=MF, MtL=MtL, r2_col=rc, r2_row=rr, kernel_ms=ms[0], products_ms=ms[1], partial_sums_ms=ms[2])

    def last_diag(self):
        a, b = C.c_uint64(), C.c_uint64()
        self._ck(self._lib.ebpmf_ctx_last_diag(self._h, C.byref(a), C.byref(b)))
        return dict(topup_units=a.value, unit_evals=b.value)

    def probe_fp64_tflops(self):
        out = C.c_double()
        self._ck(self._lib.ebpmf_probe_fp64_tflops(self._h, C.byref(out)))
        return out.value


class VgaGroup:
    """One process driving several GPUs (what a single R session does): owns an ``ebpmf_group`` --
    one context per GPU, one host thread per GPU per call, rows dealt in 256-row blocks, the
    library's NCCL combine of the update count and of v_sum / ELBO partials."""

    def __init__(self, devices):
        self._lib = _lib.load()
        devs = (C.c_int * len(devices))(*devices)
        h = C.c_void_p()
        _lib.check(self._lib.ebpmf_group_create(len(devices), devs, C.byref(h)))
        self._h = h
        self.n = self.p = None

    def _ck(self, rc):
        if rc != 0:
            msg = self._lib.ebpmf_group_last_error(self._h)
            raise EbpmfError(rc, msg.decode() if msg else "")

    def close(self):
        if getattr(self, "_h", None):
            self._lib.ebpmf_group_destroy(self._h)
            self._h = None

    def size(self):
        return int(self._lib.ebpmf_group_size(self._h))

    def set_option(self, name, value):
        self._ck(self._lib.ebpmf_group_set_option(self._h, name.encode(), float(value)))

    def set_y(self, Y):
        n, p = Y.shape
        cp, ri, x = _csc_slots(Y if _is_sparse(Y) else _sp.csc_matrix(np.asarray(Y)))
        self._ck(self._lib.ebpmf_group_set_y_csc(self._h, n, p, _ptr(cp), _ptr(ri), _ptr(x)))
        self.n, self.p = n, p

    def sum_lfactorial(self):
        out = C.c_double()
        self._ck(self._lib.ebpmf_group_sum_lfactorial(self._h, C.byref(out)))
        return out.value

    def set_m(self, M):
        M = _f(M)
        if M.shape != (self.n, self.p):
            raise ValueError("non-conformable arrays")
        self._ck(self._lib.ebpmf_group_set_m(self._h, self.n, self.p, _ptr(M)))

    def vga_step_host(self, M, EL, EF, sigma2, var_type, maxiter=10, tol=1e-5, return_V=False, M_out=None):
        """``M`` None: the shards' resident M (``set_m`` once, then every step's output) is the input."""
        if var_type not in VAR_TYPE:
            raise EbpmfError(4, "Non-supported var type")
        M = None if M is None else _f(M)
        if M is not None and M.shape != (self.n, self.p):
            raise ValueError("non-conformable arrays")
        EL = _f(np.asarray(EL, dtype=np.float64).reshape(self.n, -1))
        EF = _f(np.asarray(EF, dtype=np.float64).reshape(self.p, -1))
        s2 = _vec(sigma2)
        ln = {"by_col": self.p, "by_row": self.n, "constant": 1}[var_type]
        if s2.size != ln:
            raise ValueError("sigma2 has the wrong length for var_type %r" % var_type)
        Mo = np.empty((self.n, self.p), order="F") if M_out is None else M_out
        Vo = np.empty((self.n, self.p), order="F") if return_V else None
        vs = np.empty(ln)
        info = SolveInfo()
        self._ck(self._lib.ebpmf_group_vga_step_host(self._h, self.n, self.p, _ptr(M), EL.shape[1], _ptr(EL), _ptr(EF),
                                                     VAR_TYPE[var_type], _ptr(s2), int(maxiter), float(tol), _ptr(Mo),
                                                     _ptr(Vo), _ptr(vs), C.byref(info)))
        return dict(M=Mo, V=Vo, v_sum=(vs if var_type != "constant" else float(vs[0])), sym=info.sym,
                    ssexp=info.ssexp, slogv=info.slogv, n_updates=info.n_updates,
                    converged=bool(info.converged), passes=info.passes, kernel_ms=info.kernel_ms)


def r_batches(n, batch_size):
    """Row ranges of ebpmf_log's batches: ``n_batch = ceiling(n/batch_size)``; ``split(1:n, cut(seq_along(1:n),
    n_batch, labels=FALSE))`` (R/ebpmf_log.R:195-197).  base::cut with a break count uses n_batch+1 equally
    spaced breaks on [1, n], the outer two widened by (n-1)/1000, intervals right-closed; the batches are
    therefore contiguous.  Returns the int64 array of n_batch+1 batch starts (0-based, last = n); empty bins
    are dropped, as split() drops them."""
    n = int(n)
    n_batch = int(np.ceil(n / float(batch_size)))
    if n_batch <= 1:
        return np.array([0, n], dtype=np.int64)
    dx = float(n - 1) if n > 1 else 1.0
    breaks = np.linspace(1.0, float(n), n_batch + 1)
    breaks[0] = 1.0 - dx / 1000
    breaks[-1] = float(n) + dx / 1000
    code = np.searchsorted(breaks, np.arange(1, n + 1, dtype=np.float64), side="left")   # x in (b_k, b_{k+1}]
    starts = [0] + [int(i) for i in np.nonzero(np.diff(code))[0] + 1] + [n]
    return np.asarray(starts, dtype=np.int64)


class VgaBatched:
    """ebpmf_log's batched branch (``general_control$batch_size < n``, R/ebpmf_log.R:195-203, :254-300): one
    solve per contiguous row batch with the stop test scoped to the batch, batches pipelined through
    ``nslots`` worker contexts so host<->device copies overlap the solves; M lives on the host."""

    def __init__(self, device=0, nslots=3):
        """``device``: one device index, or a list of them (batches are dealt round-robin over
        ``len(devices) * nslots`` worker contexts; no collective is needed, batches are independent)."""
        self._lib = _lib.load()
        h = C.c_void_p()
        if isinstance(device, (list, tuple)):
            devs = (C.c_int * len(device))(*device)
            _lib.check(self._lib.ebpmf_batched_create_multi(len(device), devs, int(nslots), C.byref(h)))
        else:
            _lib.check(self._lib.ebpmf_batched_create(int(device), int(nslots), C.byref(h)))
        self._h = h
        self.n = self.p = self.batch_row0 = None

    def _ck(self, rc):
        if rc != 0:
            msg = self._lib.ebpmf_batched_last_error(self._h)
            raise EbpmfError(rc, msg.decode() if msg else "")

    def close(self):
        if getattr(self, "_h", None):
            self._lib.ebpmf_batched_destroy(self._h)
            self._h = None

    def set_option(self, name, value):
        self._ck(self._lib.ebpmf_batched_set_option(self._h, name.encode(), float(value)))

    def set_y(self, Y, batch_size=None, batch_row0=None):
        """``batch_size`` as in general_control (rows per batch, cut as R cuts them) or explicit ``batch_row0``."""
        n, p = Y.shape
        row0 = r_batches(n, batch_size) if batch_row0 is None else np.ascontiguousarray(batch_row0, dtype=np.int64)
        cp, ri, x = _csc_slots(Y if _is_sparse(Y) else _sp.csc_matrix(np.asarray(Y)))
        self._ck(self._lib.ebpmf_batched_set_y_csc(self._h, n, p, _ptr(cp), _ptr(ri), _ptr(x), len(row0) - 1, _ptr(row0)))
        self.n, self.p, self.batch_row0 = n, p, row0

    def sum_lfactorial(self):
        out = C.c_double()
        self._ck(self._lib.ebpmf_batched_sum_lfactorial(self._h, C.byref(out)))
        return out.value

    def vga_step_host(self, M, EL, EF, sigma2, var_type, maxiter=10, tol=1e-5, return_V=False, M_out=None):
        if var_type not in VAR_TYPE:
            raise EbpmfError(4, "Non-supported var type")
        M = _f(M)
        EL = _f(np.asarray(EL, dtype=np.float64).reshape(self.n, -1))
        EF = _f(np.asarray(EF, dtype=np.float64).reshape(self.p, -1))
        s2 = _vec(sigma2)
        ln = {"by_col": self.p, "by_row": self.n, "constant": 1}[var_type]
        if s2.size != ln:
            raise ValueError("sigma2 has the wrong length for var_type %r" % var_type)
        Mo = np.empty((self.n, self.p), order="F") if M_out is None else M_out
        Vo = np.empty((self.n, self.p), order="F") if return_V else None
        vs = np.empty(ln)
        nb = len(self.batch_row0) - 1
        nu = np.zeros(nb, dtype=np.int32)
        cv = np.zeros(nb, dtype=np.int32)
        info = SolveInfo()
        self._ck(self._lib.ebpmf_batched_vga_step_host(self._h, M.shape[0], M.shape[1], _ptr(M), EL.shape[1], _ptr(EL), _ptr(EF),
                                                       VAR_TYPE[var_type], _ptr(s2), int(maxiter), float(tol),
                                                       _ptr(Mo), _ptr(Vo), _ptr(vs), C.byref(info), _ptr(nu), _ptr(cv)))
        return dict(M=Mo, V=Vo, v_sum=(vs if var_type != "constant" else float(vs[0])), sym=info.sym,
                    ssexp=info.ssexp, slogv=info.slogv, n_updates=[int(v) for v in nu],
                    converged=[bool(v) for v in cv], kernel_ms=info.kernel_ms)


# ---------------------------------------------------------------------------
# drop-in solver signatures
# ---------------------------------------------------------------------------
def vga_pois_solver_mat_newton(M, X, S, Beta, Sigma2, maxiter=1000, tol=1e-8, return_V=True, info=None, ctx=None):
    """R/vga_solver_mat.R:6-41.  Returns ``{"M":…, "V":…}`` or ``M`` (return_V=False).

    X may be dense or scipy sparse (test_memory.R passes sparse counts).  A NaN at
    the stop test raises EbpmfError(EBPMF_ERR_NAN), R's "missing value where
    TRUE/FALSE needed".
    """
    ctx = ctx or default_context()
    M = _f(M)
    n, p = M.shape
    Sb, Sk = _operand(S, n, p, "S")
    Bb, Bk = _operand(Beta, n, p, "Beta")
    Vb, Vk = _operand(Sigma2, n, p, "Sigma2")
    Mo = np.empty((n, p), order="F")
    Vo = np.empty((n, p), order="F") if return_V else None
    si = SolveInfo()
    if _is_sparse(X):
        if X.shape != (n, p):
            raise ValueError("non-conformable arrays")
        cp, ri, x = _csc_slots(X)
        Xd = None
    else:
        Xd = _f(X)
        if Xd.shape != (n, p):
            raise ValueError("non-conformable arrays")
        cp = ri = x = None
    ctx._ck(ctx._lib.ebpmf_vga_newton(ctx._h, n, p, _ptr(M), _ptr(Xd), _ptr(cp), _ptr(ri), _ptr(x),
                                      _ptr(Sb), Sk, _ptr(Bb), Bk, _ptr(Vb), Vk, int(maxiter), float(tol),
                                      _ptr(Mo), _ptr(Vo), C.byref(si)))
    if info is not None:
        info.update(n_updates=si.n_updates, converged=bool(si.converged), passes=si.passes, kernel_ms=si.kernel_ms)
    if return_V:
        return {"M": Mo, "V": Vo}
    return Mo


def vga_pois_solver_mat_newton_fixed_iter(M, V, Y, S, Beta, Sigma2, maxiter=1000, ctx=None):
    """R/vga_solver_mat.R:44-65.  ``V is None`` <=> is.null(V)."""
    ctx = ctx or default_context()
    M = _f(M)
    n, p = M.shape
    Yd = _f(Y.toarray() if _is_sparse(Y) else Y)
    Sb, Sk = _operand(S, n, p, "S")
    Bb, Bk = _operand(Beta, n, p, "Beta")
    Vb, Vk = _operand(Sigma2, n, p, "Sigma2")
    V0 = None if V is None else _f(np.broadcast_to(np.asarray(V, dtype=np.float64), (n, p)))
    Mo = np.empty((n, p), order="F"); Vo = np.empty((n, p), order="F")
    ctx._ck(ctx._lib.ebpmf_vga_fixed_iter(ctx._h, n, p, _ptr(M), _ptr(V0), _ptr(Yd), _ptr(Sb), Sk, _ptr(Bb), Bk,
                                          _ptr(Vb), Vk, int(maxiter), _ptr(Mo), _ptr(Vo)))
    return {"M": Mo, "V": Vo}


def vga_pois_solver_mat_newton_low_memory(M, Y, l0, f0, EL, EF, sigma2, var_type, maxiter=1000, tol=1e-8,
                                          info=None, ctx=None):
    """R/vga_solver_mat.R:71-109.  Returns M only (:108).  Uses (and replaces) the
    context's resident state."""
    ctx = ctx or default_context()
    if var_type not in VAR_TYPE:
        raise EbpmfError(4, "Non-supported var type")
    M = _f(M)
    n, p = M.shape
    EL = _f(np.asarray(EL, dtype=np.float64).reshape(n, -1)); EF = _f(np.asarray(EF, dtype=np.float64).reshape(p, -1))
    l0 = _vec(np.broadcast_to(np.asarray(l0, dtype=np.float64), (n,)))
    f0 = _vec(np.broadcast_to(np.asarray(f0, dtype=np.float64), (p,)))
    s2 = _vec(sigma2)
    if s2.size != {"by_col": p, "by_row": n, "constant": 1}[var_type]:
        raise ValueError("sigma2 has the wrong length for var_type %r" % var_type)
    if _is_sparse(Y):
        cp, ri, x = _csc_slots(Y); Yd = None
    else:
        Yd = _f(Y); cp = ri = x = None
    Mo = np.empty((n, p), order="F")
    si = SolveInfo()
    ctx._ck(ctx._lib.ebpmf_vga_low_memory(ctx._h, n, p, EL.shape[1], _ptr(M), _ptr(Yd), _ptr(cp), _ptr(ri), _ptr(x),
                                          _ptr(l0), _ptr(f0), _ptr(EL), _ptr(EF), _ptr(s2), VAR_TYPE[var_type],
                                          int(maxiter), float(tol), _ptr(Mo), C.byref(si)))
    ctx.n, ctx.p, ctx.var_type = n, p, var_type
    if info is not None:
        info.update(n_updates=si.n_updates, converged=bool(si.converged), passes=si.passes, kernel_ms=si.kernel_ms)
    return Mo


def ebpmf_log_update_sigma2(fit_flash, sigma2, v_sum, var_type, cap_var_mean_ratio, a0, b0, n, p, ctx=None):
    """R/ebpmf_log.R:413-447.  ``fit_flash``: mapping with ``R2`` and (for the cap gate) ``EL``/``EF``."""
    ctx = ctx or default_context()
    if var_type not in VAR_TYPE:
        raise EbpmfError(4, "Non-supported var type")
    s2 = _vec(sigma2); vs = _vec(v_sum); R2 = _vec(fit_flash["R2"])
    if var_type == "constant":
        R2 = _vec(np.sum(R2))
    EL = EF = None; K = 0
    if cap_var_mean_ratio > 0:
        EL = _f(np.asarray(fit_flash["EL"], dtype=np.float64).reshape(n, -1))
        EF = _f(np.asarray(fit_flash["EF"], dtype=np.float64).reshape(p, -1)); K = EL.shape[1]
    out = np.empty_like(s2)
    ctx._ck(ctx._lib.ebpmf_update_sigma2(ctx._h, VAR_TYPE[var_type], n, p, _ptr(s2), _ptr(vs), _ptr(R2), float(a0),
                                         float(b0), float(cap_var_mean_ratio), K, _ptr(EL), _ptr(EF), _ptr(out)))
    return out if var_type != "constant" else float(out[0])


def calc_ebpmf_log_obj(n, p, sym, ssexp, slogv, sv, sigma2, R2, KL_LF, const, ss, a0, b0, ctx=None):
    """R/ebpmf_log.R:406-410."""
    ctx = ctx or default_context()
    sv = _vec(sv); s2 = _vec(sigma2); R2 = _vec(R2)     # each keeps its own length: R sums every term at its own length
    out = C.c_double()
    ctx._ck(ctx._lib.ebpmf_calc_obj(ctx._h, n, p, float(sym), float(ssexp), float(slogv), sv.size, _ptr(sv), s2.size,
                                    _ptr(s2), R2.size, _ptr(R2), float(KL_LF), float(const), float(ss), float(a0),
                                    float(b0), C.byref(out)))
    return out.value


def calc_split_PMF_obj_flashier(Y, S, sigma2, M, V, fit_flash, KL_LF, const, var_type, ctx=None):
    """inst/code/Poisson_MF_splitting.R:373-391 (the archived driver's un-fused ELBO).  ``fit_flash``: mapping with ``R2``."""
    ctx = ctx or default_context()
    if var_type not in VAR_TYPE:
        raise EbpmfError(4, "Non-supported var type")
    M = _f(M)
    n, p = M.shape
    Yd = _f(Y.toarray() if _is_sparse(Y) else Y); Vd = _f(V)
    if Yd.shape != (n, p) or Vd.shape != (n, p):
        raise ValueError("non-conformable arrays")
    Sb, Sk = _operand(S, n, p, "S")
    s2 = _vec(sigma2); R2 = _vec(fit_flash["R2"])
    out = C.c_double()
    ctx._ck(ctx._lib.ebpmf_calc_split_obj(ctx._h, n, p, _ptr(Yd), _ptr(Sb), Sk, _ptr(M), _ptr(Vd), VAR_TYPE[var_type], s2.size,
                                          _ptr(s2), R2.size, _ptr(R2), float(KL_LF), float(const), C.byref(out)))
    return out.value


def sum_lfactorial_sparseMat(Y, ctx=None):
    """R/ebpmf_log.R:470-472: sum(lfactorial(Y@x)); a dense Y gives sum(lfactorial(Y)) (:175)."""
    ctx = ctx or default_context()
    x = _vec(Y.tocsc().data) if _is_sparse(Y) else _vec(Y)
    out = C.c_double()
    ctx._ck(ctx._lib.ebpmf_sum_lfactorial(ctx._h, x.size, _ptr(x), C.byref(out)))
    return out.value


def adjust_var_shape(sigma2, var_type, n, p):
    """R/ebpmf_log.R:475-486.  Shape helper only (the kernels never materialise this)."""
    sigma2 = np.asarray(sigma2, dtype=np.float64)
    if var_type == "constant":
        return np.full((n, p), float(sigma2.reshape(-1)[0]), order="F")
    if var_type == "by_row":
        return np.asfortranarray(np.broadcast_to(sigma2.reshape(n, 1), (n, p)))
    if var_type == "by_col":
        return np.asfortranarray(np.broadcast_to(sigma2.reshape(1, p), (n, p)))
    raise EbpmfError(4, "Non-supported var type")
