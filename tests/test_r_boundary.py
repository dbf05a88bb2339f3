"""The R side of the boundary, RUN without R (tests/r_mock.py): the R wrappers of ebpmf.alpha_b200/r/R/vga_cuda.R are
evaluated by oracle/mini_r.py, their `.Call`s go through the compiled, unmodified r/src/r_glue.c (linked against a
functional mock of the R C API, tests/r_api_stub/mock_r.c) into libebpmf_b200.so.  On a GPU box that is the whole
product path of an R user -- wrapper -> glue -> C ABI -> CUDA -- and its results are held to the outputs of the
reference's own R source (tests/golden/minir_reference.npz) to 1e-8, update counts exact.  Without a GPU the same
chain is exercised up to the point where the library refuses to work without one ("no CPU fallback" as an R error).

Checked on the way, per `.Call`: the routine exists in the glue's own registration table with the arity used
(R's own check), no argument is written to, PROTECT/UNPROTECT balance, Rf_error unwinds cleanly."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

import r_mock
from oracle import mini_r as R, standin_flash, vga_numpy as vn

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "golden", "vga_golden.npz"))
MR = np.load(os.path.join(HERE, "golden", "minir_reference.npz"))
CASES = {"def": (10, 1e-5), "exh": (2, 1e-5), "tight": (1000, 1e-8)}
VTS = ["by_col", "by_row", "constant"]
RTOL = 1e-8
Y = np.asarray(G["Y"], dtype=np.float64)
N, P = Y.shape


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64).ravel(), np.asarray(b, dtype=np.float64).ravel()
    assert a.shape == b.shape
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


def relM(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b)) / np.maximum(np.abs(b), 1.0)))


def state(vt):
    return tuple(np.asarray(G["%s_%s" % (vt, k)], dtype=np.float64) for k in ("M0", "EL", "EF", "sigma2", "R2"))


def fit_flash(R2, EL, EF):
    return dict(flash_fit=dict(R2=R2, EF=[np.asfortranarray(EL), np.asfortranarray(EF)]))


# ---------------------------------------------------------------------------------------------------------------
# no GPU needed
# ---------------------------------------------------------------------------------------------------------------
def test_glue_compiles_against_the_mock_and_registers_its_routines():
    import test_r_glue
    assert r_mock.routines() == test_r_glue.registered() and len(r_mock.routines()) == 21


def test_wrappers_are_defined_with_the_reference_signatures():
    s = r_mock.RSession()
    sig = lambda name: [p for p, _ in s.it.get(name).params]                                    # noqa: E731
    assert sig("vga_pois_solver_mat_newton") == ["M", "X", "S", "Beta", "Sigma2", "maxiter", "tol", "return_V"]            # R/vga_solver_mat.R:6
    assert sig("vga_pois_solver_mat_newton_fixed_iter") == ["M", "V", "Y", "S", "Beta", "Sigma2", "maxiter"]             # :44
    assert sig("vga_pois_solver_mat_newton_low_memory") == ["M", "Y", "l0", "f0", "EL", "EF", "sigma2", "var_type", "maxiter", "tol"]   # :71
    assert sig("ebpmf_log_update_sigma2") == ["fit_flash", "sigma2", "v_sum", "var_type", "cap_var_mean_ratio", "a0", "b0", "n", "p"]   # R/ebpmf_log.R:413
    assert sig("calc_ebpmf_log_obj") == ["n", "p", "sym", "ssexp", "slogv", "sv", "sigma2", "R2", "KL_LF", "const", "ss", "a0", "b0"]   # :406
    assert sig("sum_lfactorial_sparseMat") == ["Y"]                                                                      # :470
    assert sig("calc_split_PMF_obj_flashier") == ["Y", "S", "sigma2", "M", "V", "fit_flash", "KL_LF", "const", "var_type"]
    d = {p: dflt for p, dflt in s.it.get("vga_pois_solver_mat_newton").params}
    assert d["maxiter"] == ("num", 1000.0) and d["tol"] == ("num", 1e-8) and d["return_V"] == ("name", "TRUE")


@pytest.mark.skipif(os.path.exists("/root/reference/R/vga_solver_mat.R") is False, reason="the reference checkout is not present")
def test_wrapper_signatures_equal_the_reference_files():
    """Parameter names AND default expressions of every replaced function, parsed from both sources."""
    from oracle import minir_reference as mr
    ref = mr.load_reference("/root/reference")
    s = r_mock.RSession()
    for name in ("vga_pois_solver_mat_newton", "vga_pois_solver_mat_newton_fixed_iter", "vga_pois_solver_mat_newton_low_memory",
                 "ebpmf_log_update_sigma2", "calc_ebpmf_log_obj", "sum_lfactorial_sparseMat"):
        assert s.it.get(name).params == ref.get(name).params, name
    arch = R.Interp()
    arch.source("/root/reference/inst/code/Poisson_MF_splitting.R")
    assert s.it.get("calc_split_PMF_obj_flashier").params == arch.get("calc_split_PMF_obj_flashier").params


def test_dot_call_checks_symbol_and_arity_like_r():
    s = r_mock.RSession()
    with pytest.raises(r_mock.RCallError, match=r"Incorrect number of arguments \(1\), expecting 2 for 'C_ebpmf_ctx_set_m'"):
        s.dot_call("C_ebpmf_ctx_set_m", np.zeros(3))
    with pytest.raises(r_mock.RCallError, match="not in the registration table"):
        s.dot_call("C_no_such_routine")


def test_glue_errors_unwind_as_r_errors():
    s = r_mock.RSession()
    with pytest.raises(r_mock.RCallError, match="invalid or closed GPU context"):
        s.dot_call("C_ebpmf_ctx_set_m", np.zeros(3), np.zeros((2, 2)))        # not an external pointer
    with pytest.raises(r_mock.RCallError, match="invalid or closed multi-GPU group"):
        s.dot_call("C_ebpmf_group_sum_lfactorial", None)
    live = r_mock.lib().mockr_live_objects()
    for _ in range(3):                                                          # the error path leaks nothing of ours
        with pytest.raises(r_mock.RCallError):
            s.dot_call("C_ebpmf_ctx_set_m", np.zeros(3), np.zeros((2, 2)))
    assert r_mock.lib().mockr_live_objects() == live


def test_without_a_gpu_the_r_call_fails_loudly():
    """ebpmf_b200_ctx() -> .Call(C_ebpmf_ctx_new) -> ebpmf_ctx_create: an R error, never a CPU fallback."""
    s = r_mock.RSession()
    try:
        ctx = s.call("ebpmf_b200_ctx")
    except r_mock.RCallError as e:
        assert "no CPU fallback" in str(e) and str(e).startswith("ebpmf_b200: ")
    else:                                                                       # a GPU box: the context is real
        assert isinstance(ctx, r_mock.RExt)
        s.close()


def test_values_cross_the_mock_api_unchanged():
    m = np.asfortranarray(np.arange(6, dtype=np.float64).reshape(2, 3))
    for v in (m, np.array([1.5, -2.0]), np.array([True, False, True]), R.strs("by_col"), r_mock.as_int([3, 4]), None):
        back = r_mock.from_sexp(r_mock.to_sexp(v))
        if v is None:
            assert back is None
        else:
            assert type(back) is type(v) and back.shape == v.shape and np.array_equal(back, v), v
    lst = r_mock.from_sexp(r_mock.to_sexp(R.RList([m, None, R.strs("x")], ["M", "V", "s"])))
    assert lst.names == ["M", "V", "s"] and np.array_equal(lst.get("M"), m) and lst.get("V") is None
    L = r_mock.lib()
    d = r_mock.to_sexp(R.RSparse(np.array([[0.0, 2.0], [3.0, 0.0], [0.0, 4.0]])))     # dgCMatrix slots, 0-based, column-major
    slot = lambda nm: r_mock.from_sexp(L.mockr_get_attr(d, nm))                       # noqa: E731
    assert slot(b"p").tolist() == [0, 1, 3] and slot(b"i").tolist() == [1, 0, 2] and slot(b"x").tolist() == [3.0, 2.0, 4.0]
    assert slot(b"Dim").tolist() == [3, 2] and isinstance(slot(b"i"), r_mock.RInt) and list(slot(b"class")) == ["dgCMatrix"]


# ---------------------------------------------------------------------------------------------------------------
# GPU: R wrapper -> glue -> C ABI -> CUDA, against the reference's own source
# ---------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def rs():
    s = r_mock.RSession()
    yield s
    s.close()


def check_step_list(res, k):
    sc = MR[k + "scal"]
    assert int(res.get("n_updates")[0]) == int(sc[3]) and isinstance(res.get("n_updates"), r_mock.RInt)
    assert relM(res.get("M"), MR[k + "M"]) < RTOL and rel(res.get("V"), MR[k + "V"]) < RTOL
    assert rel(res.get("v_sum"), MR[k + "v_sum"]) < RTOL
    for i, name in enumerate(("sym", "ssexp", "slogv")):
        assert abs(res.get(name)[0] - sc[i]) <= RTOL * max(1.0, abs(sc[i])), name
    assert res.get("converged").dtype == bool


@pytest.mark.gpu
@pytest.mark.parametrize("vt", VTS)
@pytest.mark.parametrize("tag", sorted(CASES))
def test_r_solver_and_step_wrappers(rs, vt, tag):
    M0, EL, EF, s2, R2 = state(vt)
    mi, tol = CASES[tag]
    k = "fn/%s_%s_" % (vt, tag)
    S2 = vn.adjust_var_shape(s2, vt, N, P)
    for X in (Y, R.RSparse(Y)):                                   # dense counts and a dgCMatrix (tests/testthat/test_memory.R passes sparse)
        res = rs.call("vga_pois_solver_mat_newton", M0, X, 1, EL @ EF.T, S2, maxiter=mi, tol=tol, return_V=True)
        assert res.names == ["M", "V"]                            # list(M=M,V=Sigma2/temp), R/vga_solver_mat.R:36
        assert relM(res.get("M"), MR[k + "M"]) < RTOL and rel(res.get("V"), MR[k + "V"]) < RTOL
    m_only = rs.call("vga_pois_solver_mat_newton", M0, Y, 1, EL @ EF.T, S2, mi, tol, False)      # return(M), :38
    assert isinstance(m_only, np.ndarray) and m_only.shape == (N, P) and relM(m_only, MR[k + "M"]) < RTOL
    rs.call("ebpmf_b200_set_y", R.RSparse(Y))
    check_step_list(rs.call("ebpmf_b200_vga_step", M0, EL, EF, s2, vt, mi, tol, True), k)
    rs.call("ebpmf_b200_set_m", M0)
    res = rs.call("ebpmf_b200_vga_step_resident", N, P, EL, EF, s2, vt, mi, tol, want_V=True)
    check_step_list(res, k)
    if tag == "def":
        for cap, key in ((0.0, "sigma2_new"), (2.0, "sigma2_new_cap"), (1e9, "sigma2_new_cap_big")):
            s = rs.call("ebpmf_log_update_sigma2", fit_flash(R2, EL, EF), s2, res.get("v_sum"), vt, cap, 1, 1, N, P)
            assert rel(s, MR[k + key]) < RTOL
        s = rs.call("ebpmf_log_update_sigma2", fit_flash(R2, EL, EF), s2, res.get("v_sum"), vt, 0, 1, 1, N, P)
        ss = {"by_col": N, "by_row": P, "constant": N * P}[vt]
        obj = rs.call("calc_ebpmf_log_obj", N, P, res.get("sym"), res.get("ssexp"), res.get("slogv"), res.get("v_sum"), s, R2, -123.4, -5.5, ss, 1, 1)
        assert abs(obj[0] - MR[k + "obj"][0]) <= RTOL * abs(MR[k + "obj"][0])
    assert all(depth <= 8 for _, depth in rs.calls) and len(rs.calls) >= 6


@pytest.mark.gpu
def test_r_other_wrappers(rs):
    M0, EL, EF, s2, _ = state("by_col")
    S2 = vn.adjust_var_shape(s2, "by_col", N, P)
    for mi in (1, 10):
        fi = rs.call("vga_pois_solver_mat_newton_fixed_iter", M0, None, Y, 1, EL @ EF.T, S2, maxiter=mi)
        assert relM(fi.get("M"), MR["fn/fixed%d_M" % mi]) < RTOL and rel(fi.get("V"), MR["fn/fixed%d_V" % mi]) < RTOL
    for vt in VTS:
        _, EL, EF, s2, _ = state(vt)
        key = "fn/lowmem_M" if vt == "by_col" else "fn/lowmem_%s_M" % vt
        for Yin in (Y, R.RSparse(Y)):
            lm = rs.call("vga_pois_solver_mat_newton_low_memory", np.log(Y + 0.5), Yin, G["L0"], G["F0"], EL[:, 2:], EF[:, 2:], s2, vt,
                         maxiter=1000, tol=1e-8)
            assert relM(lm, MR[key]) < RTOL
    lf = rs.call("sum_lfactorial_sparseMat", R.RSparse(Y))
    assert abs(lf[0] - MR["fn/lfact"][0]) < RTOL * MR["fn/lfact"][0]
    with pytest.raises(r_mock.RCallError, match="Non-supported var type"):                                  # R/ebpmf_log.R:444
        rs.call("ebpmf_log_update_sigma2", fit_flash(np.ones(P), EL, EF), np.ones(P), np.ones(P), "by_banana", 0, 1, 1, N, P)
    bad = M0.copy()
    bad[3, 4] = np.nan
    with pytest.raises(r_mock.RCallError, match="missing value where TRUE/FALSE needed"):                   # R/vga_solver_mat.R:26
        rs.call("vga_pois_solver_mat_newton", bad, Y, 1, EL @ EF.T, 0.5, maxiter=5, tol=1e-8)
    with pytest.raises(r_mock.RCallError, match="non-conformable"):
        rs.call("vga_pois_solver_mat_newton", M0, Y, np.ones(7), EL @ EF.T, 0.5)


@pytest.mark.gpu
@pytest.mark.parametrize("vt", VTS)
def test_r_batched_wrappers(rs, vt):
    """batches cut by R's own expression (R/ebpmf_log.R:197), evaluated by the interpreter."""
    M0, EL, EF, s2, R2 = state(vt)
    k = "blk/%s/25/" % vt
    batches = rs.it.run("split(1:%d, cut(seq_along(1:%d), %d, labels = FALSE))" % (N, N, int(np.ceil(N / 25))))
    h = rs.call("ebpmf_b200_batched_new", R.RSparse(Y), batches, nslots=2)
    try:
        assert int(h.attrs["n_batch"][0]) == 4
        res = rs.call("ebpmf_b200_batched_vga_step", h, M0, EL, EF, s2, vt, 10, 1e-5, True)
    finally:
        r_mock.lib().mockr_free(h.sexp)                          # the finalizer R's collector would run
    assert relM(res.get("M"), MR[k + "cap0/M"]) < RTOL and rel(res.get("v_sum"), MR[k + "cap0/v_sum"]) < RTOL
    for name in ("sym", "ssexp", "slogv"):
        assert abs(res.get(name)[0] - MR[k + "cap0/" + name]) <= RTOL * max(1.0, abs(MR[k + "cap0/" + name]))
    assert res.attrs["n_updates_per_batch"].shape == (4,) and int(res.get("n_updates")[0]) == int(res.attrs["n_updates_per_batch"].max())
    for cap in (0.0, 2.0):
        s = rs.call("ebpmf_log_update_sigma2", fit_flash(R2, EL, EF), s2, res.get("v_sum"), vt, cap, 1, 1, N, P)
        assert rel(s, MR[k + "cap%g/sigma2" % cap]) < RTOL


@pytest.mark.gpu
@pytest.mark.parametrize("vt", VTS)
def test_r_session_through_the_whole_driver(rs, vt):
    """ebpmf_log's outer loop with every hot-path call made the way the R package makes it (resident M), against
    ebpmf_log() itself run from the reference's source by the evaluator."""
    l0 = np.log(Y.mean(axis=1) + 0.1)
    f0 = np.log((Y.sum(axis=0) + 0.1) / np.exp(l0).sum())
    M_init = np.asfortranarray(np.log(Y + 0.5))
    s2 = {"by_col": np.full(P, 0.5), "by_row": np.full(N, 0.5), "constant": np.array([0.5])}[vt]
    rs.call("ebpmf_b200_set_y", Y)                                # as.matrix(Y): the dense branch of C_ebpmf_ctx_set_y
    const = -rs.dot_call("C_ebpmf_sum_lfactorial", rs.call("ebpmf_b200_ctx"), None)[0]
    rs.call("ebpmf_b200_set_m", M_init)

    def step(M, EL, EF, sig):
        r = rs.call("ebpmf_b200_vga_step_resident", N, P, EL, EF, sig, vt, 10, 1e-5, want_V=True)
        return dict(M=r.get("M"), V=r.get("V"), v_sum=r.get("v_sum"), sym=r.get("sym")[0], ssexp=r.get("ssexp")[0], slogv=r.get("slogv")[0])

    out = standin_flash.outer_loop(
        Y, l0, f0, M_init, s2, vt, 3, 4, step=step,
        update_sigma2=lambda fit, sig, vs: rs.call("ebpmf_log_update_sigma2", fit_flash(fit["R2"], fit["EL"], fit["EF"]), sig, vs, vt, 0, 1, 1, N, P),
        calc_obj=lambda *a: rs.call("calc_ebpmf_log_obj", *a)[0], const=const)
    k = "drv/%s/inf/" % vt
    assert rel(out["elbo_trace"][1:], MR[k + "elbo_trace"][1:]) < RTOL and rel(out["sigma2"], MR[k + "sigma2"]) < RTOL
    assert relM(out["M"], MR[k + "M"]) < RTOL and rel(out["V"], MR[k + "V"]) < RTOL


@pytest.mark.gpu
def test_r_products_generator_and_pool_allocator(rs):
    M0, EL, EF, s2, _ = state("by_col")
    rs.call("ebpmf_b200_set_y", R.RSparse(Y))
    rs.call("ebpmf_b200_set_m", M0)
    pr = rs.call("ebpmf_b200_m_products", N, P, EL[:, :3], EF[:, :3])                 # N1, on the resident M
    resid = M0 - EL[:, :3] @ EF[:, :3].T
    assert pr.names == ["MF", "MtL", "r2_col", "r2_row"]
    assert rel(pr.get("MF"), M0 @ EF[:, :3]) < 1e-10 and rel(pr.get("MtL"), M0.T @ EL[:, :3]) < 1e-10
    assert rel(pr.get("r2_col"), (resid ** 2).sum(axis=0)) < 1e-10 and rel(pr.get("r2_row"), (resid ** 2).sum(axis=1)) < 1e-10
    sim = rs.call("ebpmf_b200_sim_data_log", 300, 64, 3, seed=7, log_offset=-1.0)      # N4: generated on the device
    Ys = sim.get("Y")
    assert isinstance(Ys, R.RSparse) and Ys.dense.shape == (300, 64) and (Ys.dense >= 0).all() and Ys.dense.sum() > 0
    assert sim.get("EL").shape == (300, 5) and sim.get("EF").shape == (64, 5) and sim.get("sigma2").shape == (64,)
    # a result of 64 MB and more is allocated through allocVector3 + the R_allocator_t over the library's host pool
    n, p = 4096, 2100
    rng = np.random.default_rng(3)
    Yb = sp.random(n, p, density=0.02, random_state=5, data_rvs=lambda k: rng.integers(1, 5, k).astype(float)).toarray()
    ELb, EFb = rng.normal(0, 0.3, (n, 3)), rng.normal(0, 0.3, (p, 3))
    rs.call("ebpmf_b200_set_y", R.RSparse(Yb))
    res = rs.call("ebpmf_b200_vga_step", np.zeros((n, p)), ELb, EFb, np.full(p, 0.5), "by_col", 10, 1e-5)
    assert rs.last_used_allocator and res.get("V") is None and np.isfinite(res.get("M")).all()
    ref = vn.ebpmf_log_vga_step(np.zeros((512, p)), Yb[:512], ELb[:512], EFb, np.full(p, 0.5), "by_col", int(res.get("n_updates")[0]), 0.0)
    assert relM(res.get("M")[:512], ref["M"]) < RTOL


@pytest.mark.gpu
def test_r_group_wrappers_on_one_device(rs):
    """The one-process multi-GPU route of the R shim (EBPMF_B200_NGPU), here with a group of one device."""
    M0, EL, EF, s2, _ = state("by_col")
    rs.call("ebpmf_b200_group", r_mock.as_int([0]))
    rs.call("ebpmf_b200_group_set_y", R.RSparse(Y))
    lf = rs.call("ebpmf_b200_group_sum_lfactorial")
    assert abs(lf[0] - MR["fn/lfact"][0]) < RTOL * MR["fn/lfact"][0]
    check_step_list(rs.call("ebpmf_b200_group_vga_step", M0, N, P, EL, EF, s2, "by_col", 10, 1e-5, True), "fn/by_col_def_")
    rs.call("ebpmf_b200_group_set_m", M0)
    check_step_list(rs.call("ebpmf_b200_group_vga_step", None, N, P, EL, EF, s2, "by_col", 10, 1e-5, True), "fn/by_col_def_")


@pytest.mark.gpu
@pytest.mark.parametrize("vt", ["by_col", "by_row"])
def test_r_group_wrappers_on_two_devices(pkg, vt):
    """EBPMF_B200_NGPU = 2 from one R session: rows sharded over two GPUs inside the library (NCCL for the update count
    and the column sums), full host matrices in and out through the R wrappers; against the C oracle on the whole
    matrix.  Skipped on single-GPU boxes (profiles/r02_r_boundary_group_n2.log is a 2-GPU run)."""
    import ctypes
    from oracle import c_oracle, simdata
    cnt = ctypes.c_int()
    pkg.load().ebpmf_device_count(ctypes.byref(cnt))
    if cnt.value < 2:
        pytest.skip("needs 2 GPUs")
    sim = simdata.sim_data_log(1500, 211, K=4, seed=12345, log_offset=simdata.offset_for_density(0.10, K=4))
    st = simdata.solver_state(sim, start="cold", var_type=vt, seed=12346)
    ref = c_oracle.vga_step(st["M0"], sim["Y"], st["EL"], st["EF"], st["sigma2"], vt, 10, 1e-5)
    s = r_mock.RSession()
    try:
        s.call("ebpmf_b200_group", r_mock.as_int([0, 1]))
        s.call("ebpmf_b200_group_set_y", R.RSparse(sim["Y"]))
        assert abs(s.call("ebpmf_b200_group_sum_lfactorial")[0] - c_oracle.sum_lfactorial(sim["Y"])) < 1e-8 * c_oracle.sum_lfactorial(sim["Y"])
        s.call("ebpmf_b200_group_set_m", st["M0"])
        for M in (st["M0"], None):                     # M in and out; then the shards' resident M (rewound by set_m above)
            if M is None:
                s.call("ebpmf_b200_group_set_m", st["M0"])
            res = s.call("ebpmf_b200_group_vga_step", M, 1500, 211, st["EL"], st["EF"], st["sigma2"], vt, 10, 1e-5, True)
            assert int(res.get("n_updates")[0]) == ref["n_updates"]
            assert relM(res.get("M"), ref["M"]) < RTOL and rel(res.get("V"), ref["V"]) < RTOL and rel(res.get("v_sum"), ref["v_sum"]) < RTOL
            for name in ("sym", "ssexp", "slogv"):
                assert abs(res.get(name)[0] - ref[name]) <= RTOL * max(1.0, abs(ref[name]))
    finally:
        s.close()


# ---------------------------------------------------------------------------------------------------------------
# ebpmf_b200_vga_block: the ONE statement of ebpmf_log's loop that changes
# ---------------------------------------------------------------------------------------------------------------
PATCH_INIT = """
blk = ebpmf_b200_vga_block_init(Y_as_given, fit_flash$flash_fit$Y, if(n_batch>1) batches else NULL)
const = - blk$sum_lfactorial
"""
PATCH_BLOCK = """
{
  out = ebpmf_b200_vga_block(blk, fit_flash, sigma2, var_type, vga_control, sigma2_control, general_control$save_latent_M)
  fit_flash$flash_fit$Y = out$M
  if(general_control$save_latent_M){ V = out$V }
  sym = out$sym ; ssexp = out$ssexp ; slogv = out$slogv ; v_sum = out$v_sum ; sigma2 = out$sigma2
}
"""


def install_drop_in(it):
    """Splices the statements of INTEGRATION.md section 2 into the parsed body of the reference's ebpmf_log():
    `Y_as_given = Y` before the data is split, the init after the flash initialisation, the block in place of the
    `if(n_batch >1){...}else{...}` statement of the loop.  Nothing else of the function is touched."""
    from oracle import minir_reference as mr
    fn = it.get("ebpmf_log")
    body = fn.body[1]
    nb_stmt, split_if, vga_if = mr.loop_statements(it)
    body.insert(body.index(nb_stmt), R.parse("Y_as_given = Y")[0])
    i_init = mr._find(body, lambda s: s[0] == "assign" and s[1] == ("name", "fit_flash") and s[2][0] == "call"
                      and s[2][1] == ("name", "ebpmf_log_flash_init"))
    body[i_init + 1:i_init + 1] = R.parse(PATCH_INIT)
    loop = body[mr._find(body, lambda s: s[0] == "for" and s[1] == "iter")]
    lbody = loop[3][1]
    lbody[lbody.index(vga_if)] = R.parse(PATCH_BLOCK)[0]


@pytest.mark.skipif(os.path.exists("/root/reference/R/ebpmf_log.R") is False, reason="the reference checkout is not present")
@pytest.mark.parametrize("vt", VTS)
@pytest.mark.parametrize("bs", ["inf", "25"])
def test_reference_ebpmf_log_with_the_drop_in_spliced_in(vt, bs):
    """The reference's own ebpmf_log(), its VGA block replaced by ebpmf_b200_vga_block as INTEGRATION.md says, run by
    the evaluator with `.Call` answered by a test double (the CPU oracle; there is no GPU where the checkout is):
    the R-level plumbing -- which values travel, in which storage mode, in which order, what comes back under which
    name -- reproduces the traces of the untouched ebpmf_log()."""
    import r_fake_calls
    from oracle import minir_reference as mr
    s = r_mock.RSession(reference_root="/root/reference")
    fake = r_fake_calls.OracleBackedCalls(s)
    s.dot_call = fake
    install_drop_in(s.it)
    l0 = np.log(Y.mean(axis=1) + 0.1)
    f0 = np.log((Y.sum(axis=0) + 0.1) / np.exp(l0).sum())
    s2 = {"by_col": np.full(P, 0.5), "by_row": np.full(N, 0.5), "constant": np.array([0.5])}[vt]
    out = mr.run_driver(s.it, Y, l0, f0, np.asfortranarray(np.log(Y + 0.5)), s2, vt, K=3, outer_iters=4,
                        batch_size=np.inf if bs == "inf" else 25, sparse=bs != "inf")
    k = "drv/%s/%s/" % (vt, bs)
    assert rel(out["elbo_trace"][1:], MR[k + "elbo_trace"][1:]) < 1e-12 and rel(out["sigma2"], MR[k + "sigma2"]) < 1e-12
    if bs == "inf":
        assert relM(out["M"], MR[k + "M"]) < 1e-12 and rel(out["V"], MR[k + "V"]) < 1e-12
    step = "C_ebpmf_vga_step_resident" if bs == "inf" else "C_ebpmf_batched_vga_step"
    assert fake.log.count(step) == 4 and fake.log.count("C_ebpmf_log_update_sigma2") == 4 and fake.log.count("C_calc_ebpmf_log_obj") == 4
    assert "C_vga_pois_solver_mat_newton" not in fake.log          # the per-call dense signature is no longer on the loop's path


@pytest.mark.skipif(os.path.exists("/root/reference/R/ebpmf_log.R") is False, reason="the reference checkout is not present")
def test_reference_ebpmf_log_with_the_drop_in_on_the_group_route(monkeypatch):
    """EBPMF_B200_NGPU = 2: the block takes the one-process multi-GPU route (ebpmf_b200_group_*).  Same spliced
    ebpmf_log(), `.Call` answered by the test double: the R-level plumbing of that route (the 2-GPU run of the group
    wrappers themselves is tests/test_r_boundary.py::test_r_group_wrappers_on_two_devices)."""
    import r_fake_calls
    from oracle import minir_reference as mr
    monkeypatch.setenv("EBPMF_B200_NGPU", "2")
    s = r_mock.RSession(reference_root="/root/reference")
    fake = r_fake_calls.OracleBackedCalls(s)
    s.dot_call = fake
    install_drop_in(s.it)
    l0 = np.log(Y.mean(axis=1) + 0.1)
    f0 = np.log((Y.sum(axis=0) + 0.1) / np.exp(l0).sum())
    out = mr.run_driver(s.it, Y, l0, f0, np.asfortranarray(np.log(Y + 0.5)), np.full(P, 0.5), "by_col", K=3, outer_iters=4, sparse=True)
    k = "drv/by_col/inf/"
    assert rel(out["elbo_trace"][1:], MR[k + "elbo_trace"][1:]) < 1e-12 and relM(out["M"], MR[k + "M"]) < 1e-12
    assert fake.group_devices == [0, 1] and fake.log.count("C_ebpmf_group_vga_step") == 4 and "C_ebpmf_vga_step_resident" not in fake.log


@pytest.mark.gpu
@pytest.mark.parametrize("vt", VTS)
@pytest.mark.parametrize("bs", ["inf", "25"])
def test_r_block_through_the_whole_driver(rs, vt, bs):
    """ebpmf_b200_vga_block over the compiled glue and the GPU, four outer iterations, against ebpmf_log() itself."""
    l0 = np.log(Y.mean(axis=1) + 0.1)
    f0 = np.log((Y.sum(axis=0) + 0.1) / np.exp(l0).sum())
    M_init = np.asfortranarray(np.log(Y + 0.5))
    s2 = {"by_col": np.full(P, 0.5), "by_row": np.full(N, 0.5), "constant": np.array([0.5])}[vt]
    batches = None if bs == "inf" else rs.it.run("split(1:%d, cut(seq_along(1:%d), 4, labels = FALSE))" % (N, N))
    blk = rs.call("ebpmf_b200_vga_block_init", Y if bs == "inf" else R.RSparse(Y), M_init, batches)
    assert R.r_string(blk.get("route")) == ("resident" if bs == "inf" else "batched")
    vga_control = dict(maxiter_vga=10, vga_tol=1e-5)
    sigma2_control = dict(est_sigma2=True, a0=1, b0=1, cap_var_mean_ratio=0)

    def block(M, fit, sig):
        ff = dict(flash_fit=dict(Y=np.asfortranarray(M), R2=fit["R2"], EF=[fit["EL"], fit["EF"]]))
        o = rs.call("ebpmf_b200_vga_block", blk, ff, sig, vt, vga_control, sigma2_control, bs == "inf")
        return dict(M=o.get("M"), V=o.get("V"), v_sum=o.get("v_sum"), sym=o.get("sym")[0], ssexp=o.get("ssexp")[0], slogv=o.get("slogv")[0],
                    sigma2=o.get("sigma2"))

    try:
        out = standin_flash.outer_loop(Y, l0, f0, M_init, s2, vt, 3, 4, block=block, calc_obj=lambda *a: rs.call("calc_ebpmf_log_obj", *a)[0],
                                       const=-blk.get("sum_lfactorial")[0])
    finally:
        if blk.get("handle") is not None:
            r_mock.lib().mockr_free(blk.get("handle").sexp)
    k = "drv/%s/%s/" % (vt, bs)
    assert rel(out["elbo_trace"][1:], MR[k + "elbo_trace"][1:]) < RTOL and rel(out["sigma2"], MR[k + "sigma2"]) < RTOL
    if bs == "inf":
        assert relM(out["M"], MR[k + "M"]) < RTOL and rel(out["V"], MR[k + "V"]) < RTOL
