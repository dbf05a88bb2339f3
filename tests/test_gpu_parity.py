"""GPU parity tests proper: the CUDA path, called through the C ABI (ctypes), against the CPU
oracle on the same seeded inputs.  Tolerance: 1e-8 relative for M, V, sigma2, ELBO terms
(BASELINE.json north_star) and the global update count u EXACTLY equal.

`M` is compared as |dM| / max(|M|, 1): M is a log-scale quantity that crosses zero, and the
north star's 1e-8 is meant relative to the scale of the iterates."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import c_oracle, simdata, vga_numpy

pytestmark = pytest.mark.gpu

RTOL = 1e-8


def rel(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


def relM(a, b):
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1.0)))


def make_case(n, p, K, density, start="cold", var_type="by_col", seed=12345):
    c = simdata.offset_for_density(density, K=K)
    sim = simdata.sim_data_log(n, p, K=K, seed=seed, log_offset=c)
    st = simdata.solver_state(sim, start=start, var_type=var_type, seed=seed + 1)
    return sim, st


def check_step(res, ref, var_type):
    assert res["n_updates"] == ref["n_updates"][0]
    assert res["converged"] == ref["converged"][0]
    assert relM(res["M"], ref["M"]) < RTOL
    if res["V"] is not None:
        assert rel(res["V"], ref["V"]) < RTOL
    assert rel(res["v_sum"], ref["v_sum"]) < RTOL
    for k in ("sym", "ssexp", "slogv"):
        assert abs(res[k] - ref[k]) <= RTOL * max(abs(ref[k]), 1.0), (k, res[k], ref[k])


# ---------------------------------------------------------------------------------------
# fused step (K1): a1 + a2 + a3 + a4 + a5
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("var_type", ["by_col", "by_row", "constant"])
@pytest.mark.parametrize("maxiter,tol", [(10, 1e-5), (1000, 1e-8), (3, 1e-5)])
def test_fused_step_vs_oracle(gpu_ctx, var_type, maxiter, tol):
    sim, st = make_case(1500, 333, 5, 0.10, var_type=var_type)
    gpu_ctx.set_y(sp.csc_matrix(sim["Y"]))
    res = gpu_ctx.vga_step_host(st["M0"], st["EL"], st["EF"], st["sigma2"], var_type, maxiter, tol, return_V=True)
    ref = vga_numpy.ebpmf_log_vga_step(st["M0"], sim["Y"], st["EL"], st["EF"], st["sigma2"], var_type, maxiter, tol)
    check_step(res, ref, var_type)
    if maxiter == 3:
        assert not res["converged"] and res["n_updates"] == 3      # exhaustion: stale V (F6)


def test_fused_c1_readme_dense_y(gpu_ctx):
    """C1: README recipe shapes, dense Y (as.matrix(Y), R/ebpmf_log.R:202), l0 = f0 = 0."""
    sim = simdata.readme_example()
    st = simdata.solver_state(sim, start="cold")
    gpu_ctx.set_y(sim["Y"])
    res = gpu_ctx.vga_step_host(st["M0"], st["EL"], st["EF"], st["sigma2"], "by_col", 10, 1e-5, return_V=True)
    ref = vga_numpy.ebpmf_log_vga_step(st["M0"], sim["Y"], st["EL"], st["EF"], st["sigma2"], "by_col", 10, 1e-5)
    check_step(res, ref, "by_col")


def test_fused_c2_full_size(gpu_ctx):
    """C2: sim_data_log 10k x 2k, K = 5, ~10 % non-zero, package defaults; oracle = C transcription."""
    sim, st = make_case(10000, 2000, 5, 0.10)
    gpu_ctx.set_y(sp.csc_matrix(sim["Y"]))
    res = gpu_ctx.vga_step_host(st["M0"], st["EL"], st["EF"], st["sigma2"], "by_col", 10, 1e-5, return_V=True)
    ref = c_oracle.vga_step(st["M0"], sim["Y"], st["EL"], st["EF"], st["sigma2"], "by_col", 10, 1e-5)
    ref = dict(ref, n_updates=[ref["n_updates"]], converged=[ref["converged"]])
    check_step(res, ref, "by_col")


@pytest.mark.parametrize("n,p", [(1, 1), (31, 17), (257, 15), (513, 33), (1000, 100)])
def test_fused_ragged_shapes(gpu_ctx, n, p):
    sim, st = make_case(n, p, 3, 0.3, seed=n * 1000 + p)
    gpu_ctx.set_y(sp.csc_matrix(sim["Y"]))
    res = gpu_ctx.vga_step_host(st["M0"], st["EL"], st["EF"], st["sigma2"], "by_col", 50, 1e-8, return_V=True)
    ref = vga_numpy.ebpmf_log_vga_step(st["M0"], sim["Y"], st["EL"], st["EF"], st["sigma2"], "by_col", 50, 1e-8)
    check_step(res, ref, "by_col")


def test_fused_all_zero_y(gpu_ctx):
    """Empty sparse matrix: no non-zeros at all."""
    sim, st = make_case(300, 40, 3, 0.1)
    Y0 = np.zeros_like(sim["Y"])
    gpu_ctx.set_y(sp.csc_matrix(Y0))
    M0 = st["EL"] @ st["EF"].T
    res = gpu_ctx.vga_step_host(M0, st["EL"], st["EF"], st["sigma2"], "by_col", 30, 1e-8, return_V=True)
    ref = vga_numpy.ebpmf_log_vga_step(M0, Y0, st["EL"], st["EF"], st["sigma2"], "by_col", 30, 1e-8)
    check_step(res, ref, "by_col")


def test_fused_warm_start_many_iterations(gpu_ctx):
    sim, st = make_case(900, 200, 5, 0.10, start="warm")
    gpu_ctx.set_y(sp.csc_matrix(sim["Y"]))
    res = gpu_ctx.vga_step_host(st["M0"], st["EL"], st["EF"], st["sigma2"], "by_col", 1000, 1e-8, return_V=True)
    ref = vga_numpy.ebpmf_log_vga_step(st["M0"], sim["Y"], st["EL"], st["EF"], st["sigma2"], "by_col", 1000, 1e-8)
    check_step(res, ref, "by_col")
    assert res["n_updates"] > 10


def test_resident_step_rewind_and_sigma2_update(gpu_ctx, pkg):
    """Resident-state API: two identical steps after rewind give identical results; K5 + a7."""
    sim, st = make_case(1200, 256, 4, 0.08)
    n, p = sim["Y"].shape
    gpu_ctx.set_y(sp.csc_matrix(sim["Y"]))
    gpu_ctx.set_m(st["M0"]); gpu_ctx.set_factors(st["EL"], st["EF"]); gpu_ctx.set_sigma2(st["sigma2"], "by_col")
    i1 = gpu_ctx.vga_step(10, 1e-5)
    m1 = gpu_ctx.get_m(); v1 = gpu_ctx.get_v_sum()
    gpu_ctx.rewind_m()
    i2 = gpu_ctx.vga_step(10, 1e-5)
    m2 = gpu_ctx.get_m(); v2 = gpu_ctx.get_v_sum()
    assert i1.n_updates == i2.n_updates and np.array_equal(m1, m2) and np.array_equal(v1, v2)
    assert (i1.sym, i1.ssexp, i1.slogv) == (i2.sym, i2.ssexp, i2.slogv)      # reductions are deterministic
    ref = vga_numpy.ebpmf_log_vga_step(st["M0"], sim["Y"], st["EL"], st["EF"], st["sigma2"], "by_col", 10, 1e-5)
    fit = dict(R2=st["R2"], EL=st["EL"], EF=st["EF"])
    for cap in (0.0, 5.0):
        gpu_ctx.set_sigma2(st["sigma2"], "by_col")
        s_gpu = gpu_ctx.update_sigma2(st["R2"], 1.0, 1.0, cap)
        s_ref = vga_numpy.ebpmf_log_update_sigma2(fit, st["sigma2"], ref["v_sum"], "by_col", cap, 1.0, 1.0, n, p)
        assert rel(s_gpu, s_ref) < RTOL
    const = -vga_numpy.sum_lfactorial_sparseMat(sp.csc_matrix(sim["Y"]))
    assert abs(-gpu_ctx.sum_lfactorial() - const) <= RTOL * abs(const)
    o_ref = vga_numpy.calc_ebpmf_log_obj(n, p, ref["sym"], ref["ssexp"], ref["slogv"], ref["v_sum"], s_ref,
                                         st["R2"], -123.4, const, n, 1.0, 1.0)
    o_gpu = pkg.calc_ebpmf_log_obj(n, p, i1.sym, i1.ssexp, i1.slogv, v1, s_gpu, st["R2"], -123.4,
                                   -gpu_ctx.sum_lfactorial(), n, 1.0, 1.0, ctx=gpu_ctx)
    assert abs(o_gpu - o_ref) <= RTOL * abs(o_ref)


@pytest.mark.parametrize("var_type", ["by_row", "constant"])
def test_update_sigma2_other_var_types(gpu_ctx, pkg, var_type):
    sim, st = make_case(400, 90, 3, 0.2, var_type=var_type)
    n, p = sim["Y"].shape
    ref = vga_numpy.ebpmf_log_vga_step(st["M0"], sim["Y"], st["EL"], st["EF"], st["sigma2"], var_type, 10, 1e-5)
    fit = dict(R2=st["R2"], EL=st["EL"], EF=st["EF"])
    changed = []
    for cap in (0.0, 3.0, 1e9):        # 1e9: a threshold near 20, so the gate selects (by_row: against Rfast::rowMaxs' POSITIONS)
        s_ref = vga_numpy.ebpmf_log_update_sigma2(fit, st["sigma2"], ref["v_sum"], var_type, cap, 1.0, 1.0, n, p)
        s_gpu = pkg.ebpmf_log_update_sigma2(fit, st["sigma2"], ref["v_sum"], var_type, cap, 1.0, 1.0, n, p, ctx=gpu_ctx)
        assert rel(s_gpu, s_ref) < RTOL
        changed.append(int(np.sum(np.atleast_1d(s_ref) != np.atleast_1d(st["sigma2"]))))
    if var_type == "by_row":
        assert 0 < changed[2] < changed[1] == n


def test_nan_raises_like_r(gpu_ctx, pkg):
    """A NaN at the stop test is an error in R (`if(NA)`), R/vga_solver_mat.R:26."""
    sim, st = make_case(200, 50, 3, 0.2)
    M0 = st["M0"].copy(); M0[7, 3] = np.nan
    EL = st["EL"].copy(); EL[11, 2] = np.nan
    gpu_ctx.set_y(sp.csc_matrix(sim["Y"]))
    with pytest.raises(pkg.EbpmfError) as ei:
        gpu_ctx.vga_step_host(st["M0"], EL, st["EF"], st["sigma2"], "by_col", 10, 1e-5)
    assert ei.value.code == 3
    with pytest.raises(vga_numpy.RError):
        vga_numpy.ebpmf_log_vga_step(st["M0"], sim["Y"], EL, st["EF"], st["sigma2"], "by_col", 10, 1e-5)
    # a NaN in the START matrix survives the clamp M = Pmin(M, const0-1) (Rfast: std::min(M, cap) = (cap < M) ? cap : M),
    # reaches f and stops R the same way -- on the fused path, the dense signature and _low_memory alike
    with pytest.raises(pkg.EbpmfError) as ei:
        gpu_ctx.vga_step_host(M0, st["EL"], st["EF"], st["sigma2"], "by_col", 10, 1e-5)
    assert ei.value.code == 3
    with pytest.raises(pkg.EbpmfError) as ei:
        pkg.vga_pois_solver_mat_newton(M0, sim["Y"], 1, st["EL"] @ st["EF"].T, 0.5, maxiter=10, tol=1e-5, ctx=gpu_ctx)
    assert ei.value.code == 3
    with pytest.raises(pkg.EbpmfError) as ei:
        pkg.vga_pois_solver_mat_newton_low_memory(M0, sim["Y"], sim["L0"], sim["F0"], st["EL"][:, 2:], st["EF"][:, 2:], st["sigma2"], "by_col",
                                                  10, 1e-5, ctx=gpu_ctx)
    assert ei.value.code == 3
    with pytest.raises(vga_numpy.RError):
        vga_numpy.ebpmf_log_vga_step(M0, sim["Y"], st["EL"], st["EF"], st["sigma2"], "by_col", 10, 1e-5)
    with pytest.raises(c_oracle.OracleNaN):
        c_oracle.vga_step(M0, sim["Y"], st["EL"], st["EF"], st["sigma2"], "by_col", 10, 1e-5)


def test_bad_var_type_and_call_order(pkg):
    ctx = pkg.VgaContext(0)
    with pytest.raises(pkg.EbpmfError) as ei:
        ctx.vga_step(10, 1e-5)
    assert ei.value.code == 5
    with pytest.raises(pkg.EbpmfError) as ei:
        pkg.adjust_var_shape(np.ones(3), "by_banana", 3, 3)
    assert ei.value.code == 4
    ctx.close()


# ---------------------------------------------------------------------------------------
# drop-in dense signatures (K2, K3, K4)
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("sigma_kind", ["matrix", "rowvec", "scalar"])
@pytest.mark.parametrize("sparse_x", [False, True])
def test_dense_solver_signature(gpu_ctx, pkg, sigma_kind, sparse_x):
    sim, st = make_case(700, 150, 4, 0.15)
    n, p = sim["Y"].shape
    rng = np.random.default_rng(5)
    Beta = st["EL"] @ st["EF"].T
    S = np.exp(rng.normal(0, 0.1, size=(n, p)))                    # matrix S = tcrossprod(l0,f0)-like
    Sigma2 = {"matrix": vga_numpy.adjust_var_shape(st["sigma2"], "by_col", n, p),
              "rowvec": rng.uniform(0.1, 1.0, size=n), "scalar": 0.37}[sigma_kind]
    X = sp.csc_matrix(sim["Y"]) if sparse_x else sim["Y"]
    gi, oi = {}, {}
    res = pkg.vga_pois_solver_mat_newton(st["M0"], X, S, Beta, Sigma2, maxiter=1000, tol=1e-8, info=gi, ctx=gpu_ctx)
    ref = vga_numpy.vga_pois_solver_mat_newton(st["M0"], sim["Y"], S, Beta, Sigma2, maxiter=1000, tol=1e-8, info=oi)
    assert gi["n_updates"] == oi["n_updates"] and gi["converged"] == oi["converged"]
    assert relM(res["M"], ref["M"]) < RTOL and rel(res["V"], ref["V"]) < RTOL
    m_only = pkg.vga_pois_solver_mat_newton(st["M0"], X, 1, Beta, Sigma2, maxiter=4, tol=1e-8, return_V=False,
                                            ctx=gpu_ctx)
    ref2 = vga_numpy.vga_pois_solver_mat_newton(st["M0"], sim["Y"], 1, Beta, Sigma2, maxiter=4, tol=1e-8,
                                                return_V=False)
    assert isinstance(m_only, np.ndarray) and relM(m_only, ref2) < RTOL


@pytest.mark.parametrize("maxiter", [1, 10, 60])
def test_fixed_iter(gpu_ctx, pkg, maxiter):
    """a9; maxiter = 1 is its only call site (inst/code/Poisson_MF_splitting.R:212-214)."""
    sim, st = make_case(600, 120, 3, 0.2)
    n, p = sim["Y"].shape
    Beta = st["EL"] @ st["EF"].T
    Sigma2 = vga_numpy.adjust_var_shape(st["sigma2"], "by_col", n, p)
    for V0 in (None, np.full((n, p), 0.3)):
        res = pkg.vga_pois_solver_mat_newton_fixed_iter(st["M0"], V0, sim["Y"], 1, Beta, Sigma2, maxiter, ctx=gpu_ctx)
        ref = vga_numpy.vga_pois_solver_mat_newton_fixed_iter(st["M0"], V0, sim["Y"], 1, Beta, Sigma2, maxiter)
        assert relM(res["M"], ref["M"]) < RTOL and rel(res["V"], ref["V"]) < RTOL


@pytest.mark.parametrize("var_type", ["by_col", "by_row", "constant"])
@pytest.mark.parametrize("sparse_y", [False, True])
def test_low_memory(gpu_ctx, pkg, var_type, sparse_y):
    """a10 with its quirks (F8): multiplicative l0*f0, clamp to Beta, sigma2-less derivative."""
    sim, st = make_case(800, 130, 3, 0.15, var_type=var_type)
    n, p = sim["Y"].shape
    l0 = sim["L0"]; f0 = sim["F0"]
    EL = st["EL"][:, 2:]; EF = st["EF"][:, 2:]                      # l0, f0 are multiplicative here
    M0 = np.log(sim["Y"] + 0.5) - np.log(l0)[:, None] - np.log(f0)[None, :]
    Y = sp.csc_matrix(sim["Y"]) if sparse_y else sim["Y"]
    for maxiter, tol in ((1000, 1e-8), (3, 1e-5)):
        gi, oi = {}, {}
        res = pkg.vga_pois_solver_mat_newton_low_memory(M0, Y, l0, f0, EL, EF, st["sigma2"], var_type, maxiter, tol,
                                                        info=gi, ctx=gpu_ctx)
        ref = vga_numpy.vga_pois_solver_mat_newton_low_memory(M0, sim["Y"], l0, f0, EL, EF, st["sigma2"], var_type,
                                                              maxiter, tol, info=oi)
        assert gi["n_updates"] == oi["n_updates"] and gi["converged"] == oi["converged"]
        assert relM(res, ref) < RTOL


def test_sum_lfactorial(gpu_ctx, pkg):
    rng = np.random.default_rng(3)
    x = rng.poisson(3.0, size=200001).astype(np.float64)
    x[:5] = [0, 1, 2, 170, 25000]
    got = pkg.sum_lfactorial_sparseMat(x, ctx=gpu_ctx)
    ref = c_oracle.sum_lfactorial(x)
    assert abs(got - ref) <= RTOL * abs(ref)


# ---------------------------------------------------------------------------------------
# device-generated inputs (bench path) read back and checked against the oracle
# ---------------------------------------------------------------------------------------
def test_device_sim_data_matches_oracle(gpu_ctx):
    n, p, K = 3000, 640, 6
    nnz = gpu_ctx.sim_data_log(n, p, K, seed=12345, log_offset=-3.2)
    cp, ri, x = gpu_ctx.get_y_csc()
    Y = sp.csc_matrix((x, ri, cp), shape=(n, p))
    assert nnz == Y.nnz and 0.02 < nnz / (n * p) < 0.4
    assert np.all(x > 0) and np.all(x == np.round(x))
    EL, EF = gpu_ctx.get_factors()
    s2 = gpu_ctx.get_sigma2()
    M0 = gpu_ctx.get_m()
    info = gpu_ctx.vga_step(10, 1e-5, want_v=True)
    ref = c_oracle.vga_step(M0, Y.toarray(), EL, EF, s2, "by_col", 10, 1e-5)
    assert info.n_updates == ref["n_updates"] and bool(info.converged) == ref["converged"]
    assert relM(gpu_ctx.get_m(), ref["M"]) < RTOL
    assert rel(gpu_ctx.get_v(), ref["V"]) < RTOL
    assert rel(gpu_ctx.get_v_sum(), ref["v_sum"]) < RTOL
    for k in ("sym", "ssexp", "slogv"):
        assert abs(getattr(info, k) - ref[k]) <= RTOL * max(abs(ref[k]), 1.0)
    rows = gpu_ctx.get_rows(1000, 257)
    assert np.array_equal(rows, gpu_ctx.get_m()[1000:1257])


# ---------------------------------------------------------------------------------------
# options: delta mode (frozen units), periodic fast-forward ("freeze"), history order of the first pass
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("freeze,history", [(0, 0), (1, 0), (0, 1), (1, 1)])
@pytest.mark.parametrize("delta", [0.0, 1e-14])
@pytest.mark.parametrize("var_type,maxiter,tol", [("by_col", 10, 1e-5), ("by_row", 1000, 1e-8), ("constant", 3, 1e-5)])
def test_options_delta_freeze_history(pkg, freeze, history, delta, var_type, maxiter, tol):
    """Every option keeps the 1e-8 parity and R's update count; "freeze" (skipping the repeats of a period-2
    iterate sequence) and "history" (hardest tiles first) are BIT-identical to plain iteration in natural
    order; delta mode stays within ~delta of exact mode.  The second call runs with the first call's per-tile
    counts as order."""
    sim, st = make_case(1300, 301, 5, 0.10, var_type=var_type)
    ctx, plain = pkg.VgaContext(0), pkg.VgaContext(0)
    try:
        ctx.set_option("freeze", freeze); ctx.set_option("history", history); ctx.set_option("delta", delta)
        plain.set_option("freeze", 0); plain.set_option("history", 0)
        for c in (ctx, plain):
            c.set_y(sp.csc_matrix(sim["Y"]))
        ref = vga_numpy.ebpmf_log_vga_step(st["M0"], sim["Y"], st["EL"], st["EF"], st["sigma2"], var_type, maxiter, tol)
        exact = plain.vga_step_host(st["M0"], st["EL"], st["EF"], st["sigma2"], var_type, maxiter, tol, return_V=True)
        for call in range(2):
            res = ctx.vga_step_host(st["M0"], st["EL"], st["EF"], st["sigma2"], var_type, maxiter, tol, return_V=True)
            check_step(res, ref, var_type)
            assert ctx.counters()["have_history"] == 1.0
            if delta == 0.0:
                assert np.array_equal(res["M"], exact["M"]) and np.array_equal(res["V"], exact["V"])   # bit-identical
                assert np.array_equal(np.atleast_1d(res["v_sum"]), np.atleast_1d(exact["v_sum"]))
                assert (res["sym"], res["ssexp"], res["slogv"]) == (exact["sym"], exact["ssexp"], exact["slogv"])
            else:
                assert np.max(np.abs(res["M"] - exact["M"])) < 1e-12 and rel(res["V"], exact["V"]) < 1e-12
    finally:
        ctx.close(); plain.close()


def test_option_validation(pkg):
    ctx = pkg.VgaContext(0)
    with pytest.raises(pkg.EbpmfError):
        ctx.set_option("delta", 1e-3)          # would break the 1e-8 contract: refused
    with pytest.raises(pkg.EbpmfError):
        ctx.set_option("no_such_option", 1)
    ctx.close()


# ---------------------------------------------------------------------------------------
# large size: size-independent properties + an oracle row slab with the global count forced
# ---------------------------------------------------------------------------------------
def test_large_size_properties(pkg):
    """60 000 x 6 000 (3.6e8 entries, device-generated, ~5 % non-zero): the reductions returned by the step
    equal the same sums recomputed from the downloaded M, V (linearity / checksum-of-checksums), V obeys
    V*(sigma2*Y + Beta + 1 - M) = sigma2, |f| < tol everywhere, and a row slab matches the oracle run with
    the global update count forced."""
    ctx = pkg.VgaContext(0)
    try:
        n, p, K = 60000, 6000, 10
        nnz = ctx.sim_data_log(n, p, K, seed=2024, log_offset=-4.644)
        assert 0.03 < nnz / (n * p) < 0.08
        info = ctx.vga_step(10, 1e-5, want_v=True)
        assert info.converged and 3 <= info.n_updates <= 10
        M = ctx.get_m(); V = ctx.get_v(); vs = ctx.get_v_sum()
        s2 = ctx.get_sigma2()
        cp, ri, x = ctx.get_y_csc()
        Y = sp.csc_matrix((x, ri, cp), shape=(n, p))
        # checksums: column sums of V, scalar partials
        assert rel(vs, V.sum(axis=0, dtype=np.longdouble).astype(np.float64)) < 1e-10
        ssexp = float(np.sum(np.exp(M + V / 2), dtype=np.longdouble))
        slogv = float(np.sum(np.log(V) / 2 + 0.9189385, dtype=np.longdouble))
        cols = np.repeat(np.arange(p), np.diff(cp))
        sym = float(np.sum(x * M[ri, cols], dtype=np.longdouble))
        assert abs(info.ssexp - ssexp) < 1e-10 * abs(ssexp)
        assert abs(info.slogv - slogv) < 1e-10 * abs(slogv)
        assert abs(info.sym - sym) < 1e-10 * abs(sym)
        # a row slab against the oracle with the GLOBAL update count forced (tol = 0 => exactly u updates)
        rows = slice(12345, 12345 + 700)
        EL, EF = ctx.get_factors()
        Beta = EL[rows] @ EF.T
        S2 = np.repeat(s2[None, :], 700, axis=0)
        Yr = Y[rows].toarray()
        ctx.rewind_m(); M0 = ctx.get_rows(12345, 700)
        ref = vga_numpy.vga_pois_solver_mat_newton(M0, Yr, 1, Beta, S2, info.n_updates, 0.0)["M"]
        assert relM(M[rows], ref) < RTOL
        # invariants on the slab: fresh V and stationarity at the global count
        temp = S2 * Yr + Beta + 1 - M[rows]
        assert rel(V[rows] * temp, S2) < 1e-12
        f = Yr - np.exp(M[rows] + V[rows] / 2) - (M[rows] - Beta) / S2
        assert np.max(np.abs(f)) < 1e-5
    finally:
        ctx.close()


def test_idempotence_second_step_needs_no_update(gpu_ctx):
    """A converged M is a fixed point of the step: the second call breaks at iteration 1 (u = 0) and
    returns M unchanged, V recomputed from it -- as R would."""
    sim, st = make_case(2000, 257, 4, 0.10)
    gpu_ctx.set_y(sp.csc_matrix(sim["Y"]))
    r1 = gpu_ctx.vga_step_host(st["M0"], st["EL"], st["EF"], st["sigma2"], "by_col", 100, 1e-7, return_V=True)
    r2 = gpu_ctx.vga_step_host(r1["M"], st["EL"], st["EF"], st["sigma2"], "by_col", 100, 1e-7, return_V=True)
    assert r1["converged"] and r2["converged"] and r2["n_updates"] == 0
    assert np.array_equal(r2["M"], r1["M"])
    ref = vga_numpy.ebpmf_log_vga_step(r1["M"], sim["Y"], st["EL"], st["EF"], st["sigma2"], "by_col", 100, 1e-7)
    check_step(r2, ref, "by_col")


def test_maxiter_one(gpu_ctx):
    """maxiter = 1: one evaluation; either it breaks (u = 0) or one update is applied and V is stale."""
    sim, st = make_case(700, 90, 3, 0.2)
    gpu_ctx.set_y(sp.csc_matrix(sim["Y"]))
    res = gpu_ctx.vga_step_host(st["M0"], st["EL"], st["EF"], st["sigma2"], "by_col", 1, 1e-5, return_V=True)
    ref = vga_numpy.ebpmf_log_vga_step(st["M0"], sim["Y"], st["EL"], st["EF"], st["sigma2"], "by_col", 1, 1e-5)
    check_step(res, ref, "by_col")
    assert res["n_updates"] == 1 and not res["converged"]


@pytest.mark.parametrize("K,freeze", [(31, 1), (45, 1), (45, 0), (70, 1), (70, 0)])
def test_many_factors(pkg, K, freeze):
    """K_tot = K + 2 above the 32-column register tile: Beta is accumulated over chunks of 32 factor columns
    (flash_control$Kmax defaults to 30, R/ebpmf_log_controls.R:74, but greedy additions can exceed it)."""
    c = simdata.offset_for_density(0.10, K=4)
    sim = simdata.sim_data_log(700, 150, K=4, seed=5, log_offset=c)
    rng = np.random.default_rng(K)
    st = simdata.solver_state(sim, start="cold")
    EL = np.hstack([st["EL"], 0.05 * rng.normal(size=(700, K - 4))])
    EF = np.hstack([st["EF"], 0.05 * rng.normal(size=(150, K - 4))])
    ctx = pkg.VgaContext(0)
    try:
        ctx.set_option("freeze", freeze)
        ctx.set_y(sp.csc_matrix(sim["Y"]))
        ref = vga_numpy.ebpmf_log_vga_step(st["M0"], sim["Y"], EL, EF, st["sigma2"], "by_col", 100, 1e-8)
        for call in range(2):                      # the second call takes the tiles in the first call's hardness order
            res = ctx.vga_step_host(st["M0"], EL, EF, st["sigma2"], "by_col", 100, 1e-8, return_V=True)
            check_step(res, ref, "by_col")
    finally:
        ctx.close()


def test_fast_math_vs_exact_math(pkg, tmp_path):
    """The fast reciprocal / exp / log of vga_math.cuh against the same library built with
    EBPMF_EXACT_MATH=1 (IEEE division, libdevice exp/log): same update count, results within 1e-12."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exact = os.path.join(os.path.dirname(pkg.LIB_PATH), "libebpmf_b200_exactmath.so")
    assert os.path.exists(exact), "build with make -C ebpmf.alpha_b200/csrc (or __graft_entry__.build())"
    outs = []
    for tag, lib in (("fast", pkg.LIB_PATH), ("exact", exact)):
        o = str(tmp_path / (tag + ".npz"))
        env = dict(os.environ, EBPMF_B200_LIB=lib)
        subprocess.run([sys.executable, os.path.join(root, "scripts", "run_case_dump.py"), o], check=True, env=env)
        outs.append(np.load(o))
    a, b = outs
    assert a["s"][3] == b["s"][3]
    assert np.max(np.abs(a["M"] - b["M"])) < 1e-12 and rel(a["V"], b["V"]) < 1e-12
    assert rel(a["v_sum"], b["v_sum"]) < 1e-12 and rel(a["s"][:3], b["s"][:3]) < 1e-12


def test_resident_m_loop_equals_full_roundtrip(gpu_ctx):
    """ebpmf_log's loop with M resident (set_m once, then factors/sigma2 in, M out) gives exactly what
    passing M in and out every iteration gives: three outer iterations with changing factors."""
    sim, st = make_case(1100, 190, 4, 0.10)
    rng = np.random.default_rng(11)
    gpu_ctx.set_y(sp.csc_matrix(sim["Y"]))
    facs = [(st["EL"] + 0.02 * t * rng.normal(size=st["EL"].shape), st["EF"] + 0.02 * t * rng.normal(size=st["EF"].shape),
             st["sigma2"] * (1 + 0.1 * t)) for t in range(3)]
    M = st["M0"]; full = []
    for EL, EF, s2 in facs:
        r = gpu_ctx.vga_step_host(M, EL, EF, s2, "by_col", 10, 1e-5)
        M = r["M"]; full.append(r)
    gpu_ctx.set_m(st["M0"])
    for (EL, EF, s2), rf in zip(facs, full):
        r = gpu_ctx.vga_step_resident_host(EL, EF, s2, "by_col", 10, 1e-5)
        assert r["n_updates"] == rf["n_updates"] and np.array_equal(r["M"], rf["M"])
        assert np.array_equal(r["v_sum"], rf["v_sum"]) and r["sym"] == rf["sym"]


@pytest.mark.parametrize("var_type", ["by_col", "by_row"])
def test_freeze_and_history_edge_cases(pkg, var_type):
    """Periodic fast-forward + history order against plain iteration in natural order: u = 0 (R breaks at
    i = 1, M_0 final), maxiter = 1, maxiter = 2 (exhaustion: the last update is performed by the top-up pass
    for every unit), an exhausted tight tolerance, NaN; always bit-identical, always the oracle's count."""
    sim, st = make_case(1500, 200, 4, 0.10, var_type=var_type)
    a, b = pkg.VgaContext(0), pkg.VgaContext(0)
    try:
        a.set_option("freeze", 0); a.set_option("history", 0)
        for c in (a, b):
            c.set_y(sp.csc_matrix(sim["Y"]))
        ra = a.vga_step_host(st["M0"], st["EL"], st["EF"], st["sigma2"], var_type, 100, 1e-7, return_V=True)
        rb = b.vga_step_host(st["M0"], st["EL"], st["EF"], st["sigma2"], var_type, 100, 1e-7, return_V=True)
        assert ra["n_updates"] == rb["n_updates"] and np.array_equal(ra["M"], rb["M"]) and np.array_equal(ra["V"], rb["V"])
        for mi, tol, M0 in ((100, 1e-7, ra["M"]), (1, 1e-5, st["M0"]), (2, 1e-5, st["M0"]), (3, 1e-12, st["M0"]),
                            (7, 1e-13, st["M0"]), (40, 1e-13, st["M0"])):
            xa = a.vga_step_host(M0, st["EL"], st["EF"], st["sigma2"], var_type, mi, tol, return_V=True)
            xb = b.vga_step_host(M0, st["EL"], st["EF"], st["sigma2"], var_type, mi, tol, return_V=True)
            ref = vga_numpy.ebpmf_log_vga_step(M0, sim["Y"], st["EL"], st["EF"], st["sigma2"], var_type, mi, tol)
            check_step(xb, ref, var_type)
            assert xa["n_updates"] == xb["n_updates"] and xa["converged"] == xb["converged"]
            assert np.array_equal(xa["M"], xb["M"]) and np.array_equal(xa["V"], xb["V"])
            assert np.array_equal(np.atleast_1d(xa["v_sum"]), np.atleast_1d(xb["v_sum"]))
            assert (xa["sym"], xa["ssexp"], xa["slogv"]) == (xb["sym"], xb["ssexp"], xb["slogv"])
        EL = st["EL"].copy(); EL[5, 1] = np.nan
        with pytest.raises(pkg.EbpmfError) as ei:
            b.vga_step_host(st["M0"], EL, st["EF"], st["sigma2"], var_type, 10, 1e-5)
        assert ei.value.code == 3
    finally:
        a.close(); b.close()


def test_config5_variants_medium_size(gpu_ctx, pkg):
    """BASELINE config 5 (fixed_iter vs low_memory variants, K = 10) at 20 000 x 2 000: tolerance and
    iteration-count parity against the C transcription of the oracle."""
    c = simdata.offset_for_density(0.05, K=10)
    sim = simdata.sim_data_log(20000, 2000, K=10, seed=12345, log_offset=c)
    st = simdata.solver_state(sim, start="cold")
    n, p = sim["Y"].shape
    Ys = sp.csc_matrix(sim["Y"])
    Beta = st["EL"] @ st["EF"].T
    S2 = vga_numpy.adjust_var_shape(st["sigma2"], "by_col", n, p)
    for mi in (1, 10):        # maxiter = 1 is _fixed_iter's only call site (inst/code/Poisson_MF_splitting.R:212-214)
        g = pkg.vga_pois_solver_mat_newton_fixed_iter(st["M0"], None, sim["Y"], 1, Beta, S2, mi, ctx=gpu_ctx)
        r = c_oracle.vga_fixed_iter(st["M0"], None, sim["Y"], 1, Beta, S2, mi)
        assert relM(g["M"], r["M"]) < RTOL and rel(g["V"], r["V"]) < RTOL
    l0 = sim["L0"] * np.exp(sim["log_offset"]); f0 = sim["F0"]
    EL, EF = st["EL"][:, 2:], st["EF"][:, 2:]
    M0 = np.log(sim["Y"] + 0.5) - np.log(l0)[:, None] - np.log(f0)[None, :]
    for vt, s2 in (("by_col", st["sigma2"]), ("by_row", np.random.default_rng(3).uniform(0.05, 1, n)), ("constant", np.array([0.4]))):
        gi = {}
        g = pkg.vga_pois_solver_mat_newton_low_memory(M0, Ys, l0, f0, EL, EF, s2, vt, 1000, 1e-5, info=gi, ctx=gpu_ctx)
        r = c_oracle.vga_low_memory(M0, sim["Y"], l0, f0, EL, EF, s2, vt, 1000, 1e-5)
        assert gi["n_updates"] == r["n_updates"] and gi["converged"] == r["converged"]
        assert relM(g, r["M"]) < RTOL
    # a1 on the same data: same root as a10 (l0, f0 as S), different update count (F8)
    i1 = {}
    pkg.vga_pois_solver_mat_newton(np.minimum(M0, EL @ EF.T), Ys, np.outer(l0, f0), EL @ EF.T, S2, 1000, 1e-5,
                                   return_V=False, info=i1, ctx=gpu_ctx)
    r1 = c_oracle.vga_newton(np.minimum(M0, EL @ EF.T), sim["Y"], np.outer(l0, f0), EL @ EF.T, S2, 1000, 1e-5)
    assert i1["n_updates"] == r1["n_updates"]


def test_randomised_configurations(pkg):
    """40 random small configurations (shape, K, density, var_type, maxiter, tol, start, freeze, history, delta):
    every one must match the oracle to 1e-8 with the same update count."""
    rng = np.random.default_rng(20261019)
    ctx = pkg.VgaContext(0)
    try:
        for trial in range(40):
            n = int(rng.integers(1, 700)); p = int(rng.integers(1, 90)); K = int(rng.integers(1, 41))
            vt = ["by_col", "by_row", "constant"][int(rng.integers(3))]
            maxiter = int(rng.choice([1, 2, 3, 5, 10, 50, 1000])); tol = float(rng.choice([1e-3, 1e-5, 1e-8, 1e-10]))
            dens = float(rng.choice([0.0, 0.02, 0.1, 0.5, 0.9]))
            sim = simdata.sim_data_log(n, p, K=K, seed=1000 + trial, pi0=0.9,
                                       log_offset=(simdata.offset_for_density(dens, K=K, pi0=0.9) if dens > 0 else -40.0))
            st = simdata.solver_state(sim, start=["cold", "warm"][int(rng.integers(2))], var_type=vt, seed=trial, noise=0.1)
            ctx.set_option("freeze", int(rng.integers(2)))
            ctx.set_option("history", int(rng.integers(2)))
            ctx.set_option("delta", float(rng.choice([0.0, 0.0, 1e-14])))
            ctx.set_y(sp.csc_matrix(sim["Y"]) if rng.integers(2) else sim["Y"])
            res = ctx.vga_step_host(st["M0"], st["EL"], st["EF"], st["sigma2"], vt, maxiter, tol, return_V=True)
            ref = vga_numpy.ebpmf_log_vga_step(st["M0"], sim["Y"], st["EL"], st["EF"], st["sigma2"], vt, maxiter, tol)
            try:
                check_step(res, ref, vt)
            except AssertionError as e:
                raise AssertionError("trial %d: n=%d p=%d K=%d %s maxiter=%d tol=%g dens=%g: %s"
                                     % (trial, n, p, K, vt, maxiter, tol, dens, e))
    finally:
        ctx.close()


@pytest.mark.parametrize("freeze", [0, 1])
def test_retry_loop_when_verification_fails(pkg, freeze):
    """The host's K*+1 retry loop (pass 2 repeated until every unit is below tol) is exercised with a test
    hook: pass 1 finds K* for tol, pass 2 verifies against tol*1e-8, fails, and the loop must walk up to the
    count R would reach with tol*1e-8 -- results equal the oracle's for that tolerance."""
    sim, st = make_case(1800, 220, 4, 0.10)
    ctx = pkg.VgaContext(0)
    try:
        ctx.set_option("freeze", freeze)
        ctx.set_option("debug_verify_scale", 1e-8)
        ctx.set_y(sp.csc_matrix(sim["Y"]))
        res = ctx.vga_step_host(st["M0"], st["EL"], st["EF"], st["sigma2"], "by_col", 50, 1e-2, return_V=True)
        loose = vga_numpy.ebpmf_log_vga_step(st["M0"], sim["Y"], st["EL"], st["EF"], st["sigma2"], "by_col", 50, 1e-2)
        ref = vga_numpy.ebpmf_log_vga_step(st["M0"], sim["Y"], st["EL"], st["EF"], st["sigma2"], "by_col", 50, 1e-10)
        assert ref["n_updates"][0] > loose["n_updates"][0]          # the hook really forces extra passes
        assert res["passes"] == 2 + ref["n_updates"][0] - loose["n_updates"][0]
        check_step(res, ref, "by_col")
        # exhaustion reached through the retry loop: maxiter just above K*
        mi = loose["n_updates"][0] + 1
        res = ctx.vga_step_host(st["M0"], st["EL"], st["EF"], st["sigma2"], "by_col", mi, 1e-2, return_V=True)
        ref = vga_numpy.ebpmf_log_vga_step(st["M0"], sim["Y"], st["EL"], st["EF"], st["sigma2"], "by_col", mi, 1e-10)
        check_step(res, ref, "by_col")
        assert not res["converged"] and res["n_updates"] == mi
    finally:
        ctx.close()


def test_history_order_many_tiles(pkg):
    """Several waves of tiles (device-generated 40 000 x 6 000): with the previous step's per-tile counts as the
    order of the first pass (and the periodic fast-forward) the results are bit-identical to plain iteration in
    natural order, and the top-up pass has (almost) nothing left to do -- also when the factors moved in between."""
    n, p, K = 40000, 6000, 10
    a, b = pkg.VgaContext(0), pkg.VgaContext(0)
    try:
        a.set_option("freeze", 0); a.set_option("history", 0)
        for c in (a, b):
            c.sim_data_log(n, p, K, seed=77, log_offset=-4.644)
        units = ((n + 31) // 32) * ((p + 1) // 2)
        for step in range(3):
            ia, ib = a.vga_step(10, 1e-5), b.vga_step(10, 1e-5)
            assert ia.n_updates == ib.n_updates and ia.converged == ib.converged
            assert ia.sym == ib.sym and ia.ssexp == ib.ssexp and ia.slogv == ib.slogv
            assert np.array_equal(a.get_v_sum(), b.get_v_sum())
            for r0 in (0, 17000, n - 999):
                assert np.array_equal(a.get_rows(r0, 999, "M"), b.get_rows(r0, 999, "M"))
            ca, cb = a.counters(), b.counters()
            if step > 0 and ia.n_updates >= 3:
                assert cb["topup_units"] < 0.05 * units, (step, cb)
                assert cb["topup_units"] <= ca["topup_units"]
            if step > 0:
                assert cb["unit_evals"] < ca["unit_evals"]              # the fast-forward skips repeated iterates
            for c in (a, b):
                c.rewind_m()
            if step == 1:                      # move the factors: the kept order is now only a prediction
                EL, EF = a.get_factors()
                rng = np.random.default_rng(5)
                EL[:, 2:] += 0.05 * rng.normal(size=EL[:, 2:].shape); EF[:, 2:] += 0.05 * rng.normal(size=EF[:, 2:].shape)
                for c in (a, b):
                    c.set_factors(EL, EF)
    finally:
        a.close(); b.close()


def test_host_step_chunked_pipeline(pkg):
    """The host-buffer step above the pipelining threshold (6 000 x 1 100 = 53 MB; odd and even row counts, so both
    the bulk-copy and the per-thread slab paths run) with plain pageable NumPy arrays: column chunks, speculative
    copies, repeated copies of the chunks the top-up pass touched -- bit-identical to set_m + vga_step + get_m, with
    and without speculation, M in and resident."""
    for n in (6000, 6001):
        p, K = 1100, 6
        c = simdata.offset_for_density(0.08, K=K)
        sim = simdata.sim_data_log(n, p, K=K, seed=11, log_offset=c)
        st = simdata.solver_state(sim, start="cold")
        plain, piped = pkg.VgaContext(0), pkg.VgaContext(0)
        try:
            plain.set_option("freeze", 0); plain.set_option("history", 0)
            piped.set_option("pipe_chunks", 7)
            for cx in (plain, piped):
                cx.set_y(sp.csc_matrix(sim["Y"]))
            plain.set_factors(st["EL"], st["EF"]); plain.set_sigma2(st["sigma2"], "by_col")
            M = st["M0"]
            piped.set_m(M)
            for it in range(4):
                plain.set_m(M)
                info = plain.vga_step(10, 1e-5, want_v=True)
                Mref, Vref, vref = plain.get_m(), plain.get_v(), plain.get_v_sum()
                piped.set_option("speculate", 0 if it == 3 else 1)
                if it % 2 == 0:
                    r = piped.vga_step_host(M, st["EL"], st["EF"], st["sigma2"], "by_col", 10, 1e-5, return_V=(it == 2))
                else:
                    r = piped.vga_step_resident_host(st["EL"], st["EF"], st["sigma2"], "by_col", 10, 1e-5)
                assert r["n_updates"] == info.n_updates and r["converged"] == bool(info.converged)
                assert np.array_equal(r["M"], Mref), (n, it)
                assert np.array_equal(r["v_sum"], vref) and r["sym"] == info.sym and r["ssexp"] == info.ssexp
                if r["V"] is not None:
                    assert np.array_equal(r["V"], Vref)
                cnt = piped.counters()
                if it in (1, 2):
                    assert cnt["spec_bytes"] == 8.0 * n * p         # history exists: everything went out speculatively
                    assert cnt["recopy_bytes"] < cnt["spec_bytes"]
                M = Mref
        finally:
            plain.close(); piped.close()
    with pytest.raises(pkg.EbpmfError) as ei:        # host matrices must have the shape of the resident Y
        ctx = pkg.VgaContext(0)
        ctx.set_y(sp.csc_matrix(sim["Y"]))
        ctx.vga_step_host(st["M0"][:100], st["EL"][:100], st["EF"], st["sigma2"], "by_col", 10, 1e-5)
    assert ei.value.code == 1
    ctx.close()


@pytest.mark.parametrize("n,p,Kw", [(3000, 700, 5), (4097, 333, 22), (2048, 64, 32), (1500, 90, 41)])
def test_m_products_vs_numpy(pkg, n, p, Kw):
    """N1: M %*% EF, t(M) %*% EL and the column / row sums of (M - EL EF')^2 on the resident M against NumPy,
    to 1e-10 relative (even and odd row counts: bulk-copy and per-thread tile paths; Kw > 32: two passes)."""
    rng = np.random.default_rng(n + Kw)
    sim, st = make_case(n, p, 3, 0.10)
    M = np.asfortranarray(st["M0"] + 0.1 * rng.normal(size=(n, p)))
    EL = np.asfortranarray(rng.normal(size=(n, Kw))); EF = np.asfortranarray(rng.normal(size=(p, Kw)) * 0.3)
    ctx = pkg.VgaContext(0)
    try:
        ctx.set_y(sp.csc_matrix(sim["Y"])); ctx.set_m(M)
        r = ctx.m_products(EL, EF)
        R = M - EL @ EF.T
        for got, want in ((r["MF"], M @ EF), (r["MtL"], M.T @ EL), (r["r2_col"], np.sum(R * R, axis=0)),
                          (r["r2_row"], np.sum(R * R, axis=1))):
            assert np.max(np.abs(got - want)) <= 1e-10 * max(1.0, np.max(np.abs(want)))
        assert r["kernel_ms"] > 0
        r2 = ctx.m_products(EL, EF)                     # run-to-run bit-reproducible
        assert np.array_equal(r["MF"], r2["MF"]) and np.array_equal(r["MtL"], r2["MtL"]) and np.array_equal(r["r2_col"], r2["r2_col"])
    finally:
        ctx.close()


def test_malformed_dgcmatrix_is_refused(pkg):
    """@p not non-decreasing, a row index out of range or unsorted within a column: EBPMF_ERR_ARG, not device memory
    corruption (the kernels scatter through these indices)."""
    ctx = pkg.VgaContext(0)
    try:
        cp = np.array([0, 2, 4], dtype=np.int32); x = np.ones(4)
        for bad_cp, bad_ri in ((np.array([0, 3, 2], dtype=np.int32), np.array([0, 1, 0, 1], dtype=np.int32)),
                               (cp, np.array([0, 5, 0, 1], dtype=np.int32)),
                               (cp, np.array([1, 0, 0, 1], dtype=np.int32)),
                               (cp, np.array([0, -1, 0, 1], dtype=np.int32))):
            with pytest.raises(pkg.EbpmfError) as ei:
                ctx.set_y_csc(3, 2, bad_cp, bad_ri, x)
            assert ei.value.code == 1
        ctx.set_y_csc(3, 2, cp, np.array([0, 2, 1, 2], dtype=np.int32), x)
    finally:
        ctx.close()


def test_low_memory_does_not_leave_foreign_state(pkg):
    """vga_pois_solver_mat_newton_low_memory takes over the context's resident Y / M: afterwards a resident step on
    the same context must fail with EBPMF_ERR_STATE instead of running on the wrong matrices."""
    sim, st = make_case(400, 60, 3, 0.10)
    ctx = pkg.VgaContext(0)
    try:
        ctx.set_y(sp.csc_matrix(sim["Y"])); ctx.set_m(st["M0"]); ctx.set_factors(st["EL"], st["EF"])
        ctx.set_sigma2(st["sigma2"], "by_col")
        ctx.vga_step(10, 1e-5)
        pkg.vga_pois_solver_mat_newton_low_memory(st["M0"][:100, :20], sim["Y"][:100, :20], np.ones(100), np.ones(20),
                                                  st["EL"][:100], st["EF"][:20], st["sigma2"][:20], "by_col", 50, 1e-6, ctx=ctx)
        with pytest.raises(pkg.EbpmfError) as ei:
            ctx.vga_step(10, 1e-5)
        assert ei.value.code == 5
    finally:
        ctx.close()


# ---------------------------------------------------------------------------------------
# batch_size-faithful host-streamed step (R/ebpmf_log.R:195-203, :254-300)
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("var_type", ["by_col", "by_row", "constant"])
@pytest.mark.parametrize("nslots", [1, 3])
def test_batched_step_vs_oracle_batched_branch(pkg, var_type, nslots):
    """Per-batch stop scope: every batch gets ITS OWN update count; sums add up in batch order; by_row v_sum
    concatenates.  Compared with the oracle's restatement of the batched branch, batches cut as R cuts them."""
    sim, st = make_case(1500, 211, 4, 0.10, var_type=var_type)
    n = 1500
    h = pkg.VgaBatched(0, nslots)
    try:
        h.set_y(sp.csc_matrix(sim["Y"]), batch_size=400)
        assert list(h.batch_row0) == [0, 375, 750, 1125, 1500]
        batches = vga_numpy.r_cut_batches(n, 4)
        for maxiter, tol in ((10, 1e-5), (100, 1e-9), (2, 1e-5)):
            res = h.vga_step_host(st["M0"], st["EL"], st["EF"], st["sigma2"], var_type, maxiter, tol, return_V=True)
            ref = vga_numpy.ebpmf_log_vga_step(st["M0"], sim["Y"], st["EL"], st["EF"], st["sigma2"], var_type,
                                               maxiter, tol, batches=batches)
            assert res["n_updates"] == list(ref["n_updates"]) and res["converged"] == [bool(c) for c in ref["converged"]]
            assert relM(res["M"], ref["M"]) < RTOL and rel(res["V"], ref["V"]) < RTOL
            assert rel(res["v_sum"], ref["v_sum"]) < RTOL
            for k in ("sym", "ssexp", "slogv"):
                assert abs(res[k] - ref[k]) <= RTOL * max(abs(ref[k]), 1.0), (k, res[k], ref[k])
        assert abs(h.sum_lfactorial() - vga_numpy.sum_lfactorial_sparseMat(sp.csc_matrix(sim["Y"]))) < 1e-8 * 1e4
    finally:
        h.close()


def test_batched_with_a_one_row_batch(pkg):
    """n = 5 in 3 batches is cut 2/1/2.  The reference itself stops there with "non-conformable arrays" (a one-row
    batch drops to a vector, tests/test_minir_reference.py); the CUDA path solves it like any other batch, as the
    oracle does."""
    sim = simdata.sim_data_log(5, 7, K=2, seed=5, log_offset=-1.0)
    st = simdata.solver_state(sim, start="cold", var_type="by_col", seed=6)
    batches = vga_numpy.r_cut_batches(5, 3)
    h = pkg.VgaBatched(0, 2)
    try:
        h.set_y(sp.csc_matrix(sim["Y"]), batch_size=2)
        assert list(h.batch_row0) == [0, 2, 3, 5]
        res = h.vga_step_host(st["M0"], st["EL"], st["EF"], st["sigma2"], "by_col", 10, 1e-5, return_V=True)
    finally:
        h.close()
    ref = vga_numpy.ebpmf_log_vga_step(st["M0"], sim["Y"], st["EL"], st["EF"], st["sigma2"], "by_col", 10, 1e-5, batches=batches)
    assert res["n_updates"] == list(ref["n_updates"])
    assert relM(res["M"], ref["M"]) < RTOL and rel(res["V"], ref["V"]) < RTOL and rel(res["v_sum"], ref["v_sum"]) < RTOL


def test_batched_equals_one_context_per_batch(pkg):
    """The batched handle against the plain context run batch by batch on row slices (bit-identical), M_out
    aliasing M_in, an uneven last batch, and a batch count that is not a multiple of the slot count."""
    sim, st = make_case(1203, 97, 3, 0.15)
    Y = sp.csc_matrix(sim["Y"])
    row0 = np.array([0, 256, 300, 811, 812, 1203], dtype=np.int64)
    h = pkg.VgaBatched(0, 2)
    ctx = pkg.VgaContext(0)
    try:
        h.set_y(Y, batch_row0=row0)
        M = np.array(st["M0"], order="F")
        res = h.vga_step_host(M, st["EL"], st["EF"], st["sigma2"], "by_col", 10, 1e-5, M_out=M)
        assert res["M"] is M
        vs = np.zeros(97)
        for b in range(len(row0) - 1):
            lo, hi = int(row0[b]), int(row0[b + 1])
            ctx.set_y(sp.csc_matrix(Y[lo:hi, :]))
            r = ctx.vga_step_host(st["M0"][lo:hi], st["EL"][lo:hi], st["EF"], st["sigma2"], "by_col", 10, 1e-5)
            assert r["n_updates"] == res["n_updates"][b]
            assert np.array_equal(r["M"], M[lo:hi])
            vs += r["v_sum"]
        assert np.array_equal(vs, res["v_sum"])
        # NaN in one batch -> error naming the batch
        EL = st["EL"].copy(); EL[900, 1] = np.nan
        with pytest.raises(pkg.EbpmfError) as ei:
            h.vga_step_host(st["M0"], EL, st["EF"], st["sigma2"], "by_col", 10, 1e-5)
        assert ei.value.code == 3 and "batch 4" in str(ei.value)
    finally:
        h.close(); ctx.close()


def test_batched_over_two_gpus_equals_one(pkg):
    """Batches dealt over two GPUs (no collective: batches are independent) give bit-identical results to one
    GPU.  Skipped on single-GPU boxes; `scripts/mgpu_batched_check.py` is the same check as a script."""
    import ctypes
    cnt = ctypes.c_int()
    pkg.load().ebpmf_device_count(ctypes.byref(cnt))
    if cnt.value < 2:
        pytest.skip("needs 2 GPUs")
    sim, st = make_case(2100, 150, 4, 0.10)
    a, b = pkg.VgaBatched(0, 2), pkg.VgaBatched([0, 1], 2)
    try:
        for h in (a, b):
            h.set_y(sp.csc_matrix(sim["Y"]), batch_size=300)
        ra = a.vga_step_host(st["M0"], st["EL"], st["EF"], st["sigma2"], "by_col", 10, 1e-5, return_V=True)
        rb = b.vga_step_host(st["M0"], st["EL"], st["EF"], st["sigma2"], "by_col", 10, 1e-5, return_V=True)
        assert ra["n_updates"] == rb["n_updates"] and np.array_equal(ra["M"], rb["M"]) and np.array_equal(ra["V"], rb["V"])
        assert np.array_equal(ra["v_sum"], rb["v_sum"]) and ra["sym"] == rb["sym"] and ra["slogv"] == rb["slogv"]
    finally:
        a.close(); b.close()


def test_plain_c_demo_runs(pkg, tmp_path):
    """examples/vga_step_demo.c built with gcc against the shared library: a non-Python, non-R host."""
    import os
    import re
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "vga_step_demo")
    libdir = os.path.dirname(pkg.LIB_PATH)
    r = subprocess.run(["/usr/bin/gcc", "-std=c99", "-I", os.path.join(root, "include"),
                        os.path.join(root, "examples", "vga_step_demo.c"), "-L", libdir, "-lebpmf_b200",
                        "-Wl,-rpath," + libdir, "-lm", "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    # the same data replayed through the NumPy oracle (same xorshift stream): u = 5, sym = -9672.454278
    assert "5 Newton updates (converged 1)" in r.stdout, r.stdout
    sym = float(re.search(r"sym (-?[0-9.]+)", r.stdout).group(1))
    assert abs(sym - (-9672.454278)) < 1e-4


@pytest.mark.gpu
def test_calc_obj_takes_every_sum_at_its_own_length(gpu_ctx, pkg):
    """R/ebpmf_log.R:407: each sum() runs over its own operands.  With var_type 'constant' sigma2 is a scalar
    while R2 (fit_flash$flash_fit$R2) and sv may be vectors; the sigma2-only terms must be counted once."""
    rng = np.random.default_rng(3)
    n, p = 40, 17
    for ls2, lsv, lr2 in [(1, 1, p), (1, p, p), (p, p, p), (p, 1, 1), (n, n, 1)]:
        s2 = rng.uniform(0.1, 2.0, ls2); sv = rng.uniform(0.0, 5.0, lsv); R2 = rng.uniform(0.0, 30.0, lr2)
        ref = vga_numpy.calc_ebpmf_log_obj(n, p, 1.5, 2.5, -3.5, sv, s2, R2, -7.0, -11.0, float(n), 1.0, 1.0)
        got = pkg.calc_ebpmf_log_obj(n, p, 1.5, 2.5, -3.5, sv, s2, R2, -7.0, -11.0, float(n), 1.0, 1.0, ctx=gpu_ctx)
        assert abs(got - ref) <= 1e-12 * max(1.0, abs(ref)), (ls2, lsv, lr2, got, ref)


@pytest.mark.gpu
def test_step_input_stays_readable_after_the_step(gpu_ctx):
    """get_rows / get_block with which = 'M_in': a step never writes its input, so the matrix it started from can
    be read back next to the result (what the slab-scale parity check of bench.py uses)."""
    sim, st = make_case(700, 90, 3, 0.1, seed=21)
    gpu_ctx.set_y(sp.csc_matrix(sim["Y"]))
    gpu_ctx.set_m(st["M0"]); gpu_ctx.set_factors(st["EL"], st["EF"]); gpu_ctx.set_sigma2(st["sigma2"], "by_col")
    with pytest.raises(Exception):
        gpu_ctx.get_rows(0, 10, "M_in")                 # no step yet
    gpu_ctx.vga_step(10, 1e-5, want_v=True)
    assert np.array_equal(gpu_ctx.get_rows(100, 300, "M_in"), st["M0"][100:400])
    assert np.array_equal(gpu_ctx.get_block(5, 600, 7, 11, "M_in"), st["M0"][5:605, 7:18])
    assert np.array_equal(gpu_ctx.get_block(5, 600, 7, 11, "M"), gpu_ctx.get_m()[5:605, 7:18])
    assert np.array_equal(gpu_ctx.get_block(0, 700, 89, 1, "V"), gpu_ctx.get_v()[:, 89:90])
    gpu_ctx.rewind_m()
    assert np.array_equal(gpu_ctx.get_m(), st["M0"])


@pytest.mark.gpu
@pytest.mark.parametrize("var_type", ["by_col", "by_row", "constant"])
@pytest.mark.parametrize("s_kind", ["scalar", "matrix"])
def test_calc_split_obj_unfused_elbo(gpu_ctx, pkg, var_type, s_kind):
    """N4: calc_split_PMF_obj_flashier (inst/code/Poisson_MF_splitting.R:373-391) against the oracle's restatement, on
    the output of a real VGA step (ragged shape: 700 x 90), scalar S and the matrix S = tcrossprod(l0, f0) of the
    archived driver."""
    sim, st = make_case(700, 90, 3, 0.15, var_type=var_type, seed=33)
    n, p = sim["Y"].shape
    ref = vga_numpy.ebpmf_log_vga_step(st["M0"], sim["Y"], st["EL"], st["EF"], st["sigma2"], var_type, 10, 1e-5)
    S = 1.0 if s_kind == "scalar" else np.asfortranarray(np.outer(np.exp(0.1 * np.arange(n) / n), np.exp(-0.1 * np.arange(p) / p)))
    want = vga_numpy.calc_split_PMF_obj_flashier(sim["Y"], S, st["sigma2"], ref["M"], ref["V"], st["R2"], -12.5, -1234.5, var_type)
    got = pkg.calc_split_PMF_obj_flashier(sim["Y"], S, st["sigma2"], ref["M"], ref["V"], dict(R2=st["R2"]), -12.5, -1234.5,
                                          var_type, ctx=gpu_ctx)
    assert abs(got - want) <= 1e-10 * abs(want), (got, want)
    with pytest.raises(pkg.EbpmfError):
        pkg.calc_split_PMF_obj_flashier(sim["Y"], S, st["sigma2"], ref["M"], ref["V"], dict(R2=st["R2"]), 0.0, 0.0, "by_both", ctx=gpu_ctx)
