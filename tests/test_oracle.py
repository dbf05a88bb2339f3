"""CPU tests of the oracle itself (the reference ships no golden vectors: "parity unpinned").
What pins it: two independent transcriptions agree, mathematical invariants of
the solve, and an mpmath high-precision root."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import c_oracle, simdata, vga_numpy


def rel(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b)) / np.maximum(np.abs(b), 1e-300)))


@pytest.fixture(scope="module")
def case():
    sim = simdata.sim_data_log(400, 90, K=4, seed=12345, log_offset=simdata.offset_for_density(0.12, K=4))
    st = simdata.solver_state(sim, start="cold")
    Beta = st["EL"] @ st["EF"].T
    Sigma2 = vga_numpy.adjust_var_shape(st["sigma2"], "by_col", 400, 90)
    return sim, st, Beta, Sigma2


@pytest.mark.parametrize("maxiter,tol", [(10, 1e-5), (1000, 1e-8), (3, 1e-5), (1, 1e-5)])
def test_numpy_and_c_transcriptions_agree(case, maxiter, tol):
    sim, st, Beta, Sigma2 = case
    i1 = {}
    a = vga_numpy.vga_pois_solver_mat_newton(st["M0"], sim["Y"], 1, Beta, Sigma2, maxiter, tol, info=i1)
    b = c_oracle.vga_newton(st["M0"], sim["Y"], 1, Beta, Sigma2, maxiter, tol)
    c = c_oracle.vga_newton_rstyle(st["M0"], sim["Y"], 1, Beta, Sigma2, maxiter, tol)
    assert i1["n_updates"] == b["n_updates"] == c["n_updates"]
    assert i1["converged"] == b["converged"] == c["converged"]
    assert np.max(np.abs(a["M"] - b["M"])) < 1e-13 and rel(a["V"], b["V"]) < 1e-13
    assert np.array_equal(b["M"], c["M"]) and np.array_equal(b["V"], c["V"])     # same libm, same order


def test_recycling_of_S_and_Sigma2(case):
    sim, st, Beta, _ = case
    n, p = sim["Y"].shape
    rng = np.random.default_rng(1)
    s2row = rng.uniform(0.1, 1, n)
    S = np.exp(rng.normal(0, 0.1, (n, p)))
    a = vga_numpy.vga_pois_solver_mat_newton(st["M0"], sim["Y"], S, Beta, s2row, 50, 1e-8)
    b = vga_numpy.vga_pois_solver_mat_newton(st["M0"], sim["Y"], S, Beta, np.repeat(s2row[:, None], p, 1), 50, 1e-8)
    c = c_oracle.vga_newton(st["M0"], sim["Y"], S, Beta, s2row, 50, 1e-8)
    assert np.array_equal(a["M"], b["M"]) and np.max(np.abs(a["M"] - c["M"])) < 1e-13
    # sparse X is densified, as R's arithmetic would
    d = vga_numpy.vga_pois_solver_mat_newton(st["M0"], sp.csc_matrix(sim["Y"]), S, Beta, s2row, 50, 1e-8)
    assert np.array_equal(a["M"], d["M"])


def test_break_invariants(case):
    """On break: max|f| < tol and V * (Sigma2*X + Beta + 1 - M) == Sigma2 (fresh temp)."""
    sim, st, Beta, Sigma2 = case
    info = {}
    r = vga_numpy.vga_pois_solver_mat_newton(st["M0"], sim["Y"], 1, Beta, Sigma2, 1000, 1e-8, info=info)
    assert info["converged"] and info["maxabs_f"][-1] < 1e-8 and all(v >= 1e-8 for v in info["maxabs_f"][:-1])
    temp = Sigma2 * sim["Y"] + Beta + 1 - r["M"]
    assert rel(r["V"] * temp, Sigma2) < 1e-14
    f = sim["Y"] - np.exp(r["M"] + r["V"] / 2) - (r["M"] - Beta) / Sigma2
    assert np.max(np.abs(f)) < 1e-8
    assert np.all(r["V"] > 0) and np.all(r["V"] <= Sigma2 + 1e-15)       # clamp keeps temp >= 1


def test_exhaustion_returns_stale_V(case):
    """F6: without a break, V is Sigma2/temp of the PREVIOUS iterate."""
    sim, st, Beta, Sigma2 = case
    r3 = vga_numpy.vga_pois_solver_mat_newton(st["M0"], sim["Y"], 1, Beta, Sigma2, 3, 1e-12)
    r2 = vga_numpy.vga_pois_solver_mat_newton(st["M0"], sim["Y"], 1, Beta, Sigma2, 2, 1e-12)
    const0 = Sigma2 * sim["Y"] + Beta + 1
    stale = Sigma2 / (const0 - r2["M"])
    fresh = Sigma2 / (const0 - r3["M"])
    assert np.array_equal(r3["V"], stale)
    assert rel(fresh, stale) > 1e-6


def test_global_stop_differs_from_per_entry_stop(case):
    """F5: every entry gets the same number of updates; per-entry stopping gives another answer."""
    sim, st, Beta, Sigma2 = case
    info = {}
    g = vga_numpy.vga_pois_solver_mat_newton(st["M0"], sim["Y"], 1, Beta, Sigma2, 10, 1e-5, info=info)
    col = 5
    i2 = {}
    one = vga_numpy.vga_pois_solver_mat_newton(st["M0"][:, [col]], sim["Y"][:, [col]], 1, Beta[:, [col]],
                                               Sigma2[:, [col]], 10, 1e-5, info=i2)
    assert i2["n_updates"] <= info["n_updates"]
    forced = vga_numpy.vga_pois_solver_mat_newton(st["M0"][:, [col]], sim["Y"][:, [col]], 1, Beta[:, [col]],
                                                  Sigma2[:, [col]], info["n_updates"], 0.0)
    assert np.array_equal(forced["M"], g["M"][:, [col]])                # tol = 0 forces exactly u updates


def test_mpmath_root(case):
    import mpmath as mp
    mp.mp.dps = 40
    sim, st, Beta, Sigma2 = case
    r = vga_numpy.vga_pois_solver_mat_newton(st["M0"], sim["Y"], 1, Beta, Sigma2, 1000, 1e-10)
    rng = np.random.default_rng(0)
    for _ in range(12):
        i, j = int(rng.integers(400)), int(rng.integers(90))
        x, b, s2 = mp.mpf(sim["Y"][i, j]), mp.mpf(Beta[i, j]), mp.mpf(Sigma2[i, j])
        f = lambda m: x - mp.e ** (m + (s2 / 2) / (s2 * x + b + 1 - m)) - (m - b) / s2
        root = mp.findroot(f, mp.mpf(r["M"][i, j]))
        assert abs(float(root) - r["M"][i, j]) < 1e-9 * max(1.0, abs(float(root)))
        v = s2 / (s2 * x + b + 1 - root)
        assert abs(float(v) - r["V"][i, j]) < 1e-8 * float(v)


def test_fixed_iter_reaches_same_fixed_point(case):
    sim, st, Beta, Sigma2 = case
    a = vga_numpy.vga_pois_solver_mat_newton(st["M0"], sim["Y"], 1, Beta, Sigma2, 1000, 1e-11)
    b = vga_numpy.vga_pois_solver_mat_newton_fixed_iter(st["M0"], None, sim["Y"], 1, Beta, Sigma2, 80)
    c = c_oracle.vga_fixed_iter(st["M0"], None, sim["Y"], 1, Beta, Sigma2, 80)
    assert np.max(np.abs(a["M"] - b["M"])) < 1e-8 and rel(b["V"], a["V"]) < 1e-8
    assert np.max(np.abs(b["M"] - c["M"])) < 1e-12 and rel(c["V"], b["V"]) < 1e-12


@pytest.mark.parametrize("var_type", ["by_col", "by_row", "constant"])
def test_low_memory_transcriptions_and_quirks(case, var_type):
    sim, _, _, _ = case
    st = simdata.solver_state(sim, var_type=var_type)
    n, p = sim["Y"].shape
    EL, EF = st["EL"][:, 2:], st["EF"][:, 2:]
    l0, f0 = sim["L0"] * np.exp(sim["log_offset"]), sim["F0"]
    M0 = np.log(sim["Y"] + 0.5) - np.log(l0)[:, None] - np.log(f0)[None, :]
    ia = {}
    a = vga_numpy.vga_pois_solver_mat_newton_low_memory(M0, sim["Y"], l0, f0, EL, EF, st["sigma2"], var_type,
                                                        1000, 1e-8, info=ia)
    b = c_oracle.vga_low_memory(M0, sim["Y"], l0, f0, EL, EF, st["sigma2"], var_type, 1000, 1e-8)
    assert ia["n_updates"] == b["n_updates"] and ia["converged"] and b["converged"]
    assert np.max(np.abs(a - b["M"])) < 1e-12
    # F8: same root as a1 (with S = l0 f0^T), different trajectory
    S = np.outer(l0, f0)
    Sigma2 = vga_numpy.adjust_var_shape(st["sigma2"], var_type, n, p)
    i1 = {}
    r1 = vga_numpy.vga_pois_solver_mat_newton(np.minimum(M0, EL @ EF.T), sim["Y"], S, EL @ EF.T, Sigma2, 1000, 1e-8,
                                              info=i1)
    assert np.max(np.abs(r1["M"] - a)) < 1e-7
    # an unknown var_type falls through both blocks: clamp only (:78,:108)
    z = vga_numpy.vga_pois_solver_mat_newton_low_memory(M0, sim["Y"], l0, f0, EL, EF, st["sigma2"], "nope", 5, 1e-8)
    assert np.array_equal(z, np.minimum(M0, EL @ EF.T))


def test_nan_is_an_error_like_r(case):
    sim, st, Beta, Sigma2 = case
    B = Beta.copy(); B[3, 4] = np.nan
    with pytest.raises(vga_numpy.RError):
        vga_numpy.vga_pois_solver_mat_newton(st["M0"], sim["Y"], 1, B, Sigma2, 10, 1e-5)
    with pytest.raises(c_oracle.OracleNaN):
        c_oracle.vga_newton(st["M0"], sim["Y"], 1, B, Sigma2, 10, 1e-5)
    with pytest.raises(ValueError):
        vga_numpy.vga_pois_solver_mat_newton(st["M0"], sim["Y"], 1, Beta, Sigma2, 0, 1e-5)


@pytest.mark.parametrize("var_type", ["by_col", "by_row", "constant"])
def test_vga_step_partials_and_elbo_recomposition(case, var_type):
    """a4/a5/a7: fused partials recompose to the un-fused ELBO of inst/code/Poisson_MF_splitting.R:389
    (up to the truncated constant 0.9189385 vs log(2*pi)/2, F9)."""
    sim, _, _, _ = case
    st = simdata.solver_state(sim, var_type=var_type)
    n, p = sim["Y"].shape
    r = vga_numpy.ebpmf_log_vga_step(st["M0"], sim["Y"], st["EL"], st["EF"], st["sigma2"], var_type, 10, 1e-5)
    c = c_oracle.vga_step(st["M0"], sim["Y"], st["EL"], st["EF"], st["sigma2"], var_type, 10, 1e-5)
    assert r["n_updates"][0] == c["n_updates"]
    for k in ("sym", "ssexp", "slogv"):
        assert abs(r[k] - c[k]) <= 1e-12 * max(1.0, abs(c[k]))
    assert rel(r["v_sum"], c["v_sum"]) < 1e-12
    ss = {"by_col": n, "by_row": p, "constant": n * p}[var_type]
    KL, const = -12.5, -vga_numpy.sum_lfactorial_sparseMat(sp.csc_matrix(sim["Y"]))
    assert abs(const + vga_numpy._rsum(__import__("scipy.special").special.gammaln(sim["Y"] + 1))) < 1e-6
    sig = vga_numpy.ebpmf_log_update_sigma2(dict(R2=st["R2"]), st["sigma2"], r["v_sum"], var_type, 0, 1, 1, n, p)
    fused = vga_numpy.calc_ebpmf_log_obj(n, p, r["sym"], r["ssexp"], r["slogv"], r["v_sum"], sig, st["R2"], KL,
                                         const, ss, 0.0, 0.0)
    M, V = r["M"], r["V"]
    S2 = np.asarray(sig, dtype=float)
    unfused = (np.sum(sim["Y"] * M - np.exp(M + V / 2) + np.log(2 * np.pi * V) / 2 + 0.5)
               - np.sum(ss * np.log(2 * np.pi * S2) / 2) - np.sum(np.asarray(r["v_sum"]) / 2 / S2)
               - np.sum(st["R2"] / 2 / S2) + const + KL - np.sum(np.log(S2)))     # (a0+1)*log with a0 = 0
    trunc = (0.5 * np.log(2 * np.pi) - vga_numpy.SLOGV_CONST) * n * p
    assert abs(fused - (unfused - trunc)) < 1e-9 * abs(unfused)
    # the oracle's restatement of calc_split_PMF_obj_flashier (:373-391) is that formula without the prior term
    split = vga_numpy.calc_split_PMF_obj_flashier(sim["Y"], 1.0, sig, M, V, st["R2"], KL, const, var_type)
    assert abs(split - (unfused + np.sum(np.log(S2)))) < 1e-10 * abs(unfused)


def test_sigma2_update_formulas():
    rng = np.random.default_rng(2)
    n, p = 50, 20
    v, R2 = rng.uniform(1, 2, p), rng.uniform(0, n, p)
    s = vga_numpy.ebpmf_log_update_sigma2(dict(R2=R2), np.ones(p), v, "by_col", 0, 1.0, 1.0, n, p)
    assert np.allclose(s, ((v + R2) / 2 + 1) / (n / 2 + 2), rtol=1e-15)
    EL, EF = rng.normal(size=(n, 3)), rng.normal(size=(p, 3))
    s0 = rng.uniform(0.1, 1, p)
    capped = vga_numpy.ebpmf_log_update_sigma2(dict(R2=R2, EL=EL, EF=EF), s0, v, "by_col", 2.0, 1.0, 1.0, n, p)
    gate = (EL @ EF.T).max(0) > np.log(2.0) - np.log(np.exp(s0) - 1) - s0 / 2
    assert np.array_equal(capped[~gate], s0[~gate]) and np.allclose(capped[gate], s[gate])
    with pytest.raises(ValueError):
        vga_numpy.ebpmf_log_update_sigma2(dict(R2=R2), s0, v, "diag", 0, 1, 1, n, p)


def test_row_shards_sum_to_whole(case):
    """§8e: with the stop count scoped to the whole matrix, shard partial sums add up."""
    sim, st, Beta, Sigma2 = case
    whole = c_oracle.vga_step(st["M0"], sim["Y"], st["EL"], st["EF"], st["sigma2"], "by_col", 10, 1e-5)
    u = whole["n_updates"]
    acc = dict(sym=0.0, ssexp=0.0, slogv=0.0, v=np.zeros(90))
    for lo, hi in ((0, 150), (150, 400)):
        r = vga_numpy.vga_pois_solver_mat_newton(st["M0"][lo:hi], sim["Y"][lo:hi], 1, Beta[lo:hi], Sigma2[lo:hi], u, 0.0)
        V = Sigma2[lo:hi] / (Sigma2[lo:hi] * sim["Y"][lo:hi] + Beta[lo:hi] + 1 - r["M"])     # fresh temp (break)
        s = vga_numpy.elbo_partials(sim["Y"][lo:hi], r["M"], V)
        acc["sym"] += s[0]; acc["ssexp"] += s[1]; acc["slogv"] += s[2]; acc["v"] += V.sum(0)
        assert np.max(np.abs(r["M"] - whole["M"][lo:hi])) < 1e-13
    for k in ("sym", "ssexp", "slogv"):
        assert abs(acc[k] - whole[k]) < 1e-12 * abs(whole[k])
    assert rel(acc["v"], whole["v_sum"]) < 1e-13


def test_batched_branch_scopes_stop_per_batch(case):
    """R/ebpmf_log.R:254-293 + a11: cut() batches, per-batch stop, by_row v_sum concatenated."""
    sim, _, _, _ = case
    n = 400
    b = vga_numpy.r_cut_batches(n, 3)
    assert sum(len(x) for x in b) == n and np.array_equal(np.concatenate(b), np.arange(n))
    assert [len(x) for x in b] == [133, 134, 133] or max(len(x) for x in b) - min(len(x) for x in b) <= 1
    st = simdata.solver_state(sim, var_type="by_row")
    r = vga_numpy.ebpmf_log_vga_step(st["M0"], sim["Y"], st["EL"], st["EF"], st["sigma2"], "by_row", 10, 1e-5,
                                     batches=b)
    assert len(r["n_updates"]) == 3 and r["v_sum"].shape == (n,)
    w = vga_numpy.ebpmf_log_vga_step(st["M0"], sim["Y"], st["EL"], st["EF"], st["sigma2"], "by_row", 10, 1e-5)
    assert max(r["n_updates"]) <= w["n_updates"][0]


def test_sum_lfactorial():
    x = np.array([0, 1, 2, 3, 10, 170.0])
    expect = sum(np.log(np.arange(1, int(k) + 1)).sum() for k in x)
    assert abs(c_oracle.sum_lfactorial(x) - expect) < 1e-9 * expect
    assert abs(vga_numpy.sum_lfactorial_sparseMat(sp.csc_matrix(x.reshape(2, 3))) - expect) < 1e-9 * expect
