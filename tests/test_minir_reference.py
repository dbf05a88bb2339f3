"""The oracle (CPU) and the CUDA path (GPU, through the C ABI) against tests/golden/minir_reference.npz: outputs of
the reference's OWN R source (R/vga_solver_mat.R, R/ebpmf_log.R) executed statement by statement by the restricted
R evaluator oracle/mini_r.py (tests/golden/make_golden_minir.py; provenance with the files' sha256 next to the
fixture).  Three levels, as oracle/minir_reference.py describes: the path's functions, the in-place blocks of
`ebpmf_log`'s outer loop (inline sums, `cut()` batches, the per-batch stop scope, the cap gate) and `ebpmf_log()`
itself for four outer iterations with flashier stood in for (oracle/standin_flash.py).

Tolerance 1e-8 (BASELINE.json north_star), update counts EXACT.  The fixture is not GNU R output -- that is what
tests/test_golden_r.py waits for -- but it takes the hand transcription out of the chain."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import c_oracle, standin_flash, vga_numpy as vn

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


def check_step(res, u, k):
    sc = MR[k + "scal"]
    assert u == int(sc[3])
    assert relM(res["M"], MR[k + "M"]) < RTOL and rel(res["V"], MR[k + "V"]) < RTOL
    assert rel(np.atleast_1d(res["v_sum"]), MR[k + "v_sum"]) < RTOL
    for i, name in enumerate(("sym", "ssexp", "slogv")):
        assert abs(res[name] - sc[i]) <= RTOL * max(1.0, abs(sc[i])), name


def driver_inputs():
    l0 = np.log(Y.mean(axis=1) + 0.1)
    f0 = np.log((Y.sum(axis=0) + 0.1) / np.exp(l0).sum())
    return l0, f0, np.asfortranarray(np.log(Y + 0.5)), {"by_col": np.full(P, 0.5), "by_row": np.full(N, 0.5), "constant": np.array([0.5])}


def check_driver(out, vt, bs):
    k = "drv/%s/%s/" % (vt, bs)
    ref = MR[k + "elbo_trace"]
    assert out["elbo_trace"].shape == ref.shape == (5,) and np.isneginf(out["elbo_trace"][0]) and np.isneginf(ref[0])
    assert rel(out["elbo_trace"][1:], ref[1:]) < RTOL
    assert rel(out["sigma2"], MR[k + "sigma2"]) < RTOL
    if bs == "inf":                                  # with batches the reference cannot return M (see run_driver)
        assert relM(out["M"], MR[k + "M"]) < RTOL and rel(out["V"], MR[k + "V"]) < RTOL


# ---------------------------------------------------------------------------------------------------------------
# CPU: both transcriptions of the oracle
# ---------------------------------------------------------------------------------------------------------------
def test_fixture_is_what_the_generator_says():
    assert len(MR.files) == 134
    import json
    prov = json.load(open(os.path.join(HERE, "golden", "minir_reference.provenance.json")))
    assert "NOT GNU R" in prov["evaluator"] and set(prov["reference_files_sha256"]) == {"R/vga_solver_mat.R", "R/ebpmf_log.R",
                                                                                         "R/ebpmf_log_controls.R"}


@pytest.mark.parametrize("vt", VTS)
@pytest.mark.parametrize("tag", sorted(CASES))
def test_oracle_functions(vt, tag):
    M0, EL, EF, s2, R2 = state(vt)
    mi, tol = CASES[tag]
    k = "fn/%s_%s_" % (vt, tag)
    r = vn.ebpmf_log_vga_step(M0, Y, EL, EF, s2, vt, mi, tol)
    check_step(r, r["n_updates"][0], k)
    c = c_oracle.vga_step(M0, Y, EL, EF, s2, vt, mi, tol)
    check_step(c, c["n_updates"], k)
    if tag == "def":
        for cap, key in ((0.0, "sigma2_new"), (2.0, "sigma2_new_cap"), (1e9, "sigma2_new_cap_big")):
            s = vn.ebpmf_log_update_sigma2(dict(R2=R2, EL=EL, EF=EF), s2, r["v_sum"], vt, cap, 1, 1, N, P)
            assert rel(np.atleast_1d(s), MR[k + key]) < RTOL
        s = vn.ebpmf_log_update_sigma2(dict(R2=R2), s2, r["v_sum"], vt, 0, 1, 1, N, P)
        ss = {"by_col": N, "by_row": P, "constant": N * P}[vt]
        obj = vn.calc_ebpmf_log_obj(N, P, r["sym"], r["ssexp"], r["slogv"], r["v_sum"], s, R2, -123.4, -5.5, ss, 1, 1)
        assert abs(obj - MR[k + "obj"][0]) <= RTOL * abs(MR[k + "obj"][0])


def test_oracle_other_solvers():
    M0, EL, EF, s2, _ = state("by_col")
    S2 = vn.adjust_var_shape(s2, "by_col", N, P)
    for mi in (1, 10):
        for fi in (vn.vga_pois_solver_mat_newton_fixed_iter(M0, None, Y, 1, EL @ EF.T, S2, mi),
                   c_oracle.vga_fixed_iter(M0, None, Y, 1, EL @ EF.T, S2, mi)):
            assert relM(fi["M"], MR["fn/fixed%d_M" % mi]) < RTOL and rel(fi["V"], MR["fn/fixed%d_V" % mi]) < RTOL
    for vt in VTS:
        _, EL, EF, s2, _ = state(vt)
        key = "fn/lowmem_M" if vt == "by_col" else "fn/lowmem_%s_M" % vt
        lm = vn.vga_pois_solver_mat_newton_low_memory(np.log(Y + 0.5), Y, G["L0"], G["F0"], EL[:, 2:], EF[:, 2:], s2, vt, 1000, 1e-8)
        assert relM(lm, MR[key]) < RTOL
        lc = c_oracle.vga_low_memory(np.log(Y + 0.5), Y, G["L0"], G["F0"], EL[:, 2:], EF[:, 2:], s2, vt, 1000, 1e-8)
        assert relM(lc["M"], MR[key]) < RTOL
    assert abs(vn.sum_lfactorial_sparseMat(sp.csc_matrix(Y)) - MR["fn/lfact"][0]) < RTOL * MR["fn/lfact"][0]
    assert abs(c_oracle.sum_lfactorial(Y) - MR["fn/lfact"][0]) < RTOL * MR["fn/lfact"][0]


@pytest.mark.parametrize("vt", VTS)
def test_oracle_loop_blocks_in_place(vt):
    """The statements of ebpmf_log's loop body executed where they stand (R/ebpmf_log.R:254-334)."""
    M0, EL, EF, s2, R2 = state(vt)
    fit = dict(R2=R2, EL=EL, EF=EF)
    k = "blk/%s/inf/" % vt
    r = vn.ebpmf_log_vga_step(M0, Y, EL, EF, s2, vt, 10, 1e-5)
    assert rel(np.atleast_1d(r["v_sum"]), MR[k + "cap0/v_sum"]) < RTOL
    for name in ("sym", "ssexp", "slogv"):
        assert abs(r[name] - MR[k + "cap0/" + name]) <= RTOL * max(1.0, abs(MR[k + "cap0/" + name]))
    for cap in (0.0, 2.0):
        s = vn.ebpmf_log_update_sigma2(fit, s2, r["v_sum"], vt, cap, 1, 1, N, P)
        assert rel(np.atleast_1d(s), MR[k + "cap%g/sigma2" % cap]) < RTOL
    k = "blk/%s/25/" % vt                                    # general_control$batch_size = 25 -> 4 batches, per-batch stop scope
    batches = vn.r_cut_batches(N, int(np.ceil(N / 25)))
    assert [int(b[0]) for b in batches] + [N] == MR[k + "cap0/batch_row0"].astype(int).tolist() and MR[k + "cap0/n_batch"] == 4
    r = vn.ebpmf_log_vga_step(M0, Y, EL, EF, s2, vt, 10, 1e-5, batches=batches)
    assert relM(r["M"], MR[k + "cap0/M"]) < RTOL and rel(np.atleast_1d(r["v_sum"]), MR[k + "cap0/v_sum"]) < RTOL
    for name in ("sym", "ssexp", "slogv"):
        assert abs(r[name] - MR[k + "cap0/" + name]) <= RTOL * max(1.0, abs(MR[k + "cap0/" + name]))
    for cap in (0.0, 2.0):
        s = vn.ebpmf_log_update_sigma2(fit, s2, r["v_sum"], vt, cap, 1, 1, N, P)
        assert rel(np.atleast_1d(s), MR[k + "cap%g/sigma2" % cap]) < RTOL


@pytest.mark.parametrize("vt", VTS)
@pytest.mark.parametrize("bs", ["inf", "25"])
def test_oracle_through_the_whole_driver(vt, bs):
    """ebpmf_log() itself under the evaluator vs the oracle's functions in the same outer loop."""
    l0, f0, M_init, s2 = driver_inputs()
    batches = None if bs == "inf" else vn.r_cut_batches(N, int(np.ceil(N / 25)))
    out = standin_flash.outer_loop(
        Y, l0, f0, M_init, s2[vt], vt, 3, 4,
        step=lambda M, EL, EF, sig: vn.ebpmf_log_vga_step(M, Y, EL, EF, sig, vt, 10, 1e-5, batches=batches),
        update_sigma2=lambda fit, sig, vs: vn.ebpmf_log_update_sigma2(fit, sig, vs, vt, 0, 1, 1, N, P),
        calc_obj=vn.calc_ebpmf_log_obj, const=-vn.sum_lfactorial_sparseMat(Y))
    check_driver(out, vt, bs)


# ---------------------------------------------------------------------------------------------------------------
# GPU: the CUDA path through the C ABI
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("vt", VTS)
@pytest.mark.parametrize("tag", sorted(CASES))
def test_cuda_functions(gpu_ctx, pkg, vt, tag):
    M0, EL, EF, s2, R2 = state(vt)
    mi, tol = CASES[tag]
    k = "fn/%s_%s_" % (vt, tag)
    gpu_ctx.set_y(sp.csc_matrix(Y))
    res = gpu_ctx.vga_step_host(M0, EL, EF, s2, vt, mi, tol, return_V=True)
    check_step(res, res["n_updates"], k)
    if tag == "def":
        s = gpu_ctx.update_sigma2(R2, 1.0, 1.0, 0.0)                        # on the resident v_sum
        assert rel(np.atleast_1d(s), MR[k + "sigma2_new"]) < RTOL
        ss = {"by_col": N, "by_row": P, "constant": N * P}[vt]
        obj = pkg.calc_ebpmf_log_obj(N, P, res["sym"], res["ssexp"], res["slogv"], res["v_sum"], s, R2, -123.4, -5.5, ss, 1, 1, ctx=gpu_ctx)
        assert abs(obj - MR[k + "obj"][0]) <= RTOL * abs(MR[k + "obj"][0])
        for cap, key in ((0.0, "sigma2_new"), (2.0, "sigma2_new_cap"), (1e9, "sigma2_new_cap_big")):     # host signature, with the cap gate
            s = pkg.ebpmf_log_update_sigma2(dict(R2=R2, EL=EL, EF=EF), s2, res["v_sum"], vt, cap, 1, 1, N, P, ctx=gpu_ctx)
            assert rel(np.atleast_1d(s), MR[k + key]) < RTOL
    # the drop-in dense signature on the same inputs (R/vga_solver_mat.R:6-41 called directly)
    info = {}
    d = pkg.vga_pois_solver_mat_newton(M0, Y, 1, EL @ EF.T, vn.adjust_var_shape(s2, vt, N, P), maxiter=mi, tol=tol, info=info, ctx=gpu_ctx)
    assert info["n_updates"] == int(MR[k + "scal"][3])
    assert relM(d["M"], MR[k + "M"]) < RTOL and rel(d["V"], MR[k + "V"]) < RTOL


@pytest.mark.gpu
def test_cuda_other_solvers(gpu_ctx, pkg):
    M0, EL, EF, s2, _ = state("by_col")
    S2 = vn.adjust_var_shape(s2, "by_col", N, P)
    for mi in (1, 10):
        fi = pkg.vga_pois_solver_mat_newton_fixed_iter(M0, None, Y, 1, EL @ EF.T, S2, mi, ctx=gpu_ctx)
        assert relM(fi["M"], MR["fn/fixed%d_M" % mi]) < RTOL and rel(fi["V"], MR["fn/fixed%d_V" % mi]) < RTOL
    for vt in VTS:
        _, EL, EF, s2, _ = state(vt)
        key = "fn/lowmem_M" if vt == "by_col" else "fn/lowmem_%s_M" % vt
        for Yin in (Y, sp.csc_matrix(Y)):
            lm = pkg.vga_pois_solver_mat_newton_low_memory(np.log(Y + 0.5), Yin, G["L0"], G["F0"], EL[:, 2:], EF[:, 2:], s2, vt, 1000, 1e-8,
                                                           ctx=gpu_ctx)
            assert relM(lm, MR[key]) < RTOL
    lf = pkg.sum_lfactorial_sparseMat(sp.csc_matrix(Y), ctx=gpu_ctx)
    assert abs(lf - MR["fn/lfact"][0]) < RTOL * MR["fn/lfact"][0]


@pytest.mark.gpu
@pytest.mark.parametrize("vt", VTS)
def test_cuda_batched_branch_in_place(pkg, vt):
    M0, EL, EF, s2, R2 = state(vt)
    k = "blk/%s/25/" % vt
    h = pkg.VgaBatched(0, 2)
    try:
        h.set_y(sp.csc_matrix(Y), batch_size=25)
        assert list(h.batch_row0) == MR[k + "cap0/batch_row0"].astype(int).tolist()
        r = h.vga_step_host(M0, EL, EF, s2, vt, 10, 1e-5, return_V=True)
    finally:
        h.close()
    assert relM(r["M"], MR[k + "cap0/M"]) < RTOL and rel(np.atleast_1d(r["v_sum"]), MR[k + "cap0/v_sum"]) < RTOL
    for name in ("sym", "ssexp", "slogv"):
        assert abs(r[name] - MR[k + "cap0/" + name]) <= RTOL * max(1.0, abs(MR[k + "cap0/" + name]))
    for cap in (0.0, 2.0):
        s = pkg.ebpmf_log_update_sigma2(dict(R2=R2, EL=EL, EF=EF), s2, r["v_sum"], vt, cap, 1, 1, N, P)
        assert rel(np.atleast_1d(s), MR[k + "cap%g/sigma2" % cap]) < RTOL


@pytest.mark.gpu
@pytest.mark.parametrize("vt", VTS)
@pytest.mark.parametrize("bs", ["inf", "25"])
def test_cuda_through_the_whole_driver(gpu_ctx, pkg, vt, bs):
    """Four outer iterations of ebpmf_log around the CUDA step (resident M for batch_size = Inf, the batched
    handle otherwise) against ebpmf_log() itself run by the evaluator."""
    l0, f0, M_init, s2 = driver_inputs()
    h = None
    if bs == "inf":
        gpu_ctx.set_y(sp.csc_matrix(Y))
        gpu_ctx.set_m(M_init)                                               # M stays on the device between the iterations
        step = lambda M, EL, EF, sig: gpu_ctx.vga_step_resident_host(EL, EF, sig, vt, 10, 1e-5, return_V=True)      # noqa: E731
        lfact = gpu_ctx.sum_lfactorial()
    else:
        h = pkg.VgaBatched(0, 3)
        h.set_y(sp.csc_matrix(Y), batch_size=25)
        step = lambda M, EL, EF, sig: h.vga_step_host(M, EL, EF, sig, vt, 10, 1e-5, return_V=True)                  # noqa: E731
        lfact = h.sum_lfactorial()
    try:
        out = standin_flash.outer_loop(
            Y, l0, f0, M_init, s2[vt], vt, 3, 4, step=step,
            update_sigma2=lambda fit, sig, vs: pkg.ebpmf_log_update_sigma2(fit, sig, vs, vt, 0, 1, 1, N, P, ctx=gpu_ctx),
            calc_obj=lambda *a: pkg.calc_ebpmf_log_obj(*a, ctx=gpu_ctx), const=-lfact)
    finally:
        if h is not None:
            h.close()
    check_driver(out, vt, bs)


# ---------------------------------------------------------------------------------------------------------------
# differential: the oracle against the reference's source on RANDOM shapes and settings (needs the checkout)
# ---------------------------------------------------------------------------------------------------------------
REFROOT = os.environ.get("EBPMF_REFERENCE", "/root/reference")


@pytest.mark.skipif(not os.path.exists(os.path.join(REFROOT, "R", "vga_solver_mat.R")),
                    reason="the reference checkout is not present (it never is on the GPU box)")
def test_oracle_equals_the_executed_reference_on_random_cases():
    """60 random problems (1 x 1 up to 70 x 45, all var_types, dense counts, cold and warm starts, tolerances from 1e-3
    to 1e-10, 1 to 40 iterations, row batches of random size, the cap gate on and off): both transcriptions must
    reproduce the reference's source to 1e-12 (observed: bit for bit except where a matrix product is summed in
    another order -- `tcrossprod` on a row batch against the same rows of the whole product, one ulp).
    The fixture pins one problem; this pins the transcriptions themselves."""
    from oracle import mini_r, minir_reference as mr, simdata
    it = mr.load_reference(REFROOT)
    rng = np.random.default_rng(2024)
    n_batched = exact = n_broken = 0
    TOL = 1e-12
    for case in range(60):
        n, p, K = int(rng.integers(1, 71)), int(rng.integers(1, 46)), int(rng.integers(1, 4))
        vt = VTS[case % 3]
        sim = simdata.sim_data_log(n, p, K=K, seed=1000 + case, log_offset=float(rng.uniform(-2.5, 0.5)))
        st = simdata.solver_state(sim, start="warm" if case % 4 == 0 else "cold", var_type=vt, seed=2000 + case)
        mi, tol = int(rng.integers(1, 41)), float(10.0 ** rng.uniform(-10, -3))
        cap = float(rng.choice([0.0, 0.5, 3.0, 1e9]))
        bs = np.inf if case % 2 == 0 or n < 4 else int(rng.integers(2, max(3, n // 2 + 1)))
        Yc, M0, EL, EF, s2, R2 = sim["Y"], st["M0"], st["EL"], st["EF"], st["sigma2"], st["R2"]
        n_batch = 0 if np.isinf(bs) else int(np.ceil(n / bs))
        if n_batch > 1 and p > 1 and min(len(b) for b in vn.r_cut_batches(n, n_batch)) == 1:
            # a ONE-row batch breaks the reference: Y[b,] drops to a vector, as.matrix() makes it p x 1 against the
            # 1 x p Beta and Sigma2 (R/ebpmf_log.R:199,260-266).  The CUDA path and the oracle accept such batches.
            with pytest.raises(mini_r.RError, match="non-conformable"):
                mr.run_loop_blocks(it, Yc, M0, EL, EF, s2, R2, vt, batch_size=bs, maxiter_vga=mi, vga_tol=tol, cap_var_mean_ratio=cap)
            n_broken += 1
            continue
        ref = mr.run_loop_blocks(it, Yc, M0, EL, EF, s2, R2, vt, batch_size=bs, maxiter_vga=mi, vga_tol=tol, cap_var_mean_ratio=cap)
        batches = None
        if ref["n_batch"] > 1:
            n_batched += 1
            batches = vn.r_cut_batches(n, ref["n_batch"])
            assert [int(b[0]) for b in batches] + [n] == ref["batch_row0"].tolist(), (case, n, bs)
        o = vn.ebpmf_log_vga_step(M0, Yc, EL, EF, s2, vt, mi, tol, batches=batches)
        what = (case, n, p, vt, mi, tol, bs)
        exact += int(np.array_equal(o["M"], ref["M"]))
        assert relM(o["M"], ref["M"]) < TOL and rel(np.atleast_1d(o["v_sum"]), ref["v_sum"]) < TOL, what
        for name in ("sym", "ssexp", "slogv"):
            assert abs(o[name] - ref[name]) <= TOL * max(1.0, abs(ref[name])), what
        s_new = vn.ebpmf_log_update_sigma2(dict(R2=R2, EL=EL, EF=EF), s2, o["v_sum"], vt, cap, 1, 1, n, p)
        assert rel(np.atleast_1d(s_new), ref["sigma2"]) < TOL, what
        c = c_oracle.vga_step(M0, Yc, EL, EF, s2, vt, mi, tol) if batches is None else None
        if c is not None:
            assert relM(c["M"], ref["M"]) < TOL and rel(np.atleast_1d(c["v_sum"]), ref["v_sum"]) < TOL, what
        # the two other solver signatures on the same problem
        S2 = vn.adjust_var_shape(s2, vt, n, p)
        fi_ref = it.call("vga_pois_solver_mat_newton_fixed_iter", M0, None, Yc, 1, EL @ EF.T, S2, maxiter=min(mi, 5))
        fi = vn.vga_pois_solver_mat_newton_fixed_iter(M0, None, Yc, 1, EL @ EF.T, S2, min(mi, 5))
        assert relM(fi["M"], fi_ref.get("M")) < TOL and rel(fi["V"], fi_ref.get("V")) < TOL, what
        if EL.shape[1] > 2:
            Ml = np.log(Yc + 0.5)
            lm_ref = it.call("vga_pois_solver_mat_newton_low_memory", Ml, Yc, sim["L0"], sim["F0"], EL[:, 2:], EF[:, 2:], s2, vt, maxiter=mi, tol=tol)
            lm = vn.vga_pois_solver_mat_newton_low_memory(Ml, Yc, sim["L0"], sim["F0"], EL[:, 2:], EF[:, 2:], s2, vt, mi, tol)
            assert relM(lm, np.asarray(lm_ref)) < TOL, what
    assert n_batched >= 15 and exact >= 45, (n_batched, exact, n_broken)


@pytest.mark.skipif(not os.path.exists(os.path.join(REFROOT, "R", "vga_solver_mat.R")),
                    reason="the reference checkout is not present (it never is on the GPU box)")
def test_reference_breaks_on_a_one_row_batch_and_the_oracle_does_not():
    """n = 5 in 3 batches is cut 2/1/2 (R/ebpmf_log.R:197).  For the one-row batch `Y[b,]` and `flash_fit$Y[b,]` drop
    to vectors, `as.matrix()` turns the counts into a p x 1 matrix, and `Sigma2*X` meets a 1 x p matrix: the reference
    stops with "non-conformable arrays".  The oracle and the CUDA path treat a one-row batch like any other -- a
    superset of the reference's behaviour, recorded here so that nobody mistakes it for a parity gap."""
    from oracle import mini_r, minir_reference as mr, simdata
    it = mr.load_reference(REFROOT)
    sim = simdata.sim_data_log(5, 7, K=2, seed=5, log_offset=-1.0)
    st = simdata.solver_state(sim, start="cold", var_type="by_col", seed=6)
    batches = vn.r_cut_batches(5, 3)
    assert [len(b) for b in batches] == [2, 1, 2]
    with pytest.raises(mini_r.RError, match="non-conformable"):
        mr.run_loop_blocks(it, sim["Y"], st["M0"], st["EL"], st["EF"], st["sigma2"], st["R2"], "by_col", batch_size=2)
    o = vn.ebpmf_log_vga_step(st["M0"], sim["Y"], st["EL"], st["EF"], st["sigma2"], "by_col", 10, 1e-5, batches=batches)
    assert np.isfinite(o["M"]).all() and len(o["n_updates"]) == 3


@pytest.mark.skipif(not os.path.exists(os.path.join(REFROOT, "R", "vga_solver_mat.R")),
                    reason="the reference checkout is not present (it never is on the GPU box)")
def test_oracle_equals_the_executed_reference_operand_kinds_and_elbo():
    """The stand-alone signatures with every operand shape R would recycle (scalar / length-n vector / matrix for S,
    Beta and Sigma2; V given or NULL; dense or dgCMatrix counts), the ELBO with operands of unequal length (each sum()
    at its own length, R/ebpmf_log.R:407), the constant-variance sigma2 update with a vector R2, and the archived
    driver's un-fused ELBO (inst/code/Poisson_MF_splitting.R:373-391) -- oracle vs the reference's source."""
    from oracle import mini_r, minir_reference as mr, simdata
    it = mr.load_reference(REFROOT)
    it.source(os.path.join(REFROOT, "inst", "code", "Poisson_MF_splitting.R"))     # archived driver: calc_split_PMF_obj_flashier
    for f in mr.FILES:                                                             # ... whose adjust_var_shape etc. must not shadow the package's
        it.source(os.path.join(REFROOT, f))
    rng = np.random.default_rng(7)
    TOL = 1e-12
    for case in range(24):
        n, p = int(rng.integers(2, 40)), int(rng.integers(2, 30))
        sim = simdata.sim_data_log(n, p, K=2, seed=300 + case, log_offset=float(rng.uniform(-2.0, 0.5)))
        st = simdata.solver_state(sim, start="cold", var_type="by_col", seed=400 + case)
        Yc, M0 = sim["Y"], st["M0"]
        kinds = {"scalar": lambda lo, hi: float(rng.uniform(lo, hi)), "rowvec": lambda lo, hi: rng.uniform(lo, hi, size=n),
                 "matrix": lambda lo, hi: np.asfortranarray(rng.uniform(lo, hi, size=(n, p)))}
        ks, kb, kv = (list(kinds)[int(rng.integers(0, 3))] for _ in range(3))
        S, Beta, Sigma2 = kinds[ks](0.5, 1.5), kinds[kb](-1.0, 1.0), kinds[kv](0.1, 1.2)
        mi, tol = int(rng.integers(1, 30)), float(10.0 ** rng.uniform(-9, -4))
        X = mini_r.RSparse(Yc) if case % 2 else Yc
        ref = it.call("vga_pois_solver_mat_newton", M0, X, S, Beta, Sigma2, maxiter=mi, tol=tol)
        o = vn.vga_pois_solver_mat_newton(M0, Yc, S, Beta, Sigma2, mi, tol)
        what = (case, n, p, ks, kb, kv, mi, tol)
        assert relM(o["M"], ref.get("M")) < TOL and rel(o["V"], ref.get("V")) < TOL, what
        c = c_oracle.vga_newton(M0, Yc, S, Beta, Sigma2, mi, tol)
        assert relM(c["M"], ref.get("M")) < TOL and rel(c["V"], ref.get("V")) < TOL, what
        V0 = None if case % 3 == 0 else np.asfortranarray(rng.uniform(0.05, 0.4, size=(n, p)))
        ref = it.call("vga_pois_solver_mat_newton_fixed_iter", M0, V0, Yc, S, Beta, Sigma2, maxiter=int(rng.integers(1, 6)) if V0 is None else 3)
        with np.errstate(all="ignore"):                     # exp() overflows on the way for some draws, in R as here
            o = vn.vga_pois_solver_mat_newton_fixed_iter(M0, V0, Yc, S, Beta, Sigma2, 3) if V0 is not None else None
        if o is not None:
            assert relM(o["M"], ref.get("M")) < TOL and rel(o["V"], ref.get("V")) < TOL, what
        # ELBO with operands of unequal length, sigma2 update 'constant' with a vector R2, un-fused ELBO
        r = vn.ebpmf_log_vga_step(M0, Yc, st["EL"], st["EF"], st["sigma2"], "by_col", 10, 1e-5)
        lens = [(p, p, p), (1, 1, p), (p, 1, n), (1, p, 1)][case % 4]                          # len(sv), len(sigma2), len(R2)
        sv, sg, R2 = rng.uniform(1, 5, lens[0]), rng.uniform(0.2, 1.5, lens[1]), rng.uniform(0.5, 9, lens[2])
        a = (n, p, r["sym"], r["ssexp"], r["slogv"], sv, sg, R2, -12.5, -3.25, float(n), 1.0, 1.0)
        assert abs(vn.calc_ebpmf_log_obj(*a) - it.call("calc_ebpmf_log_obj", *a)[0]) <= TOL * abs(vn.calc_ebpmf_log_obj(*a)), what
        R2v = rng.uniform(0.5, 9, p)
        s_ref = it.call("ebpmf_log_update_sigma2", {"flash_fit": {"R2": R2v}}, 0.7, 33.0, "constant", 0, 1, 1, n, p)
        s_o = vn.ebpmf_log_update_sigma2(dict(R2=R2v), np.array([0.7]), 33.0, "constant", 0, 1, 1, n, p)
        assert rel(np.atleast_1d(s_o), s_ref) < TOL, what
        for vt in VTS:
            sg = {"by_col": rng.uniform(0.2, 1.5, p), "by_row": rng.uniform(0.2, 1.5, n), "constant": rng.uniform(0.2, 1.5, 1)}[vt]
            R2s = rng.uniform(0.5, 9, sg.size)
            ref_v = it.call("calc_split_PMF_obj_flashier", Yc, S if ks != "rowvec" else 1.0, sg, r["M"], r["V"], {"flash.fit": {"R2": R2s}}, -2.5, -77.0, vt)
            o_v = vn.calc_split_PMF_obj_flashier(Yc, S if ks != "rowvec" else 1.0, sg, r["M"], r["V"], R2s, -2.5, -77.0, vt)
            assert abs(o_v - ref_v[0]) <= TOL * abs(o_v), (what, vt)
    assert abs(vn.sum_lfactorial_sparseMat(sp.csc_matrix(Yc)) - it.call("sum_lfactorial_sparseMat", mini_r.RSparse(Yc))[0]) <= TOL * 1e3
