"""REFERENCE-made golden vectors: tests/golden/make_golden.R (run by a maintainer who has R and the reference
package) writes the reference's own outputs for the committed inputs of tests/golden/r_inputs/ into
tests/golden/r_reference/.  When that directory exists these tests pin the oracle (CPU) and the CUDA path (GPU)
against it -- M, V, v_sum, sym/ssexp/slogv, sigma2, the ELBO to 1e-8 and the update count exactly.  R is not
installed in this repository's images, so here the directory is absent and the tests SKIP, saying so: parity
stays "unpinned" until someone runs the script.  What is always checked: the exported inputs are the inputs of
the oracle-made fixtures, byte for byte, in the layout the R script reads."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import c_oracle, vga_numpy as vn

HERE = os.path.dirname(os.path.abspath(__file__))
INP = os.path.join(HERE, "golden", "r_inputs")
REF = os.path.join(HERE, "golden", "r_reference")
G = np.load(os.path.join(HERE, "golden", "vga_golden.npz"))
CASES = {"def": (10, 1e-5), "exh": (2, 1e-5), "tight": (1000, 1e-8)}
needs_r = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "manifest.tsv")),
                             reason="tests/golden/r_reference/ absent: nobody has run tests/golden/make_golden.R "
                                    "(needs R + the reference package); parity remains unpinned")


def read_dir(d):
    out = {}
    with open(os.path.join(d, "manifest.tsv")) as f:
        next(f)
        for line in f:
            nm, nr, nc = line.split()
            a = np.fromfile(os.path.join(d, nm + ".f64"), dtype="<f8")
            assert a.size == int(nr) * int(nc), nm
            out[nm] = a.reshape((int(nc), int(nr))).T          # column-major on disk
    return out


def rel(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b)) / np.maximum(np.abs(b), 1e-300)))


def test_exported_inputs_are_the_fixture_inputs():
    d = read_dir(INP)
    assert len(d) == 18
    for nm, a in d.items():
        g = np.asarray(G[nm], dtype=np.float64)
        g = g.reshape(g.shape[0], -1) if g.ndim else g.reshape(1, 1)
        assert np.array_equal(a, g), nm
    assert open(os.path.join(HERE, "golden", "make_golden.R")).read().count("source(file.path(ref") == 2


@needs_r
@pytest.mark.parametrize("vt", ["by_col", "by_row", "constant"])
@pytest.mark.parametrize("tag", sorted(CASES))
def test_oracle_matches_the_reference_outputs(vt, tag):
    check_oracle_steps(read_dir(REF), vt, tag)


def check_oracle_steps(R, vt, tag):
    mi, tol = CASES[tag]
    k = "%s_%s_" % (vt, tag)
    for res, u in ((lambda r: (r, r["n_updates"][0]))(vn.ebpmf_log_vga_step(G[vt + "_M0"], G["Y"], G[vt + "_EL"], G[vt + "_EF"],
                                                                           G[vt + "_sigma2"], vt, mi, tol)),
                   (lambda c: (c, c["n_updates"]))(c_oracle.vga_step(G[vt + "_M0"], G["Y"], G[vt + "_EL"], G[vt + "_EF"],
                                                                     G[vt + "_sigma2"], vt, mi, tol))):
        sc = R[k + "scal"].ravel()
        assert u == int(sc[3])
        assert np.max(np.abs(res["M"] - R[k + "M"]) / np.maximum(np.abs(R[k + "M"]), 1.0)) < 1e-8
        assert rel(res["V"], R[k + "V"]) < 1e-8 and rel(np.atleast_1d(res["v_sum"]), R[k + "v_sum"].ravel()) < 1e-8
        for i, name in enumerate(("sym", "ssexp", "slogv")):
            assert abs(res[name] - sc[i]) <= 1e-8 * max(1.0, abs(sc[i]))
        if tag == "def":
            s2 = vn.ebpmf_log_update_sigma2(dict(R2=G[vt + "_R2"]), G[vt + "_sigma2"], res["v_sum"], vt, 0, 1, 1, *G["Y"].shape)
            assert rel(np.atleast_1d(s2), R[k + "sigma2_new"].ravel()) < 1e-8
            for cap, key in ((2.0, "sigma2_new_cap"), (1e9, "sigma2_new_cap_big")):      # the cap gate; by_row: Rfast::rowMaxs' positions
                sc2 = vn.ebpmf_log_update_sigma2(dict(R2=G[vt + "_R2"], EL=G[vt + "_EL"], EF=G[vt + "_EF"]), G[vt + "_sigma2"], res["v_sum"], vt,
                                                 cap, 1, 1, *G["Y"].shape)
                assert rel(np.atleast_1d(sc2), R[k + key].ravel()) < 1e-8
            n, p = G["Y"].shape
            ss = {"by_col": n, "by_row": p, "constant": n * p}[vt]
            obj = vn.calc_ebpmf_log_obj(n, p, res["sym"], res["ssexp"], res["slogv"], res["v_sum"], s2, G[vt + "_R2"], -123.4, -5.5,
                                        ss, 1, 1)
            assert abs(obj - R[k + "obj"].ravel()[0]) <= 1e-8 * abs(R[k + "obj"].ravel()[0])


@needs_r
def test_oracle_other_solvers_match_the_reference_outputs():
    check_oracle_other_solvers(read_dir(REF))


def check_oracle_other_solvers(R):
    EL, EF, s2 = G["by_col_EL"], G["by_col_EF"], G["by_col_sigma2"]
    S2 = vn.adjust_var_shape(s2, "by_col", *G["Y"].shape)
    fi = c_oracle.vga_fixed_iter(G["by_col_M0"], None, G["Y"], 1, EL @ EF.T, S2, 1)
    assert np.max(np.abs(fi["M"] - R["fixed1_M"])) < 1e-8 and rel(fi["V"], R["fixed1_V"]) < 1e-8
    lm = c_oracle.vga_low_memory(np.log(G["Y"] + 0.5), G["Y"], G["L0"], G["F0"], EL[:, 2:], EF[:, 2:], s2, "by_col", 1000, 1e-8)
    assert np.max(np.abs(lm["M"] - R["lowmem_M"]) / np.maximum(np.abs(R["lowmem_M"]), 1.0)) < 1e-8
    assert abs(c_oracle.sum_lfactorial(G["Y"]) - R["lfact"].ravel()[0]) < 1e-8 * R["lfact"].ravel()[0]


@needs_r
@pytest.mark.gpu
@pytest.mark.parametrize("vt", ["by_col", "by_row", "constant"])
@pytest.mark.parametrize("tag", sorted(CASES))
def test_cuda_matches_the_reference_outputs(gpu_ctx, pkg, vt, tag):
    R = read_dir(REF)
    mi, tol = CASES[tag]
    k = "%s_%s_" % (vt, tag)
    gpu_ctx.set_y(sp.csc_matrix(G["Y"]))
    res = gpu_ctx.vga_step_host(G[vt + "_M0"], G[vt + "_EL"], G[vt + "_EF"], G[vt + "_sigma2"], vt, mi, tol, return_V=True)
    sc = R[k + "scal"].ravel()
    assert res["n_updates"] == int(sc[3])
    assert np.max(np.abs(res["M"] - R[k + "M"]) / np.maximum(np.abs(R[k + "M"]), 1.0)) < 1e-8
    assert rel(res["V"], R[k + "V"]) < 1e-8 and rel(np.atleast_1d(res["v_sum"]), R[k + "v_sum"].ravel()) < 1e-8
    for i, name in enumerate(("sym", "ssexp", "slogv")):
        assert abs(res[name] - sc[i]) <= 1e-8 * max(1.0, abs(sc[i]))
    if tag == "def":
        s2 = gpu_ctx.update_sigma2(G[vt + "_R2"], 1.0, 1.0, 0.0)
        assert rel(np.atleast_1d(s2), R[k + "sigma2_new"].ravel()) < 1e-8
        for cap, key in ((2.0, "sigma2_new_cap"), (1e9, "sigma2_new_cap_big")):
            sc2 = pkg.ebpmf_log_update_sigma2(dict(R2=G[vt + "_R2"], EL=G[vt + "_EL"], EF=G[vt + "_EF"]), G[vt + "_sigma2"], res["v_sum"], vt,
                                              cap, 1, 1, *G["Y"].shape, ctx=gpu_ctx)
            assert rel(np.atleast_1d(sc2), R[k + key].ravel()) < 1e-8
        n, p = G["Y"].shape
        ss = {"by_col": n, "by_row": p, "constant": n * p}[vt]
        obj = pkg.calc_ebpmf_log_obj(n, p, res["sym"], res["ssexp"], res["slogv"], res["v_sum"], s2, G[vt + "_R2"], -123.4, -5.5,
                                     ss, 1, 1, ctx=gpu_ctx)
        assert abs(obj - R[k + "obj"].ravel()[0]) <= 1e-8 * abs(R[k + "obj"].ravel()[0])


# ---------------------------------------------------------------------------------------------------------------
# make_golden.R itself, run by the restricted R evaluator (oracle/mini_r.py) instead of GNU R
# ---------------------------------------------------------------------------------------------------------------
REFROOT = os.environ.get("EBPMF_REFERENCE", "/root/reference")


@pytest.mark.skipif(not os.path.exists(os.path.join(REFROOT, "R", "vga_solver_mat.R")),
                    reason="the reference checkout is not present (it never is on the GPU box)")
def test_make_golden_r_runs_under_the_evaluator_and_writes_what_the_loader_expects(tmp_path):
    """The script a maintainer is asked to run has never seen GNU R here.  Executed by oracle/mini_r.py -- base R's
    file I/O (readBin, writeBin, read.delim, data.frame rows, write.table) stood in for by oracle/minir_io.py -- it
    sources the reference's files, runs to the end and leaves a directory that the loader above reads and that both
    transcriptions of the oracle match.  (Its outputs are NOT committed as tests/golden/r_reference/: that name is
    reserved for files written by GNU R.)"""
    from oracle import minir_io
    out = minir_io.run_make_golden(os.path.join(HERE, "golden", "make_golden.R"), REFROOT, str(tmp_path / "r_reference"))
    R = read_dir(str(tmp_path / "r_reference"))
    assert len(R) == 52 and "wrote 52 arrays" in out
    for vt in ("by_col", "by_row", "constant"):
        for tag in sorted(CASES):
            check_oracle_steps(R, vt, tag)
    check_oracle_other_solvers(R)
    MR = np.load(os.path.join(HERE, "golden", "minir_reference.npz"))              # same evaluator, other driver: identical numbers
    for k in ("by_col_def_M", "by_row_tight_V", "constant_exh_scal", "fixed1_V", "lowmem_M", "lfact", "by_row_def_sigma2_new_cap_big",
              "by_col_def_sigma2_new_cap"):
        assert np.array_equal(R[k].ravel(), MR["fn/" + k].ravel()), k
