"""Generates tests/golden/vga_golden.npz from the CPU oracle (NOT from the reference: R cannot run
here -- "parity unpinned", see oracle/vga_numpy.py).  The fixtures freeze the oracle's outputs on
small seeded cases so that later edits of the oracle or of the CUDA path are caught as drift.
Run from the repo root:  python tests/golden/make_golden.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
from oracle import simdata, vga_numpy as vn

out = {}
sim = simdata.sim_data_log(96, 40, K=3, seed=12345, log_offset=-2.0)
for vt in ("by_col", "by_row", "constant"):
    st = simdata.solver_state(sim, start="cold", var_type=vt, seed=77)
    for tag, (mi, tol) in (("def", (10, 1e-5)), ("exh", (2, 1e-5)), ("tight", (1000, 1e-8))):
        r = vn.ebpmf_log_vga_step(st["M0"], sim["Y"], st["EL"], st["EF"], st["sigma2"], vt, mi, tol)
        k = "%s_%s_" % (vt, tag)
        out[k + "M"] = r["M"]; out[k + "V"] = r["V"]; out[k + "v_sum"] = np.atleast_1d(r["v_sum"])
        out[k + "scal"] = np.array([r["sym"], r["ssexp"], r["slogv"], r["n_updates"][0], float(r["converged"][0])])
    out[vt + "_sigma2"] = st["sigma2"]; out[vt + "_R2"] = st["R2"]
    out[vt + "_EL"] = st["EL"]; out[vt + "_EF"] = st["EF"]; out[vt + "_M0"] = st["M0"]
out["Y"] = sim["Y"]; out["L0"] = sim["L0"]; out["F0"] = sim["F0"]
st = simdata.solver_state(sim, start="cold", var_type="by_col", seed=77)
Beta = st["EL"] @ st["EF"].T; S2 = vn.adjust_var_shape(st["sigma2"], "by_col", 96, 40)
fi = vn.vga_pois_solver_mat_newton_fixed_iter(st["M0"], None, sim["Y"], 1, Beta, S2, 1)
out["fixed1_M"] = fi["M"]; out["fixed1_V"] = fi["V"]
lm = vn.vga_pois_solver_mat_newton_low_memory(np.log(sim["Y"] + 0.5), sim["Y"], sim["L0"], sim["F0"], st["EL"][:, 2:],
                                              st["EF"][:, 2:], st["sigma2"], "by_col", 1000, 1e-8)
out["lowmem_M"] = lm
out["lfact"] = np.array([vn.sum_lfactorial_sparseMat(sim["Y"])])
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "vga_golden.npz"), **out)
print("wrote", len(out), "arrays")

# ---- the batched branch (R/ebpmf_log.R:254-300): general_control$batch_size = 25 -> 4 batches of 24 rows ----
bout = {}
batches = vn.r_cut_batches(96, int(np.ceil(96 / 25)))
bout["batch_row0"] = np.array([int(b[0]) for b in batches] + [96], dtype=np.int64)
for vt in ("by_col", "by_row", "constant"):
    st = simdata.solver_state(sim, start="cold", var_type=vt, seed=77)
    r = vn.ebpmf_log_vga_step(st["M0"], sim["Y"], st["EL"], st["EF"], st["sigma2"], vt, 10, 1e-5, batches=batches)
    bout[vt + "_M"] = r["M"]; bout[vt + "_V"] = r["V"]; bout[vt + "_v_sum"] = np.atleast_1d(r["v_sum"])
    bout[vt + "_scal"] = np.array([r["sym"], r["ssexp"], r["slogv"]])
    bout[vt + "_n_updates"] = np.array(r["n_updates"], dtype=np.int64)
    bout[vt + "_converged"] = np.array(r["converged"], dtype=np.int64)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "vga_golden_batched.npz"), **bout)
print("wrote", len(bout), "batched arrays")
