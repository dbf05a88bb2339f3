"""Two-GPU check of the batched handle (run on a box with >= 2 GPUs):
    python scripts/mgpu_batched_check.py
Batches dealt over the GPUs must reproduce the oracle's batched branch and the one-GPU result bit for bit."""
import os
import sys

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402
from oracle import simdata, vga_numpy  # noqa: E402

pkg = ge.load_package()
n, p, K = 4200, 260, 5
sim = simdata.sim_data_log(n, p, K=K, seed=9, log_offset=simdata.offset_for_density(0.1, K=K))
st = simdata.solver_state(sim, start="cold")
one, two = pkg.VgaBatched(0, 2), pkg.VgaBatched([0, 1], 2)
for h in (one, two):
    h.set_y(sp.csc_matrix(sim["Y"]), batch_size=500)
nb = len(one.batch_row0) - 1
ref = vga_numpy.ebpmf_log_vga_step(st["M0"], sim["Y"], st["EL"], st["EF"], st["sigma2"], "by_col", 10, 1e-5,
                                   batches=vga_numpy.r_cut_batches(n, nb))
r1 = one.vga_step_host(st["M0"], st["EL"], st["EF"], st["sigma2"], "by_col", 10, 1e-5, return_V=True)
r2 = two.vga_step_host(st["M0"], st["EL"], st["EF"], st["sigma2"], "by_col", 10, 1e-5, return_V=True)
err = float(np.max(np.abs(r2["M"] - ref["M"]) / np.maximum(np.abs(ref["M"]), 1.0)))
ok = (r2["n_updates"] == list(ref["n_updates"]) and err < 1e-8 and np.array_equal(r1["M"], r2["M"])
      and np.array_equal(r1["v_sum"], r2["v_sum"]) and r1["sym"] == r2["sym"])
print("batches", nb, "n_updates", r2["n_updates"], "max rel err vs oracle %.2e" % err, "two == one:", np.array_equal(r1["M"], r2["M"]),
      "OK" if ok else "FAIL")
one.close(); two.close()
sys.exit(0 if ok else 1)
