"""Run one seeded fused VGA step with the library in $EBPMF_B200_LIB and dump M, V, sums to an .npz
(used by tests to compare builds, e.g. fast math vs EBPMF_EXACT_MATH)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, scipy.sparse as sp
import __graft_entry__ as ge
from oracle import simdata
pkg = ge.load_package()
out = sys.argv[1]
c = simdata.offset_for_density(0.10, K=5)
sim = simdata.sim_data_log(2000, 400, K=5, seed=4242, log_offset=c)
st = simdata.solver_state(sim, start="cold", seed=4243)
ctx = pkg.VgaContext(0)
ctx.set_y(sp.csc_matrix(sim["Y"]))
r = ctx.vga_step_host(st["M0"], st["EL"], st["EF"], st["sigma2"], "by_col", 1000, 1e-8, return_V=True)
np.savez(out, M=r["M"], V=r["V"], v_sum=r["v_sum"], s=np.array([r["sym"], r["ssexp"], r["slogv"], r["n_updates"]]))
ctx.close()
