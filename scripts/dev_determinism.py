import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, scipy.sparse as sp
import __graft_entry__ as ge
from oracle import simdata
pkg = ge.load_package()
ctx = pkg.VgaContext(0)
c = simdata.offset_for_density(0.08, K=4)
sim = simdata.sim_data_log(1200, 256, K=4, seed=12345, log_offset=c)
st = simdata.solver_state(sim, start="cold", seed=12346)
ctx.set_y(sp.csc_matrix(sim["Y"]))
ctx.set_m(st["M0"]); ctx.set_factors(st["EL"], st["EF"]); ctx.set_sigma2(st["sigma2"], "by_col")
ms = []
for rep in range(4):
    i = ctx.vga_step(10, 1e-5, want_v=True)
    ms.append((ctx.get_m(), ctx.get_v(), i.n_updates, i.passes, ctx.last_diag()))
    ctx.rewind_m()
for r in range(1, 4):
    d = ms[r][0] != ms[0][0]
    dv = ms[r][1] != ms[0][1]
    print("rep", r, "u", ms[r][2], "passes", ms[r][3], ms[r][4], "ndiff M", d.sum(), "ndiff V", dv.sum())
    if d.sum():
        ii, jj = np.nonzero(d)
        print("  rows", np.unique(ii // 32)[:20], "cols", np.unique(jj)[:40])
        k = 0
        print("  sample", ii[k], jj[k], repr(ms[0][0][ii[k], jj[k]]), repr(ms[r][0][ii[k], jj[k]]), "y", sim["Y"][ii[k], jj[k]])
        rel = np.abs(ms[r][0][d] - ms[0][0][d]) / np.abs(ms[0][0][d])
        print("  max rel diff", rel.max())
