"""Single-process multi-GPU check (the R-session shape): one ebpmf_group over all visible GPUs, full
host matrices in, full host matrices out, compared with the CPU oracle on the whole matrix.
   python scripts/mgpu_group_check.py [ngpu]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, scipy.sparse as sp
import __graft_entry__ as ge
from oracle import simdata, c_oracle
pkg = ge.load_package()
import ctypes
cnt = ctypes.c_int(); pkg.load().ebpmf_device_count(ctypes.byref(cnt))
ng = int(sys.argv[1]) if len(sys.argv) > 1 else cnt.value
grp = pkg.VgaGroup(list(range(ng)))
ok = True
for (n, p, K, vt, mi, tol) in [(5000, 700, 5, "by_col", 10, 1e-5), (3001, 333, 3, "by_row", 1000, 1e-8),
                                (4100, 512, 4, "constant", 3, 1e-5), (40000, 4000, 8, "by_col", 10, 1e-5)]:
    sim = simdata.sim_data_log(n, p, K=K, seed=12345, log_offset=simdata.offset_for_density(0.08, K=K))
    st = simdata.solver_state(sim, start="cold", var_type=vt)
    grp.set_y(sp.csc_matrix(sim["Y"]))
    t = time.perf_counter()
    res = grp.vga_step_host(st["M0"], st["EL"], st["EF"], st["sigma2"], vt, mi, tol, return_V=True)
    dt = time.perf_counter() - t
    ref = c_oracle.vga_step(st["M0"], sim["Y"], st["EL"], st["EF"], st["sigma2"], vt, mi, tol)
    eM = float(np.max(np.abs(res["M"] - ref["M"]) / np.maximum(np.abs(ref["M"]), 1.0)))
    eV = float(np.max(np.abs(res["V"] - ref["V"]) / np.abs(ref["V"])))
    eS = float(np.max(np.abs(np.atleast_1d(res["v_sum"]) - np.atleast_1d(ref["v_sum"])) / np.abs(np.atleast_1d(ref["v_sum"]))))
    eE = max(abs(res[k] - ref[k]) / max(1.0, abs(ref[k])) for k in ("sym", "ssexp", "slogv"))
    good = eM < 1e-8 and eV < 1e-8 and eS < 1e-8 and eE < 1e-8 and res["n_updates"] == ref["n_updates"]
    ok = ok and good
    print("group x%d" % ng, (n, p, K, vt, mi, tol), "u", res["n_updates"], ref["n_updates"],
          "errM %.1e errV %.1e errVsum %.1e errELBO %.1e  %.3f s" % (eM, eV, eS, eE, dt), "OK" if good else "FAIL", flush=True)
grp.close()
sys.exit(0 if ok else 1)
