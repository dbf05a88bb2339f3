"""Time the resident VGA step for the library in $EBPMF_B200_LIB (A/B of kernel variants / options).

    python scripts/ab_run.py [c4slab|c3|c2] [freeze=0|1] [history=0|1] [perturb=0|1]
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import __graft_entry__ as ge
pkg = ge.load_package()
ctx = pkg.VgaContext(0)
wl = sys.argv[1] if len(sys.argv) > 1 else "c4slab"
opts = dict(a.split("=") for a in sys.argv[2:])
cfg = {"c4slab": (125000, 30000, 20, -5.3828), "c3": (100000, 20000, 10, -4.6440), "c2": (10000, 2000, 5, -3.4256)}[wl]
n, p, K, off = cfg
nnz = ctx.sim_data_log(n, p, K, seed=12345, log_offset=off)
for k in ("freeze", "history"):
    if k in opts:
        ctx.set_option(k, float(opts[k]))
units = ((n + 31) // 32) * ((p + 1) // 2)
for rep in range(5):
    i = ctx.vga_step(10, 1e-5); ctx.rewind_m()
    c = ctx.counters()
    print(wl, opts, "rep", rep, dict(u=i.n_updates, ms=round(i.kernel_ms, 2), p1=round(i.pass1_ms, 2), p2=round(i.pass2_ms, 2),
                                     rd=round(i.reduce_ms, 2), ge_s=round(n * p / i.kernel_ms / 1e6, 2)),
          "dens=%.4f topup_frac=%.4f topup_tiles=%d/%d evals/unit=%.2f" % (nnz / n / p, c["topup_units"] / units, c["topup_tiles"],
                                                                        c["tiles"], c["unit_evals"] / units), flush=True)
    if rep == 2 and opts.get("perturb", "0") == "1":
        EL, EF = ctx.get_factors()
        rng = np.random.default_rng(1)
        EL[:, 2:] += 0.05 * rng.normal(size=EL[:, 2:].shape); EF[:, 2:] += 0.05 * rng.normal(size=EF[:, 2:].shape)
        ctx.set_factors(EL, EF)
        print("factors perturbed (sd 0.05)")
ctx.close()
