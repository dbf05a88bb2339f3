"""Time the resident VGA step for the library in $EBPMF_B200_LIB (A/B of kernel variants / options)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
pkg = ge.load_package()
ctx = pkg.VgaContext(0)
wl = sys.argv[1] if len(sys.argv) > 1 else "c4slab"
cfg = {"c4slab": (125000, 30000, 20, -5.3828), "c3": (100000, 20000, 10, -4.6440), "c2": (10000, 2000, 5, -3.4256)}[wl]
n, p, K, off = cfg
nnz = ctx.sim_data_log(n, p, K, seed=12345, log_offset=off)
res = []
for (mi, tol) in ((10, 1e-5),):
    for rep in range(4):
        i = ctx.vga_step(mi, tol); ctx.rewind_m()
        res.append(dict(u=i.n_updates, ms=round(i.kernel_ms, 2), p1=round(i.pass1_ms, 2), p2=round(i.pass2_ms, 2),
                        rd=round(i.reduce_ms, 2), ge_s=round(n * p / i.kernel_ms / 1e6, 2)))
d = ctx.last_diag()
units = ((n + 31) // 32) * ((p + 1) // 2)
print(os.environ.get("EBPMF_B200_LIB", "default").split("/")[-1], wl, "pipeline=%s delta=%s" % (os.environ.get("EBPMF_B200_PIPELINE", "fused"), os.environ.get("EBPMF_B200_DELTA", "0")),
      "dens=%.4f" % (nnz / n / p), res[-1], "topup_frac=%.3f evals/unit=%.2f" % (d["topup_units"] / units, d["unit_evals"] / units))
