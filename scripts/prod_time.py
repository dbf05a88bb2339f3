"""Time ebpmf_ctx_m_products on the config-4 slab for a few factor counts (A/B of kernel variants via $EBPMF_B200_LIB)."""
import sys; sys.path.insert(0, ".")
import __graft_entry__ as ge, numpy as np
pkg = ge.load_package(); ctx = pkg.VgaContext(0)
ctx.sim_data_log(125000, 30000, 20, seed=12345, log_offset=-5.3828)
EL, EF = ctx.get_factors()
Ks = [int(k) for k in sys.argv[1:]] or [22, 8, 16, 32]
for K in Ks:
    A = EL[:, :K] if K <= 22 else np.hstack([EL, EL[:, :K - 22]]); B = EF[:, :K] if K <= 22 else np.hstack([EF, EF[:, :K - 22]])
    for _ in range(3):
        r = ctx.m_products(A, B)
    print("K", K, "ms", round(r["kernel_ms"], 2), "kernel", round(r["products_ms"], 2), "sums", round(r["partial_sums_ms"], 2), "TF",
          round(4 * K * 3.75e9 / r["products_ms"] / 1e9, 2), "GB/s", round(30e3 / r["products_ms"], 1), flush=True)
