"""Generates tests/golden/minir_reference.npz: outputs of the REFERENCE'S OWN R SOURCE (R/vga_solver_mat.R,
R/ebpmf_log.R, R/ebpmf_log_controls.R, read from the checkout given as argv[1], default /root/reference), executed
by the restricted R evaluator oracle/mini_r.py on the committed inputs of tests/golden/vga_golden.npz (the same
bytes tests/golden/r_inputs/ holds for GNU R).  See oracle/minir_reference.py for the three levels (functions,
in-place loop blocks, the whole `ebpmf_log()` driver with flashier/vebpm stood in for).

This is NOT output of GNU R (tests/golden/make_golden.R is the script for that; R is not installed here): the
arithmetic is NumPy's under R's evaluation rules.  What it removes is the hand transcription between the
reference's text and the numbers the oracle and the CUDA path are held to.

Run from the repo root:  python tests/golden/make_golden_minir.py [/path/to/ebpmf.alpha]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np                                    # noqa: E402
from oracle import minir_reference as mr              # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def driver_inputs(Y):
    """Offsets and starting values of the driver runs (any deterministic function of Y would do)."""
    n, p = Y.shape
    l0 = np.log(Y.mean(axis=1) + 0.1)
    f0 = np.log((Y.sum(axis=0) + 0.1) / np.exp(l0).sum())
    return l0, f0, np.asfortranarray(np.log(Y + 0.5)), {"by_col": np.full(p, 0.5), "by_row": np.full(n, 0.5), "constant": np.array([0.5])}


def generate(ref_root):
    G = np.load(os.path.join(HERE, "vga_golden.npz"))
    printed = []
    it = mr.load_reference(ref_root, echo=printed.append)
    out = {"fn/" + k: v for k, v in mr.run_functions(it, G).items()}
    Y = np.asarray(G["Y"], dtype=np.float64)
    for vt in ("by_col", "by_row", "constant"):
        state = [np.asarray(G["%s_%s" % (vt, k)], dtype=np.float64) for k in ("M0", "EL", "EF", "sigma2", "R2")]
        for bs in (np.inf, 25):
            for cap in (0.0, 2.0):
                r = mr.run_loop_blocks(it, Y, *state, vt, batch_size=bs, cap_var_mean_ratio=cap)
                if np.isinf(bs):              # the unbatched block IS the function call of run_functions: keep one copy
                    assert np.array_equal(r.pop("M"), out["fn/%s_def_M" % vt]) and np.array_equal(r.pop("V"), out["fn/%s_def_V" % vt])
                if cap > 0:                   # the cap gate only changes sigma2
                    r = {"sigma2": r["sigma2"]}
                for k, v in r.items():
                    out["blk/%s/%s/cap%g/%s" % (vt, "inf" if np.isinf(bs) else int(bs), cap, k)] = np.asarray(v, dtype=np.float64)
    l0, f0, M_init, s2 = driver_inputs(Y)
    for vt in ("by_col", "by_row", "constant"):
        for bs in (np.inf, 25):
            r = mr.run_driver(it, Y, l0, f0, M_init, s2[vt], vt, K=3, outer_iters=4, batch_size=bs, sparse=not np.isinf(bs))
            for k, v in r.items():
                out["drv/%s/%s/%s" % (vt, "inf" if np.isinf(bs) else int(bs), k)] = np.asarray(v, dtype=np.float64)
    return out, printed, it.warnings


if __name__ == "__main__":
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    out, printed, warnings = generate(ref)
    np.savez_compressed(os.path.join(HERE, "minir_reference.npz"), **out)
    with open(os.path.join(HERE, "minir_reference.provenance.json"), "w") as f:
        json.dump({"evaluator": "oracle/mini_r.py (restricted R evaluator on NumPy; NOT GNU R)", "numpy": np.__version__,
                   "reference_files_sha256": mr.provenance(ref), "arrays": len(out),
                   "lines_printed_by_the_reference": len(printed), "warnings": warnings}, f, indent=1)
    print("wrote", len(out), "arrays;", len(printed), "lines printed by the reference's own print() calls")
