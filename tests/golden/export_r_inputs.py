"""Writes the INPUTS of tests/golden/vga_golden.npz as raw little-endian float64 files that base R can read
(readBin), for tests/golden/make_golden.R -- the script a maintainer with R and the reference package runs to
produce REFERENCE outputs for exactly these inputs (the only route from "parity unpinned" to pinned parity:
R cannot run in this repository's build or GPU images).
Run from the repo root:  python tests/golden/export_r_inputs.py"""
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "vga_golden.npz"))
OUT = os.path.join(HERE, "r_inputs")
os.makedirs(OUT, exist_ok=True)
names = ["Y", "L0", "F0"] + ["%s_%s" % (vt, k) for vt in ("by_col", "by_row", "constant") for k in ("M0", "EL", "EF", "sigma2", "R2")]
with open(os.path.join(OUT, "manifest.tsv"), "w") as mf:
    mf.write("name\tnrow\tncol\n")
    for nm in names:
        a = np.asarray(G[nm], dtype="<f8")
        a2 = a.reshape(a.shape[0], -1) if a.ndim else a.reshape(1, 1)
        np.asfortranarray(a2).T.tofile(os.path.join(OUT, nm + ".f64"))      # column-major bytes
        mf.write("%s\t%d\t%d\n" % (nm, a2.shape[0], a2.shape[1]))
print("wrote", len(names), "arrays to", OUT)
