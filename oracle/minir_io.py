"""Base R's file I/O and data frames for oracle/mini_r.py -- just enough to run tests/golden/make_golden.R, the script
a maintainer with GNU R is asked to run, under the evaluator instead (TEST INFRASTRUCTURE ONLY).

Stood in for: commandArgs, grep(value = TRUE), sub, normalizePath, dirname, file.path, dir.create, read.delim,
data.frame (columns of equal length; `df$col`, `df[rows, ]`, `df[k, ] <- list(...)`, nrow), readBin / writeBin for
doubles, write.table, cat(file =), R.version.string, packageVersion.  Anything else raises."""
import os
import re

import numpy as np

from . import mini_r as R


class RDataFrame:
    def __init__(self, names, cols):
        self.names, self.cols = list(names), [np.asarray(c) for c in cols]

    def nrow(self):
        return len(self.cols[0]) if self.cols else 0

    def r_dim(self):
        return np.array([self.nrow(), len(self.cols)], dtype=np.float64)

    def r_dollar(self, name):
        return self.cols[self.names.index(name)] if name in self.names else None

    def r_index_get(self, idx, drop=True):
        if len(idx) != 2 or idx[1] is not R._MISSING:
            raise R.RError("unsupported: data frame subscript other than df[rows, ]")
        rows = R.resolve_index(idx[0], self.nrow())
        return RDataFrame(self.names, [c[rows] for c in self.cols])

    def r_index_set(self, idx, value):
        if len(idx) != 2 or idx[1] is not R._MISSING or not isinstance(value, R.RList) or len(value) != len(self.cols):
            raise R.RError("unsupported: data frame assignment other than df[k, ] <- list(one value per column)")
        k = int(R.scalar(idx[0])) - 1
        cols = []
        for c, v in zip(self.cols, value.vals):
            x = R.scalar(v)
            c = c.astype(object) if isinstance(x, str) else c.astype(np.float64)
            if k >= len(c):
                c = np.concatenate([c, np.array([None] * (k + 1 - len(c)), dtype=object) if c.dtype == object else np.full(k + 1 - len(c), np.nan)])
            else:
                c = c.copy()
            c[k] = x
            cols.append(c)
        return RDataFrame(self.names, cols)


def install(it, script_path, args):
    d = it.def_builtin
    s = lambda v: v if isinstance(v, str) else R.r_string(v)      # noqa: E731

    def command_args(it_, trailingOnly=False):
        if R.as_logical_scalar(trailingOnly):
            return R.strs(*args)
        return R.strs("R", "--no-echo", "--file=" + script_path, "--args", *args)

    def grep(it_, pattern, x, value=False, **kw):
        hits = [i for i, v in enumerate(R.flat(R.dense(x))) if re.search(s(pattern), v)]
        return R.strs(*[R.flat(R.dense(x))[i] for i in hits]) if R.as_logical_scalar(value) else np.array([i + 1.0 for i in hits])

    def read_delim(it_, file, stringsAsFactors=False, **kw):
        with open(s(file)) as f:
            rows = [line.rstrip("\n").split("\t") for line in f if line.strip()]
        names, rows = rows[0], rows[1:]
        cols = []
        for j in range(len(names)):
            col = [r[j] for r in rows]
            try:
                cols.append(np.array([float(v) for v in col]))
            except ValueError:
                cols.append(R.strs(*col))
        return RDataFrame(names, cols)

    def data_frame(it_, stringsAsFactors=False, **cols):
        return RDataFrame(list(cols), [R.dense(v) for v in cols.values()])

    def read_bin(it_, con, what="numeric", n=1, size=8, endian="little", **kw):
        if s(what) not in ("double", "numeric") or int(R.scalar(size)) != 8:
            raise R.RError("unsupported: readBin of anything but 8-byte doubles")
        a = np.fromfile(s(con), dtype="<f8" if s(endian) == "little" else ">f8")
        return np.asarray(a[:int(R.scalar(n))], dtype=np.float64)

    def write_bin(it_, object, con, size=8, endian="little", **kw):
        if int(R.scalar(size)) != 8:
            raise R.RError("unsupported: writeBin of anything but 8-byte doubles")
        R.flat(R.dense(object)).astype("<f8" if s(endian) == "little" else ">f8").tofile(s(con))
        return None

    def write_table(it_, x, file, sep=" ", quote=True, row_names=True, **kw):
        if R.as_logical_scalar(quote) or R.as_logical_scalar(row_names):
            raise R.RError("unsupported: write.table with quotes or row names")
        with open(s(file), "w") as f:
            f.write(s(sep).join(x.names) + "\n")
            for i in range(x.nrow()):
                f.write(s(sep).join(v if isinstance(v, str) else R.fmt_num(v) for v in (c[i] for c in x.cols)) + "\n")
        return None

    def cat(it_, *a, file=None, sep=" ", **kw):
        text = s(sep).join(" ".join(v if isinstance(v, str) else R.fmt_num(v) for v in R.flat(R.dense(x))) for x in a if x is not None)
        if file is not None:
            with open(s(file), "w") as f:
                f.write(text)
        else:
            it_.echo(text)
        return None

    d("commandArgs", command_args)
    d("grep", grep)
    d("sub", lambda it_, pattern, replacement, x, **kw: R.strs(*[re.sub(s(pattern), s(replacement), v, count=1) for v in R.flat(R.dense(x))]))
    d("normalizePath", lambda it_, path, **kw: R.strs(os.path.abspath(s(path))))
    d("dirname", lambda it_, path: R.strs(os.path.dirname(s(path))))
    d("file.path", lambda it_, *parts: R.strs(os.path.join(*[s(x) for x in parts])))
    d("dir.create", lambda it_, path, **kw: os.makedirs(s(path), exist_ok=True))
    d("read.delim", read_delim)
    d("data.frame", data_frame)
    d("readBin", read_bin)
    d("writeBin", write_bin)
    d("write.table", write_table)
    d("cat", cat)
    d("packageVersion", lambda it_, pkg: R.strs("(stand-in: oracle/mini_r.py)"))
    it.set("R.version.string", "oracle/mini_r.py (restricted R evaluator, NOT GNU R)")


def run_make_golden(script, ref_root, outdir):
    """Runs tests/golden/make_golden.R (unmodified) against the checkout `ref_root`; its outputs land in `outdir`
    (the script writes next to itself, so it is run from a link placed beside `outdir`).  Returns what it printed."""
    work = os.path.dirname(os.path.abspath(outdir))
    if os.path.basename(os.path.abspath(outdir)) != "r_reference":
        raise ValueError("the script names its output directory r_reference")
    os.makedirs(work, exist_ok=True)
    link = os.path.join(work, os.path.basename(script))
    for src, dst in ((os.path.abspath(script), link), (os.path.join(os.path.dirname(os.path.abspath(script)), "r_inputs"), os.path.join(work, "r_inputs"))):
        if not os.path.lexists(dst):
            os.symlink(src, dst)
    printed = []
    it = R.Interp(echo=printed.append)
    install(it, link, [ref_root])
    it.def_builtin("normalizePath", lambda it_, path, **kw: R.strs(os.path.join(work, os.path.basename(R.r_string(path)))))   # keep the link
    it.source(link)
    return "\n".join(printed)
