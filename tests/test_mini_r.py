"""The restricted R evaluator (oracle/mini_r.py, test infrastructure) is itself pinned here:

* known answers of base-R semantics the hot path leans on -- operator precedence, column-major recycling, 1-based /
  negative / logical subscripts, replacement through `$`/`[[`/`[`, lazy arguments, `if (NA)`, long-double `sum()`,
  `cut()`/`split()` -- written down from R's documented behaviour (?Syntax, ?Arithmetic, ?Extract, ?cut);
* every R file of the reference parses (when the checkout is present);
* re-running the reference's source through it reproduces the committed tests/golden/minir_reference.npz (when the
  checkout is present: /root/reference does not exist on the GPU box)."""
import glob
import importlib.util
import os

import numpy as np
import pytest

from oracle import mini_r as R

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("EBPMF_REFERENCE", "/root/reference")
needs_ref = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "R", "vga_solver_mat.R")),
                               reason="the reference checkout is not present (it never is on the GPU box)")


def ev(src, **vars_):
    it = R.Interp()
    for k, v in vars_.items():
        it.set(k, v)
    return it.run(src)


def vals(src, **vars_):
    return R.dense(ev(src, **vars_)).tolist()


@pytest.mark.parametrize("src,expect", [
    ("-2^2", [-4.0]),                        # unary minus binds below ^
    ("2^3^2", [512.0]),                      # ^ is right-associative
    ("-1:3", [-1.0, 0.0, 1.0, 2.0, 3.0]),    # ... and above :
    ("1:3*2", [2.0, 4.0, 6.0]),              # : above *
    ("2*3%%2", [2.0]),                       # %any% above *
    ("1:0", [1.0, 0.0]),                     # the 1:0 trap: counts DOWN
    ("5 %% Inf", [5.0]),                     # iter %% save_fit_every with save_fit_every = Inf
    ("7 %/% 2", [3.0]),
    ("!TRUE & FALSE", [False]),              # ! below comparison, above &
    ("1 < 2 & 3 > 4 | TRUE", [True]),
    ("2^-1", [0.5]),
    ("x = y = 3; x + y", [6.0]),
    ("f = function(a, b = a * 2) a + b; f(2)", [6.0]),                       # a default sees the other arguments
    ("f = function(n, ...) g(...); g = function(x, y) x - y; f(0, 5, y = 2)", [3.0]),
    ("f = function(nrow, ncol) nrow - ncol; f(nc = 1, 5)", [4.0]),           # partial argument names
    ("s = 0; for (i in 1:10) { if (i == 3) next; if (i > 5) break; s = s + i }; s", [12.0]),
    ("if (FALSE) 1 else if (FALSE) 2 else 3", [3.0]),
    ("sum(c(1e16, 1, -1e16))", [1.0]),       # sum() accumulates in long double
    ("max(c(1, NaN, 3))", None),             # NaN propagates through max()
    ("x = c(5, 6, 7); x[-1]", [6.0, 7.0]),
    ("x = c(5, 6, 7); x[c(TRUE, FALSE)]", [5.0, 7.0]),                       # a logical subscript is recycled
    ("x = c(5, 6, 7); x[5]", None),          # past the end: NA
    ("x = c(5, 6, 7); x[5] = 1; length(x)", [5.0]),                          # ... and assignment extends
    ("x = c(); x[2] = 4; x[1]", None),
    ("x = -Inf; x[3] = 2; x[2]", None),
    ("which(c(1, 5, 2) > 1.5)", [2.0, 3.0]),
    ("x = c(1, 2, 3, 4); x[which(x > 2)] = (x * 10)[which(x > 2)]; x", [1.0, 2.0, 30.0, 40.0]),
    ("length(NULL)", [0.0]),
    ("is.null(list(a = 1)$b)", [True]),
    ("l = list(alpha = 1, beta = 2); l$al", [1.0]),                          # $ matches partially
    ("l = list(a = 1); l$b$c = 5; l$b$c", [5.0]),                            # nested replacement creates the path
    ("l = list(a = 1, b = 2); l$a = NULL; length(l)", [1.0]),
    ("l = list(1, 2); l[[2]] = c(7, 8); l[[2]][2]", [8.0]),
    ("l = modifyList(list(a = 1, b = 2), list(b = 3, c = 4)); c(l$a, l$b, l$c)", [1.0, 3.0, 4.0]),
    ("ceiling(96 / Inf)", [0.0]),            # n_batch for the default batch_size
    ("round(2.567, 2)", [2.57]),
])
def test_base_r_known_answers(src, expect):
    got = vals(src)
    if expect is None:
        assert len(got) == 1 and np.isnan(got[0])
    else:
        assert got == expect


def test_matrices_are_column_major_and_vectors_recycle_down_columns():
    m = ev("matrix(1:6, nrow = 2, ncol = 3)")
    assert m.tolist() == [[1, 3, 5], [2, 4, 6]]
    assert ev("matrix(1:6, nrow = 2, ncol = 3, byrow = TRUE)").tolist() == [[1, 2, 3], [4, 5, 6]]
    assert ev("matrix(c(10, 20, 30), nrow = 2, ncol = 3, byrow = T)").tolist() == [[10, 20, 30], [10, 20, 30]]   # adjust_var_shape by_col
    assert ev("matrix(c(10, 20), nrow = 2, ncol = 3, byrow = F)").tolist() == [[10, 10, 10], [20, 20, 20]]       # ... by_row
    assert ev("matrix(1:6, 2, 3) * c(10, 100)").tolist() == [[10, 30, 50], [200, 400, 600]]       # a length-n vector scales ROWS
    assert ev("c(10, 100) / matrix(1, 2, 3)").tolist() == [[10, 10, 10], [100, 100, 100]]
    assert ev("m = matrix(1:6, 2, 3); m[2, ]").tolist() == [2, 4, 6]
    assert ev("m = matrix(1:6, 2, 3); m[, 2, drop = FALSE]").shape == (2, 1)
    assert ev("m = matrix(1:6, 2, 3); m[, -c(1, 2)]").tolist() == [5, 6]
    assert ev("m = matrix(0, 3, 2); m[c(1, 3), ] = matrix(1:4, 2, 2); m").tolist() == [[1, 3], [0, 0], [2, 4]]
    assert ev("colSums(matrix(1:6, 2, 3))").tolist() == [3, 7, 11] and ev("rowSums(matrix(1:6, 2, 3))").tolist() == [9, 12]
    assert ev("tcrossprod(matrix(1:4, 2, 2), matrix(1:6, 3, 2))").tolist() == [[13, 17, 21], [18, 24, 30]]
    # do_matprod's rules for vector operands: tcrossprod(EL[b,], EF) with a one-row batch takes the vector as a ROW
    assert ev("tcrossprod(c(1, 2), matrix(1:6, 3, 2))").tolist() == [[9, 12, 15]]
    assert ev("tcrossprod(c(1, 2, 3), c(1, 10))").tolist() == [[1, 10], [2, 20], [3, 30]]       # two vectors: outer product
    assert ev("crossprod(c(1, 2, 3), matrix(1:6, 3, 2))").tolist() == [[14, 32]]
    with pytest.raises(R.RError, match="non-conformable"):
        ev("tcrossprod(c(1, 2, 3), matrix(1:6, 3, 2))")
    with pytest.raises(R.RError, match="non-conformable"):
        ev("matrix(1, 2, 3) + matrix(1, 3, 2)")


def test_matmul_binds_tighter_than_times_and_diagonal_scales_exactly():
    """`l0*exp(..)%*%Diagonal(p,f0)` is l0 * (exp(..) %*% D): R/vga_solver_mat.R:83,97."""
    got = ev("c(2, 3) * matrix(1, 2, 2) %*% Diagonal(2, c(10, 100))")
    assert got.tolist() == [[20, 200], [30, 300]]
    x = np.array([[0.1, 0.2], [0.3, 0.7]])
    assert np.array_equal(ev("x %*% Diagonal(2, c(3, 1/3))", x=x), x * np.array([3, 1 / 3])[None, :])      # no summation


def test_arguments_are_lazy_like_r():
    """my_ifelse(var_type=='by_row', sigma2[b], adjust_var_shape(...)) (R/ebpmf_log.R:266) only works because the
    branch not taken is never evaluated."""
    assert vals("my_ifelse = function(test, yes, no) { if (test) { return(yes) } else { return(no) } }; my_ifelse(TRUE, 1, stop('boom'))") == [1.0]
    with pytest.raises(R.RError, match="boom"):
        ev("f = function(a) a; f(stop('boom'))")


def test_if_on_a_missing_value_is_an_error_like_r():
    with pytest.raises(R.RError, match="missing value where TRUE/FALSE needed"):
        ev("f = c(1, NaN); if (max(abs(f)) < 1e-5) 1 else 2")


def test_cut_and_split_make_rs_row_batches():
    """split(1:n, cut(seq_along(1:n), n_batch, labels = FALSE)), R/ebpmf_log.R:197: equal-width intervals on
    [1, n] widened by 0.1 % at both ends, right-closed.  n = 10 in 3 -> 4/3/3 (R prints 1 1 1 1 2 2 2 3 3 3)."""
    assert vals("cut(seq_along(1:10), 3, labels = FALSE)") == [1, 1, 1, 1, 2, 2, 2, 3, 3, 3]
    b = ev("split(1:96, cut(seq_along(1:96), 4, labels = FALSE))")
    assert [len(v) for v in b.vals] == [24, 24, 24, 24] and b.names == ["1", "2", "3", "4"] and b.vals[1][0] == 25
    b = ev("split(1:1500, cut(seq_along(1:1500), 4, labels = FALSE))")
    assert [int(v[0]) for v in b.vals] == [1, 376, 751, 1126]


def test_unsupported_constructs_fail_loudly():
    for src in ("quote(x)", "x = 1; x@y", "library(stats); nosuchfunction(1)", "'a' < 'b'"):
        with pytest.raises(R.RError):
            ev(src)


@needs_ref
def test_every_r_file_of_the_reference_parses():
    files = sorted(glob.glob(os.path.join(REF, "R", "*.R")) + glob.glob(os.path.join(REF, "inst", "code", "*.R")) +
                   glob.glob(os.path.join(REF, "tests", "testthat", "*.R")))
    assert len(files) >= 40
    for f in files:
        R.parse(open(f).read())


@needs_ref
def test_rerunning_the_reference_reproduces_the_committed_fixture():
    spec = importlib.util.spec_from_file_location("make_golden_minir", os.path.join(HERE, "golden", "make_golden_minir.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    out, printed, warnings = mod.generate(REF)
    committed = np.load(os.path.join(HERE, "golden", "minir_reference.npz"))
    assert sorted(out) == sorted(committed.files)
    for k, v in out.items():
        a, b = np.asarray(v, dtype=np.float64), committed[k]
        assert a.shape == b.shape, k
        assert np.allclose(a, b, rtol=1e-12, atol=1e-12, equal_nan=True), k
    assert len(printed) == 20 and not warnings      # the print(range(f)) of _low_memory's by_row/constant branch (R/vga_solver_mat.R:99)
    from oracle import minir_reference as mr
    import json
    prov = json.load(open(os.path.join(HERE, "golden", "minir_reference.provenance.json")))
    assert prov["reference_files_sha256"] == mr.provenance(REF)


def test_r_snippets_of_integration_md_are_valid_r():
    """The ```r blocks of INTEGRATION.md (loop excerpts: braces left open are closed here) parse as R."""
    import re
    text = open(os.path.join(os.path.dirname(HERE), "INTEGRATION.md")).read()
    blocks = re.findall(r"```r\n(.*?)```", text, flags=re.S)
    assert len(blocks) >= 4
    for b in blocks:
        code = "\n".join(line.split("#")[0] for line in b.splitlines())
        R.parse(code + "\n" + "}" * (code.count("{") - code.count("}")))
