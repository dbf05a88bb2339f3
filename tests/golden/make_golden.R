# make_golden.R -- REFERENCE outputs for the committed inputs of tests/golden/r_inputs/.
#
# This repository's images have no R, so its oracle (oracle/) is pinned only by internal evidence
# ("parity unpinned", DESIGN.md section 5).  A maintainer who has R and the reference package closes that gap:
#
#     Rscript tests/golden/make_golden.R /path/to/ebpmf.alpha        # the reference checkout (DongyueXie/ebpmf.alpha)
#
# It sources the reference's OWN files (R/vga_solver_mat.R, R/ebpmf_log.R -- needs Rfast and Matrix, as the
# package does), runs the functions of the hot path on the inputs in tests/golden/r_inputs/ and writes their
# results as raw little-endian doubles to tests/golden/r_reference/ (+ manifest.tsv).  tests/test_golden_r.py then
# checks the oracle (CPU) and the CUDA path (GPU) against those files to 1e-8, update counts exactly; without
# the directory the test is skipped and says why.  Nothing here is a re-implementation: every number comes out
# of the reference's functions, except sym/ssexp/slogv/v_sum, which the reference computes inline in ebpmf_log
# (R/ebpmf_log.R:316-327) and which are restated here line by line with the citation.
#
# Nobody here has GNU R, but the script is not untested: tests/test_golden_r.py runs it, unmodified, under the restricted
# R evaluator oracle/mini_r.py (base R's file I/O stood in for by oracle/minir_io.py) against the reference checkout and
# feeds what it writes to the same loader.
args = commandArgs(trailingOnly = TRUE)
if (length(args) < 1) stop("usage: Rscript tests/golden/make_golden.R /path/to/ebpmf.alpha")
ref = args[1]
here = dirname(normalizePath(sub("--file=", "", grep("--file=", commandArgs(FALSE), value = TRUE)[1])))
suppressPackageStartupMessages({ library(Matrix); library(Rfast); library(matrixStats) })
source(file.path(ref, "R", "vga_solver_mat.R"))
source(file.path(ref, "R", "ebpmf_log.R"))

fitted.standin = function(object, ...) tcrossprod(object$L_pm, object$F_pm)     # see the cap gate below
colMaxs = matrixStats::colMaxs                                                  # as the package's NAMESPACE resolves them:
rowMaxs = Rfast::rowMaxs                                                        # importFrom(matrixStats,colMaxs), importFrom(Rfast,rowMaxs)
inp = file.path(here, "r_inputs"); outd = file.path(here, "r_reference")
dir.create(outd, showWarnings = FALSE)
man = read.delim(file.path(inp, "manifest.tsv"), stringsAsFactors = FALSE)
rd = function(nm) {
  r = man[man$name == nm, ]
  x = readBin(file.path(inp, paste0(nm, ".f64")), what = "double", n = r$nrow * r$ncol, size = 8, endian = "little")
  if (r$ncol == 1) x else matrix(x, r$nrow, r$ncol)
}
out_man = data.frame(name = character(), nrow = integer(), ncol = integer(), stringsAsFactors = FALSE)
wr = function(nm, x) {
  x = as.matrix(x); storage.mode(x) = "double"
  writeBin(as.vector(x), file.path(outd, paste0(nm, ".f64")), size = 8, endian = "little")
  out_man[nrow(out_man) + 1, ] <<- list(nm, nrow(x), ncol(x))
}

Y = rd("Y"); n = nrow(Y); p = ncol(Y)
cases = list(def = c(10, 1e-5), exh = c(2, 1e-5), tight = c(1000, 1e-8))
for (vt in c("by_col", "by_row", "constant")) {
  M0 = rd(paste0(vt, "_M0")); EL = rd(paste0(vt, "_EL")); EF = rd(paste0(vt, "_EF"))
  sigma2 = rd(paste0(vt, "_sigma2")); R2 = rd(paste0(vt, "_R2"))
  Beta = tcrossprod(EL, EF)                                     # fitted(fit_flash), R/ebpmf_log.R:305
  S2 = adjust_var_shape(sigma2, vt, n, p)                       # R/ebpmf_log.R:306, :475-486
  for (tag in names(cases)) {
    maxiter = cases[[tag]][1]; tol = cases[[tag]][2]
    res = vga_pois_solver_mat_newton(M0, Y, 1, Beta, S2, maxiter = maxiter, tol = tol, return_V = TRUE)   # :302-308
    # the update count is not returned by the reference; it is the smallest k whose maxiter = k run ends at the same M
    u = maxiter
    for (k in seq_len(maxiter)) {
      Mk = vga_pois_solver_mat_newton(M0, Y, 1, Beta, S2, maxiter = k, tol = tol, return_V = FALSE)
      if (identical(Mk, res$M)) { u = k; break }
    }
    M1 = vga_pois_solver_mat_newton(M0, Y, 1, Beta, S2, maxiter = 1, tol = Inf, return_V = FALSE)         # tol = Inf: breaks at i = 1
    if (identical(M1, res$M)) u = 0
    sym = sum(Y * res$M)                                        # R/ebpmf_log.R:316
    ssexp = sum(exp(res$M + res$V / 2))                         # :317
    slogv = sum(log(res$V) / 2 + 0.9189385)                     # :318
    v_sum = if (vt == "by_col") colSums(res$V) else if (vt == "by_row") rowSums(res$V) else sum(res$V)    # :321-327
    k0 = paste0(vt, "_", tag, "_")
    wr(paste0(k0, "M"), res$M); wr(paste0(k0, "V"), res$V); wr(paste0(k0, "v_sum"), v_sum)
    wr(paste0(k0, "scal"), c(sym, ssexp, slogv, u))
    if (tag == "def") {
      fit_flash = list(flash_fit = list(R2 = R2))
      s2new = ebpmf_log_update_sigma2(fit_flash, sigma2, v_sum, vt, 0, 1, 1, n, p)                         # :329-332, :413-447
      ss = if (vt == "by_col") n else if (vt == "by_row") p else n * p                                     # var_offset_for_obj, :178-192
      obj = calc_ebpmf_log_obj(n, p, sym, ssexp, slogv, v_sum, s2new, R2, -123.4, -5.5, ss, 1, 1)          # :355, :406-410
      wr(paste0(k0, "sigma2_new"), s2new); wr(paste0(k0, "obj"), obj)
      # the cap gate (:415-419, :424-429, :435-440) calls fitted(fit_flash): give the stand-in list a class with a fitted()
      # method = what fitted.flash returns, L_pm %*% t(F_pm).  by_row goes through Rfast::rowMaxs (positions!), by_col
      # through matrixStats::colMaxs (values); 1e9 puts the threshold near 20 so that the by_row gate selects rows.
      fit_cap = list(flash_fit = list(R2 = R2), L_pm = EL, F_pm = EF); class(fit_cap) = "standin"
      wr(paste0(k0, "sigma2_new_cap"), ebpmf_log_update_sigma2(fit_cap, sigma2, v_sum, vt, 2, 1, 1, n, p))
      wr(paste0(k0, "sigma2_new_cap_big"), ebpmf_log_update_sigma2(fit_cap, sigma2, v_sum, vt, 1e9, 1, 1, n, p))
    }
  }
}
# the two other solver signatures and the lfactorial constant (by_col state)
M0 = rd("by_col_M0"); EL = rd("by_col_EL"); EF = rd("by_col_EF"); sigma2 = rd("by_col_sigma2")
Beta = tcrossprod(EL, EF); S2 = adjust_var_shape(sigma2, "by_col", n, p)
fi = vga_pois_solver_mat_newton_fixed_iter(M0, NULL, Y, 1, Beta, S2, maxiter = 1)                          # R/vga_solver_mat.R:44-65
wr("fixed1_M", fi$M); wr("fixed1_V", fi$V)
lm = suppressMessages(vga_pois_solver_mat_newton_low_memory(log(Y + 0.5), Matrix(Y, sparse = TRUE), rd("L0"), rd("F0"),
                      EL[, -(1:2)], EF[, -(1:2)], sigma2, "by_col", maxiter = 1000, tol = 1e-8))           # :71-109
wr("lowmem_M", as.matrix(lm))
wr("lfact", sum_lfactorial_sparseMat(Matrix(Y, sparse = TRUE)))                                            # R/ebpmf_log.R:470-472
write.table(out_man, file.path(outd, "manifest.tsv"), sep = "\t", quote = FALSE, row.names = FALSE)
cat("wrote", nrow(out_man), "arrays to", outd, "\n")
cat("R", R.version.string, "| Rfast", as.character(packageVersion("Rfast")), "| Matrix", as.character(packageVersion("Matrix")), "\n",
    file = file.path(outd, "PROVENANCE.txt"))
