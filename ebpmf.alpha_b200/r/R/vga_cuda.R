# vga_cuda.R -- R wrappers with the reference's UNCHANGED signatures over the .Call shim
# (src/r_glue.c -> include/ebpmf_b200.h).  Drop this file into the package's R/ directory in
# place of the bodies in R/vga_solver_mat.R:6-109 and R/ebpmf_log.R:406-447,470-472.
# There is no CPU fallback: without a usable B200 the .Call raises an R error.
#
# NOT run in this repository (R is not installed); the Python mirror ebpmf.alpha_b200/solvers.py
# marshals exactly the same arguments into the same C entry points and IS tested.

.ebpmf_b200_env = new.env(parent = emptyenv())

#' one GPU context per R session (external pointer with a finalizer)
ebpmf_b200_ctx = function(device = as.integer(Sys.getenv("EBPMF_B200_DEVICE", "0"))){
  if(is.null(.ebpmf_b200_env$ctx)){
    .ebpmf_b200_env$ctx = .Call(C_ebpmf_ctx_new, as.integer(device))
  }
  .ebpmf_b200_env$ctx
}

.as_num_or_dgC = function(Y){
  if(is(Y,'dgCMatrix')) return(Y)
  if(is(Y,'sparseMatrix')) return(as(Y,'dgCMatrix'))
  Y = as.matrix(Y); storage.mode(Y) = 'double'; Y
}

#'@title a matrix version of the vga solver using Newton's method   (R/vga_solver_mat.R:6)
vga_pois_solver_mat_newton = function(M,X,S,Beta,Sigma2,maxiter=1000,tol=1e-8,return_V = TRUE){
  storage.mode(M) = 'double'
  .Call(C_vga_pois_solver_mat_newton, ebpmf_b200_ctx(), M, .as_num_or_dgC(X), as.double(S),
        as.double(Beta), as.double(Sigma2), as.integer(maxiter), as.double(tol), as.logical(return_V))
}

#'@title vga solver for fixed iteration, not necessary to convergence   (R/vga_solver_mat.R:44)
vga_pois_solver_mat_newton_fixed_iter = function(M,V,Y,S,Beta,Sigma2,maxiter=1000){
  storage.mode(M) = 'double'
  .Call(C_vga_pois_solver_mat_newton_fixed_iter, ebpmf_b200_ctx(), M,
        if(is.null(V)) NULL else matrix(as.double(V),nrow(M),ncol(M)), as.matrix(Y) + 0,
        as.double(S), as.double(Beta), as.double(Sigma2), as.integer(maxiter))
}

#'@title low memory mode   (R/vga_solver_mat.R:71)
vga_pois_solver_mat_newton_low_memory = function(M,Y,l0,f0,EL,EF,
                                                 sigma2,var_type,
                                                 maxiter=1000,tol=1e-8){
  storage.mode(M) = 'double'
  .Call(C_vga_pois_solver_mat_newton_low_memory, ebpmf_b200_ctx(), M, .as_num_or_dgC(Y),
        rep_len(as.double(l0), nrow(M)), rep_len(as.double(f0), ncol(M)),
        matrix(as.double(EL), nrow(M)), matrix(as.double(EF), ncol(M)),
        as.double(sigma2), var_type, as.integer(maxiter), as.double(tol))
}

#'@title Calc elbo   (R/ebpmf_log.R:406)
#' sv, sigma2 and R2 go down with their own lengths: each sum() of R/ebpmf_log.R:407 is taken at its own length
calc_ebpmf_log_obj = function(n,p,sym,ssexp,slogv,sv,sigma2,R2,KL_LF,const,ss,a0,b0){
  .Call(C_calc_ebpmf_log_obj, ebpmf_b200_ctx(), n, p, sym, ssexp, slogv, as.double(sv),
        as.double(sigma2), as.double(R2), KL_LF, const, ss, a0, b0)
}

#'@title calc splitting PMF objective function.   (inst/code/Poisson_MF_splitting.R:373, archived driver)
calc_split_PMF_obj_flashier = function(Y,S,sigma2,M,V,fit_flash,KL_LF,const,var_type){
  Y = as.matrix(Y); storage.mode(Y) = 'double'; storage.mode(M) = 'double'; storage.mode(V) = 'double'
  .Call(C_calc_split_PMF_obj_flashier, ebpmf_b200_ctx(), Y, as.double(S), as.double(sigma2), M, V,
        as.double(fit_flash$flash.fit$R2), KL_LF, const, var_type)
}

#'@title update sigma2   (R/ebpmf_log.R:413)
ebpmf_log_update_sigma2 = function(fit_flash,sigma2,v_sum,var_type,cap_var_mean_ratio,a0,b0,n,p){
  R2 = fit_flash$flash_fit$R2
  if(var_type=='constant') R2 = sum(R2)
  EL = NULL; EF = NULL
  if(cap_var_mean_ratio>0){ EL = fit_flash$flash_fit$EF[[1]]; EF = fit_flash$flash_fit$EF[[2]] }
  .Call(C_ebpmf_log_update_sigma2, ebpmf_b200_ctx(), as.double(sigma2), as.double(v_sum), as.double(R2),
        var_type, as.double(cap_var_mean_ratio), as.double(a0), as.double(b0), n, p, EL, EF)
}

#'@title calculate sum of log factorial of all elements of a sparse Matrix   (R/ebpmf_log.R:470)
sum_lfactorial_sparseMat = function(Y){
  .Call(C_ebpmf_sum_lfactorial, ebpmf_b200_ctx(), as.double(Y@x))
}

#' The fused VGA step for ebpmf_log's outer loop: replaces R/ebpmf_log.R:302-327 (see INTEGRATION.md).
#' Y must have been uploaded once with ebpmf_b200_set_y(Y) (it is constant across iterations).
ebpmf_b200_set_y = function(Y){ invisible(.Call(C_ebpmf_ctx_set_y, ebpmf_b200_ctx(), .as_num_or_dgC(Y))) }
ebpmf_b200_vga_step = function(M,EL,EF,sigma2,var_type,maxiter,tol,want_V=FALSE){
  storage.mode(M) = 'double'
  .Call(C_ebpmf_vga_step, ebpmf_b200_ctx(), M, EL + 0, EF + 0, as.double(sigma2), var_type,
        as.integer(maxiter), as.double(tol), as.logical(want_V))
}

#' Resident-M variant used by ebpmf_log's loop: upload M_init once, then every iteration sends only
#' the factors and sigma2 and receives res$M (R feeds res$M back unchanged, R/ebpmf_log.R:309).
ebpmf_b200_set_m = function(M){ storage.mode(M) = 'double'; invisible(.Call(C_ebpmf_ctx_set_m, ebpmf_b200_ctx(), M)) }
ebpmf_b200_vga_step_resident = function(n,p,EL,EF,sigma2,var_type,maxiter,tol,want_V=FALSE){
  .Call(C_ebpmf_vga_step_resident, ebpmf_b200_ctx(), as.integer(c(n,p)), EL + 0, EF + 0, as.double(sigma2), var_type,
        as.integer(maxiter), as.double(tol), as.logical(want_V))
}

#' Batched branch of ebpmf_log's loop (general_control$batch_size < n): replaces R/ebpmf_log.R:254-293.
#' `batches` is the list made at R/ebpmf_log.R:197 (contiguous ranges); Y is the ORIGINAL dgCMatrix, uploaded
#' once; every call streams M through the GPU batch by batch with the stop test scoped to each batch.
ebpmf_b200_batched_new = function(Y, batches, device = 0L, nslots = 3L){
  row0 = c(vapply(batches, function(b) b[1] - 1, 0), nrow(Y))
  h = .Call(C_ebpmf_batched_new, as(Y,'dgCMatrix'), as.double(row0), as.integer(device), as.integer(nslots))
  attr(h, 'n_batch') = length(batches)
  h
}
ebpmf_b200_batched_vga_step = function(h,M,EL,EF,sigma2,var_type,maxiter,tol,want_V=FALSE){
  storage.mode(M) = 'double'
  .Call(C_ebpmf_batched_vga_step, h, M, EL + 0, EF + 0, as.double(sigma2), var_type,
        as.integer(maxiter), as.double(tol), as.logical(want_V), as.integer(attr(h,'n_batch')))
}

#' N1: what flashier needs from the RESIDENT M, computed on the GPU (M need not come back to the host):
#' list(MF = M %*% EF_w, MtL = t(M) %*% EL_w, r2_col, r2_row) with r2 = col/row sums of (M - EL_w EF_w')^2,
#' i.e. flash_fit$R2 for var_type by_col / by_row (R/ebpmf_log_flash_update.R:35,55-64).
ebpmf_b200_m_products = function(n,p,EL_w,EF_w){
  .Call(C_ebpmf_m_products, ebpmf_b200_ctx(), as.integer(c(n,p)), EL_w + 0, EF_w + 0)
}

#' N4: sim_data_log (R/simulate_poisson_matrix.R:9-61) generated on the GPU, straight into the context's
#' resident state; returns the dgCMatrix Y and the solver state (the start M stays on the device).
ebpmf_b200_sim_data_log = function(n,p,K,seed=12345,log_offset=0,background_unif_range=c(1,2),pi0=0.8,noise=0.2,
                                   warm_start=FALSE){
  r = .Call(C_ebpmf_sim_data_log, ebpmf_b200_ctx(), n, p, K, seed, as.double(log_offset),
            as.double(background_unif_range), as.double(pi0), as.double(noise), as.logical(warm_start))
  nnz = r$nnz
  list(Y = new('dgCMatrix', i = r$i[seq_len(nnz)], p = r$p, x = r$x[seq_len(nnz)], Dim = r$Dim),
       EL = r$EL, EF = r$EF, sigma2 = r$sigma2)
}

#' Several GPUs from ONE R session (EBPMF_B200_NGPU > 1): rows sharded over the devices, the update count and
#' v_sum / ELBO partials combined over NCCL inside the library.  Same calls, same results as the single-GPU
#' wrappers above; `ebpmf_b200_use_group()` tells ebpmf_log's loop which set to call (INTEGRATION.md section 3).
ebpmf_b200_ngpu = function() as.integer(Sys.getenv("EBPMF_B200_NGPU", "1"))
ebpmf_b200_use_group = function() ebpmf_b200_ngpu() > 1L
ebpmf_b200_group = function(devices = seq_len(ebpmf_b200_ngpu()) - 1L){
  if(is.null(.ebpmf_b200_env$group)){
    .ebpmf_b200_env$group = .Call(C_ebpmf_group_new, as.integer(devices))
  }
  .ebpmf_b200_env$group
}
ebpmf_b200_group_set_y = function(Y){ invisible(.Call(C_ebpmf_group_set_y, ebpmf_b200_group(), as(Y,'dgCMatrix'))) }
ebpmf_b200_group_sum_lfactorial = function(){ .Call(C_ebpmf_group_sum_lfactorial, ebpmf_b200_group()) }
ebpmf_b200_group_set_m = function(M){ storage.mode(M) = 'double'; invisible(.Call(C_ebpmf_group_set_m, ebpmf_b200_group(), M)) }
#' M = NULL: the shards already hold the input (set_m once, then every step's output, R/ebpmf_log.R:309)
ebpmf_b200_group_vga_step = function(M,n,p,EL,EF,sigma2,var_type,maxiter,tol,want_V=FALSE){
  if(!is.null(M)) storage.mode(M) = 'double'
  .Call(C_ebpmf_group_vga_step, ebpmf_b200_group(), M, as.integer(c(n,p)), EL + 0, EF + 0, as.double(sigma2), var_type,
        as.integer(maxiter), as.double(tol), as.logical(want_V))
}
