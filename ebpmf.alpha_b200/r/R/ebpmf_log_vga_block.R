# ebpmf_log_vga_block.R -- the VGA block of ebpmf_log's outer loop (R/ebpmf_log.R:254-334) on the GPU, as ONE call.
#
# ebpmf_log() keeps its signature and everything around this block (flashier fit, KL, ELBO test, traces).  The
# maintainer replaces the `if(n_batch >1){ ... }else{ ... }` statement of the loop body by the four lines shown in
# INTEGRATION.md section 2; the route -- one GPU with the latent M resident, several GPUs of this R session
# (EBPMF_B200_NGPU > 1), or row batches streamed through the GPU (general_control$batch_size < n) -- is chosen
# here, once, from the same quantities the reference uses.
#
# Run in this repository without GNU R: tests/test_r_boundary.py evaluates this file with oracle/mini_r.py over the
# compiled glue, on the GPU box against the outputs of the reference's own ebpmf_log().

#' Once, before the loop: upload Y (and the start M), pick the route.
#' @param Y        the count matrix as ebpmf_log received it (matrix or sparse Matrix), NOT the per-batch list
#' @param M_init   fit_flash$flash_fit$Y after ebpmf_log_flash_init (R/ebpmf_log.R:216-226)
#' @param batches  the list made at R/ebpmf_log.R:197 when n_batch > 1, else NULL
#' @return list(route, handle, n, p, sum_lfactorial): `const = - blk$sum_lfactorial` replaces R/ebpmf_log.R:172-176
ebpmf_b200_vga_block_init = function(Y, M_init, batches = NULL){
  n = nrow(M_init)
  p = ncol(M_init)
  if(!is.null(batches)){
    h = ebpmf_b200_batched_new(Y, batches)
    # the batched handle keeps Y per batch; the constant is the same sum over all stored counts (R/ebpmf_log.R:173, :470)
    slf = .Call(C_ebpmf_sum_lfactorial, ebpmf_b200_ctx(), as.double(as(Y,'dgCMatrix')@x))
    return(list(route = 'batched', handle = h, n = n, p = p, sum_lfactorial = slf))
  }
  if(ebpmf_b200_use_group()){
    ebpmf_b200_group_set_y(Y)
    ebpmf_b200_group_set_m(M_init)
    return(list(route = 'group', handle = NULL, n = n, p = p, sum_lfactorial = ebpmf_b200_group_sum_lfactorial()))
  }
  ebpmf_b200_set_y(Y)
  ebpmf_b200_set_m(M_init)
  slf = .Call(C_ebpmf_sum_lfactorial, ebpmf_b200_ctx(), NULL)
  list(route = 'resident', handle = NULL, n = n, p = p, sum_lfactorial = slf)
}

#' Every outer iteration: the solve for every entry, the ELBO partial sums, v_sum and the sigma2 update.
#' @param M  only read on the batched route (M lives on the host there: fit_flash$flash_fit$Y); the other routes keep
#'           M on the device, where the previous step left it (R feeds res$M back unchanged, R/ebpmf_log.R:309)
#' @return list(M, V, sym, ssexp, slogv, v_sum, sigma2, n_updates)
ebpmf_b200_vga_block = function(blk, fit_flash, sigma2, var_type, vga_control, sigma2_control, want_V = FALSE){
  EL = fit_flash$flash_fit$EF[[1]]
  EF = fit_flash$flash_fit$EF[[2]]
  if(blk$route == 'batched'){
    res = ebpmf_b200_batched_vga_step(blk$handle, fit_flash$flash_fit$Y, EL, EF, sigma2, var_type,
                                      vga_control$maxiter_vga, vga_control$vga_tol, want_V)
  }else if(blk$route == 'group'){
    res = ebpmf_b200_group_vga_step(NULL, blk$n, blk$p, EL, EF, sigma2, var_type,
                                    vga_control$maxiter_vga, vga_control$vga_tol, want_V)
  }else{
    res = ebpmf_b200_vga_step_resident(blk$n, blk$p, EL, EF, sigma2, var_type,
                                       vga_control$maxiter_vga, vga_control$vga_tol, want_V)
  }
  if(sigma2_control$est_sigma2){
    sigma2 = ebpmf_log_update_sigma2(fit_flash, sigma2, res$v_sum, var_type,
                                     sigma2_control$cap_var_mean_ratio, sigma2_control$a0, sigma2_control$b0, blk$n, blk$p)
  }
  list(M = res$M, V = res$V, sym = res$sym, ssexp = res$ssexp, slogv = res$slogv, v_sum = res$v_sum,
       sigma2 = sigma2, n_updates = res$n_updates)
}
