"""ebpmf.alpha_b200 -- B200-native VGA Newton hot path of ebpmf.alpha's ``ebpmf_log``.

Host-side mirror of the reference's R operator interface over the plain C ABI in
``include/ebpmf_b200.h`` (CUDA kernels in ``csrc/``, built into ``lib/``).  The
directory name contains a dot, so load it with ``__graft_entry__.load_package()``
(registers it as module ``ebpmf_alpha_b200``).
"""
from ._lib import EbpmfError, SolveInfo, VAR_TYPE, LIB_PATH, SIGNATURES, load
from .sharding import shard_rows, shard_csc
from .solvers import (host_matrix, host_pool_stats, host_pool_trim, VgaContext, VgaGroup, VgaBatched, r_batches, default_context, vga_pois_solver_mat_newton,
                      vga_pois_solver_mat_newton_fixed_iter, vga_pois_solver_mat_newton_low_memory,
                      ebpmf_log_update_sigma2, calc_ebpmf_log_obj, calc_split_PMF_obj_flashier, sum_lfactorial_sparseMat,
                      adjust_var_shape)

__all__ = ["EbpmfError", "SolveInfo", "VAR_TYPE", "LIB_PATH", "SIGNATURES", "load", "VgaContext", "VgaGroup",
           "VgaBatched", "r_batches", "host_matrix", "host_pool_stats", "host_pool_trim",
           "default_context", "vga_pois_solver_mat_newton", "vga_pois_solver_mat_newton_fixed_iter",
           "vga_pois_solver_mat_newton_low_memory", "ebpmf_log_update_sigma2", "calc_ebpmf_log_obj",
           "calc_split_PMF_obj_flashier", "sum_lfactorial_sparseMat", "adjust_var_shape", "shard_rows", "shard_csc"]
