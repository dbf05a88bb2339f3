"""CPU tests of the host-side mirror: marshalling of R-recycled operands, dgCMatrix slots,
row sharding, and the multi-rank combine logic over gloo (world_size 2)."""
import os
import socket
import sys

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import c_oracle, simdata, vga_numpy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_operand_kinds(pkg):
    from ebpmf_alpha_b200 import solvers, _lib
    n, p = 5, 3
    assert solvers._operand(1, n, p, "S")[1] == _lib.OP_SCALAR
    assert solvers._operand(np.ones(n), n, p, "S")[1] == _lib.OP_ROWVEC
    buf, kind = solvers._operand(np.arange(15.0).reshape(n, p), n, p, "S")
    assert kind == _lib.OP_MATRIX and buf.flags.f_contiguous and buf[2, 1] == 7.0
    with pytest.raises(ValueError):
        solvers._operand(np.ones(p), n, p, "S")          # R would recycle with a warning; we refuse
    with pytest.raises(ValueError):
        solvers._operand(np.ones((n, p + 1)), n, p, "S")


def test_csc_slots_are_dgcmatrix_shaped(pkg):
    from ebpmf_alpha_b200 import solvers
    rng = np.random.default_rng(0)
    Y = rng.poisson(0.3, size=(40, 7)).astype(float)
    cp, ri, x = solvers._csc_slots(sp.coo_matrix(Y))
    assert cp.dtype == np.int32 and ri.dtype == np.int32 and x.dtype == np.float64
    assert cp[0] == 0 and cp[-1] == np.count_nonzero(Y)
    for j in range(7):
        rows = ri[cp[j]:cp[j + 1]]
        assert np.all(np.diff(rows) > 0)
        assert np.array_equal(Y[rows, j], x[cp[j]:cp[j + 1]])


def test_adjust_var_shape_matches_oracle(pkg):
    s = np.array([1.0, 2.0, 3.0])
    for vt, n, p in (("by_col", 4, 3), ("by_row", 3, 5), ("constant", 2, 2)):
        assert np.array_equal(pkg.adjust_var_shape(s if vt != "constant" else s[:1], vt, n, p),
                              vga_numpy.adjust_var_shape(s if vt != "constant" else s[:1], vt, n, p))


@pytest.mark.parametrize("n,world", [(1000, 1), (1000, 2), (125000, 8), (100, 8), (257, 2), (1000000, 8)])
def test_shard_rows(pkg, n, world):
    sh = pkg.shard_rows(n, world)
    assert len(sh) == world and sh[0][0] == 0
    assert sum(r for _, r in sh) == n
    for (a, ra), (b, _) in zip(sh, sh[1:]):
        assert a + ra == b
    assert all(r % 256 == 0 for _, r in sh[:-1] if r and _ + r < n)


@pytest.mark.parametrize("n,batch_size", [(1000, 300), (1000, 1000), (1000, 5000), (7, 2), (100, 33), (12345, 1000),
                                          (5, 1), (2, 1), (513, 256), (4097, 64)])
def test_r_batches_equals_r_cut(pkg, n, batch_size):
    """Host mirror of `split(1:n, cut(seq_along(1:n), n_batch, labels=FALSE))` (R/ebpmf_log.R:195-197): contiguous
    ranges whose starts equal the oracle's restatement of base::cut."""
    row0 = pkg.r_batches(n, batch_size)
    n_batch = int(np.ceil(n / batch_size))
    assert row0[0] == 0 and row0[-1] == n and np.all(np.diff(row0) > 0)
    if n_batch <= 1:
        assert list(row0) == [0, n]
        return
    ref = vga_numpy.r_cut_batches(n, n_batch)
    assert [int(b[0]) for b in ref] + [n] == list(row0)
    assert all(np.array_equal(b, np.arange(b[0], b[-1] + 1)) for b in ref)


def test_shard_csc_equals_row_slice(pkg):
    rng = np.random.default_rng(1)
    Y = sp.csc_matrix(rng.poisson(0.2, size=(700, 31)).astype(float))
    Y.sort_indices()
    for row0, nrows in pkg.shard_rows(700, 3):
        cp, ri, x = pkg.shard_csc(Y.indptr, Y.indices, Y.data, row0, nrows)
        got = sp.csc_matrix((x, ri, cp), shape=(nrows, 31)).toarray()
        assert np.array_equal(got, Y[row0:row0 + nrows].toarray())


# ---- world_size-2 gloo: the combine that the library does over NCCL ------------------------
def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close(); return port


def _rank_main(rank, world, port, q):
    import torch.distributed as dist
    import torch
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    pkg = ge.load_package()
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    sim = simdata.sim_data_log(600, 64, K=3, seed=12345, log_offset=-3.0)
    st = simdata.solver_state(sim)
    row0, nrows = pkg.shard_rows(600, world, align=32)[rank]
    sl = slice(row0, row0 + nrows)
    Beta = st["EL"][sl] @ st["EF"].T
    Sigma2 = vga_numpy.adjust_var_shape(st["sigma2"], "by_col", nrows, 64)
    # pass 1: the shard's own stop count; MAX over ranks = lower bound K* of the global count
    info = {}
    vga_numpy.vga_pois_solver_mat_newton(st["M0"][sl], sim["Y"][sl], 1, Beta, Sigma2, 10, 1e-5, info=info)
    k = torch.tensor([info["n_updates"]], dtype=torch.int32)
    dist.all_reduce(k, op=dist.ReduceOp.MAX)
    # pass 2: every shard does exactly K* updates, verifies max|f| < tol at K* (MAX of the fail flag)
    K = int(k.item())
    r = vga_numpy.vga_pois_solver_mat_newton(st["M0"][sl], sim["Y"][sl], 1, Beta, Sigma2, K, 0.0) if K > 0 else None
    M = r["M"] if r is not None else np.minimum(st["M0"][sl], Sigma2 * sim["Y"][sl] + Beta)
    temp = Sigma2 * sim["Y"][sl] + Beta + 1 - M
    V = Sigma2 / temp
    f = sim["Y"][sl] - np.exp(M + V / 2) - (M - Beta) / Sigma2
    fail = torch.tensor([int(not (np.max(np.abs(f)) < 1e-5))], dtype=torch.int32)
    dist.all_reduce(fail, op=dist.ReduceOp.MAX)
    s = vga_numpy.elbo_partials(sim["Y"][sl], M, V)
    red = torch.tensor(np.concatenate([V.sum(0), s]))
    dist.all_reduce(red, op=dist.ReduceOp.SUM)
    if rank == 0:
        q.put(dict(K=K, fail=int(fail.item()), red=red.numpy()))
    dist.destroy_process_group()


def test_two_rank_combine_equals_single_rank():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    got = q.get(timeout=120)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    sim = simdata.sim_data_log(600, 64, K=3, seed=12345, log_offset=-3.0)
    st = simdata.solver_state(sim)
    whole = c_oracle.vga_step(st["M0"], sim["Y"], st["EL"], st["EF"], st["sigma2"], "by_col", 10, 1e-5)
    assert got["K"] == whole["n_updates"] and got["fail"] == 0
    assert np.max(np.abs(got["red"][:64] - whole["v_sum"]) / whole["v_sum"]) < 1e-12
    for i, k in enumerate(("sym", "ssexp", "slogv")):
        assert abs(got["red"][64 + i] - whole[k]) < 1e-12 * abs(whole[k])
