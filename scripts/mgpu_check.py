"""Multi-GPU parity check, run under torchrun on N GPUs of one box:
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/mgpu_check.py
Rows are sharded over the ranks (one ebpmf_ctx per GPU, the library's own NCCL communicator);
rank 0 compares the gathered M, the all-reduced v_sum / ELBO partials and the global update count
with the CPU oracle on the WHOLE matrix."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, scipy.sparse as sp, torch, torch.distributed as dist
import __graft_entry__ as ge
from oracle import simdata, c_oracle

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
pkg = ge.load_package()
ctx = pkg.VgaContext(local)
uid = [pkg.VgaContext.nccl_unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
ctx.comm_init(rank, world, uid[0])
print("rank", rank, "peers mapped over NVLink:", ctx.peer_count(), flush=True)

ok = True
for (n, p, K, var_type, mi, tol) in [(5000, 700, 5, "by_col", 10, 1e-5), (3001, 333, 3, "by_col", 1000, 1e-8),
                                      (4096, 512, 4, "constant", 10, 1e-5), (2500, 300, 4, "by_row", 3, 1e-5)]:
    sim = simdata.sim_data_log(n, p, K=K, seed=12345, log_offset=simdata.offset_for_density(0.08, K=K))
    st = simdata.solver_state(sim, start="cold", var_type=var_type)
    row0, nrows = pkg.shard_rows(n, world)[rank]
    sl = slice(row0, row0 + nrows)
    Yc = sp.csc_matrix(sim["Y"]); Yc.sort_indices()
    cp, ri, x = pkg.shard_csc(Yc.indptr, Yc.indices, Yc.data, row0, nrows)
    ctx.set_y_csc(nrows, p, cp, ri, x)
    s2 = st["sigma2"][sl] if var_type == "by_row" else st["sigma2"]
    res = ctx.vga_step_host(st["M0"][sl], st["EL"][sl], st["EF"], s2, var_type, mi, tol)
    parts = [None] * world
    dist.all_gather_object(parts, (row0, res["M"], res["v_sum"], res["sym"], res["ssexp"], res["slogv"], res["n_updates"]))
    if rank == 0:
        ref = c_oracle.vga_step(st["M0"], sim["Y"], st["EL"], st["EF"], st["sigma2"], var_type, mi, tol)
        M = np.vstack([q[1] for q in sorted(parts, key=lambda q: q[0])])
        eM = float(np.max(np.abs(M - ref["M"]) / np.maximum(np.abs(ref["M"]), 1.0)))
        if var_type == "by_row":
            vs = np.concatenate([q[2] for q in sorted(parts, key=lambda q: q[0])])
        else:
            vs = np.asarray(parts[0][2])
            assert all(np.array_equal(np.asarray(q[2]), vs) for q in parts), "v_sum differs between ranks"
        eV = float(np.max(np.abs(vs - ref["v_sum"]) / np.abs(ref["v_sum"])))
        eS = max(abs(parts[0][3 + i] - ref[k]) / max(1.0, abs(ref[k])) for i, k in enumerate(("sym", "ssexp", "slogv")))
        same_u = all(q[6] == ref["n_updates"] for q in parts)
        good = eM < 1e-8 and eV < 1e-8 and eS < 1e-8 and same_u
        ok = ok and good
        print("mgpu", world, (n, p, K, var_type, mi, tol), "u", parts[0][6], ref["n_updates"], "errM %.2e errV %.2e errS %.2e" % (eM, eV, eS),
              "OK" if good else "FAIL", flush=True)
flag = torch.tensor([int(ok)], device="cuda")
dist.broadcast(flag, src=0)
ctx.close()
dist.destroy_process_group()
sys.exit(0 if flag.item() else 1)
