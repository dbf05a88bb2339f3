"""Developer probe run under gpurun: hardware facts, FP64 peak, a first timing of the fused step."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import __graft_entry__ as ge

pkg = ge.load_package()
out = {}
try:
    import psutil
    out["host"] = dict(cpus=os.cpu_count(), ram_gb=psutil.virtual_memory().total / 2**30)
except Exception as e:
    out["host"] = str(e)
ctx = pkg.VgaContext(0)
out["fp64_tflops"] = [ctx.probe_fp64_tflops() for _ in range(3)]
for (n, p, K, off) in [(20000, 4000, 10, -3.4), (125000, 30000, 20, -4.0)]:
    try:
        t = time.time(); nnz = ctx.sim_data_log(n, p, K, seed=12345, log_offset=off); tg = time.time() - t
        rec = dict(n=n, p=p, K=K, nnz=nnz, density=nnz / (n * p), gen_s=tg, runs=[])
        for (mi, tol) in [(10, 1e-5), (1000, 1e-8)]:
            for rep in range(3):
                info = ctx.vga_step(mi, tol)
                rec["runs"].append(dict(maxiter=mi, tol=tol, u=info.n_updates, conv=info.converged, passes=info.passes,
                                        ms=info.kernel_ms, gentries_s=n * p / info.kernel_ms / 1e6))
                ctx.rewind_m()
        out["%dx%d" % (n, p)] = rec
    except Exception as e:
        out["%dx%d" % (n, p)] = "ERR " + str(e)
print(json.dumps(out, indent=1))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "dev_check.json"), "w"), indent=1)
