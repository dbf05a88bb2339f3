#!/usr/bin/env python
"""bench.py -- VGA Newton entries/sec (fp64, N x p) of the ebpmf_log hot path on B200.

One "step" = one full VGA step of ebpmf_log's outer loop (R/ebpmf_log.R:302-332) on one batch
of synthetic sim_data_log counts: a1 (Newton solve with R's GLOBAL stop rule) + a2 (Beta rebuilt
from the factors) + a4 (ELBO partial sums) + a5 (column sums of V), package defaults
maxiter_vga = 10, vga_tol = 1e-5 (R/ebpmf_log_controls.R:33-36).

Workload (config.workload): the per-GPU row slab of BASELINE.json config 4 -- 1M x 30k, K = 20,
~5 % non-zeros, row-sharded over 8 GPUs = 125 000 x 30 000 per GPU.  The full matrix (240 GB of
M) does not fit one GPU, so N = 1 runs one slab and N GPUs run N slabs (weak scaling; N = 8 is
exactly config 4), with the library's NCCL combine of the update count and of v_sum[p] + ELBO
partials in every step.  Inputs are generated ON THE DEVICE (ebpmf_ctx_sim_data_log) and are
30 GB per GPU, far larger than L2, so no L2 flush is needed between steps.

  value      entries/s over all ranks, inputs resident in HBM, device time (CUDA events on the
             library's stream), max over ranks
  e2e        the same metric through the host-buffer C-ABI call the R .Call shim makes every outer
             iteration, ebpmf_vga_step_resident_host: factors and sigma2 in from pinned host memory,
             step, D2H of M and v_sum (the step's input M is the previous step's output and stays on
             the device, as in R/ebpmf_log.R:309), on a bounded row sample of the same workload
             (PCIe-bound, independent of the row count).  Next to it: e2e_full_roundtrip
             (ebpmf_vga_step_host, M in AND out) and e2e_batched (ebpmf_batched_vga_step_host, the
             batch_size < n branch of R/ebpmf_log.R:254-300, batches pipelined)
  delta_mode, scored_pipeline   the same step with the two options of DESIGN.md section 3b / 4 (never
             the headline value)
  roofline   HBM line for the dominant kernel (first pass); the FP64 pipe is the binding unit
             (DESIGN.md §4), so an "fp64" object with the measured DFMA peak sits beside it
  cpu_baseline  the oracle's C transcription of the R code (OpenMP, all host cores) on a bounded
             row sample -- a restatement of the R code, NOT the R code (R is not installed)

`--impl reference` times that CPU restatement alone (the reference cannot run here).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout carries exactly one JSON line.  Libraries (NCCL's version banner) write to fd 1 behind
# Python's back, so fd 1 is pointed at stderr while the benchmark runs and restored for the result.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    os.dup2(_REAL_STDOUT, 1)
    print(line, flush=True)
    os.dup2(2, 1)


# per-GPU slab of BASELINE.json config 4; log offsets calibrated so that mean(1-exp(-lambda)) hits
# the target density (oracle/simdata.offset_for_density on a 4096^2 sample): K=20 @5 %: -5.3828
WORKLOADS = {
    "c4slab": dict(n=125000, p=30000, K=20, log_offset=-5.3828, density_target=0.05,
                   label="BASELINE config 4 slab: 125000 x 30000 per GPU (1M x 30k over 8 GPUs), K=20 (+2 offset cols)"),
    "c3": dict(n=100000, p=20000, K=10, log_offset=-4.6440, density_target=0.05,
               label="BASELINE config 3: 100000 x 20000, K=10 (+2 offset cols)"),
    "c2": dict(n=10000, p=2000, K=5, log_offset=-3.4256, density_target=0.10,
               label="BASELINE config 2: 10000 x 2000, K=5 (+2 offset cols)"),
}
MAXITER_VGA, VGA_TOL = 10, 1e-5
DELTA = 1e-14
SEED = 12345


def read_peaks():
    try:
        d = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock, power and throttle reasons sampled DURING the timed region (NVML every 50 ms;
    nvidia-smi polling as a fallback)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.stop = index, [], threading.Event()
        self.t = threading.Thread(target=self.run, daemon=True)
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nvml = None

    def sample_nvml(self):
        n = self.nvml
        sm = n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)
        pw = n.nvmlDeviceGetPowerUsage(self.h) / 1000.0
        try:
            r = n.nvmlDeviceGetCurrentClocksEventReasons(self.h)
        except Exception:
            r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        bit = lambda name: "Active" if (r & getattr(n, name, 0)) else "Not Active"
        return [str(sm), str(self.max), str(pw), bit("nvmlClocksThrottleReasonHwSlowdown"),
                bit("nvmlClocksThrottleReasonHwThermalSlowdown"), bit("nvmlClocksThrottleReasonSwThermalSlowdown"),
                bit("nvmlClocksThrottleReasonSwPowerCap")]

    def run(self):
        while not self.stop.is_set():
            try:
                if self.nvml is not None:
                    self.rows.append(self.sample_nvml())
                else:
                    o = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                        "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                    f = [x.strip() for x in o.strip().split(",")]
                    if len(f) >= 7:
                        self.rows.append(f)
            except Exception:
                pass
            self.stop.wait(0.05 if self.nvml is not None else 0.2)

    def __enter__(self):
        self.t.start(); return self

    def __exit__(self, *a):
        self.stop.set(); self.t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["no samples"])
        sm = sorted(float(r[0]) for r in self.rows)
        reasons = []
        for i, name in ((3, "hw_slowdown"), (4, "hw_thermal_slowdown"), (5, "sw_thermal_slowdown"), (6, "sw_power_cap")):
            if any(r[i].lower().startswith("active") for r in self.rows):
                reasons.append(name)
        return dict(sm_mhz=sm[len(sm) // 2], sm_max_mhz=float(self.rows[0][1]), reasons=reasons,
                    power_w_max=max(float(r[2]) for r in self.rows), samples=len(self.rows),
                    source="nvml" if self.nvml is not None else "nvidia-smi")


def host_cores():
    """All host cores this process may use (torchrun exports OMP_NUM_THREADS=1, so ask the OS)."""
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def numa_local_cpus(device):
    """CPUs of the NUMA node the GPU hangs off (sysfs), or None.  Pinned host buffers are placed by first
    touch, so allocating them from a NUMA-local CPU keeps each rank's PCIe traffic on its own socket."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(device)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        return (node, cpus) if cpus else None
    except Exception:
        return None


def cpu_sample_inputs(ctx, rows):
    """Rows [0, rows) of the device-generated workload, pulled to the host for the CPU arm."""
    import scipy.sparse as sp
    n, p, K, nnz = ctx.shape()
    cp, ri, x = ctx.get_y_csc()
    Y = sp.csc_matrix((x, ri, cp), shape=(n, p))[:rows].toarray(order="F")
    EL, EF = ctx.get_factors()
    return dict(Y=Y, EL=np.asfortranarray(EL[:rows]), EF=EF, sigma2=ctx.get_sigma2(), M0=ctx.get_rows(0, rows))


def cpu_baseline(sample, reps=2):
    """The oracle's C transcription of R/ebpmf_log.R:302-327 (OpenMP, all cores) on the sample."""
    from oracle import c_oracle
    rows, p = sample["Y"].shape
    cores = host_cores()
    best, out = None, None
    for _ in range(reps):
        t = time.perf_counter()
        out = c_oracle.vga_step(sample["M0"], sample["Y"], sample["EL"], sample["EF"], sample["sigma2"], "by_col",
                                MAXITER_VGA, VGA_TOL, nthreads=cores)
        dt = time.perf_counter() - t
        best = dt if best is None else min(best, dt)
    # what the R interpreter does: one pass + one temporary per operator, single thread (smaller sample)
    r2 = max(64, rows // 8)
    Beta = sample["EL"][:r2] @ sample["EF"].T
    S2 = np.repeat(sample["sigma2"][None, :], r2, axis=0)
    t = time.perf_counter()
    c_oracle.vga_newton_rstyle(sample["M0"][:r2], sample["Y"][:r2], 1.0, Beta, S2, MAXITER_VGA, VGA_TOL)
    t_r = time.perf_counter() - t
    return dict(value=rows * p / best, unit="entries/s", cores=cores, kind="port",
                sample="rows [0,%d) x %d cols of the same workload, %d Newton updates (slab-local stop); "
                       "C transcription of the R code, OpenMP; R itself is not installed" % (rows, p, out["n_updates"]),
                rstyle_1thread_entries_per_s=r2 * p / t_r,
                rstyle_sample="rows [0,%d), operator-by-operator with temporaries, 1 thread (a1 only)" % r2), out


def run_reference(args, wl, rank, world):
    """--impl reference: the CPU restatement alone, on rank 0 (the R reference cannot run here)."""
    if rank != 0:
        return
    from oracle import c_oracle, simdata
    rows = args.cpu_rows
    p, K = wl["p"], wl["K"]
    sim = simdata.sim_data_log(rows, p, K=K, seed=SEED, log_offset=wl["log_offset"])
    st = simdata.solver_state(sim, start="cold")
    cores = host_cores()
    times = []
    for it in range(args.warmup + args.steps):
        t = time.perf_counter()
        out = c_oracle.vga_step(st["M0"], sim["Y"], st["EL"], st["EF"], st["sigma2"], "by_col", MAXITER_VGA, VGA_TOL,
                                nthreads=cores)
        if it >= args.warmup:
            times.append(time.perf_counter() - t)
    ms = 1e3 * sum(times) / len(times)
    val = rows * p / (ms * 1e-3)
    sample = ("%d x %d row sample of the workload per step (NumPy-generated sim_data_log, same DGP), u=%d; C "
              "transcription of the R code with OpenMP -- R is not installed, this is not the R interpreter"
              % (rows, p, out["n_updates"]))
    emit(json.dumps({
        "impl": "reference", "metric": "VGA Newton entries/sec (fp64, N x p)", "value": val, "unit": "entries/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["label"], "maxiter_vga": MAXITER_VGA, "vga_tol": VGA_TOL, "var_type": "by_col"},
        "cpu_baseline": {"value": val, "unit": "entries/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "entries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c4slab", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-rows", type=int, default=2048, help="rows of the bounded CPU-baseline sample")
    ap.add_argument("--e2e-rows", type=int, default=16384, help="rows of the bounded host-buffer (e2e) sample")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, wl, rank, world)
        return

    import torch
    import __graft_entry__ as ge
    pkg = ge.load_package()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ctx = pkg.VgaContext(local_rank)
    if world > 1:   # the library's own NCCL communicator; torch.distributed only ships the id
        uid = [pkg.VgaContext.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(rank, world, uid[0])

    n, p, K = wl["n"], wl["p"], wl["K"]
    nnz = ctx.sim_data_log(n, p, K, row0=rank * n, seed=SEED, log_offset=wl["log_offset"], warm_start=False)
    density = nnz / (n * p)
    fp64_tf = max(ctx.probe_fp64_tflops() for _ in range(3))

    # ---- resident-state steps -----------------------------------------------------------------
    infos = []
    for _ in range(args.warmup):
        ctx.vga_step(MAXITER_VGA, VGA_TOL)
        ctx.rewind_m()
    barrier()
    with ClockSampler(local_rank) as clk:
        t0 = time.perf_counter()
        for _ in range(args.steps):
            infos.append(ctx.vga_step(MAXITER_VGA, VGA_TOL))
            ctx.rewind_m()
        barrier()
        wall = time.perf_counter() - t0
    dev_ms = sum(i.kernel_ms for i in infos)
    dev_ms = allmax(dev_ms)
    wall = allmax(wall)
    ms_per_step = dev_ms / args.steps
    entries = n * p * world
    value = entries / (ms_per_step * 1e-3)
    u = infos[-1].n_updates
    diag = ctx.last_diag()
    p1 = sum(i.pass1_ms for i in infos) / args.steps
    p2 = sum(i.pass2_ms for i in infos) / args.steps
    rd = sum(i.reduce_ms for i in infos) / args.steps

    # ---- the same step in delta mode (option, not the default) ----------------
    ctx.set_option("delta", DELTA)
    dinfos = []
    for it in range(1 + args.steps):
        inf = ctx.vga_step(MAXITER_VGA, VGA_TOL)
        ctx.rewind_m()
        if it > 0:
            dinfos.append(inf)
    barrier()
    ctx.set_option("delta", 0.0)
    d_ms = allmax(sum(i.kernel_ms for i in dinfos)) / args.steps
    delta_mode = {"delta": DELTA, "value": entries / (d_ms * 1e-3), "unit": "entries/s", "ms_per_step": d_ms,
                  "n_updates": dinfos[-1].n_updates,
                  "pass_ms": {"first": sum(i.pass1_ms for i in dinfos) / args.steps,
                              "topup": sum(i.pass2_ms for i in dinfos) / args.steps},
                  "note": "units whose next Newton update is <= delta (and |f| < tol/16) are frozen; M, V within "
                          "~delta of the exact-count result, n_updates identical; NOT the headline value"}

    # ---- the same step with the scored pipeline (option, not the default) --------
    # exact mode, bit-identical results: an FP32 tensor-core scoring pass orders the tiles of pass 1
    ctx.set_option("pipeline", 3)
    sinfos = []
    for it in range(1 + args.steps):
        inf = ctx.vga_step(MAXITER_VGA, VGA_TOL)
        ctx.rewind_m()
        if it > 0:
            sinfos.append(inf)
    barrier()
    ctx.set_option("pipeline", 0)
    s_ms = allmax(sum(i.kernel_ms for i in sinfos)) / args.steps
    scored_mode = {"value": entries / (s_ms * 1e-3), "unit": "entries/s", "ms_per_step": s_ms,
                   "n_updates": sinfos[-1].n_updates,
                   "pass_ms": {"score_and_first": sum(i.pass1_ms for i in sinfos) / args.steps,
                               "topup": sum(i.pass2_ms for i in sinfos) / args.steps},
                   "note": "pipeline = 3: tiles of pass 1 taken by decreasing predicted hardness (tf32 scoring pass); "
                           "same results bit for bit, exact update count; NOT the headline value"}

    # ---- roofline (HBM line of the dominant kernel = first pass) --------------------------------
    peak, peak_src = read_peaks()
    alg_bytes_per_entry = 16.0 + 12.0 * density            # read M 8 + write M 8 + CSC (x 8 + i 4) * density
    alg_bytes = alg_bytes_per_entry * n * p                 # per launch of the first-pass kernel (one slab)
    achieved = alg_bytes / (p1 * 1e-3) / 1e9
    # DRAM bytes of that kernel from the committed ncu capture (profiles/r01_ncu_summary.md), this workload only
    traffic = 95.81 if (args.workload == "c4slab" and os.environ.get("EBPMF_B200_PIPELINE", "fused") == "fused") else None
    roofline = {"bound": "hbm", "kernel": "vga_fused_kernel<MODE_FIRST>", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_unit": "GB per launch (ncu dram read+write)",
                "algorithmic_gb_per_launch": alg_bytes / 1e9, "peak_source": peak_src,
                "algorithmic_bytes_per_entry": alg_bytes_per_entry,
                "note": "the FP64 pipe, not HBM, binds this kernel for u > 0 (see fp64)"}
    # FP64 line: FP64-pipe instructions per entry, counted in the SASS of the Newton loop:
    # 18 per evaluation of f (u+1 of them, +1 for the 58 % of units redone in pass 2), 6.5 per update,
    # K_tot FMAs for Beta, ~20 in the epilogue
    dp_instr = 18.0 * (u + 1.6) + 6.5 * u + 1.0 * (K + 2) + 20.0
    fp64 = {"dfma_tflops_measured": fp64_tf, "dp_instr_per_entry_model": dp_instr,
            "line_entries_per_s": (fp64_tf * 1e12 / 2.0) / dp_instr,
            "frac_of_fp64_line": (value / world) / ((fp64_tf * 1e12 / 2.0) / dp_instr)}

    # Issue-port line (what binds pass 1): FP64 instructions hold a sub-partition's
    # issue port for 2 cycles (3 with three distinct register sources), everything else 1 cycle.  The
    # instruction counts per 32 entries come from the committed ncu capture of this workload
    # (profiles/r01_ncu_summary.md); the cycles are measured live.
    issue = None
    if args.workload == "c4slab" and os.environ.get("EBPMF_B200_PIPELINE", "fused") == "fused":
        clk_mhz = 1965.0
        cycles_per_32 = p1 * 1e-3 * clk_mhz * 1e6 * 148 * 4 / (n * p / 32.0)
        slots = 651.5 + 263.1
        issue = {"bound": "issue port (FP64 = 2 slots)", "slots_per_32_entries": slots,
                 "inst_per_32_entries": 651.5, "fp64_inst_per_32_entries": 263.1,
                 "three_register_dfma_per_32_entries": 59.4,
                 "three_register_note": "not added to slots: with them the model (974) exceeds the measured cycles, "
                                        "so their third cycle is at least partly hidden behind other warps",
                 "cycles_per_32_entries_measured": cycles_per_32, "frac": slots / cycles_per_32,
                 "sm_mhz_assumed": clk_mhz, "source": "profiles/r01_ncu_summary.md + scripts/probes"}

    out = {
        "metric": "VGA Newton entries/sec (fp64, N x p)", "value": value, "unit": "entries/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["label"], "rows_per_gpu": n, "cols": p, "K_tot": K + 2, "density": density,
                   "nnz_per_gpu": nnz, "maxiter_vga": MAXITER_VGA, "vga_tol": VGA_TOL, "var_type": "by_col",
                   "start": "M0 = log(Y + 0.5)", "l2": "inputs (%.1f GB per GPU) are larger than L2; no flush"
                   % (16.0 * n * p / 1e9), "parallelism": "row-sharded x%d, NCCL MAX(u) + SUM(v_sum[p]+3)" % world},
        "n_updates": u, "converged": bool(infos[-1].converged), "passes": infos[-1].passes,
        "pass_ms": {"first": p1, "topup": p2, "reduce": rd}, "wall_ms_per_step": 1e3 * wall / args.steps,
        "topup_units_frac": diag["topup_units"] / max(1.0, ((n + 31) // 32) * ((p + 1) // 2)),
        "peer_kstar_sharing": {"peers_mapped": ctx.peer_count() if world > 1 else 0,
                               "how": "pass-1 kernels read the other GPUs' published update count over NVLink peer memory"},
        "roofline": roofline, "fp64": fp64, "issue_roofline": issue, "delta_mode": delta_mode, "scored_pipeline": scored_mode,
        "gpu_launches": 6 * args.steps, "clocks": clk.summary(),
    }

    # ---- parity gate + CPU baseline on a bounded row sample (rank 0) -----------------------------
    if rank == 0 and not args.no_cpu:
        sample = cpu_sample_inputs(ctx, args.cpu_rows)
        cb, ref = cpu_baseline(sample)
        out["cpu_baseline"] = cb
        # parity: the oracle on the sample with the GLOBAL update count forced (tol = 0 => exactly u updates)
        from oracle import vga_numpy
        Beta = sample["EL"] @ sample["EF"].T
        S2 = np.repeat(sample["sigma2"][None, :], args.cpu_rows, axis=0)
        if u > 0:
            Mref = vga_numpy.vga_pois_solver_mat_newton(sample["M0"], sample["Y"], 1, Beta, S2, u, 0.0)["M"]
        else:
            Mref = np.minimum(sample["M0"], S2 * sample["Y"] + Beta)
        ctx.vga_step(MAXITER_VGA, VGA_TOL)
        Mgpu = ctx.get_rows(0, args.cpu_rows)
        ctx.rewind_m()
        err = float(np.max(np.abs(Mgpu - Mref) / np.maximum(np.abs(Mref), 1.0)))
        out["parity"] = {"max_rel_err_M_sample": err, "tol": 1e-8, "ok": bool(err < 1e-8),
                         "how": "oracle on rows [0,%d) with the global u forced" % args.cpu_rows}
    elif not args.no_cpu and dist is not None:
        ctx.vga_step(MAXITER_VGA, VGA_TOL)      # keep the collectives of rank 0's extra step matched
        ctx.rewind_m()

    # ---- e2e: host buffers through ebpmf_vga_step_host --------------------------------------------
    if not args.no_e2e:
        er = min(args.e2e_rows, n)
        EL, EF = ctx.get_factors()
        s2 = ctx.get_sigma2()
        cp, ri, x = ctx.get_y_csc()
        scp, sri, sx = pkg.shard_csc(cp, ri, x, 0, er)
        # pinned buffers are allocated (and first touched) from a CPU of the GPU's own NUMA node
        all_cpus = os.sched_getaffinity(0)
        numa = numa_local_cpus(local_rank)
        if numa:
            os.sched_setaffinity(0, numa[1])
        M0 = torch.empty((p, er), dtype=torch.float64).pin_memory().numpy().T       # pinned, column-major er x p
        M0[...] = ctx.get_rows(0, er)
        Mo = torch.empty((p, er), dtype=torch.float64).pin_memory().numpy().T       # pinned result buffer
        Mo[...] = 0.0
        if numa:
            os.sched_setaffinity(0, all_cpus)
        ELs = np.asfortranarray(EL[:er])
        ectx = pkg.VgaContext(local_rank)
        ectx.set_y_csc(er, p, scp, sri, sx)
        # (a) the drop-in signature: M in, M out, every step
        for _ in range(2):
            r = ectx.vga_step_host(M0, ELs, EF, s2, "by_col", MAXITER_VGA, VGA_TOL, M_out=Mo)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            r = ectx.vga_step_host(M0, ELs, EF, s2, "by_col", MAXITER_VGA, VGA_TOL, M_out=Mo)
        torch.cuda.synchronize()
        full_wall = allmax(time.perf_counter() - t0)
        # (b) ebpmf_log's loop: M is the previous step's output and stays resident (R/ebpmf_log.R:309);
        #     per step the factors and sigma2 go in from pinned host memory, M and v_sum come out
        ELp = torch.empty((K + 2, er), dtype=torch.float64).pin_memory().numpy().T; ELp[...] = ELs
        EFp = torch.empty((K + 2, p), dtype=torch.float64).pin_memory().numpy().T; EFp[...] = EF
        ectx.set_m(M0)
        for _ in range(2):
            r = ectx.vga_step_resident_host(ELp, EFp, s2, "by_col", MAXITER_VGA, VGA_TOL, M_out=Mo)
            ectx.rewind_m()                          # pointer swap: every timed step starts from the same M
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            r = ectx.vga_step_resident_host(ELp, EFp, s2, "by_col", MAXITER_VGA, VGA_TOL, M_out=Mo)
            ectx.rewind_m()
        torch.cuda.synchronize()
        e_wall = allmax(time.perf_counter() - t0)
        out["e2e"] = {"value": er * p * world / (e_wall / args.steps), "unit": "entries/s",
                      "h2d_bytes_per_step": int(8 * (er + p) * (K + 2) + 8 * p),
                      "d2h_bytes_per_step": int(8 * er * p + 8 * p + 64),
                      "call": "ebpmf_vga_step_resident_host: the per-iteration call of ebpmf_log's outer loop -- factors "
                              "and sigma2 in from pinned host memory, M and v_sum out; the step's input M is the "
                              "previous step's output and is already on the device (R feeds res$M back unchanged, "
                              "R/ebpmf_log.R:309, R/ebpmf_log_flash_update.R:35)",
                      "sample": "rows [0,%d) x %d cols per GPU per step (bounded: the full slab is 30 GB); PCIe-bound"
                                % (er, p), "n_updates": r["n_updates"], "ms_per_step": 1e3 * e_wall / args.steps,
                      "pinned_buffers_numa_node": (numa[0] if numa else None)}
        out["e2e_full_roundtrip"] = {"value": er * p * world / (full_wall / args.steps), "unit": "entries/s",
                                     "h2d_bytes_per_step": int(8 * er * p + 8 * (er + p) * (K + 2) + 8 * p),
                                     "d2h_bytes_per_step": int(8 * er * p + 8 * p + 64),
                                     "call": "ebpmf_vga_step_host: M in AND out every step (the stand-alone drop-in signature)",
                                     "ms_per_step": 1e3 * full_wall / args.steps}
        ectx.close()
        # (c) ebpmf_log's batched branch (general_control$batch_size < n, R/ebpmf_log.R:254-300): one solve per row
        #     batch with the stop test scoped to the batch, M in and out from pinned host memory, the batches
        #     pipelined through 3 worker contexts so H2D, solve and D2H of different batches overlap
        import scipy.sparse as sp
        bs = max(256, er // 8)
        bh = pkg.VgaBatched(local_rank, 3)
        bh.set_y(sp.csc_matrix((sx, sri, scp), shape=(er, p)), batch_size=bs)
        for _ in range(2):
            rb_ = bh.vga_step_host(M0, ELp, EFp, s2, "by_col", MAXITER_VGA, VGA_TOL, M_out=Mo)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            rb_ = bh.vga_step_host(M0, ELp, EFp, s2, "by_col", MAXITER_VGA, VGA_TOL, M_out=Mo)
        torch.cuda.synchronize()
        b_wall = allmax(time.perf_counter() - t0)
        out["e2e_batched"] = {"value": er * p * world / (b_wall / args.steps), "unit": "entries/s",
                              "h2d_bytes_per_step": int(8 * er * p + 8 * (er + p * (len(bh.batch_row0) - 1)) * (K + 2) + 8 * p),
                              "d2h_bytes_per_step": int(8 * er * p + 8 * p * (len(bh.batch_row0) - 1) + 64),
                              "call": "ebpmf_batched_vga_step_host: ebpmf_log's batched branch (batch_size = %d rows, %d "
                                      "batches, per-batch stop as in R/ebpmf_log.R:259-268), M in AND out, 3 pipelined "
                                      "worker contexts" % (bs, len(bh.batch_row0) - 1),
                              "n_updates_per_batch": rb_["n_updates"], "ms_per_step": 1e3 * b_wall / args.steps}
        bh.close()
    if rank == 0:
        emit(json.dumps(out))
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
