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
             library's stream), max over ranks.  Steady state of ebpmf_log's loop: the tiles of the
             first pass are taken in the order of the previous step's per-tile counts (kept by the
             library).  `first_step` is the same step without that history, `perturbed_history` a
             step whose history comes from DIFFERENT factors (the bench rewinds M between steps, so
             the plain history is a perfect predictor; the perturbed one is not)
  e2e        the same metric through the host-buffer C-ABI call the R .Call shim makes every outer
             iteration, ebpmf_vga_step_resident_host: factors and sigma2 in from PLAIN PAGEABLE host arrays
             (what R hands the shim), step, M and v_sum out into a NEW result matrix per call, allocated
             the way the glue allocates it (allocVector3 + the library's pool allocator: warm, page-locked
             blocks that R's gc returns); the full slab at N = 1.  Next to it: e2e_pool_pageable (pool
             blocks not page-locked), e2e_fresh_pages (plain Rf_allocMatrix: first-touch page faults every
             call), e2e_full_roundtrip (M in AND out, pageable), e2e_batched (the batch_size < n branch),
             host_d2h_ceiling (what the box's host side absorbs with every GPU copying at once), and
             m_products (N1: what flashier needs from M computed on the device, so that M need not leave it) and
             e2e_m_resident (the outer iteration end to end when only those products cross PCIe)
  roofline   HBM line of the dominant kernel (first pass).  The FP64 pipe / issue port binds it
             (DESIGN.md section 4), so `fp64` (measured DFMA peak) and `issue_roofline` sit beside it; the
             instruction counts and DRAM traffic come from the committed ncu capture
             profiles/r02_ncu_metrics.json (written by scripts/ncu_summary.py), the times are live
  cpu_baseline  the oracle's C transcription of the R code (OpenMP, all host cores) on a bounded
             row sample -- a restatement of the R code, NOT the R code (R is not installed)
  parity     every run: rows [0, 2048) x all columns and ALL rows x 8 columns of the slab against the oracle
             with the global update count forced (M, V, and v_sum of those columns); `--full-parity`: every
             entry of the slab, the three ELBO sums, all of v_sum, and R's update count derived from the
             oracle's global max|f| per iterate.  parity_multi: config 2 row-sharded over the N ranks
             against the oracle on the whole matrix (M, V, all-reduced v_sum, the sums, u on every rank)

`--workload c5` times the solver VARIANTS of BASELINE config 5 (dense drop-in signature, _fixed_iter,
_low_memory) on a row slab of 200k x 20k, and _low_memory at the full 200k x 20k.  `--impl reference` times the CPU restatement alone.
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
    # config 5 (200k x 20k, K=10): the dense signatures need 4-6 n x p operands on the host AND the device
    # (6 x 32 GB at full size), so the variants run on a row slab of it
    "c5": dict(n=50000, p=20000, K=10, log_offset=-4.6440, density_target=0.05,
               label="BASELINE config 5 row slab: 50000 x 20000 of 200000 x 20000, K=10 (+2 offset cols); solver variants"),
}
MAXITER_VGA, VGA_TOL = 10, 1e-5
C5_MAXITER, C5_TOL = 10, 1e-5
DELTA = 1e-14
SEED = 12345
NCU_METRICS = os.path.join(ROOT, "profiles", "r02_ncu_metrics.json")


def read_peaks():
    try:
        d = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def read_ncu_metrics():
    """Counters of the committed ncu capture of the first-pass kernel (scripts/ncu_summary.py --json)."""
    try:
        return json.load(open(NCU_METRICS))
    except Exception:
        return None


def mem_available_bytes():
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable:"):
                return int(line.split()[1]) * 1024
    except Exception:
        pass
    return 64 << 30


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


def pageable_f(rows, cols):
    """A fresh rows x cols column-major float64 matrix in plain pageable memory (np.empty -> malloc -> mmap),
    untouched: what Rf_allocMatrix hands the .Call shim."""
    return np.empty((cols, rows), dtype=np.float64).T


def cpu_sample_inputs(ctx, rows):
    """Rows [0, rows) of the device-generated workload, pulled to the host for the CPU arm."""
    import scipy.sparse as sp
    n, p, K, nnz = ctx.shape()
    cp, ri, x = ctx.get_y_csc()
    Y = sp.csc_matrix((x, ri, cp), shape=(n, p))[:rows].toarray(order="F")
    EL, EF = ctx.get_factors()
    return dict(Y=Y, EL=np.asfortranarray(EL[:rows]), EF=EF, sigma2=ctx.get_sigma2(), M0=ctx.get_rows(0, rows),
                csc=(cp, ri, x))


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


def oracle_forced(M0, Y, EL, EF, sigma2, u, converged, want_trace=False):
    """The oracle's result for a block of entries when the GLOBAL update count is u (decided by entries outside
    the block): exactly u updates (C transcription with tol = 0), then what R/vga_solver_mat.R:35-36 returns --
    V from the temp of the final M when the loop broke, from the temp before the last update when it did not.
    want_trace: also max|f| over the block at the iterates 0..u (what R's stop test looked at)."""
    from oracle import c_oracle
    Beta = np.asfortranarray(EL @ EF.T)
    S2 = np.asfortranarray(np.repeat(np.asarray(sigma2)[None, :], M0.shape[0], axis=0))
    trace = []
    if u > 0:
        r = c_oracle.vga_newton(M0, Y, 1.0, Beta, S2, u, 0.0)
        M, Vstale = r["M"], r["V"]
        trace = list(r["maxabs_f"][:u])
    else:
        M, Vstale = np.minimum(M0, S2 * Y + Beta), None
    temp = S2 * Y + Beta + 1 - M
    V = S2 / temp if (converged or Vstale is None) else Vstale
    if want_trace:
        f = Y - np.exp(M + (S2 / 2) / temp) - M * (1 / S2) + Beta / S2           # R/vga_solver_mat.R:22-25 at the final M
        trace.append(float(np.max(np.abs(f))))
        return M, V, np.asarray(trace)
    return M, V


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


def step_summary(infos, ctx, n, p):
    k = len(infos)
    c = ctx.counters()
    units = ((n + 31) // 32) * ((p + 1) // 2)
    return {"ms_per_step": sum(i.kernel_ms for i in infos) / k, "n_updates": infos[-1].n_updates, "passes": infos[-1].passes,
            "pass_ms": {"first": sum(i.pass1_ms for i in infos) / k, "topup": sum(i.pass2_ms for i in infos) / k,
                        "reduce": sum(i.reduce_ms for i in infos) / k},
            "topup_units_frac": c["topup_units"] / units, "evaluations_per_unit": c["unit_evals"] / units}


# ------------------------------------------------------------------------------------------------
# --workload c5: the solver variants (K2 dense drop-in, K3 _fixed_iter, K4 _low_memory)
# ------------------------------------------------------------------------------------------------
def run_variants(args, wl, pkg, torch, local_rank):
    import scipy.sparse as sp
    from oracle import c_oracle
    n, p, K = wl["n"], wl["p"], wl["K"]
    ctx = pkg.VgaContext(local_rank)
    nnz = ctx.sim_data_log(n, p, K, row0=0, seed=SEED, log_offset=wl["log_offset"], warm_start=False)
    peak, peak_src = read_peaks()
    cp, ri, x = ctx.get_y_csc()
    Ysp = sp.csc_matrix((x, ri, cp), shape=(n, p))
    EL, EF = ctx.get_factors()
    s2 = ctx.get_sigma2()
    M0 = ctx.get_m()
    srows = min(args.cpu_rows, n)
    Ys = Ysp[:srows].toarray(order="F")
    out = {"metric": "VGA Newton entries/sec (fp64, N x p)", "unit": "entries/s", "n_gpus": 1, "steps": args.steps,
           "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
           "data": "synthetic",
           "config": {"workload": wl["label"], "rows": n, "cols": p, "K_tot": K + 2, "density": nnz / (n * p),
                      "maxiter": C5_MAXITER, "tol": C5_TOL,
                      "note": "maxiter / tol as ebpmf_log calls the solver (R/ebpmf_log_controls.R:33-36); with the signature "
                              "defaults (1000, 1e-8) a handful of extreme counts (> 1e7) whose rounding floor of f is above "
                              "1e-8 keep the GLOBAL stop test from ever passing and the solve runs all 1000 updates -- see "
                              "`signature_defaults`"}}
    variants = {}

    def rel(a, b):
        return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1.0)))

    # --- K4: vga_pois_solver_mat_newton_low_memory (sparse Y, l0/f0 scales, F8 derivative) -------------
    rng = np.random.default_rng(5)
    l0 = np.exp(rng.normal(0, 0.1, n)); f0 = np.exp(rng.normal(0, 0.1, p))
    infos = []
    for it in range(1 + args.steps):
        info = {}
        Mo = pkg.vga_pois_solver_mat_newton_low_memory(M0, Ysp, l0, f0, EL, EF, s2, "by_col", C5_MAXITER, C5_TOL, ctx=ctx, info=info)
        if it > 0:
            infos.append(info)
    ms = sum(i["kernel_ms"] for i in infos) / len(infos)
    ref = c_oracle.vga_low_memory(M0[:srows], Ys, l0[:srows], f0, EL[:srows], EF, s2, "by_col", infos[-1]["n_updates"], 0.0)
    alg = (16.0 + 12.0 * nnz / (n * p)) * n * p
    variants["low_memory"] = {"value": n * p / (ms * 1e-3), "ms_per_step": ms, "n_updates": infos[-1]["n_updates"],
                              "algorithmic_bytes_per_entry": alg / (n * p), "hbm_frac": alg / (ms * 1e-3) / 1e9 / peak,
                              "parity_max_rel_err_M_rows": rel(Mo[:srows], ref["M"]),
                              "parity_how": "C oracle on rows [0,%d) with the global update count forced" % srows}
    info = {}
    pkg.vga_pois_solver_mat_newton_low_memory(M0, Ysp, l0, f0, EL, EF, s2, "by_col", 1000, 1e-8, ctx=ctx, info=info)
    variants["low_memory"]["signature_defaults"] = {"maxiter": 1000, "tol": 1e-8, "n_updates": info["n_updates"],
                                                    "converged": info["converged"], "ms_per_step": info["kernel_ms"],
                                                    "max_count_in_Y": float(x.max())}
    del Mo
    # --- K2: vga_pois_solver_mat_newton, dense operands (M, X, Beta, Sigma2 in; M, V out = 48 B/entry) -----
    Beta = np.asfortranarray(EL @ EF.T)
    S2 = np.asfortranarray(np.repeat(s2[None, :], n, axis=0))
    Yd = Ysp.toarray(order="F")
    infos = []
    for it in range(1 + args.steps):
        info = {}
        r = pkg.vga_pois_solver_mat_newton(M0, Yd, 1.0, Beta, S2, C5_MAXITER, C5_TOL, return_V=True, info=info, ctx=ctx)
        if it > 0:
            infos.append(info)
    ms = sum(i["kernel_ms"] for i in infos) / len(infos)
    u = infos[-1]["n_updates"]
    ref = c_oracle.vga_newton(M0[:srows], Ys, 1.0, Beta[:srows], S2[:srows], u, 0.0)
    Vref = S2[:srows] / (S2[:srows] * Ys + Beta[:srows] + 1 - ref["M"]) if infos[-1]["converged"] else ref["V"]
    variants["dense_signature"] = {"value": n * p / (ms * 1e-3), "ms_per_step": ms, "n_updates": u,
                                   "algorithmic_bytes_per_entry": 48.0, "hbm_frac": 48.0 * n * p / (ms * 1e-3) / 1e9 / peak,
                                   "parity_max_rel_err_M_rows": rel(r["M"][:srows], ref["M"]),
                                   "parity_max_rel_err_V_rows": rel(r["V"][:srows], Vref),
                                   "parity_how": "C oracle on rows [0,%d) with the global update count forced" % srows}
    # --- K3: vga_pois_solver_mat_newton_fixed_iter (maxiter = 1: the only call site, inst/code/Poisson_MF_splitting.R:212) ----
    fi = {}
    for mi in (1, 10):
        ts = []
        for it in range(1 + min(args.steps, 3)):
            r = pkg.vga_pois_solver_mat_newton_fixed_iter(M0, None, Yd, 1.0, Beta, S2, mi, ctx=ctx)
            if it > 0:
                ts.append(ctx.counters()["last_kernel_ms"])
        ms = sum(ts) / len(ts)
        ref = c_oracle.vga_fixed_iter(M0[:srows], None, Ys, 1.0, Beta[:srows], S2[:srows], mi)
        fi["maxiter_%d" % mi] = {"value": n * p / (ms * 1e-3), "ms_per_step": ms, "algorithmic_bytes_per_entry": 48.0,
                                 "hbm_frac": 48.0 * n * p / (ms * 1e-3) / 1e9 / peak,
                                 "parity_max_rel_err_M_rows": rel(r["M"][:srows], ref["M"]),
                                 "parity_max_rel_err_V_rows": rel(r["V"][:srows], ref["V"])}
    variants["fixed_iter"] = fi
    launches_slab = int(ctx.counters()["launches"])
    ctx.close()
    del Yd, Beta, S2, r
    # --- K4 at config 5's STATED size, 200 000 x 20 000 (the fused path needs only M and the CSC Y; the dense signatures
    #     would need 6 x 32 GB on the device): device time of the solve and a row-sample parity check -----------------
    if mem_available_bytes() > 130e9:
        nF = 200000
        cF = pkg.VgaContext(local_rank)
        nnzF = cF.sim_data_log(nF, p, K, row0=0, seed=SEED, log_offset=wl["log_offset"], warm_start=False)
        cpF, riF, xF = cF.get_y_csc()
        YF = sp.csc_matrix((xF, riF, cpF), shape=(nF, p))
        ELF, EFF = cF.get_factors()
        s2F = cF.get_sigma2()
        M0F = cF.get_m()
        l0F = np.exp(np.random.default_rng(6).normal(0, 0.1, nF))
        infoF = {}
        t0 = time.perf_counter()
        MoF = pkg.vga_pois_solver_mat_newton_low_memory(M0F, YF, l0F, f0, ELF, EFF, s2F, "by_col", C5_MAXITER, C5_TOL, ctx=cF, info=infoF)
        wallF = time.perf_counter() - t0
        YsF = YF[:srows].toarray(order="F")
        refF = c_oracle.vga_low_memory(M0F[:srows], YsF, l0F[:srows], f0, ELF[:srows], EFF, s2F, "by_col", infoF["n_updates"], 0.0)
        variants["low_memory_full_size"] = {"rows": nF, "cols": p, "nnz": int(nnzF), "value": nF * p / (infoF["kernel_ms"] * 1e-3),
                                            "kernel_ms": infoF["kernel_ms"], "n_updates": infoF["n_updates"],
                                            "call_wall_s_incl_64GB_of_pageable_host_copies": wallF,
                                            "parity_max_rel_err_M_rows": rel(MoF[:srows], refF["M"]),
                                            "parity_how": "C oracle on rows [0,%d) with the global update count forced" % srows}
        launches_slab += int(cF.counters()["launches"])
        cF.close()
        del MoF, M0F, YF
    out["variants"] = variants
    out["value"] = variants["dense_signature"]["value"]
    out["ms_per_step"] = variants["dense_signature"]["ms_per_step"]
    out["roofline"] = {"bound": "hbm", "kernel": "vga_dense_kernel<MODE_FIRST>", "achieved": 48.0 * n * p / (out["ms_per_step"] * 1e-3) / 1e9,
                       "peak": peak, "unit": "GB/s", "frac": variants["dense_signature"]["hbm_frac"], "traffic": None,
                       "peak_source": peak_src, "note": "whole solve (both passes) against 48 B/entry"}
    out["gpu_launches"] = launches_slab
    emit(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c4slab", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-rows", type=int, default=2048, help="rows of the bounded CPU-baseline / parity sample")
    ap.add_argument("--e2e-rows", type=int, default=0, help="rows of the host-buffer (e2e) run; 0 = the whole slab if host memory allows")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--full-parity", action="store_true", help="check M, V and the three ELBO sums of the WHOLE slab against the oracle (minutes)")
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
    if args.workload == "c5":
        if rank == 0:
            run_variants(args, wl, pkg, torch, local_rank)
        return
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

    def allmin(x):
        return -allmax(-x)

    def new_comm_ctx():
        c = pkg.VgaContext(local_rank)
        if world > 1:   # the library's own NCCL communicator; torch.distributed only ships the id
            uid = [pkg.VgaContext.nccl_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(uid, src=0)
            c.comm_init(rank, world, uid[0])
        return c

    ctx = new_comm_ctx()
    n, p, K = wl["n"], wl["p"], wl["K"]
    nnz = ctx.sim_data_log(n, p, K, row0=rank * n, seed=SEED, log_offset=wl["log_offset"], warm_start=False)
    density = nnz / (n * p)
    fp64_tf = max(ctx.probe_fp64_tflops() for _ in range(3))
    entries = n * p * world

    # ---- resident-state steps -----------------------------------------------------------------
    first = ctx.vga_step(MAXITER_VGA, VGA_TOL)          # no history yet: natural tile order
    ctx.rewind_m()
    first_step = step_summary([first], ctx, n, p)
    for _ in range(max(0, args.warmup - 1)):
        ctx.vga_step(MAXITER_VGA, VGA_TOL)
        ctx.rewind_m()
    barrier()
    infos = []
    launches0 = ctx.counters()["launches"]
    with ClockSampler(local_rank) as clk:
        t0 = time.perf_counter()
        for _ in range(args.steps):
            infos.append(ctx.vga_step(MAXITER_VGA, VGA_TOL))
            ctx.rewind_m()
        barrier()
        wall = time.perf_counter() - t0
    launches = int(ctx.counters()["launches"] - launches0)
    steady = step_summary(infos, ctx, n, p)
    dev_ms = allmax(sum(i.kernel_ms for i in infos))
    wall = allmax(wall)
    ms_per_step = dev_ms / args.steps
    value = entries / (ms_per_step * 1e-3)
    u = infos[-1].n_updates
    p1 = steady["pass_ms"]["first"]

    # ---- a step whose tile order comes from a DIFFERENT problem: factors perturbed after the history was taken -----
    EL0, EF0 = ctx.get_factors()
    rng = np.random.default_rng(1 + rank)
    ELp, EFp = EL0.copy(), EF0.copy()
    ELp[:, 2:] += 0.05 * rng.normal(size=ELp[:, 2:].shape)
    EFp[:, 2:] += 0.05 * np.random.default_rng(1).normal(size=EFp[:, 2:].shape)      # EF is replicated: same on every rank
    ctx.set_factors(ELp, EFp)
    pinfo = ctx.vga_step(MAXITER_VGA, VGA_TOL)          # history = the unperturbed step's counts
    ctx.rewind_m()
    perturbed = step_summary([pinfo], ctx, n, p)
    perturbed["ms_per_step"] = allmax(perturbed["ms_per_step"])
    perturbed["value"] = entries / (perturbed["ms_per_step"] * 1e-3)
    perturbed["note"] = "factor columns 3.. perturbed by N(0, 0.05^2) AFTER the per-tile history was recorded"
    ctx.set_factors(EL0, EF0)
    ctx.vga_step(MAXITER_VGA, VGA_TOL); ctx.rewind_m()  # restore the matching history
    del ELp, EFp

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
                  "note": "units whose next Newton update is <= delta (and |f| < tol/16) are frozen; M, V within "
                          "~delta of the exact-count result, n_updates identical; NOT the headline value"}
    ctx.vga_step(MAXITER_VGA, VGA_TOL); ctx.rewind_m()

    # ---- N1: what flashier needs from M, computed on the device (M never leaves it) -------------------
    peak, peak_src = read_peaks()
    mp = None
    for _ in range(2):
        mp = ctx.m_products(EL0, EF0)
    barrier()
    mp_ms = allmax(mp["kernel_ms"])
    flops = 4.0 * (K + 2) * n * p                      # two products, K_tot multiply-adds per entry each
    m_products = {"ms": mp_ms, "products_kernel_ms": mp["products_ms"], "partial_sums_ms": mp["partial_sums_ms"], "K": K + 2,
                  "value": entries / (mp_ms * 1e-3), "unit": "entries/s",
                  "algorithmic_bytes_per_entry": 8.0,
                  "hbm_frac_of_8B_per_entry": 8.0 * n * p / (mp["products_ms"] * 1e-3) / 1e9 / peak,
                  "fp64_tflops": flops / (mp["products_ms"] * 1e-3) / 1e12,
                  "fp64_frac_of_measured_dfma_peak": flops / (mp["products_ms"] * 1e-3) / 1e12 / fp64_tf,
                  "bound": "FP64 pipe for K_tot > ~12 (4 K_tot flop per entry on the FP64 tensor path, mma.sync.m8n8k4.f64), HBM below",
                  "d2h_bytes": int(8 * (n + p) * (K + 2) + 8 * (n + p)),
                  "what": "ebpmf_ctx_m_products: M %*% EF, t(M) %*% EL, column/row sums of (M - EL EF')^2 (flashier's R2) from the "
                          "resident M (R/ebpmf_log_flash_update.R:35,55-64); ONE streaming kernel reads M once"}

    # ---- N1 end to end: one outer iteration of ebpmf_log when M NEVER leaves the device -------------------
    # factors and sigma2 in from pageable host arrays; out: v_sum, the three ELBO sums, and what flashier takes from M
    # (M %*% EF, t(M) %*% EL, R2) instead of M itself.  Wall clock around the public calls, max over ranks.
    e2e_n1 = None
    try:
        if world > 1:
            raise RuntimeError("measured at N = 1 only (a diagnostic leg adds no collectives to the scaling runs)")
        s20 = ctx.get_sigma2()
        t0 = 0.0
        for it in range(1 + args.steps):
            if it == 1:
                barrier()
                t0 = time.perf_counter()
            ctx.set_factors(EL0, EF0)
            ctx.set_sigma2(s20, "by_col")
            ninf = ctx.vga_step(MAXITER_VGA, VGA_TOL)
            nvs = ctx.get_v_sum()
            nmp = ctx.m_products(EL0, EF0)              # of the step's result, still on the device
            ctx.rewind_m()
        barrier()
        per = allmax(time.perf_counter() - t0) / args.steps
        e2e_n1 = {"value": entries / per, "unit": "entries/s", "ms_per_step": 1e3 * per, "n_updates": ninf.n_updates,
                  "h2d_bytes_per_step": int(EL0.nbytes + EF0.nbytes + s20.nbytes),
                  "d2h_bytes_per_step": int(nvs.nbytes + nmp["MF"].nbytes + nmp["MtL"].nbytes + nmp["r2_col"].nbytes + nmp["r2_row"].nbytes + 24),
                  "call": "set_factors + set_sigma2 + vga_step + get_v_sum + m_products through the C ABI with pageable host arrays: "
                          "the per-iteration traffic of a flashier backend built on ebpmf_b200_m_products (INTEGRATION.md, N1); "
                          "NOT the reference's loop as it stands -- that needs M on the host and is `e2e`"}
        ctx.vga_step(MAXITER_VGA, VGA_TOL); ctx.rewind_m()
    except Exception as e:                                  # a diagnostic leg must never cost the bench line
        e2e_n1 = {"skipped" if world > 1 else "error": str(e)}

    # ---- roofline (HBM line of the dominant kernel = first pass) --------------------------------
    ncu = read_ncu_metrics() if args.workload == "c4slab" else None
    alg_bytes_per_entry = 16.0 + 12.0 * density            # read M 8 + write M 8 + CSC (x 8 + i 4) * density
    alg_bytes = alg_bytes_per_entry * n * p                 # per launch of the first-pass kernel (one slab)
    achieved = alg_bytes / (p1 * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "vga_fused_kernel<MODE_FIRST>", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak,
                "traffic": (ncu["dram_gb_per_launch"] if ncu else None), "traffic_unit": "GB per launch (ncu dram read+write)",
                "traffic_source": (os.path.relpath(NCU_METRICS, ROOT) if ncu else None),
                "algorithmic_gb_per_launch": alg_bytes / 1e9, "peak_source": peak_src,
                "algorithmic_bytes_per_entry": alg_bytes_per_entry,
                "step_level_frac": alg_bytes / (ms_per_step * 1e-3) / 1e9 / peak,
                "note": "the FP64 pipe / issue port, not HBM, binds this kernel for u > 0 (see fp64, issue_roofline)"}
    # FP64 line: FP64-pipe warp-instructions per 32 entries from the ncu capture x the live launch time,
    # against the DFMA rate measured live (ebpmf_probe_fp64_tflops: one DFMA warp-instruction = 64 flop)
    fp64 = {"dfma_tflops_measured": fp64_tf}
    # SURVEY.md section 8(d)'s ALGORITHMIC FP64 line: R performs u updates on every entry = (u+1) exp + (3u+2) divisions +
    # 1 log + (13u+13+K_tot) add/mul, costed there at 25 / 10 / 30 FP64 instructions -> 70u + 110 + K_tot per entry
    survey_instr = 70.0 * u + 110.0 + (K + 2)
    survey_line = (fp64_tf * 1e12 / 2.0) / survey_instr
    fp64["survey_8d_model"] = {"dp_instr_per_entry": survey_instr, "line_entries_per_s_per_gpu": survey_line,
                               "frac": (value / world) / survey_line,
                               "note": "the survey's attainable bound for a kernel that evaluates every entry u+1 times with "
                                       "library-grade exp/div/log; > 1 here because exp costs 9, a reciprocal 2 and the "
                                       "fast-forward skips the evaluations a converged unit would repeat"}
    issue = None
    if ncu:
        dfma_per_s = fp64_tf * 1e12 / 64.0                  # warp-level DFMA instructions per second, whole chip
        fp64_inst = ncu["fp64_inst_per_32_entries"] * (n * p / 32.0)
        fp64.update({"fp64_inst_per_32_entries": ncu["fp64_inst_per_32_entries"],
                     "frac_of_fp64_line": fp64_inst / (p1 * 1e-3) / dfma_per_s,
                     "line_entries_per_s": dfma_per_s * 32.0 / ncu["fp64_inst_per_32_entries"],
                     "source": os.path.relpath(NCU_METRICS, ROOT), "kernel": ncu.get("kernel"), "git": ncu.get("git")})
        # issue-port line: an FP64 instruction holds a sub-partition's issue port for 2 cycles
        clk_mhz = clk.summary().get("sm_mhz") or 1965.0
        cycles_per_32 = p1 * 1e-3 * clk_mhz * 1e6 * 148 * 4 / (n * p / 32.0)
        slots = ncu["inst_per_32_entries"] + ncu["fp64_inst_per_32_entries"]
        issue = {"bound": "issue port (FP64 = 2 slots)", "slots_per_32_entries": slots,
                 "inst_per_32_entries": ncu["inst_per_32_entries"], "fp64_inst_per_32_entries": ncu["fp64_inst_per_32_entries"],
                 "cycles_per_32_entries_measured": cycles_per_32, "frac": slots / cycles_per_32, "sm_mhz": clk_mhz,
                 "note": "the builder's own model, not a hardware roofline"}

    out = {
        "metric": "VGA Newton entries/sec (fp64, N x p)", "value": value, "unit": "entries/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["label"], "rows_per_gpu": n, "cols": p, "K_tot": K + 2, "density": density,
                   "nnz_per_gpu": nnz, "maxiter_vga": MAXITER_VGA, "vga_tol": VGA_TOL, "var_type": "by_col",
                   "start": "M0 = log(Y + 0.5)", "l2": "inputs (%.1f GB per GPU) are larger than L2; no flush"
                   % (16.0 * n * p / 1e9), "parallelism": "row-sharded x%d, NCCL MAX(u) + SUM(v_sum[p]+3)" % world,
                   "tile_order": "previous step's per-tile update counts (kept by the library across steps)"},
        "n_updates": u, "converged": bool(infos[-1].converged), "passes": infos[-1].passes,
        "pass_ms": steady["pass_ms"], "wall_ms_per_step": 1e3 * wall / args.steps,
        "topup_units_frac": steady["topup_units_frac"], "evaluations_per_unit": steady["evaluations_per_unit"],
        "first_step": first_step, "perturbed_history": perturbed,
        "peer_kstar_sharing": {"peers_mapped": ctx.peer_count() if world > 1 else 0,
                               "how": "pass-1 kernels read the other GPUs' published update count over NVLink peer memory"},
        "roofline": roofline, "fp64": fp64, "issue_roofline": issue, "delta_mode": delta_mode, "m_products": m_products, "e2e_m_resident": e2e_n1,
        "gpu_launches": launches, "gpu_launches_how": "kernel launches counted by the library inside the timed region",
        "clocks": clk.summary(),
    }

    # ---- e2e: PAGEABLE host buffers through the host-buffer C ABI -----------------------------------------
    if not args.no_e2e:
        budget = int(0.40 * allmin(float(mem_available_bytes())) / max(1, min(world, 8)))
        er = args.e2e_rows if args.e2e_rows > 0 else n
        if 2 * 8 * er * p > budget:                      # result + M0 copies on the host
            er = max(2048, (budget // (2 * 8 * p)) // 256 * 256)
        er = min(er, n)
        s2 = ctx.get_sigma2()
        if er == n:
            ectx, ELs = ctx, EL0
        else:                                           # bounded row sample in its own context (no communicator)
            cp, ri, x = ctx.get_y_csc()
            scp, sri, sx = pkg.shard_csc(cp, ri, x, 0, er)
            ELs = np.asfortranarray(EL0[:er])
            ectx = pkg.VgaContext(local_rank)
            ectx.set_y_csc(er, p, scp, sri, sx)
            ectx.set_m(ctx.get_rows(0, er))
            del cp, ri, x
        # what the host side of this box can absorb: every rank copies 2 GiB device -> PINNED host memory at the same
        # time (no library code involved); the e2e numbers below cannot exceed this x 1/8 B per entry
        hb = torch.empty(1 << 28, dtype=torch.float64).pin_memory()
        db = torch.zeros(1 << 28, dtype=torch.float64, device="cuda")
        hb.copy_(db); torch.cuda.synchronize()
        barrier()
        t0 = time.perf_counter()
        for _ in range(3):
            hb.copy_(db, non_blocking=True)
        torch.cuda.synchronize()
        tw = allmax(time.perf_counter() - t0)
        out["host_d2h_ceiling"] = {"gb_per_s_per_gpu": 3 * 8 * (1 << 28) / tw / 1e9, "gb_per_s_all_gpus": world * 3 * 8 * (1 << 28) / tw / 1e9,
                                   "entries_per_s_all_gpus": world * 3 * (1 << 28) / tw,
                                   "how": "every rank: 3 x 2 GiB cudaMemcpyAsync device -> pinned host, concurrently; max over ranks"}
        del hb, db
        # (b) ebpmf_log's loop: M is the previous step's output and stays resident (R/ebpmf_log.R:309); per step the
        #     factors and sigma2 go in (pageable NumPy arrays), M and v_sum come out.  The result matrix is a NEW
        #     matrix every call, allocated the way the .Call glue allocates it (allocVector3 with the library's
        #     allocator = ebpmf_host_alloc: a block of the warm host pool); the previous result stays alive during
        #     the call (fit_flash holds it) and goes back to the pool afterwards, as under R's garbage collector.
        def timed_resident(alloc, warm, steps):
            prev = None
            for it in range(warm + steps):
                if it == warm:
                    barrier()
                    t0 = time.perf_counter()
                cur = alloc(er, p)
                r = ectx.vga_step_resident_host(ELs, EF0, s2, "by_col", MAXITER_VGA, VGA_TOL, M_out=cur)
                ectx.rewind_m()
                prev = cur
                del cur
            torch.cuda.synchronize()
            w = allmax(time.perf_counter() - t0)
            del prev
            return w / steps, r["n_updates"]

        ks = max(1, min(args.steps, 3))
        t0 = time.perf_counter()
        per, nu = timed_resident(pkg.host_matrix, 3, ks)
        setup_s = time.perf_counter() - t0 - per * (3 + ks)
        cnt = ectx.counters()
        out["e2e"] = {"value": er * p * world / per, "unit": "entries/s",
                      "h2d_bytes_per_step": int(8 * (er + p) * (K + 2) + 8 * p),
                      "d2h_bytes_per_step": int(8 * er * p + 8 * p + 64),
                      "call": "ebpmf_vga_step_resident_host: the per-iteration call of ebpmf_log's outer loop -- factors and "
                              "sigma2 in from PAGEABLE host memory (plain NumPy arrays, as R's REAL() pointers are), M and "
                              "v_sum out; every call gets a NEW result matrix, allocated as the .Call glue allocates it "
                              "(allocVector3 + the library's allocator ebpmf_host_alloc: blocks of a warm pool, page-locked "
                              "once, returned by R's gc); the step's input M is the previous step's output and is already on "
                              "the device (R feeds res$M back unchanged, R/ebpmf_log.R:309, R/ebpmf_log_flash_update.R:35)",
                      "sample": "rows [0,%d) of %d x %d cols per GPU per step%s" % (er, n, p, "" if er == n else " (bounded by host memory)"),
                      "steps": ks, "n_updates": nu, "ms_per_step": 1e3 * per,
                      "d2h_gb_per_s_per_gpu": 8.0 * er * p / per / 1e9,
                      "one_time_pool_setup_s": setup_s,
                      "speculative_copy_bytes": cnt["spec_bytes"], "recopied_bytes": cnt["recopy_bytes"],
                      "host_pool": pkg.host_pool_stats()}
        pkg.host_pool_trim()
        # (b0) the pool with pageable blocks (EBPMF_B200_POOL_PIN=0): through the pinned ring + host memcpy threads
        os.environ["EBPMF_B200_POOL_PIN"] = "0"
        per, _ = timed_resident(pkg.host_matrix, 3, ks)
        out["e2e_pool_pageable"] = {"value": er * p * world / per, "unit": "entries/s", "ms_per_step": 1e3 * per,
                                    "d2h_gb_per_s_per_gpu": 8.0 * er * p / per / 1e9,
                                    "call": "as e2e with EBPMF_B200_POOL_PIN=0: warm but pageable result blocks"}
        del os.environ["EBPMF_B200_POOL_PIN"]
        pkg.host_pool_trim()
        # (b') a plain Rf_allocMatrix-style result: fresh, never-touched pageable pages every call
        per, _ = timed_resident(pageable_f, 1, ks)
        out["e2e_fresh_pages"] = {"value": er * p * world / per, "unit": "entries/s", "ms_per_step": 1e3 * per,
                                  "d2h_gb_per_s_per_gpu": 8.0 * er * p / per / 1e9,
                                  "call": "as e2e, but the result is a freshly mmap'ed matrix every call (np.empty = what "
                                          "Rf_allocMatrix gives without the allocator): first-touch page faults included"}
        pkg.host_pool_trim()
        # (a) the stand-alone drop-in signature: M in AND out every step.  In ebpmf_log's loop the input M IS the previous
        #     call's result (fit_flash$flash_fit$Y = res$M), i.e. a pool block, and the output is a new pool block
        Mi = ectx.get_m(out=pkg.host_matrix(er, p))     # the resident M is the bench's M0 (every step above was rewound)
        kf = max(1, min(args.steps, 2))
        for _ in range(2):                              # both pool blocks warm
            Mo = pkg.host_matrix(er, p)
            ectx.vga_step_host(Mi, ELs, EF0, s2, "by_col", MAXITER_VGA, VGA_TOL, M_out=Mo)
            del Mo
        barrier()
        t0 = time.perf_counter()
        for _ in range(kf):
            Mo = pkg.host_matrix(er, p)
            ectx.vga_step_host(Mi, ELs, EF0, s2, "by_col", MAXITER_VGA, VGA_TOL, M_out=Mo)
            del Mo
        torch.cuda.synchronize()
        w = allmax(time.perf_counter() - t0)
        out["e2e_full_roundtrip"] = {"value": er * p * world / (w / kf), "unit": "entries/s", "ms_per_step": 1e3 * w / kf,
                                     "h2d_bytes_per_step": int(8 * er * p + 8 * (er + p) * (K + 2) + 8 * p),
                                     "d2h_bytes_per_step": int(8 * er * p + 8 * p + 64),
                                     "call": "ebpmf_vga_step_host: M in AND out every step (the stand-alone drop-in signature); input and "
                                             "result matrices are pool blocks, as in ebpmf_log's loop where the input is the previous result"}
        if er == n:
            ctx.rewind_m()                              # leave ctx with the bench's M0 again
        else:
            ectx.close()
        del Mi
        pkg.host_pool_trim()
        # (c) ebpmf_log's batched branch (general_control$batch_size < n, R/ebpmf_log.R:254-300) on a bounded sample:
        #     one solve per row batch, stop test scoped to the batch, M in and out (pageable), 3 pipelined worker contexts
        import scipy.sparse as sp
        br = min(16384, n)
        cp, ri, x = ctx.get_y_csc()
        scp, sri, sx = pkg.shard_csc(cp, ri, x, 0, br)
        del cp, ri, x
        Mb = pkg.host_matrix(br, p); Mb[...] = ctx.get_rows(0, br)
        Mbo = pkg.host_matrix(br, p)
        bs = max(256, br // 8)
        bh = pkg.VgaBatched(local_rank, 3)
        bh.set_y(sp.csc_matrix((sx, sri, scp), shape=(br, p)), batch_size=bs)
        ELb = np.asfortranarray(EL0[:br])
        for _ in range(2):
            rb_ = bh.vga_step_host(Mb, ELb, EF0, s2, "by_col", MAXITER_VGA, VGA_TOL, M_out=Mbo)
        barrier()
        t0 = time.perf_counter()
        for _ in range(ks):
            rb_ = bh.vga_step_host(Mb, ELb, EF0, s2, "by_col", MAXITER_VGA, VGA_TOL, M_out=Mbo)
        torch.cuda.synchronize()
        b_wall = allmax(time.perf_counter() - t0)
        out["e2e_batched"] = {"value": br * p * world / (b_wall / ks), "unit": "entries/s",
                              "call": "ebpmf_batched_vga_step_host: ebpmf_log's batched branch (batch_size = %d rows, %d "
                                      "batches, per-batch stop as in R/ebpmf_log.R:259-268), M in AND out (pool blocks), 3 pipelined "
                                      "worker contexts; %d-row sample" % (bs, len(bh.batch_row0) - 1, br),
                              "n_updates_per_batch": rb_["n_updates"], "ms_per_step": 1e3 * b_wall / ks}
        bh.close()
        del Mb, Mbo

    # ---- parity gate at slab scale + CPU baseline (rank 0 checks; every rank runs the step: collectives) ----------
    if not args.no_cpu:
        pinfo = ctx.vga_step(MAXITER_VGA, VGA_TOL, want_v=True)
        if rank == 0:
            from oracle import c_oracle
            import scipy.sparse as sp
            rows = min(args.cpu_rows, n)
            vs_gpu = ctx.get_v_sum()
            cols = sorted(set(int(c) for c in (0, 1, p // 7, p // 2, p // 2 + 1, (3 * p) // 4, p - 2, p - 1)))
            take = lambda which: np.column_stack([ctx.get_block(0, n, c, 1, which)[:, 0] for c in cols])
            sample = cpu_sample_inputs(ctx, rows)
            sample["M0"] = ctx.get_rows(0, rows, "M_in")     # the M the step started from (still on the device)
            rel = lambda a, b: float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1.0)))
            relv = lambda a, b: float(np.max(np.abs(a - b) / np.abs(b)))
            uu, cv = pinfo.n_updates, bool(pinfo.converged)
            Mref, Vref = oracle_forced(sample["M0"], sample["Y"], sample["EL"], sample["EF"], sample["sigma2"], uu, cv)
            # whole columns: all rows of the slab for a handful of columns
            cp, ri, x = sample["csc"]
            Ysp = sp.csc_matrix((x, ri, cp), shape=(n, p))
            Yc = np.asfortranarray(Ysp[:, cols].toarray())
            Mcr, Vcr = oracle_forced(take("M_in"), Yc, EL0, EF0[cols], sample["sigma2"][cols], uu, cv)
            errs = {"M_rows": rel(ctx.get_rows(0, rows, "M"), Mref), "V_rows": relv(ctx.get_rows(0, rows, "V"), Vref),
                    "M_cols": rel(take("M"), Mcr), "V_cols": relv(take("V"), Vcr)}
            if world == 1:                              # v_sum[j] = sum over ALL rows of V[, j]: exact for whole columns
                errs["v_sum_cols"] = relv(vs_gpu[cols], Vcr.sum(axis=0))
            par = {"max_rel_err": errs, "tol": 1e-8, "n_updates": uu,
                   "how": "oracle (C transcription, global u forced) on rows [0,%d) x all columns and on all %d rows x columns %s: "
                          "M, V%s" % (rows, n, cols, ", and v_sum of those columns" if world == 1 else "")}
            if args.full_parity and world == 1:
                Ycsr = Ysp.tocsr()
                worst = {"M": 0.0, "V": 0.0}
                sums = np.zeros(3)
                vsum = np.zeros(p)
                gtrace = np.zeros(uu + 1)
                blk = 4096
                for r0 in range(0, n, blk):
                    nr = min(blk, n - r0)
                    Y_b = np.asfortranarray(Ycsr[r0:r0 + nr].toarray())
                    Mr, Vr, tr = oracle_forced(ctx.get_rows(r0, nr, "M_in"), Y_b, EL0[r0:r0 + nr], EF0, sample["sigma2"], uu, cv,
                                               want_trace=True)
                    gtrace = np.maximum(gtrace, tr)
                    worst["M"] = max(worst["M"], rel(ctx.get_rows(r0, nr, "M"), Mr))
                    worst["V"] = max(worst["V"], relv(ctx.get_rows(r0, nr, "V"), Vr))
                    pr = c_oracle.partials(Y_b, Mr, Vr)
                    sums += [pr["sym"], pr["ssexp"], pr["slogv"]]
                    vsum += pr["col_v"]
                gp = [pinfo.sym, pinfo.ssexp, pinfo.slogv]
                # R's loop: test at iterate k (k = 0, 1, ...), break at the first max|f| < tol; u = that k, or maxiter
                below = [k for k in range(min(uu, MAXITER_VGA - 1) + 1) if gtrace[k] < VGA_TOL]
                u_oracle = below[0] if below else (MAXITER_VGA if uu == MAXITER_VGA else -1)    # -1: the oracle would run on
                par["full_slab"] = {"max_rel_err_M": worst["M"], "max_rel_err_V": worst["V"],
                                    "rel_err_sym_ssexp_slogv": [abs(a - b) / abs(b) for a, b in zip(gp, sums)],
                                    "max_rel_err_v_sum": relv(vs_gpu, vsum),
                                    "oracle_max_abs_f_by_iterate": [float(v) for v in gtrace],
                                    "n_updates_oracle": int(u_oracle), "n_updates_gpu": int(uu),
                                    "n_updates_equal": bool(u_oracle == uu),
                                    "how": "every entry of the slab against the oracle in %d-row blocks; sums accumulated over the "
                                           "blocks; the oracle's global max|f| per iterate gives R's update count" % blk}
                errs["full_slab"] = max([worst["M"], worst["V"], par["full_slab"]["max_rel_err_v_sum"], 0.0 if u_oracle == uu else 1.0]
                                        + par["full_slab"]["rel_err_sym_ssexp_slogv"])
            ctx.rewind_m()
            par["ok"] = bool(max(errs.values()) < 1e-8)
            out["parity"] = par
            cb, _ = cpu_baseline(sample)
            out["cpu_baseline"] = cb
        else:
            ctx.rewind_m()

        # ---- parity of the row-sharded path against the oracle on the WHOLE matrix (config 2) -------------
        from oracle import simdata, c_oracle
        import scipy.sparse as sp
        n2, p2, K2 = 10000, 2000, 5
        sim = simdata.sim_data_log(n2, p2, K=K2, seed=SEED, log_offset=WORKLOADS["c2"]["log_offset"])
        st = simdata.solver_state(sim, start="cold")
        row0, nrows = pkg.shard_rows(n2, world)[rank]
        sl = slice(row0, row0 + nrows)
        Yc = sp.csc_matrix(sim["Y"]); Yc.sort_indices()
        cp, ri, x = pkg.shard_csc(Yc.indptr, Yc.indices, Yc.data, row0, nrows)
        c2 = new_comm_ctx()
        c2.set_y_csc(nrows, p2, cp, ri, x)
        res = c2.vga_step_host(st["M0"][sl], st["EL"][sl], st["EF"], st["sigma2"], "by_col", MAXITER_VGA, VGA_TOL, return_V=True)
        ref = c_oracle.vga_step(st["M0"], sim["Y"], st["EL"], st["EF"], st["sigma2"], "by_col", MAXITER_VGA, VGA_TOL)
        e = [float(np.max(np.abs(res["M"] - ref["M"][sl]) / np.maximum(np.abs(ref["M"][sl]), 1.0))),
             float(np.max(np.abs(res["V"] - ref["V"][sl]) / np.abs(ref["V"][sl]))),
             float(np.max(np.abs(res["v_sum"] - ref["v_sum"]) / np.abs(ref["v_sum"]))),
             max(abs(res[k] - ref[k]) / abs(ref[k]) for k in ("sym", "ssexp", "slogv")),
             0.0 if res["n_updates"] == ref["n_updates"] else 1.0]
        e = [allmax(v) for v in e]
        c2.close()
        out["parity_multi"] = {"workload": "BASELINE config 2 (10000 x 2000, K=5) row-sharded over %d GPU(s), oracle on the whole matrix" % world,
                               "max_rel_err_over_ranks": {"M_shard": e[0], "V_shard": e[1], "v_sum_allreduced": e[2], "sym_ssexp_slogv": e[3]},
                               "n_updates": res["n_updates"], "n_updates_equal_on_every_rank": e[4] == 0.0, "tol": 1e-8,
                               "ok": bool(max(e[:4]) < 1e-8 and e[4] == 0.0)}

    if rank == 0:
        emit(json.dumps(out))
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
