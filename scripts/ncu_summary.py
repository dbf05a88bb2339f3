#!/usr/bin/env python3
"""Turn an `ncu --set full --import-source on` report into the markdown tables kept under profiles/.

    python scripts/ncu_summary.py gpurun_out/r01_prof.ncu-rep --entries 3.75e9 [--min-region 8]

Reads the report with `ncu -i ... --page raw --csv` and `--page source --csv` (ncu runs here without a
GPU), prints:
  * the raw metrics the roofline lines in bench.py / DESIGN.md quote,
  * per-32-entry instruction and FP64-instruction counts and the measured SM-sub-partition cycles,
  * the issue-slot model of DESIGN.md section 4 (slots = instr + FP64 instr + 3-register DFMAs),
  * SASS regions of equal execution count with their share of instructions and stall samples.
"""
import argparse
import csv
import io
import re
import subprocess
import sys

RAW_KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__cycles_active.avg", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__waves_per_multiprocessor", "smsp__warps_active.avg.per_cycle_active",
    "smsp__warps_eligible.avg.per_cycle_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
]

FP64_OPS = ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX")


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def distinct_regs(sass):
    """Number of distinct register source operands of a DFMA (destination excluded)."""
    ops = sass.split(None, 1)[1] if " " in sass else ""
    parts = [x.strip() for x in ops.rstrip(" ;").split(",")]
    srcs = parts[1:]
    regs = set()
    for s in srcs:
        m = re.match(r"^[-|~]*(R\d+)", s)
        if m and not s.startswith("RZ"):
            regs.add(m.group(1))
    return len(regs)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--entries", type=float, required=True, help="matrix entries one launch processes")
    ap.add_argument("--min-region", type=int, default=8)
    ap.add_argument("--sms", type=int, default=148)
    ap.add_argument("--json", default=None, help="also write the counters bench.py quotes (profiles/r02_ncu_metrics.json)")
    a = ap.parse_args()

    raw = ncu_csv(a.report, "raw")
    h, u, v = raw[0], raw[1], raw[2]
    col = {n: i for i, n in enumerate(h)}
    val = lambda k: v[col[k]] if k in col else ""
    print(f"## `{val('Kernel Name')}`\n")
    print("| metric | value | unit |\n|---|---|---|")
    for k in RAW_KEYS:
        if k in col:
            print(f"| `{k}` | {v[col[k]]} | {u[col[k]]} |")
    stalls = []
    for n, i in col.items():
        m = re.match(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active\.ratio", n)
        if m:
            try:
                stalls.append((float(v[i]), m.group(1)))
            except ValueError:
                pass
    stalls.sort(reverse=True)
    if stalls:
        print("\nWarp stall reasons (warps per issue-active cycle): "
              + ", ".join(f"{n} {x:.2f}" for x, n in stalls[:9]))

    w32 = a.entries / 32.0
    inst = float(val("smsp__inst_executed.sum"))
    cyc = float(val("sm__cycles_active.avg")) * a.sms * 4
    fp64_pct = float(val("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"))
    fp64_inst = fp64_pct / 100.0 * cyc / 2.0       # the pipe's sustained peak is one instruction per 2 cycles
    print(f"\nPer 32 entries (one warp-instruction covers 32 entries): {inst / w32:.1f} instructions executed, "
          f"{fp64_inst / w32:.1f} of them on the FP64 pipe; kernel time = {cyc / w32:.1f} SM-sub-partition cycles.")

    src = ncu_csv(a.report, "source")
    hdr_i = next(i for i, r in enumerate(src) if r and r[0] == "Address")
    sh = {n: i for i, n in enumerate(src[hdr_i])}
    rows = src[hdr_i + 1:]
    ex = [float(r[sh["Instructions Executed"]] or 0) for r in rows]
    smp = [float(r[sh["# Samples"]] or 0) for r in rows]
    sass = [r[sh["Source"]].strip() for r in rows]
    tot_s = sum(smp) or 1.0
    three = sum(e for e, s in zip(ex, sass) if s.split()[0:1] == ["DFMA"] and distinct_regs(s) == 3)
    fp64_src = sum(e for e, s in zip(ex, sass) if s.split() and s.split()[-1 if s.startswith("@") else 0] and
                   any(op in s.split()[1 if s.startswith("@") else 0] for op in FP64_OPS))
    print(f"From the source page: {fp64_src / w32:.1f} FP64-pipe instructions per 32 entries, "
          f"{three / w32:.1f} DFMAs with three distinct register sources.")
    slots = (inst + fp64_inst + three) / w32
    print(f"Issue-slot model: {inst / w32:.1f} + {fp64_inst / w32:.1f} + {three / w32:.1f} = {slots:.1f} slots "
          f"vs {cyc / w32:.1f} cycles measured -> {slots / (cyc / w32):.3f}.")

    if a.json:
        import json, os
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        try:
            git = subprocess.run(["git", "-C", root, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
        except Exception:
            git = None
        gb = lambda k: float(val(k)) * {"Gbyte": 1.0, "Mbyte": 1e-3, "Kbyte": 1e-6, "byte": 1e-9}[u[col[k]]]
        json.dump({"kernel": val("Kernel Name"), "git": git, "report": os.path.basename(a.report), "entries_per_launch": a.entries,
                   "gpu_time_ms": float(val("gpu__time_duration.sum")) * {"ms": 1.0, "us": 1e-3, "s": 1e3, "ns": 1e-6}[u[col["gpu__time_duration.sum"]]],
                   "dram_read_gb": gb("dram__bytes_read.sum"), "dram_write_gb": gb("dram__bytes_write.sum"),
                   "dram_gb_per_launch": gb("dram__bytes_read.sum") + gb("dram__bytes_write.sum"),
                   "inst_per_32_entries": inst / w32, "fp64_inst_per_32_entries": fp64_inst / w32,
                   "three_register_dfma_per_32_entries": three / w32, "cycles_per_32_entries": cyc / w32,
                   "fp64_pipe_pct_of_peak": fp64_pct, "issue_active_pct": float(val("smsp__issue_active.avg.pct_of_peak_sustained_active")),
                   "registers_per_thread": int(float(val("launch__registers_per_thread"))),
                   "note": "one `ncu --set full --clock-control none` capture of the first-pass kernel of a steady-state step "
                           "(scripts/ab_run.py c4slab); per-launch times under ncu are serialised, bench.py measures its own"},
                  open(a.json, "w"), indent=1)
    print("\n| SASS range | static instr | executions per 32 entries (each) | instr per 32 entries | FP64 instr per 32 entries | share of stall samples |")
    print("|---|---|---|---|---|---|")
    i = 0
    n = len(rows)
    while i < n:
        j = i
        while j + 1 < n and ex[i] > 0 and abs(ex[j + 1] - ex[i]) <= 0.02 * ex[i]:
            j += 1
        if j - i + 1 >= a.min_region and ex[i] > 0:
            tot = sum(ex[i:j + 1])
            dp = sum(e for e, s in zip(ex[i:j + 1], sass[i:j + 1])
                     if s.split() and any(op in s.split()[1 if s.startswith("@") else 0] for op in FP64_OPS))
            print(f"| {i}-{j} | {j - i + 1} | {ex[i] / w32:.2f} | {tot / w32:.1f} | {dp / w32:.1f} | "
                  f"{100.0 * sum(smp[i:j + 1]) / tot_s:.1f} % |")
        i = j + 1


if __name__ == "__main__":
    sys.exit(main())
