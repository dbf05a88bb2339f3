"""The committed ncu counters that bench.py quotes (profiles/r02_ncu_metrics.json) must describe a kernel that
the shipped library actually contains, and carry what bench.py reads."""
import json
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_ncu_metrics_file_matches_the_built_library(pkg):
    d = json.load(open(os.path.join(ROOT, "profiles", "r02_ncu_metrics.json")))
    for key in ("kernel", "git", "dram_gb_per_launch", "inst_per_32_entries", "fp64_inst_per_32_entries", "cycles_per_32_entries"):
        assert key in d, key
    m = re.match(r"void vga_fused_kernel<(\d+), (\d+), (\d+), (\d+), (\d+)>\(FusedArgs\)", d["kernel"])
    assert m, d["kernel"]
    mangled = "_ZN5ebpmf16vga_fused_kernelILi%sELi%sELi%sELi%sELi%sEEEvNS_9FusedArgsE" % m.groups()
    out = subprocess.run(["cuobjdump", "-elf", "-fun", mangled, pkg.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0 and "not found" not in (out.stdout + out.stderr):
        pytest.skip("cuobjdump unavailable")
    syms = subprocess.run(["cuobjdump", "-symbols", pkg.LIB_PATH], capture_output=True, text=True).stdout
    assert mangled in syms, "profiles/r02_ncu_metrics.json names a kernel the library does not contain"
    assert 60.0 < d["dram_gb_per_launch"] < 75.0          # the step reads M once and writes it once (62.3 GB algorithmic)
    assert d["fp64_inst_per_32_entries"] < d["inst_per_32_entries"]
