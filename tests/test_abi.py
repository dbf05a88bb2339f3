"""CPU checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/ebpmf_b200.h declares, the ctypes prototypes cover them, and -- there being no GPU in the
build container -- the product path fails loudly instead of falling back to a CPU path."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ebpmf_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ebpmf_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_path():
    names = declared_functions()
    for must in ("ebpmf_vga_newton", "ebpmf_vga_fixed_iter", "ebpmf_vga_low_memory", "ebpmf_ctx_vga_step",
                 "ebpmf_vga_step_host", "ebpmf_ctx_update_sigma2", "ebpmf_update_sigma2", "ebpmf_calc_obj",
                 "ebpmf_ctx_sum_lfactorial", "ebpmf_ctx_comm_init", "ebpmf_ctx_set_y_csc"):
        assert must in names


def test_library_exports_every_declared_symbol(pkg):
    assert os.path.exists(pkg.LIB_PATH), "run __graft_entry__.build() first"
    lib = ctypes.CDLL(pkg.LIB_PATH)
    for name in declared_functions():
        assert hasattr(lib, name), "symbol %s declared in the header but not exported" % name


def test_ctypes_prototypes_cover_the_header(pkg):
    assert sorted(pkg.SIGNATURES) == declared_functions()
    lib = pkg.load()
    assert lib.ebpmf_version().startswith(b"ebpmf_b200")


def test_library_has_sm100a_code_only(pkg):
    out = subprocess.run(["cuobjdump", "-lelf", pkg.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    archs = set(re.findall(r"sm_(\d+a?)", out.stdout))
    assert archs == {"100a"}, archs


def test_no_link_dependency_on_torch_or_nccl(pkg):
    out = subprocess.run(["ldd", pkg.LIB_PATH], capture_output=True, text=True).stdout
    assert "torch" not in out and "nccl" not in out and "libcudart" in out


def test_product_package_does_not_import_the_oracle():
    pkg_dir = os.path.join(ROOT, "ebpmf.alpha_b200")
    for base, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".c", ".h", ".R")):
                txt = open(os.path.join(base, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "liboracle" not in txt and "vga_oracle" not in txt, f


def test_fails_loudly_without_a_gpu(pkg):
    """No CPU fallback: on a box without a GPU every entry point errors (EBPMF_ERR_CUDA)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.EbpmfError) as ei:
        pkg.VgaContext(0)
    assert ei.value.code == 2 and "no CPU fallback" in str(ei.value)
    with pytest.raises(pkg.EbpmfError):
        pkg.vga_pois_solver_mat_newton(np.zeros((2, 2)), np.zeros((2, 2)), 1, np.zeros((2, 2)), np.ones((2, 2)))


def test_plain_c_host_compiles_links_and_fails_loudly_without_a_gpu(pkg, tmp_path):
    """examples/vga_step_demo.c: the header is valid C99, the library links from plain C with no other
    dependency, and without a GPU the program stops with the library's message (exit 2) -- no CPU path."""
    exe = str(tmp_path / "vga_step_demo")
    libdir = os.path.dirname(pkg.LIB_PATH)
    cmd = ["/usr/bin/gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", "vga_step_demo.c"), "-L", libdir, "-lebpmf_b200", "-Wl,-rpath," + libdir, "-lm",
           "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the demo is run by tests/test_gpu_parity.py")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 2 and "no CPU fallback" in r.stderr


def test_host_pool_hands_out_and_reuses_warm_blocks(pkg):
    """ebpmf_host_alloc / ebpmf_host_free (the R_allocator_t behind the glue's result matrices): no GPU needed
    for pageable blocks.  A freed block comes back for the next request of a similar size; a pointer that is
    not ours is refused; host_matrix() views are column-major and return their block when they die."""
    import gc
    pkg.host_pool_trim()
    s0 = pkg.host_pool_stats()
    a = pkg.host_matrix(300, 70)
    assert a.shape == (300, 70) and a.flags.f_contiguous and a.dtype == np.float64
    a[...] = np.arange(300 * 70, dtype=np.float64).reshape(70, 300).T
    assert a[5, 3] == 3 * 300 + 5
    addr = a.ctypes.data
    s1 = pkg.host_pool_stats()
    assert s1["fresh"] == s0["fresh"] + 1 and s1["live_blocks"] == s0["live_blocks"] + 1
    del a
    gc.collect()
    s2 = pkg.host_pool_stats()
    assert s2["live_blocks"] == s0["live_blocks"] and s2["kept_bytes"] >= 300 * 70 * 8
    b = pkg.host_matrix(300, 70)                    # same size: the warm block comes back
    assert b.ctypes.data == addr and pkg.host_pool_stats()["reused"] == s2["reused"] + 1
    c = pkg.host_matrix(300, 70)                    # second live matrix: a second block (the previous result stays alive)
    assert c.ctypes.data != addr
    del b, c
    gc.collect()
    lib = pkg.load()
    junk = np.zeros(4)
    assert lib.ebpmf_host_free(ctypes.c_void_p(junk.ctypes.data)) == 1          # EBPMF_ERR_ARG: not a pool block
    pkg.host_pool_trim()
    assert pkg.host_pool_stats()["kept_bytes"] == 0
