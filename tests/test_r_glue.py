"""The R side of the boundary (ebpmf.alpha_b200/r) cannot be built or run here -- R is not installed.  What CAN be
checked without R: r_glue.c is valid C against the documented signatures of the R API subset it uses
(tests/r_api_stub, declarations only, never linked), every registered .Call routine has the arity of its
definition, and every `.Call(C_xxx, ...)` in the R wrappers passes that many arguments."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GLUE = os.path.join(ROOT, "ebpmf.alpha_b200", "r", "src", "r_glue.c")
RWRAP = os.path.join(ROOT, "ebpmf.alpha_b200", "r", "R", "vga_cuda.R")


def registered():
    src = open(GLUE).read()
    return {m.group(1): int(m.group(2)) for m in re.finditer(r'\{"(C_\w+)",\s*\(DL_FUNC\)&\1,\s*(\d+)\}', src)}


def test_glue_is_valid_c_against_the_r_api_subset():
    cmd = ["/usr/bin/gcc", "-std=c99", "-fsyntax-only", "-Wall", "-Wextra", "-Werror", "-Wno-cast-function-type",
           "-Wno-unused-parameter", "-I", os.path.join(ROOT, "tests", "r_api_stub"), "-I", os.path.join(ROOT, "include"), GLUE]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_registered_arity_matches_definitions():
    src = open(GLUE).read()
    reg = registered()
    assert len(reg) >= 13
    for name, nargs in reg.items():
        m = re.search(r"^SEXP %s\(([^)]*)\)" % name, src, re.M | re.S)
        assert m, "no definition of %s" % name
        assert m.group(1).count("SEXP ") == nargs, (name, nargs, m.group(1))


def call_arg_counts(text):
    """(routine, number of arguments after the routine) for every .Call(...) in an R source"""
    out = []
    for m in re.finditer(r"\.Call\((?=\s*C_\w+\s*,)", text):
        i = m.end(); depth = 1; commas = 0; in_str = None
        while depth and i < len(text):
            c = text[i]
            if in_str:
                in_str = None if c == in_str else in_str
            elif c in "'\"":
                in_str = c
            elif c in "([{":
                depth += 1
            elif c in ")]}":
                depth -= 1
            elif c == "," and depth == 1:
                commas += 1
            i += 1
        name = re.match(r"\s*(C_\w+)", text[m.end():]).group(1)
        out.append((name, commas))
    return out


def test_r_wrappers_pass_the_registered_number_of_arguments():
    reg = registered()
    calls = call_arg_counts(open(RWRAP).read())
    assert len(calls) >= 12
    for name, nargs in calls:
        assert name in reg, name + " is not registered in r_glue.c"
        assert reg[name] == nargs, (name, "registered", reg[name], ".Call passes", nargs)
    # the integration snippet uses the same routines
    for name, nargs in call_arg_counts(open(os.path.join(ROOT, "INTEGRATION.md")).read()):
        assert name in reg and reg[name] == nargs, (name, nargs)
