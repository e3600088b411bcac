"""The tree passes its own pre-commit checks (codestyle/): docstring rules over the library and the tools, source hygiene over the native tree,
no torch headers in kernel sources; plus the checkers' unit tests."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "codestyle"))


def _tracked(*patterns):
    out = subprocess.run(["git", "ls-files", *patterns], cwd=ROOT, capture_output=True, text=True).stdout.split()
    if out:
        return out
    found = []                                   # not a git checkout (exported tree): walk
    for base, _, files in os.walk(ROOT):
        if any(part in base for part in ("/.git", "/build", "/baseline", "/gpurun_out")):
            continue
        for f in files:
            rel = os.path.relpath(os.path.join(base, f), ROOT)
            if any(rel.endswith(p.lstrip("*")) for p in patterns):
                found.append(rel)
    return found


def test_docstring_rules_hold_over_the_library_and_tools():
    import docstring_checker as D

    files = [f for f in _tracked("*.py") if f.startswith(("paddlefleetx_b200/", "tools/", "tasks/")) or f == "bench.py"]
    assert len(files) > 150
    bad = []
    for f in files:
        with open(os.path.join(ROOT, f), encoding="utf-8") as fh:
            bad += [(f, *x) for x in D.findings(f, fh.read(), max_undocumented=40)]
    assert not bad, bad[:10]


def test_native_sources_are_clean():
    import cpp_lint as L
    import kernel_include_checker as K

    files = _tracked("*.cu", "*.cuh", "*.cpp", "*.h")
    assert len(files) >= 20
    bad = []
    for f in files:
        with open(os.path.join(ROOT, f), encoding="utf-8") as fh:
            text = fh.read()
        bad += [(f, *x) for x in L.findings(f, text, max_line=230)]
        if f.endswith((".cu", ".cuh")):
            bad += [(f, n, "torch header") for n, line in enumerate(text.split("\n"), 1) if K.BAD.search(line)]
    assert not bad, bad[:10]
    assert L.findings("x.h", "int a;\t\nusing namespace std;\n\n") and not L.findings("x.h", "#pragma once\nint a;\n")


def test_checker_unit_tests():
    import test_docstring_checker as T

    for name in dir(T):
        if name.startswith("test_"):
            getattr(T, name)()
