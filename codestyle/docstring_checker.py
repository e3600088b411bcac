"""Docstring style checker (pre-commit hook; the reference ships a pylint plugin for the same job, codestyle/docstring_checker.py — this
one is a stand-alone ``ast`` pass because the docstrings here are prose, not ``Args:`` tables).

Rules (each finding is ``path:line: CODE message``; exit status 1 if any):
  D101  a library module opens with a docstring (what it does; for re-implemented components also the reference file:line)
  D102  a docstring that fits on one line is written on one line (no lone opening / closing quote lines)
  D103  continuation lines of a docstring are indented at least as far as the line that opens it
  D104  a public function or method longer than ``--max-undocumented`` statements has a docstring
  D105  names listed under ``Args:`` / ``Parameters:`` exist in the signature (a stale table is worse than none)
  D106  a docstring is not just the function's name repeated

``check(path)`` keeps the historical boolean meaning (module docstring present) for callers that only want D101.
"""
from __future__ import annotations

import argparse
import ast
import re
import sys
from typing import List, Tuple

Finding = Tuple[int, str, str]


def _count_statements(node: ast.AST) -> int:
    return sum(isinstance(n, ast.stmt) for n in ast.walk(node)) - 1


def _doc_node(node):
    body = getattr(node, "body", None)
    if body and isinstance(body[0], ast.Expr) and isinstance(body[0].value, ast.Constant) and isinstance(body[0].value.value, str):
        return body[0]
    return None


def _arg_names(fn) -> set:
    a = fn.args
    names = {x.arg for x in a.posonlyargs + a.args + a.kwonlyargs}
    if a.vararg:
        names.add(a.vararg.arg)
    if a.kwarg:
        names.add(a.kwarg.arg)
    return names


_SECTION = re.compile(r"^\s*(Args|Arguments|Parameters)\s*:\s*$")
_ENTRY = re.compile(r"^\s*\*{0,2}([A-Za-z_][A-Za-z0-9_]*)\s*(\([^)]*\))?\s*:")
_OTHER_SECTION = re.compile(r"^\s*(Returns?|Raises|Yields?|Examples?|Notes?)\s*:\s*$")


def _documented_args(doc: str) -> List[str]:
    names, inside, indent = [], False, None
    for line in doc.splitlines():
        if _SECTION.match(line):
            inside, indent = True, None
            continue
        if inside:
            if not line.strip():
                continue
            if _OTHER_SECTION.match(line):
                inside = False
                continue
            cur = len(line) - len(line.lstrip())
            if indent is None:
                indent = cur
            if cur < indent:
                inside = False
                continue
            m = _ENTRY.match(line) if cur == indent else None
            if m:
                names.append(m.group(1))
    return names


def findings(path: str, src: str, max_undocumented: int = 40, library: bool = True) -> List[Finding]:
    out: List[Finding] = []
    try:
        tree = ast.parse(src)
    except SyntaxError as exc:
        return [(exc.lineno or 1, "D100", f"syntax error: {exc.msg}")]
    lines = src.splitlines()
    if library and src.strip() and not path.endswith("__init__.py") and ast.get_docstring(tree) is None:
        out.append((1, "D101", "missing module docstring"))
    for node in ast.walk(tree):
        if not isinstance(node, (ast.Module, ast.ClassDef, ast.FunctionDef, ast.AsyncFunctionDef)):
            continue
        dn = _doc_node(node)
        is_fn = isinstance(node, (ast.FunctionDef, ast.AsyncFunctionDef))
        if dn is None:
            if is_fn and not node.name.startswith("_") and _count_statements(node) > max_undocumented:
                out.append((node.lineno, "D104", f"public function '{node.name}' has {_count_statements(node)} statements and no docstring"))
            continue
        text = dn.value.value
        first, last = dn.lineno, dn.end_lineno
        stripped = text.strip()
        if first != last and "\n" not in stripped and len(stripped) + dn.col_offset + 6 <= 140:
            out.append((first, "D102", "one-line docstring written on several lines"))
        if first != last:
            base = len(lines[first - 1]) - len(lines[first - 1].lstrip())
            for i in range(first, last):
                ln = lines[i]
                if ln.strip() and len(ln) - len(ln.lstrip()) < base:
                    out.append((i + 1, "D103", "docstring continuation line is indented less than its opening line"))
                    break
        if is_fn:
            real = _arg_names(node)
            for name in _documented_args(text):
                if name not in real and name not in ("self", "cls"):
                    out.append((first, "D105", f"'{name}' is documented but is not a parameter of '{node.name}'"))
            if re.sub(r"[\W_]+", "", stripped.lower()) == re.sub(r"[\W_]+", "", node.name.lower()):
                out.append((first, "D106", f"docstring of '{node.name}' only repeats its name"))
    return sorted(out)


def check(path: str) -> bool:
    """``True`` when the file parses and (unless it is empty or a package ``__init__``) opens with a docstring."""
    with open(path, encoding="utf-8") as f:
        src = f.read()
    return not any(code in ("D100", "D101") for _, code, _ in findings(path, src))


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("files", nargs="*")
    ap.add_argument("--max-undocumented", type=int, default=40, help="statements a public function may have without a docstring (D104)")
    ap.add_argument("--select", default="", help="comma-separated codes to report (default: all)")
    args = ap.parse_args(argv)
    select = {c.strip() for c in args.select.split(",") if c.strip()}
    bad = 0
    for p in args.files:
        with open(p, encoding="utf-8") as f:
            src = f.read()
        for line, code, msg in findings(p, src, args.max_undocumented):
            if select and code not in select:
                continue
            print(f"{p}:{line}: {code} {msg}")
            bad += 1
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
