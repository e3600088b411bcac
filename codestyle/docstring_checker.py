"""Every library module must open with a docstring that says what it does (and, for re-implemented components, which reference
file:line it corresponds to).  Used as a pre-commit hook; exits non-zero listing offenders."""
import ast
import sys


def check(path: str) -> bool:
    with open(path, encoding="utf-8") as f:
        src = f.read()
    if not src.strip() or path.endswith("__init__.py"):
        return True
    try:
        return ast.get_docstring(ast.parse(src)) is not None
    except SyntaxError as exc:
        print(f"{path}: syntax error {exc}")
        return False


if __name__ == "__main__":
    bad = [p for p in sys.argv[1:] if not check(p)]
    for p in bad:
        print(f"{p}: missing module docstring")
    sys.exit(1 if bad else 0)
