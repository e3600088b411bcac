"""Source hygiene for the native tree (``ops/csrc``, the C++ index builder): the checks of the reference's cpplint / clang-format hooks that
do not need those tools installed (codestyle/cpplint_pre_commit.hook, clang_format.hook).

  C101  no tab characters          C102  no trailing whitespace           C103  line length <= ``--max-line`` (default 160)
  C104  headers have ``#pragma once`` or an include guard               C105  no ``using namespace`` at file scope in a header
  C106  file ends with exactly one newline

Exit status 1 if any finding; ``path:line: CODE message`` per finding.
"""
from __future__ import annotations

import argparse
import re
import sys
from typing import List, Tuple

HEADER = re.compile(r"\.(h|hpp|cuh)$")


def findings(path: str, text: str, max_line: int = 160) -> List[Tuple[int, str, str]]:
    out = []
    lines = text.split("\n")
    for i, line in enumerate(lines, 1):
        if "\t" in line:
            out.append((i, "C101", "tab character"))
        if line != line.rstrip():
            out.append((i, "C102", "trailing whitespace"))
        if len(line) > max_line:
            out.append((i, "C103", f"line is {len(line)} characters long (limit {max_line})"))
    if HEADER.search(path):
        if "#pragma once" not in text and not re.search(r"#ifndef\s+(\w+)\s*\n\s*#define\s+\1", text):
            out.append((1, "C104", "header without #pragma once / include guard"))
        for i, line in enumerate(lines, 1):
            if re.match(r"using\s+namespace\s+\w", line):
                out.append((i, "C105", "'using namespace' at file scope in a header"))
    if text and (not text.endswith("\n") or text.endswith("\n\n")):
        out.append((len(lines), "C106", "file must end with exactly one newline"))
    return out


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("files", nargs="*")
    ap.add_argument("--max-line", type=int, default=160)
    args = ap.parse_args(argv)
    bad = 0
    for p in args.files:
        with open(p, encoding="utf-8") as f:
            text = f.read()
        for line, code, msg in findings(p, text, args.max_line):
            print(f"{p}:{line}: {code} {msg}")
            bad += 1
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
