"""Every Python module of the reference package imports under the same dotted name here (``ppfleetx.<path>``), and the public names it
defines at module level exist in ours.  This is the "switch the import and keep going" contract of the ``ppfleetx`` alias package."""
import ast
import importlib
import os

import pytest

REF = "/root/reference/ppfleetx"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present")


def _reference_modules():
    for dirpath, _, files in os.walk(REF):
        for f in sorted(files):
            if f.endswith(".py"):
                rel = os.path.relpath(os.path.join(dirpath, f), os.path.dirname(REF))[:-3]
                name = rel.replace(os.sep, ".")
                yield (name[:-len(".__init__")] if name.endswith(".__init__") else name), os.path.join(dirpath, f)


def test_every_reference_module_name_imports():
    missing = []
    n = 0
    for name, _ in _reference_modules():
        n += 1
        try:
            importlib.import_module(name)
        except Exception as e:          # noqa: BLE001 - the report wants every failure, whatever its type
            missing.append(f"{name}: {type(e).__name__}: {e}")
    assert n > 150
    assert not missing, "\n".join(missing)


_SKIP_PREFIX = ("_",)
# public names of the reference not provided yet, one ``module.name`` per line; the test fails when a name outside this file goes missing AND
# when a listed name has been provided (so the file can only shrink)
_KNOWN_GAPS_FILE = os.path.join(os.path.dirname(__file__), "reference_namespace_known_gaps.txt")


def _known_gaps():
    if not os.path.exists(_KNOWN_GAPS_FILE):
        return set()
    with open(_KNOWN_GAPS_FILE, encoding="utf-8") as f:
        return {line.strip() for line in f if line.strip() and not line.startswith("#")}


def _public_defs(path):
    tree = ast.parse(open(path, encoding="utf-8").read())
    out = []
    for node in tree.body:
        if isinstance(node, (ast.ClassDef, ast.FunctionDef)) and not node.name.startswith(_SKIP_PREFIX):
            out.append(node.name)
    return out


def test_public_classes_and_functions_exist():
    absent = set()
    total = 0
    for name, path in _reference_modules():
        try:
            mod = importlib.import_module(name)
        except Exception:               # noqa: BLE001 - reported by the test above
            continue
        for sym in _public_defs(path):
            total += 1
            if not hasattr(mod, sym):
                absent.add(f"{name}.{sym}")
    assert total > 500
    known = _known_gaps()
    new_gaps, closed = sorted(absent - known), sorted(known - absent)
    assert not new_gaps, f"{len(new_gaps)} of {total} public names missing:\n" + "\n".join(new_gaps)
    assert not closed, "provided now, remove from reference_namespace_known_gaps.txt:\n" + "\n".join(closed)
