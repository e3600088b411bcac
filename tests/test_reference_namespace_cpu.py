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


# names the reference defines only as scaffolding for Paddle itself (static-graph specs, custom-op registration, private helpers of its
# file-local implementation): not part of what a user imports
_SKIP_PREFIX = ("_",)
_ALLOWED_ABSENT = {
    # module -> names that have no meaning outside Paddle
}


def _public_defs(path):
    tree = ast.parse(open(path, encoding="utf-8").read())
    out = []
    for node in tree.body:
        if isinstance(node, (ast.ClassDef, ast.FunctionDef)) and not node.name.startswith(_SKIP_PREFIX):
            out.append(node.name)
    return out


def test_public_classes_and_functions_exist():
    absent = []
    total = 0
    for name, path in _reference_modules():
        try:
            mod = importlib.import_module(name)
        except Exception:               # noqa: BLE001 - reported by the test above
            continue
        for sym in _public_defs(path):
            total += 1
            if not hasattr(mod, sym) and sym not in _ALLOWED_ABSENT.get(name, ()):
                absent.append(f"{name}.{sym}")
    assert total > 500
    assert not absent, f"{len(absent)} of {total} public names missing:\n" + "\n".join(absent)
