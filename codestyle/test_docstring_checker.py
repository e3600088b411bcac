"""Unit tests of the docstring checker, one per rule (the reference ships a unittest for its pylint docstring plugin,
codestyle/test_docstring_checker.py)."""
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from docstring_checker import check, findings, main  # noqa: E402


def _codes(src, **kw):
    return [c for _, c, _ in findings("m.py", src, **kw)]


def test_module_docstring_rule_and_boolean_entry():
    with tempfile.TemporaryDirectory() as d:
        good, bad, init = os.path.join(d, "g.py"), os.path.join(d, "b.py"), os.path.join(d, "__init__.py")
        open(good, "w").write('"""doc"""\nx = 1\n')
        open(bad, "w").write("x = 1\n")
        open(init, "w").write("x = 1\n")
        assert check(good) and not check(bad) and check(init)
        assert main([good]) == 0 and main([bad]) == 1 and main([bad, "--select", "D104"]) == 0
    assert _codes("def f(:\n") == ["D100"]


def test_one_line_docstring_on_several_lines():
    assert "D102" in _codes('"""m"""\ndef f():\n    """\n    short\n    """\n')
    assert "D102" not in _codes('"""m"""\ndef f():\n    """first\n    second"""\n')


def test_continuation_indent():
    assert "D103" in _codes('"""m"""\ndef f():\n    """first\n  second"""\n')
    assert "D103" not in _codes('"""m"""\ndef f():\n    """first\n    second\n\n        indented example"""\n')


def test_long_public_function_needs_a_docstring():
    body = "".join(f"    x{i} = {i}\n" for i in range(12))
    src = '"""m"""\ndef public():\n' + body + "def _private():\n" + body
    assert _codes(src, max_undocumented=10) == ["D104"]
    assert _codes(src, max_undocumented=20) == []


def test_documented_args_must_exist():
    src = '"""m"""\ndef f(a, *rest, key=None, **kw):\n    """Do it.\n\n    Args:\n        a (int): first\n        rest: more\n        key: k\n        gone: removed long ago\n\n    Returns:\n        nothing: really\n    """\n'
    found = findings("m.py", src)
    assert [c for _, c, _ in found] == ["D105"] and "'gone'" in found[0][2]


def test_docstring_that_repeats_the_name():
    assert "D106" in _codes('"""m"""\ndef get_world_size():\n    """Get world size."""\n')
    assert "D106" not in _codes('"""m"""\ndef get_world_size():\n    """Number of ranks in the default group."""\n')


if __name__ == "__main__":
    for name, fn in list(globals().items()):
        if name.startswith("test_"):
            fn()
    print("ok")
