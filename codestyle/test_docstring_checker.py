"""Unit test of the docstring checker (reference codestyle ships a unittest for its pylint docstring checker)."""
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from docstring_checker import check  # noqa: E402


def test_checker():
    with tempfile.TemporaryDirectory() as d:
        good, bad = os.path.join(d, "g.py"), os.path.join(d, "b.py")
        open(good, "w").write('"""doc"""\nx = 1\n')
        open(bad, "w").write("x = 1\n")
        assert check(good) and not check(bad)


if __name__ == "__main__":
    test_checker()
    print("ok")
