"""Module-level lazy re-exports (PEP 562).  The reference spreads one model over single-card / hybrid / pipeline / auto files and its users
import names from whichever file they knew; here each model has one implementation module plus a few siblings (``pipe``, ``generation``...).
``__getattr__ = lazy_exports(__name__, {...})`` lets the implementation module answer for names that live in a sibling without importing
it eagerly (the siblings import the implementation module themselves)."""
from __future__ import annotations

import importlib
from typing import Callable, Dict


def lazy_exports(module_name: str, table: Dict[str, str]) -> Callable[[str], object]:
    """``table``: exported name -> ``"relative.or.absolute.module"`` or ``"module:attribute"`` when the attribute has another name there."""
    package = module_name.rpartition(".")[0]

    def __getattr__(name: str):
        target = table.get(name)
        if target is None:
            raise AttributeError(f"module {module_name!r} has no attribute {name!r}")
        mod_name, _, attr = target.partition(":")
        mod = importlib.import_module(mod_name, package) if mod_name.startswith(".") else importlib.import_module(mod_name)
        return getattr(mod, attr or name)

    return __getattr__
