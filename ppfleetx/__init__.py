"""Drop-in import alias: ``import ppfleetx.<anything>`` resolves to ``paddlefleetx_b200.<anything>`` so that user code written
against the reference package layout (``ppfleetx.utils.config``, ``ppfleetx.core.EagerEngine``, ``ppfleetx.models.build_module``,
``ppfleetx.data.build_dataloader``, ``ppfleetx.distributed.apis.env`` …) keeps working unchanged."""
import importlib
import importlib.abc
import importlib.util
import sys

import paddlefleetx_b200 as _impl

_PREFIX, _TARGET = "ppfleetx", "paddlefleetx_b200"


class _AliasFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname == _PREFIX or not fullname.startswith(_PREFIX + "."):
            return None
        real = _TARGET + fullname[len(_PREFIX):]
        try:
            if importlib.util.find_spec(real) is None:
                return None
        except (ImportError, ValueError):
            return None
        return importlib.util.spec_from_loader(fullname, self)

    def create_module(self, spec):
        mod = importlib.import_module(_TARGET + spec.name[len(_PREFIX):])
        sys.modules[spec.name] = mod
        return mod

    def exec_module(self, module):
        pass


sys.meta_path.insert(0, _AliasFinder())
__path__ = list(_impl.__path__)
__version__ = getattr(_impl, "__version__", "0.1.0")
