"""Drop-in import alias: ``import ppfleetx.<anything>`` resolves to ``paddlefleetx_b200.<anything>`` so that user code written
against the reference package layout (``ppfleetx.utils.config``, ``ppfleetx.core.EagerEngine``, ``ppfleetx.models.build_module``,
``ppfleetx.data.build_dataloader``, ``ppfleetx.distributed.apis.env`` …) keeps working unchanged."""
import importlib
import importlib.abc
import importlib.util
import sys

import paddlefleetx_b200 as _impl

_PREFIX, _TARGET = "ppfleetx", "paddlefleetx_b200"

# Modules whose contents live under a different path here (the reference splits single-card / hybrid / auto variants of a model into
# separate files; this framework has one topology-aware implementation per model).  Keys and values are relative to the package root.
_LM, _VM, _MM = "models.language_model", "models.vision_model", "models.multimodal_model"
_MOVED = {
    f"{_LM}.gpt.dygraph": f"{_LM}.gpt", f"{_LM}.gpt.dygraph.single_model": f"{_LM}.gpt.model", f"{_LM}.gpt.dygraph.hybrid_model": f"{_LM}.gpt.model",
    f"{_LM}.gpt.dygraph.processor": f"{_LM}.gpt.processor", f"{_LM}.gpt.dygraph.sequence_parallel_utils": "parallel.tp_layers",
    f"{_LM}.gpt.auto": f"{_LM}.auto_module", f"{_LM}.gpt.auto.auto_model": f"{_LM}.gpt.model", f"{_LM}.gpt.auto.auto_module": f"{_LM}.auto_module",
    f"{_LM}.ernie.dygraph": f"{_LM}.ernie", f"{_LM}.ernie.dygraph.single_model": f"{_LM}.ernie.model", f"{_LM}.ernie.dygraph.hybrid_model": f"{_LM}.ernie.model",
    f"{_LM}.ernie.auto": f"{_LM}.ernie", f"{_LM}.ernie.auto.auto_model": f"{_LM}.ernie.model", f"{_LM}.ernie.auto.auto_module": f"{_LM}.ernie.ernie_module",
    f"{_LM}.ernie.auto.auto_transformer": f"{_LM}.ernie.model", f"{_LM}.ernie.layers": f"{_LM}.ernie.model", f"{_LM}.ernie.layers.transformer": f"{_LM}.ernie.model",
    f"{_LM}.ernie.layers.distributed_transformer": f"{_LM}.ernie.model", f"{_LM}.ernie.layers.model_outputs": f"{_LM}.ernie.model_outputs",
    f"{_LM}.ernie.layers.utils": f"{_LM}.ernie.utils",
    f"{_LM}.t5": f"{_MM}.t5", f"{_LM}.t5.modeling": f"{_MM}.t5.modeling", f"{_LM}.t5.utils": f"{_MM}.t5.utils", f"{_LM}.debertav2": f"{_MM}.debertav2", f"{_LM}.debertav2.modeling": f"{_MM}.debertav2.modeling",
    f"{_LM}.utils": f"{_LM}.language_module", f"{_LM}.auto_utils": f"{_LM}.language_module",
    f"{_LM}.moe.gate.base_gate": f"{_LM}.moe.gate.gates", f"{_LM}.moe.gate.naive_gate": f"{_LM}.moe.gate.gates", f"{_LM}.moe.gate.gshard_gate": f"{_LM}.moe.gate.gates",
    f"{_LM}.moe.gate.switch_gate": f"{_LM}.moe.gate.gates", f"{_LM}.moe.comm": f"{_LM}.moe.comm_ops",
    f"{_VM}.layers.attention": f"{_VM}.vit.vit", f"{_VM}.layers.mlp": f"{_VM}.vit.vit", f"{_VM}.layers.droppath": f"{_VM}.layers", f"{_VM}.layers.embedding": f"{_VM}.layers",
    f"{_VM}.layers.identity": f"{_VM}.layers", f"{_VM}.layers.initializer": f"{_VM}.layers", f"{_VM}.loss.cross_entropy": f"{_VM}.loss", f"{_VM}.metrics.accuracy": f"{_VM}.metrics",
    "data.tokenizers.t5_tokenization_utils": "data.tokenizers.tokenization_utils_base",
    "data.data_tools.ernie.preprocess": "data.data_tools.ernie", "data.data_tools.ernie.preprocess.create_pretraining_data": "data.data_tools.ernie.create_pretraining_data",
    "data.data_tools.ernie.preprocess.trans_to_json": "data.data_tools.ernie.trans_to_json", "data.data_tools.ernie.preprocess.words_segmentation": "data.data_tools.ernie.words_segmentation",
    "data.data_tools.ernie.preprocess.ernie_dataset": "data.dataset.ernie.ernie_dataset", "data.data_tools.ernie.preprocess.dataset_utils": "data.dataset.ernie.dataset_utils",
    "ops.topp_sampling": "ops.functional",
}


def _real_name(fullname: str) -> str:
    rel = fullname[len(_PREFIX) + 1:]
    return _TARGET + "." + _MOVED.get(rel, rel)


class _AliasFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname == _PREFIX or not fullname.startswith(_PREFIX + "."):
            return None
        real = _real_name(fullname)
        try:
            if importlib.util.find_spec(real) is None:
                return None
        except (ImportError, ValueError):
            return None
        return importlib.util.spec_from_loader(fullname, self)

    def create_module(self, spec):
        mod = importlib.import_module(_real_name(spec.name))
        rel = spec.name[len(_PREFIX) + 1:]
        if not hasattr(mod, "__path__") and any(k.startswith(rel + ".") for k in _MOVED):
            # a relocated *package* whose contents now sit in one plain module: hand out a package-shaped view of that module so its
            # (equally relocated) sub-modules can still be imported through it
            import types

            view = types.ModuleType(spec.name, mod.__doc__)
            view.__dict__.update({k: v for k, v in mod.__dict__.items() if not k.startswith("__")})
            view.__path__ = []
            mod = view
        sys.modules[spec.name] = mod
        return mod

    def exec_module(self, module):
        pass


sys.meta_path.insert(0, _AliasFinder())
__path__ = list(_impl.__path__)
__version__ = getattr(_impl, "__version__", "0.1.0")
