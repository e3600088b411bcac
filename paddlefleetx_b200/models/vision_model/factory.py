"""``build({name: X, **kw})`` looks X up in the vision namespace: models, losses, metrics, layers (reference factory.py:28-35)."""
import copy


def build(config):
    from . import layers, loss, metrics, moco, resnet, vit

    cfg = copy.deepcopy(dict(config))
    name = cfg.pop("name")
    for mod in (vit, moco, resnet, loss, metrics, layers):
        if hasattr(mod, name):
            return getattr(mod, name)(**cfg)
    raise ValueError(f"{name} is not a known vision model / loss / metric")
