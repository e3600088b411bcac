"""``transform`` / ``create_preprocess_operators`` under the reference's module path (ppfleetx/data/transforms/utils.py:18-45)."""
from .preprocess import build_transforms, transform  # noqa: F401


def create_preprocess_operators(params):
    """``[{OpName: {kwargs}}, ...]`` (the YAML ``transform_ops`` list) -> list of callables."""
    return build_transforms(params)
