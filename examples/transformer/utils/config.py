"""Config handling for the engine-less examples (reference examples/transformer/utils/config.py:36-600 re-implements the
whole config stack a second time; here the examples share the framework's loader and only add the GPT post-processing)."""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..", "..")))

from paddlefleetx_b200.utils.config import AttrDict, get_config as _get_config, override_config, parse_args, parse_config, print_config  # noqa: E402,F401


def process_configs(config):
    """Model-family defaults (ffn = 4h, recompute granularity, vocab padding, dataset sizing)."""
    from paddlefleetx_b200.models.language_model.language_module import process_configs as gpt_process

    return gpt_process(config)


def get_config(fname, overrides=None, show=False):
    cfg = _get_config(fname, overrides=overrides, show=False)
    cfg = process_configs(cfg) or cfg
    if show:
        print_config(cfg)
    return cfg
