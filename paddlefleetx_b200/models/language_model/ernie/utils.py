"""Construction helpers of the ERNIE family: recorded constructor arguments and tolerant ``forward`` patches.

Reference: ppfleetx/models/language_model/ernie/layers/utils.py:23-174.  ``InitTrackerMeta`` is the metaclass of the reference's pretrained
model classes: every instance remembers the keyword arguments it was built with (``init_config``), with optional ``_pre_init`` /
``_post_init`` hooks around ``__init__``; assigning a ``forward`` that predates the ``output_hidden_states`` / ``output_attentions`` /
``return_dict`` arguments wraps it so that callers passing those still work.
"""
from __future__ import annotations

import functools
import inspect
import warnings

import torch.nn as nn

_NEW_FORWARD_ARGS = ("output_hidden_states", "output_attentions", "return_dict")


def fn_args_to_dict(func, *args, **kwargs) -> dict:
    """Name -> value for a call ``func(*args, **kwargs)``: positionals by parameter order, then declared defaults, then keywords."""
    spec = inspect.getfullargspec(func)
    bound = dict(zip(spec.args, args))
    if spec.defaults:
        for name, default in zip(spec.args[-len(spec.defaults):], spec.defaults):
            bound.setdefault(name, default)
    bound.update(kwargs)
    return bound


def adapt_stale_fwd_patch(self, name, value):
    """``obj.forward = patch`` where ``patch`` lacks the newer keyword arguments of the current ``forward``: returns a wrapper that drops them."""
    if name != "forward" or type(value).__name__.endswith("StaticFunction") or not callable(value):
        return value
    try:
        patch_args = inspect.getfullargspec(value).args
        current_args = inspect.getfullargspec(self.forward).args
    except TypeError:
        return value
    missing = [a for a in _NEW_FORWARD_ARGS if a in current_args and a not in patch_args]
    if not missing:
        return value
    who = type(self) if isinstance(self, nn.Module) else self
    warnings.warn(f"The `forward` of {who} is patched with a function that does not take {missing}; these arguments are dropped "
                  "before the patch is called.  The patch should be updated.")
    pass_self = isinstance(self, nn.Module) and inspect.isfunction(value) and not inspect.ismethod(value)

    @functools.wraps(value)
    def wrapped(*args, **kwargs):
        for a in missing:
            kwargs.pop(a, None)
        return value(self, *args, **kwargs) if pass_self else value(*args, **kwargs)

    return wrapped


class InitTrackerMeta(type(nn.Module)):
    """Metaclass that records constructor keyword arguments on the instance as ``init_config`` (plus ``init_args`` for positionals and
    ``init_class``), running ``cls._pre_init(self, init_fn, *a, **kw)`` / ``cls._post_init(...)`` around ``__init__`` when defined."""

    def __init__(cls, name, bases, attrs):
        own_init = "__init__" in attrs
        cls.__init__ = InitTrackerMeta.init_and_track_conf(cls.__init__, getattr(cls, "_pre_init", None) if own_init else None,
                                                           getattr(cls, "_post_init", None) if own_init else None)
        super().__init__(name, bases, attrs)

    @staticmethod
    def init_and_track_conf(init_func, pre_init_func=None, post_init_func=None):
        @functools.wraps(init_func)
        def tracked(self, *args, **kwargs):
            if pre_init_func:
                pre_init_func(self, init_func, *args, **kwargs)
            init_func(self, *args, **kwargs)
            if post_init_func:
                post_init_func(self, init_func, *args, **kwargs)
            config = dict(kwargs)
            if args:
                config["init_args"] = args
            config["init_class"] = type(self).__name__
            object.__setattr__(self, "init_config", config)

        return tracked

    def __setattr__(cls, name, value):
        super().__setattr__(name, adapt_stale_fwd_patch(cls, name, value))
