"""Structured model outputs for the ERNIE family (``return_dict=True``).

Reference: ppfleetx/models/language_model/ernie/layers/model_outputs.py:35-464 — a ``ModelOutput`` is at once a dataclass, an ordered
mapping of its non-``None`` fields and a tuple of them: ``out.logits``, ``out["logits"]`` and ``out[0]`` address the same tensor, ``None``
fields disappear from the mapping / tuple views, and the mapping cannot be edited structurally (no ``del`` / ``pop`` / ``update``).

The behaviour is re-implemented on a small explicit core (``_sync`` keeps the mapping view equal to the set of non-``None`` fields) rather than
on the reference's ``__post_init__`` state machine; the field lists of the concrete outputs are the reference's.
"""
from __future__ import annotations

from collections import OrderedDict
from dataclasses import dataclass, fields
from typing import Any, Optional, Tuple

import numpy as np
import torch


def is_tensor(x) -> bool:
    return isinstance(x, (torch.Tensor, np.ndarray))


class ModelOutput(OrderedDict):
    """Base of all structured outputs.  Subclasses are ``@dataclass``es whose fields after the first default to ``None``."""

    def __post_init__(self):
        fs = fields(self)
        if not fs:
            raise ValueError(f"{type(self).__name__} has no fields.")
        if any(f.default is not None for f in fs[1:]):
            raise ValueError(f"{type(self).__name__} should not have more than one required field.")
        for f in fs:                                   # lists become tuples (per-layer collections)
            v = getattr(self, f.name)
            if isinstance(v, list):
                object.__setattr__(self, f.name, tuple(v))
        first = getattr(self, fs[0].name)
        rest_empty = all(getattr(self, f.name) is None for f in fs[1:])
        if rest_empty and not is_tensor(first) and first is not None:
            # a mapping or an iterable of (name, value) pairs passed as the only argument populates the fields by name
            pairs = first.items() if isinstance(first, dict) else first
            try:
                pairs = list(pairs)
            except TypeError:
                pairs = None
            if pairs is not None and pairs and all(isinstance(p, (tuple, list)) and len(p) == 2 and isinstance(p[0], str) for p in pairs):
                object.__setattr__(self, fs[0].name, None)
                for k, v in pairs:
                    object.__setattr__(self, k, v)
        self._sync()

    def _sync(self):
        OrderedDict.clear(self)
        for f in fields(self):
            v = getattr(self, f.name)
            if v is not None:
                OrderedDict.__setitem__(self, f.name, v)

    # ---- structural edits of the mapping view are refused (reference model_outputs.py:100-118)
    def _refuse(self, what):
        raise Exception(f"You cannot use ``{what}`` on a {type(self).__name__} instance.")

    def __delitem__(self, *a, **k): self._refuse("__delitem__")
    def setdefault(self, *a, **k): self._refuse("setdefault")
    def pop(self, *a, **k): self._refuse("pop")
    def update(self, *a, **k): self._refuse("update")

    def __getitem__(self, k):
        if isinstance(k, str):
            return OrderedDict.__getitem__(self, k)
        return self.to_tuple()[k]

    def __setattr__(self, name, value):
        object.__setattr__(self, name, value)
        if name in {f.name for f in fields(self)}:
            if value is None:
                if name in self.keys():
                    OrderedDict.__delitem__(self, name)
            else:
                OrderedDict.__setitem__(self, name, value)

    def __setitem__(self, key, value):
        OrderedDict.__setitem__(self, key, value)
        object.__setattr__(self, key, value)

    def __reduce__(self):            # OrderedDict's pickling would call __init__ without the dataclass arguments
        return (type(self), tuple(getattr(self, f.name) for f in fields(self)))

    def to_tuple(self) -> Tuple[Any, ...]:
        """All non-``None`` fields, in declaration order."""
        return tuple(OrderedDict.__getitem__(self, k) for k in self.keys())


TupleOfTensors = Optional[Tuple[torch.Tensor, ...]]


@dataclass
class ErnieForPreTrainingOutput(ModelOutput):
    """``loss`` (MLM + sentence-order, when labels were given), ``prediction_logits`` [b, s, vocab] (or [n_masked, vocab]),
    ``seq_relationship_logits`` [b, 2], per-layer ``hidden_states`` (embedding output first) and ``attentions`` [b, heads, s, s]."""
    loss: Optional[torch.Tensor] = None
    prediction_logits: Optional[torch.Tensor] = None
    seq_relationship_logits: Optional[torch.Tensor] = None
    hidden_states: TupleOfTensors = None
    attentions: TupleOfTensors = None


@dataclass
class BaseModelOutputWithPastAndCrossAttentions(ModelOutput):
    """Encoder stack output: ``last_hidden_state`` [b, s, h], ``past_key_values`` (per layer ``(k, v)`` of [b, heads, s, d]),
    ``hidden_states``, ``attentions``, ``cross_attentions`` (always ``None`` for the encoder-only ERNIE)."""
    last_hidden_state: Optional[torch.Tensor] = None
    past_key_values: Optional[Tuple[Tuple[torch.Tensor, ...], ...]] = None
    hidden_states: TupleOfTensors = None
    attentions: TupleOfTensors = None
    cross_attentions: TupleOfTensors = None


@dataclass
class BaseModelOutputWithPoolingAndCrossAttentions(ModelOutput):
    """``ErnieModel`` output: the encoder fields plus ``pooler_output`` [b, h] (tanh-dense of the first token)."""
    last_hidden_state: Optional[torch.Tensor] = None
    pooler_output: Optional[torch.Tensor] = None
    past_key_values: Optional[Tuple[Tuple[torch.Tensor, ...], ...]] = None
    hidden_states: TupleOfTensors = None
    attentions: TupleOfTensors = None
    cross_attentions: TupleOfTensors = None


@dataclass
class SequenceClassifierOutput(ModelOutput):
    loss: Optional[torch.Tensor] = None
    logits: Optional[torch.Tensor] = None
    hidden_states: TupleOfTensors = None
    attentions: TupleOfTensors = None


@dataclass
class TokenClassifierOutput(ModelOutput):
    loss: Optional[torch.Tensor] = None
    logits: Optional[torch.Tensor] = None
    hidden_states: TupleOfTensors = None
    attentions: TupleOfTensors = None


@dataclass
class QuestionAnsweringModelOutput(ModelOutput):
    loss: Optional[torch.Tensor] = None
    start_logits: Optional[torch.Tensor] = None
    end_logits: Optional[torch.Tensor] = None
    hidden_states: TupleOfTensors = None
    attentions: TupleOfTensors = None


@dataclass
class MultipleChoiceModelOutput(ModelOutput):
    loss: Optional[torch.Tensor] = None
    logits: Optional[torch.Tensor] = None
    hidden_states: TupleOfTensors = None
    attentions: TupleOfTensors = None


@dataclass
class MaskedLMOutput(ModelOutput):
    loss: Optional[torch.Tensor] = None
    logits: Optional[torch.Tensor] = None
    hidden_states: TupleOfTensors = None
    attentions: TupleOfTensors = None


@dataclass
class CausalLMOutputWithCrossAttentions(ModelOutput):
    loss: Optional[torch.Tensor] = None
    logits: Optional[torch.Tensor] = None
    past_key_values: Optional[Tuple[Tuple[torch.Tensor, ...], ...]] = None
    hidden_states: TupleOfTensors = None
    attentions: TupleOfTensors = None
    cross_attentions: TupleOfTensors = None
