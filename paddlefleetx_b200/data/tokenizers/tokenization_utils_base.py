"""Shared tokenizer machinery: ``BatchEncoding``, padding / truncation strategies, ``SpecialTokensMixin`` and the
``PreTrainedTokenizer`` base every vocabulary-file tokenizer here derives from (reference
ppfleetx/data/tokenizers/tokenization_utils_base.py:622-1852 and t5_tokenization_utils.py:338-984 — HF-derived, ~2.8 kLoC;
this is one compact implementation of the same public surface: ``__call__`` / ``encode_plus`` / ``batch_encode_plus`` /
``prepare_for_model`` / ``truncate_sequences`` / ``pad`` / ``decode`` / ``from_pretrained`` / ``save_pretrained``).

Offline by design: ``from_pretrained`` resolves a *local* directory (argument, ``PFX_TOKENIZER_DIR``, ``~/.cache/ppfleetx/<name>``)
and never downloads.  ``return_tensors`` accepts ``"pt"`` (torch), ``"np"`` and, for recipe compatibility, ``"paddle"`` / ``"pd"`` as
aliases of ``"pt"``.
"""
from __future__ import annotations

import json
import os
from enum import Enum
from typing import Any, Dict, Iterable, List, Optional, Sequence, Tuple, Union

VERY_LARGE_INTEGER = int(1e30)
SPECIAL_TOKENS_MAP_FILE = "special_tokens_map.json"
ADDED_TOKENS_FILE = "added_tokens.json"
TOKENIZER_CONFIG_FILE = "tokenizer_config.json"


def is_sentencepiece_available() -> bool:
    import importlib.util

    return importlib.util.find_spec("sentencepiece") is not None


def is_tokenizers_available() -> bool:
    import importlib.util

    return importlib.util.find_spec("tokenizers") is not None


SPECIAL_MODEL_TYPE_TO_MODULE_NAME = {"openai-gpt": "openai", "data2vec-audio": "data2vec", "data2vec-text": "data2vec", "data2vec-vision": "data2vec"}


def model_type_to_module_name(key: str) -> str:
    """``config.json`` model type -> module name (dashes become underscores, a few irregular names are listed)."""
    return SPECIAL_MODEL_TYPE_TO_MODULE_NAME.get(key, key.replace("-", "_"))


class ExplicitEnum(str, Enum):
    @classmethod
    def _missing_(cls, value):
        raise ValueError(f"{value!r} is not a valid {cls.__name__}; choose one of {[m.value for m in cls]}")


class PaddingStrategy(ExplicitEnum):
    LONGEST = "longest"
    MAX_LENGTH = "max_length"
    DO_NOT_PAD = "do_not_pad"


class TruncationStrategy(ExplicitEnum):
    ONLY_FIRST = "only_first"
    ONLY_SECOND = "only_second"
    LONGEST_FIRST = "longest_first"
    DO_NOT_TRUNCATE = "do_not_truncate"


class TensorType(ExplicitEnum):
    PYTORCH = "pt"
    NUMPY = "np"


def _tensor_type(kind) -> Optional[TensorType]:
    if kind is None:
        return None
    if isinstance(kind, TensorType):
        return kind
    return TensorType({"paddle": "pt", "pd": "pt", "torch": "pt", "numpy": "np"}.get(kind, kind))


class BatchEncoding(dict):
    """Dictionary of model inputs with attribute access and tensor conversion."""

    def __init__(self, data: Optional[Dict[str, Any]] = None, tensor_type=None, prepend_batch_axis: bool = False):
        super().__init__(data or {})
        self.convert_to_tensors(tensor_type, prepend_batch_axis)

    def __getattr__(self, item):
        try:
            return self[item]
        except KeyError as e:
            raise AttributeError(item) from e

    def convert_to_tensors(self, tensor_type=None, prepend_batch_axis: bool = False) -> "BatchEncoding":
        kind = _tensor_type(tensor_type)
        if kind is None:
            return self
        if kind is TensorType.PYTORCH:
            import torch

            make = lambda v: torch.as_tensor(v)  # noqa: E731
            is_t = torch.is_tensor
        else:
            import numpy as np

            make, is_t = np.asarray, lambda v: isinstance(v, np.ndarray)
        for k, v in list(self.items()):
            if is_t(v):
                continue
            if prepend_batch_axis:
                v = [v]
            try:
                self[k] = make(v)
            except (ValueError, TypeError) as e:
                raise ValueError(f"cannot build a rectangular tensor for {k!r}: enable padding=True / truncation=True") from e
        return self

    def to(self, device) -> "BatchEncoding":
        for k, v in list(self.items()):
            if hasattr(v, "to"):
                self[k] = v.to(device)
        return self


class AddedToken(str):
    """A string that remembers it is a special token (kept so ``isinstance(tok, str)`` call sites need no branch)."""

    def __new__(cls, content: str, lstrip=False, rstrip=False, single_word=False, normalized=True):
        obj = super().__new__(cls, content)
        obj.lstrip, obj.rstrip, obj.single_word, obj.normalized = lstrip, rstrip, single_word, normalized
        return obj


class Trie:
    """Longest-match splitter over a small set of literal tokens (special / added tokens never reach the sub-word model)."""

    def __init__(self, words: Iterable[str] = ()):
        self.root: Dict[str, dict] = {}
        for w in words:
            self.add(w)

    def add(self, word: str) -> None:
        if not word:
            return
        node = self.root
        for ch in word:
            node = node.setdefault(ch, {})
        node[""] = True

    def split(self, text: str) -> List[str]:
        if not self.root:
            return [text] if text else []
        out, start, i, n = [], 0, 0, len(text)
        while i < n:
            node, j, end = self.root, i, -1
            while j < n and text[j] in node:
                node = node[text[j]]
                j += 1
                if "" in node:
                    end = j
            if end > 0:
                if i > start:
                    out.append(text[start:i])
                out.append(text[i:end])
                i = start = end
            else:
                i += 1
        if start < n:
            out.append(text[start:])
        return out


class SpecialTokensMixin:
    SPECIAL_TOKENS_ATTRIBUTES = ["bos_token", "eos_token", "unk_token", "sep_token", "pad_token", "cls_token", "mask_token",
                                 "additional_special_tokens"]

    def __init__(self, **kwargs):
        self._bos_token = self._eos_token = self._unk_token = self._sep_token = None
        self._pad_token = self._cls_token = self._mask_token = None
        self._pad_token_type_id = 0
        self._additional_special_tokens: List[str] = []
        for key in self.SPECIAL_TOKENS_ATTRIBUTES:
            value = kwargs.pop(key, None)
            if value is None:
                continue
            if key == "additional_special_tokens":
                assert isinstance(value, (list, tuple)) and all(isinstance(t, str) for t in value), "additional_special_tokens must be a list of str"
                self._additional_special_tokens = list(value)
            else:
                setattr(self, "_" + key, value if isinstance(value, str) else str(value.get("content", value)) if isinstance(value, dict) else str(value))
        self._unused_kwargs = kwargs

    def _tok(name):  # noqa: N805 — property factory
        def getter(self):
            return getattr(self, "_" + name)

        def setter(self, value):
            setattr(self, "_" + name, value)

        return property(getter, setter)

    def _tok_id(name):  # noqa: N805
        def getter(self):
            tok = getattr(self, "_" + name)
            return None if tok is None else self.convert_tokens_to_ids(tok)

        return property(getter)

    bos_token, eos_token, unk_token, sep_token = _tok("bos_token"), _tok("eos_token"), _tok("unk_token"), _tok("sep_token")
    pad_token, cls_token, mask_token = _tok("pad_token"), _tok("cls_token"), _tok("mask_token")
    bos_token_id, eos_token_id, unk_token_id, sep_token_id = _tok_id("bos_token"), _tok_id("eos_token"), _tok_id("unk_token"), _tok_id("sep_token")
    pad_token_id, cls_token_id, mask_token_id = _tok_id("pad_token"), _tok_id("cls_token"), _tok_id("mask_token")
    del _tok, _tok_id

    @property
    def additional_special_tokens(self) -> List[str]:
        return list(self._additional_special_tokens)

    @additional_special_tokens.setter
    def additional_special_tokens(self, value):
        self._additional_special_tokens = list(value)

    @property
    def additional_special_tokens_ids(self) -> List[int]:
        return self.convert_tokens_to_ids(self.additional_special_tokens)

    @property
    def pad_token_type_id(self) -> int:
        return self._pad_token_type_id

    @property
    def special_tokens_map(self) -> Dict[str, Union[str, List[str]]]:
        out: Dict[str, Union[str, List[str]]] = {}
        for key in self.SPECIAL_TOKENS_ATTRIBUTES:
            v = getattr(self, "_" + key)
            if v:
                out[key] = list(v) if isinstance(v, list) else str(v)
        return out

    @property
    def all_special_tokens(self) -> List[str]:
        seen, out = set(), []
        for v in self.special_tokens_map.values():
            for t in (v if isinstance(v, list) else [v]):
                if t not in seen:
                    seen.add(t)
                    out.append(t)
        return out

    @property
    def all_special_ids(self) -> List[int]:
        return self.convert_tokens_to_ids(self.all_special_tokens)

    def add_special_tokens(self, special_tokens: Dict[str, Union[str, List[str]]]) -> int:
        added = 0
        for key, value in special_tokens.items():
            assert key in self.SPECIAL_TOKENS_ATTRIBUTES, f"{key} is not a special-token attribute"
            if key == "additional_special_tokens":
                new = [t for t in value if t not in self._additional_special_tokens]
                self._additional_special_tokens.extend(new)
                added += self.add_tokens(new, special_tokens=True)
            else:
                setattr(self, "_" + key, value)
                added += self.add_tokens([value], special_tokens=True)
        return added

    def add_tokens(self, new_tokens, special_tokens: bool = False) -> int:
        raise NotImplementedError


def _resolve_dir(name_or_dir: str, required: Sequence[str]) -> str:
    cands = [name_or_dir, os.path.join(os.environ.get("PFX_TOKENIZER_DIR", ""), name_or_dir) if os.environ.get("PFX_TOKENIZER_DIR") else "",
             os.environ.get("PFX_TOKENIZER_DIR", ""), os.path.expanduser(os.path.join("~/.cache/ppfleetx", name_or_dir))]
    for d in cands:
        if d and os.path.isdir(d) and all(os.path.isfile(os.path.join(d, f)) for f in required):
            return d
    if os.path.isfile(name_or_dir) and len(required) == 1:
        return os.path.dirname(os.path.abspath(name_or_dir))
    raise FileNotFoundError(
        f"tokenizer files {list(required)} for {name_or_dir!r} not found locally (looked in {[c for c in cands if c]}). This machine is "
        "offline: put the vocabulary files in a directory and pass it, or set PFX_TOKENIZER_DIR.")


class PreTrainedTokenizer(SpecialTokensMixin):
    """Base class: subclasses provide the sub-word model (``_tokenize`` / ``_convert_token_to_id`` / ``_convert_id_to_token`` /
    ``convert_tokens_to_string`` / ``vocab_size``) and the sequence template (``build_inputs_with_special_tokens`` …)."""

    vocab_files_names: Dict[str, str] = {}
    model_input_names: List[str] = ["input_ids", "token_type_ids", "attention_mask"]
    padding_side: str = "right"
    truncation_side: str = "right"

    def __init__(self, model_max_length: Optional[int] = None, padding_side: Optional[str] = None, truncation_side: Optional[str] = None, **kwargs):
        super().__init__(**kwargs)
        self.model_max_length = int(model_max_length) if model_max_length else VERY_LARGE_INTEGER
        if padding_side:
            self.padding_side = padding_side
        if truncation_side:
            self.truncation_side = truncation_side
        assert self.padding_side in ("left", "right") and self.truncation_side in ("left", "right")
        self.added_tokens_encoder: Dict[str, int] = {}
        self.added_tokens_decoder: Dict[int, str] = {}
        self._no_split = Trie()
        self._init_kwargs: Dict[str, Any] = {}

    # ------------------------------------------------------------------ abstract sub-word model
    @property
    def vocab_size(self) -> int:
        raise NotImplementedError

    def _tokenize(self, text: str) -> List[str]:
        raise NotImplementedError

    def _convert_token_to_id(self, token: str) -> int:
        raise NotImplementedError

    def _convert_id_to_token(self, index: int) -> str:
        raise NotImplementedError

    def convert_tokens_to_string(self, tokens: List[str]) -> str:
        return " ".join(tokens)

    def get_vocab(self) -> Dict[str, int]:
        vocab = {self._convert_id_to_token(i): i for i in range(self.vocab_size)}
        vocab.update(self.added_tokens_encoder)
        return vocab

    def __len__(self) -> int:
        return self.vocab_size + len(self.added_tokens_encoder)

    # ------------------------------------------------------------------ sequence template (single-sequence default)
    def build_inputs_with_special_tokens(self, token_ids_0: List[int], token_ids_1: Optional[List[int]] = None) -> List[int]:
        return list(token_ids_0) + (list(token_ids_1) if token_ids_1 is not None else [])

    def create_token_type_ids_from_sequences(self, token_ids_0: List[int], token_ids_1: Optional[List[int]] = None) -> List[int]:
        return [0] * len(self.build_inputs_with_special_tokens(token_ids_0, token_ids_1))

    def get_special_tokens_mask(self, token_ids_0: List[int], token_ids_1: Optional[List[int]] = None, already_has_special_tokens: bool = False) -> List[int]:
        if already_has_special_tokens:
            special = set(self.all_special_ids)
            return [1 if t in special else 0 for t in token_ids_0]
        built = self.build_inputs_with_special_tokens(token_ids_0, token_ids_1)
        # positions that are not taken from either input, found by a two-pointer walk
        mask, seq, k = [], list(token_ids_0) + (list(token_ids_1) if token_ids_1 else []), 0
        for t in built:
            if k < len(seq) and t == seq[k]:
                mask.append(0)
                k += 1
            else:
                mask.append(1)
        return mask

    def num_special_tokens_to_add(self, pair: bool = False) -> int:
        return len(self.build_inputs_with_special_tokens([], [] if pair else None))

    # ------------------------------------------------------------------ added tokens
    def add_tokens(self, new_tokens, special_tokens: bool = False) -> int:
        if isinstance(new_tokens, str):
            new_tokens = [new_tokens]
        added = 0
        for tok in new_tokens:
            tok = str(tok)
            known = tok in self.added_tokens_encoder
            if not known:
                try:
                    idx = self._convert_token_to_id(tok)
                    known = idx is not None and (self.unk_token is None or idx != self._convert_token_to_id(self.unk_token) or tok == self.unk_token)
                except (KeyError, NotImplementedError):
                    known = False
            if not known:
                idx = len(self)
                self.added_tokens_encoder[tok] = idx
                self.added_tokens_decoder[idx] = tok
                added += 1
            self._no_split.add(tok)
        return added

    def _refresh_no_split(self) -> None:
        self._no_split = Trie(list(self.all_special_tokens) + list(self.added_tokens_encoder))

    # ------------------------------------------------------------------ text -> tokens -> ids
    def prepare_for_tokenization(self, text: str, **kwargs) -> Tuple[str, Dict[str, Any]]:
        return text, kwargs

    def tokenize(self, text: str, **kwargs) -> List[str]:
        text, _ = self.prepare_for_tokenization(text, **kwargs)
        self._refresh_no_split()
        literal = set(self.all_special_tokens) | set(self.added_tokens_encoder)
        out: List[str] = []
        for piece in self._no_split.split(text):
            if piece in literal:
                out.append(piece)
            elif piece.strip() or piece:
                out.extend(self._tokenize(piece))
        return out

    def convert_tokens_to_ids(self, tokens):
        if tokens is None:
            return None
        if isinstance(tokens, str):
            return self.added_tokens_encoder[tokens] if tokens in self.added_tokens_encoder else self._convert_token_to_id(tokens)
        return [self.convert_tokens_to_ids(t) for t in tokens]

    def convert_ids_to_tokens(self, ids, skip_special_tokens: bool = False):
        if isinstance(ids, int):
            return self.added_tokens_decoder[ids] if ids in self.added_tokens_decoder else self._convert_id_to_token(ids)
        special = set(self.all_special_ids) if skip_special_tokens else ()
        return [self.convert_ids_to_tokens(int(i)) for i in ids if int(i) not in special]

    def encode(self, text, text_pair=None, add_special_tokens: bool = True, **kwargs) -> List[int]:
        return self.encode_plus(text, text_pair, add_special_tokens=add_special_tokens, return_attention_mask=False, return_token_type_ids=False,
                                **kwargs)["input_ids"]

    def decode(self, token_ids, skip_special_tokens: bool = False, clean_up_tokenization_spaces: bool = True) -> str:
        if hasattr(token_ids, "tolist"):
            token_ids = token_ids.tolist()
        tokens = self.convert_ids_to_tokens(list(token_ids), skip_special_tokens=skip_special_tokens)
        # sub-word runs go through the model's detokeniser; literal (added / special) tokens are spliced in verbatim
        literal = set(self.added_tokens_encoder) | set(self.all_special_tokens)
        parts, run = [], []
        for t in tokens:
            if t in literal:
                if run:
                    parts.append(self.convert_tokens_to_string(run))
                    run = []
                parts.append(t)
            else:
                run.append(t)
        if run:
            parts.append(self.convert_tokens_to_string(run))
        text = " ".join(p for p in parts if p != "") if len(parts) > 1 else (parts[0] if parts else "")
        return self.clean_up_tokenization(text) if clean_up_tokenization_spaces else text

    def batch_decode(self, sequences, **kwargs) -> List[str]:
        return [self.decode(s, **kwargs) for s in sequences]

    @staticmethod
    def clean_up_tokenization(text: str) -> str:
        for a, b in ((" .", "."), (" ?", "?"), (" !", "!"), (" ,", ","), (" ' ", "'"), (" n't", "n't"), (" 'm", "'m"), (" 's", "'s"), (" 've", "'ve"),
                     (" 're", "'re")):
            text = text.replace(a, b)
        return text

    # ------------------------------------------------------------------ strategies
    def _get_padding_truncation_strategies(self, padding=False, truncation=False, max_length=None) -> Tuple[PaddingStrategy, TruncationStrategy, Optional[int]]:
        if padding is True:
            pad = PaddingStrategy.LONGEST
        elif padding in (False, None):
            pad = PaddingStrategy.DO_NOT_PAD
        else:
            pad = PaddingStrategy(padding)
        if truncation is True:
            trunc = TruncationStrategy.LONGEST_FIRST
        elif truncation in (False, None):
            trunc = TruncationStrategy.DO_NOT_TRUNCATE
        else:
            trunc = TruncationStrategy(truncation)
        if max_length is None and (pad is PaddingStrategy.MAX_LENGTH or trunc is not TruncationStrategy.DO_NOT_TRUNCATE):
            if self.model_max_length >= VERY_LARGE_INTEGER:
                if pad is PaddingStrategy.MAX_LENGTH:
                    pad = PaddingStrategy.DO_NOT_PAD
                trunc = TruncationStrategy.DO_NOT_TRUNCATE
            else:
                max_length = self.model_max_length
        if pad is not PaddingStrategy.DO_NOT_PAD and self.pad_token is None:
            raise ValueError("padding requested but the tokenizer has no pad_token (tokenizer.pad_token = tokenizer.eos_token is a common choice)")
        return pad, trunc, max_length

    def truncate_sequences(self, ids: List[int], pair_ids: Optional[List[int]] = None, num_tokens_to_remove: int = 0,
                           truncation_strategy="longest_first", stride: int = 0) -> Tuple[List[int], Optional[List[int]], List[int]]:
        if num_tokens_to_remove <= 0:
            return ids, pair_ids, []
        strat = TruncationStrategy(truncation_strategy) if not isinstance(truncation_strategy, TruncationStrategy) else truncation_strategy
        left = self.truncation_side == "left"

        def cut(seq: List[int], n: int) -> Tuple[List[int], List[int]]:
            n = min(n, len(seq))
            window = min(len(seq), stride + n)
            return (seq[n:], seq[:window]) if left else (seq[: len(seq) - n], seq[len(seq) - window:])

        overflow: List[int] = []
        if strat is TruncationStrategy.ONLY_FIRST or (strat is TruncationStrategy.LONGEST_FIRST and pair_ids is None):
            if len(ids) <= num_tokens_to_remove and strat is TruncationStrategy.ONLY_FIRST:
                raise ValueError(f"cannot remove {num_tokens_to_remove} tokens from a first sequence of length {len(ids)}")
            ids, overflow = cut(ids, num_tokens_to_remove)
        elif strat is TruncationStrategy.ONLY_SECOND:
            if pair_ids is None or len(pair_ids) <= num_tokens_to_remove:
                raise ValueError("only_second truncation needs a second sequence longer than the number of tokens to remove")
            pair_ids, overflow = cut(pair_ids, num_tokens_to_remove)
        elif strat is TruncationStrategy.LONGEST_FIRST:
            for _ in range(num_tokens_to_remove):
                if len(ids) > len(pair_ids):
                    ids = ids[1:] if left else ids[:-1]
                else:
                    pair_ids = pair_ids[1:] if left else pair_ids[:-1]
        return ids, pair_ids, overflow

    # ------------------------------------------------------------------ ids -> model inputs
    def prepare_for_model(self, ids: List[int], pair_ids: Optional[List[int]] = None, add_special_tokens: bool = True, padding=False, truncation=False,
                          max_length: Optional[int] = None, stride: int = 0, pad_to_multiple_of: Optional[int] = None, return_tensors=None,
                          return_token_type_ids: Optional[bool] = None, return_attention_mask: Optional[bool] = None,
                          return_overflowing_tokens: bool = False, return_special_tokens_mask: bool = False, return_length: bool = False,
                          prepend_batch_axis: bool = False, **unused) -> BatchEncoding:
        pad, trunc, max_length = self._get_padding_truncation_strategies(padding, truncation, max_length)
        pair = pair_ids is not None
        if return_token_type_ids is None:
            return_token_type_ids = "token_type_ids" in self.model_input_names
        if return_attention_mask is None:
            return_attention_mask = "attention_mask" in self.model_input_names
        total = len(ids) + (len(pair_ids) if pair else 0) + (self.num_special_tokens_to_add(pair) if add_special_tokens else 0)
        enc: Dict[str, Any] = {}
        if trunc is not TruncationStrategy.DO_NOT_TRUNCATE and max_length and total > max_length:
            ids, pair_ids, overflow = self.truncate_sequences(ids, pair_ids, total - max_length, trunc, stride)
            if return_overflowing_tokens:
                enc["overflowing_tokens"], enc["num_truncated_tokens"] = overflow, total - max_length
        if add_special_tokens:
            seq = self.build_inputs_with_special_tokens(ids, pair_ids)
            types = self.create_token_type_ids_from_sequences(ids, pair_ids)
        else:
            seq = list(ids) + (list(pair_ids) if pair else [])
            types = [0] * len(ids) + ([1] * len(pair_ids) if pair else [])
        enc["input_ids"] = seq
        if return_token_type_ids:
            enc["token_type_ids"] = types
        if return_special_tokens_mask:
            enc["special_tokens_mask"] = self.get_special_tokens_mask(ids, pair_ids) if add_special_tokens else [0] * len(seq)
        if pad is not PaddingStrategy.DO_NOT_PAD or return_attention_mask:
            enc = self._pad(enc, max_length=max_length, padding_strategy=pad, pad_to_multiple_of=pad_to_multiple_of, return_attention_mask=return_attention_mask)
        if return_length:
            enc["length"] = len(enc["input_ids"])
        return BatchEncoding(enc, tensor_type=return_tensors, prepend_batch_axis=prepend_batch_axis)

    def _pad(self, enc: Dict[str, Any], max_length: Optional[int] = None, padding_strategy=PaddingStrategy.DO_NOT_PAD,
             pad_to_multiple_of: Optional[int] = None, return_attention_mask: Optional[bool] = None) -> Dict[str, Any]:
        if return_attention_mask is None:
            return_attention_mask = "attention_mask" in self.model_input_names
        ids = enc[self.model_input_names[0]]
        if padding_strategy is PaddingStrategy.LONGEST:
            max_length = len(ids)
        if max_length is not None and pad_to_multiple_of and max_length % pad_to_multiple_of:
            max_length = (max_length // pad_to_multiple_of + 1) * pad_to_multiple_of
        need = padding_strategy is not PaddingStrategy.DO_NOT_PAD and max_length is not None and len(ids) < max_length
        if return_attention_mask and "attention_mask" not in enc:
            enc["attention_mask"] = [1] * len(ids)
        if not need:
            return enc
        diff = max_length - len(ids)
        fill = {"attention_mask": 0, "token_type_ids": self.pad_token_type_id, "special_tokens_mask": 1, self.model_input_names[0]: self.pad_token_id}
        for key, value in fill.items():
            if key in enc:
                enc[key] = (list(enc[key]) + [value] * diff) if self.padding_side == "right" else ([value] * diff + list(enc[key]))
        return enc

    def pad(self, encoded_inputs, padding=True, max_length: Optional[int] = None, pad_to_multiple_of: Optional[int] = None,
            return_attention_mask: Optional[bool] = None, return_tensors=None) -> BatchEncoding:
        """Pad one encoding, a dict of batched lists, or a list of encodings (collate-function use)."""
        if isinstance(encoded_inputs, (list, tuple)) and encoded_inputs and isinstance(encoded_inputs[0], dict):
            encoded_inputs = {k: [e[k] for e in encoded_inputs] for k in encoded_inputs[0]}
        main = self.model_input_names[0]
        assert main in encoded_inputs, f"pad() needs {main!r}"
        first = encoded_inputs[main]
        if hasattr(first, "tolist"):
            encoded_inputs = {k: (v.tolist() if hasattr(v, "tolist") else v) for k, v in encoded_inputs.items()}
            first = encoded_inputs[main]
        pad, _, max_length = self._get_padding_truncation_strategies(padding=padding, max_length=max_length)
        if first and not isinstance(first[0], (list, tuple)):
            return BatchEncoding(self._pad(dict(encoded_inputs), max_length, pad, pad_to_multiple_of, return_attention_mask), tensor_type=return_tensors)
        if pad is PaddingStrategy.LONGEST:
            max_length, pad = max(len(x) for x in first), PaddingStrategy.MAX_LENGTH
        out: Dict[str, List[Any]] = {}
        for i in range(len(first)):
            one = self._pad({k: v[i] for k, v in encoded_inputs.items()}, max_length, pad, pad_to_multiple_of, return_attention_mask)
            for k, v in one.items():
                out.setdefault(k, []).append(v)
        return BatchEncoding(out, tensor_type=return_tensors)

    # ------------------------------------------------------------------ user-facing encode calls
    def _ids_of(self, text) -> List[int]:
        if isinstance(text, str):
            return self.convert_tokens_to_ids(self.tokenize(text))
        if isinstance(text, (list, tuple)) and text and isinstance(text[0], str):
            return self.convert_tokens_to_ids(list(text))  # pre-tokenised
        if isinstance(text, (list, tuple)) and all(isinstance(t, int) for t in text):
            return list(text)
        raise ValueError("input must be a string, a list of tokens or a list of ids")

    def encode_plus(self, text, text_pair=None, add_special_tokens: bool = True, padding=False, truncation=False, max_length: Optional[int] = None,
                    return_tensors=None, **kwargs) -> BatchEncoding:
        ids = self._ids_of(text)
        pair = self._ids_of(text_pair) if text_pair is not None else None
        return self.prepare_for_model(ids, pair, add_special_tokens=add_special_tokens, padding=padding, truncation=truncation, max_length=max_length,
                                      return_tensors=return_tensors, prepend_batch_axis=return_tensors is not None, **kwargs)

    def batch_encode_plus(self, batch_text_or_text_pairs, add_special_tokens: bool = True, padding=False, truncation=False,
                          max_length: Optional[int] = None, pad_to_multiple_of: Optional[int] = None, return_tensors=None,
                          return_attention_mask: Optional[bool] = None, **kwargs) -> BatchEncoding:
        pad, trunc, max_length = self._get_padding_truncation_strategies(padding, truncation, max_length)
        rows: List[Dict[str, Any]] = []
        for item in batch_text_or_text_pairs:
            a, b = item if isinstance(item, tuple) and len(item) == 2 else (item, None)  # a tuple is a (text, text_pair) item
            rows.append(dict(self.prepare_for_model(self._ids_of(a), self._ids_of(b) if b is not None else None, add_special_tokens=add_special_tokens,
                                                    padding=False, truncation=trunc.value, max_length=max_length, return_attention_mask=False,
                                                    return_tensors=None, **kwargs)))
        batch = {k: [r[k] for r in rows] for k in rows[0]} if rows else {self.model_input_names[0]: []}
        if not rows:
            return BatchEncoding(batch)
        padded = self.pad(batch, padding=pad.value if pad is not PaddingStrategy.DO_NOT_PAD else False, max_length=max_length,
                          pad_to_multiple_of=pad_to_multiple_of, return_attention_mask=return_attention_mask)
        return BatchEncoding(dict(padded), tensor_type=return_tensors)

    def __call__(self, text, text_pair=None, **kwargs) -> BatchEncoding:
        batched = isinstance(text, (list, tuple)) and (not text or isinstance(text[0], (str, list, tuple)))
        if batched and text and isinstance(text[0], str) and kwargs.pop("is_split_into_words", False):
            batched = False
        if batched:
            items = list(zip(text, text_pair)) if text_pair is not None else list(text)
            return self.batch_encode_plus(items, **kwargs)
        return self.encode_plus(text, text_pair, **kwargs)

    # ------------------------------------------------------------------ persistence
    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path: str, *init_inputs, **kwargs):
        required = list(cls.vocab_files_names.values())
        d = _resolve_dir(pretrained_model_name_or_path, required)
        init_kwargs: Dict[str, Any] = {}
        cfg_file = os.path.join(d, TOKENIZER_CONFIG_FILE)
        if os.path.isfile(cfg_file):
            with open(cfg_file, encoding="utf-8") as f:
                init_kwargs.update({k: v for k, v in json.load(f).items() if k not in ("tokenizer_class", "name_or_path")})
        stm = os.path.join(d, SPECIAL_TOKENS_MAP_FILE)
        if os.path.isfile(stm):
            with open(stm, encoding="utf-8") as f:
                for k, v in json.load(f).items():
                    init_kwargs.setdefault(k, v["content"] if isinstance(v, dict) else v)
        init_kwargs.update(kwargs)
        for key, fname in cls.vocab_files_names.items():
            init_kwargs[key] = os.path.join(d, fname)
        tok = cls(*init_inputs, **init_kwargs)
        added = os.path.join(d, ADDED_TOKENS_FILE)
        if os.path.isfile(added):
            with open(added, encoding="utf-8") as f:
                for t, i in sorted(json.load(f).items(), key=lambda kv: kv[1]):
                    tok.added_tokens_encoder[t], tok.added_tokens_decoder[int(i)] = int(i), t
        tok.name_or_path = pretrained_model_name_or_path
        return tok

    def save_vocabulary(self, save_directory: str, filename_prefix: Optional[str] = None) -> Tuple[str, ...]:
        raise NotImplementedError

    def save_pretrained(self, save_directory: str, filename_prefix: Optional[str] = None) -> Tuple[str, ...]:
        os.makedirs(save_directory, exist_ok=True)
        pre = (filename_prefix + "-") if filename_prefix else ""
        cfg = dict(self._init_kwargs)
        if self.model_max_length < VERY_LARGE_INTEGER:
            cfg["model_max_length"] = self.model_max_length
        cfg["tokenizer_class"] = type(self).__name__
        files = []
        for name, payload in ((TOKENIZER_CONFIG_FILE, cfg), (SPECIAL_TOKENS_MAP_FILE, self.special_tokens_map)):
            path = os.path.join(save_directory, pre + name)
            with open(path, "w", encoding="utf-8") as f:
                json.dump(payload, f, ensure_ascii=False, indent=1)
            files.append(path)
        if self.added_tokens_encoder:
            path = os.path.join(save_directory, pre + ADDED_TOKENS_FILE)
            with open(path, "w", encoding="utf-8") as f:
                json.dump(self.added_tokens_encoder, f, ensure_ascii=False)
            files.append(path)
        return tuple(files) + tuple(self.save_vocabulary(save_directory, filename_prefix))


PreTrainedTokenizerBase = PreTrainedTokenizer
