"""T5 sentencepiece tokenizer (reference ppfleetx/data/tokenizers/t5_tokenizer.py:86-1905; used by Imagen's T5 text tower,
models/multimodal_model/imagen/modeling.py:230).

Vocabulary = the sentencepiece unigram model (``spiece.model``) followed by ``extra_ids`` sentinel tokens ``<extra_id_0>`` …
which occupy the *top* ids in reverse order (``<extra_id_0>`` is the last id).  Sequences end with ``</s>``; pairs are
``A </s> B </s>``; there are no token-type ids.  ``from_pretrained`` takes a local directory holding ``spiece.model``.
"""
from __future__ import annotations

import os
import re
import shutil
from typing import List, Optional, Tuple

from .tokenization_utils_base import PreTrainedTokenizer

DEFAULT_T5_NAME = "t5-small"
MAX_LENGTH = 256
SPIECE_UNDERLINE = "▁"


class T5Tokenizer(PreTrainedTokenizer):
    vocab_files_names = {"vocab_file": "spiece.model"}
    model_input_names = ["input_ids", "attention_mask"]
    padding_side = "right"
    truncation_side = "right"

    def __init__(self, vocab_file: str, eos_token: str = "</s>", unk_token: str = "<unk>", pad_token: str = "<pad>", extra_ids: int = 100,
                 additional_special_tokens: Optional[List[str]] = None, model_max_length: Optional[int] = 512, **kwargs):
        import sentencepiece as spm

        sentinels = [f"<extra_id_{i}>" for i in range(extra_ids)]
        if additional_special_tokens is None:
            additional_special_tokens = sentinels
        elif extra_ids > 0:
            given = [t for t in additional_special_tokens if re.fullmatch(r"<extra_id_\d+>", t)]
            if not given:
                additional_special_tokens = list(additional_special_tokens) + sentinels
            elif len(set(given)) != extra_ids:
                raise ValueError(f"extra_ids={extra_ids} but additional_special_tokens carries {len(set(given))} sentinel tokens")
        super().__init__(eos_token=eos_token, unk_token=unk_token, pad_token=pad_token, additional_special_tokens=additional_special_tokens,
                         model_max_length=model_max_length, **kwargs)
        self.vocab_file, self._extra_ids = vocab_file, extra_ids
        self.sp_model = spm.SentencePieceProcessor()
        self.sp_model.Load(vocab_file)
        self._init_kwargs = {"extra_ids": extra_ids}

    @property
    def vocab_size(self) -> int:
        return self.sp_model.get_piece_size() + self._extra_ids

    def _tokenize(self, text: str) -> List[str]:
        return self.sp_model.encode(text, out_type=str)

    def _convert_token_to_id(self, token: str) -> int:
        m = re.fullmatch(r"<extra_id_(\d+)>", token)
        if m and int(m.group(1)) < self._extra_ids:
            return self.vocab_size - 1 - int(m.group(1))
        return self.sp_model.piece_to_id(token)

    def _convert_id_to_token(self, index: int) -> str:
        if index < self.sp_model.get_piece_size():
            return self.sp_model.IdToPiece(index)
        return f"<extra_id_{self.vocab_size - 1 - index}>"

    def convert_tokens_to_string(self, tokens: List[str]) -> str:
        return self.sp_model.decode_pieces(list(tokens))

    # -- sequence template:  X </s>   |   A </s> B </s>
    def _with_eos(self, ids: List[int]) -> List[int]:
        ids = list(ids)
        return ids if ids and ids[-1] == self.eos_token_id else ids + [self.eos_token_id]

    def build_inputs_with_special_tokens(self, token_ids_0: List[int], token_ids_1: Optional[List[int]] = None) -> List[int]:
        out = self._with_eos(token_ids_0)
        return out if token_ids_1 is None else out + self._with_eos(token_ids_1)

    def get_special_tokens_mask(self, token_ids_0, token_ids_1=None, already_has_special_tokens: bool = False) -> List[int]:
        if already_has_special_tokens:
            return super().get_special_tokens_mask(token_ids_0, token_ids_1, True)
        mask = [0] * len(token_ids_0) + [1]
        return mask if token_ids_1 is None else mask + [0] * len(token_ids_1) + [1]

    def num_special_tokens_to_add(self, pair: bool = False) -> int:
        return 2 if pair else 1

    def save_vocabulary(self, save_directory: str, filename_prefix: Optional[str] = None) -> Tuple[str]:
        out = os.path.join(save_directory, ((filename_prefix + "-") if filename_prefix else "") + self.vocab_files_names["vocab_file"])
        if os.path.abspath(out) != os.path.abspath(self.vocab_file):
            if os.path.isfile(self.vocab_file):
                shutil.copyfile(self.vocab_file, out)
            else:
                with open(out, "wb") as f:
                    f.write(self.sp_model.serialized_model_proto())
        return (out,)

    def __getstate__(self):
        state = self.__dict__.copy()
        state["sp_model"] = None
        return state

    def __setstate__(self, state):
        import sentencepiece as spm

        self.__dict__ = state
        self.sp_model = spm.SentencePieceProcessor()
        self.sp_model.Load(self.vocab_file)


def get_t5_tokenizer(name: str = DEFAULT_T5_NAME) -> T5Tokenizer:
    return T5Tokenizer.from_pretrained(name)


def t5_tokenize(texts: List[str], tokenizer: T5Tokenizer, max_length: int = MAX_LENGTH):
    """-> (input_ids [b, L], attention_mask [b, L]) padded to the longest text, truncated to ``max_length``."""
    enc = tokenizer.batch_encode_plus(list(texts), return_tensors="pt", padding="longest", max_length=max_length, truncation=True)
    return enc.input_ids, enc.attention_mask
