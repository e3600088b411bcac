"""DeBERTa-v2 sentencepiece tokenizer (reference ppfleetx/data/tokenizers/debertav2_tokenizer.py:55-2163; Imagen's DeBERTa
text tower, models/multimodal_model/imagen/modeling.py:238).

``SPMTokenizer`` wraps ``spm.model`` (whose first pieces are ``[PAD] [CLS] [SEP] [UNK]``) and appends the special tokens the
model file lacks (``[MASK]``); ``DebertaV2Tokenizer`` is the BERT-style template on top: ``[CLS] A [SEP]`` /
``[CLS] A [SEP] B [SEP]`` with token-type ids 0 / 1.
"""
from __future__ import annotations

import os
import shutil
import unicodedata
from typing import Dict, List, Optional, Tuple

from .tokenization_utils_base import PreTrainedTokenizer

MAX_LENGTH = 256


def _is_punct(ch: str) -> bool:
    cp = ord(ch)
    return (33 <= cp <= 47) or (58 <= cp <= 64) or (91 <= cp <= 96) or (123 <= cp <= 126) or unicodedata.category(ch).startswith("P")


def convert_to_unicode(text):
    """``str`` as it is, ``bytes`` decoded as UTF-8 (undecodable bytes dropped)."""
    if isinstance(text, str):
        return text
    if isinstance(text, bytes):
        return text.decode("utf-8", "ignore")
    raise ValueError(f"Unsupported string type: {type(text)}")


class SPMTokenizer:
    def __init__(self, vocab_file: str, special_tokens: List[str], split_by_punct: bool = False):
        import sentencepiece as spm

        if not os.path.isfile(vocab_file):
            raise FileNotFoundError(vocab_file)
        self.vocab_file, self.split_by_punct = vocab_file, split_by_punct
        self.spm = spm.SentencePieceProcessor()
        self.spm.Load(vocab_file)
        n = self.spm.get_piece_size()
        self.vocab: Dict[str, int] = {self.spm.IdToPiece(i): i for i in range(n)}
        self.ids_to_tokens: List[str] = [self.spm.IdToPiece(i) for i in range(n)]
        self.special_tokens: List[str] = []
        for tok in special_tokens:
            self.add_special_token(tok)

    def add_special_token(self, token: str) -> int:
        if token not in self.special_tokens:
            self.special_tokens.append(token)
        if token not in self.vocab:
            self.vocab[token] = len(self.ids_to_tokens)
            self.ids_to_tokens.append(token)
        return self.vocab[token]

    def _split_punct(self, text: str) -> List[str]:
        out, cur = [], ""
        for ch in text:
            if _is_punct(ch):
                if cur:
                    out.append(cur)
                out.append(ch)
                cur = ""
            else:
                cur += ch
        if cur:
            out.append(cur)
        return out

    def tokenize(self, text: str) -> List[str]:
        if self.split_by_punct:
            return [p for word in self._split_punct(text) for p in self.spm.encode(word, out_type=str)]
        return self.spm.encode(text, out_type=str)

    def decode(self, tokens: List[str]) -> str:
        return self.spm.decode_pieces([t for t in tokens if t not in self.special_tokens])

    def id(self, token: str) -> int:
        return self.vocab.get(token, self.vocab.get("[UNK]", 0))

    def token(self, index: int) -> str:
        return self.ids_to_tokens[index] if 0 <= index < len(self.ids_to_tokens) else "[UNK]"


class DebertaV2Tokenizer(PreTrainedTokenizer):
    vocab_files_names = {"vocab_file": "spm.model"}
    model_input_names = ["input_ids", "token_type_ids", "attention_mask"]

    def __init__(self, vocab_file: str, do_lower_case: bool = False, split_by_punct: bool = False, bos_token: str = "[CLS]", eos_token: str = "[SEP]",
                 unk_token: str = "[UNK]", sep_token: str = "[SEP]", pad_token: str = "[PAD]", cls_token: str = "[CLS]", mask_token: str = "[MASK]",
                 model_max_length: Optional[int] = 512, **kwargs):
        super().__init__(bos_token=bos_token, eos_token=eos_token, unk_token=unk_token, sep_token=sep_token, pad_token=pad_token, cls_token=cls_token,
                         mask_token=mask_token, model_max_length=model_max_length, **kwargs)
        self.do_lower_case, self.vocab_file = do_lower_case, vocab_file
        self._tokenizer = SPMTokenizer(vocab_file, self.all_special_tokens, split_by_punct=split_by_punct)
        self._init_kwargs = {"do_lower_case": do_lower_case, "split_by_punct": split_by_punct}

    @property
    def vocab_size(self) -> int:
        return len(self._tokenizer.ids_to_tokens)

    @property
    def vocab(self) -> Dict[str, int]:
        return self._tokenizer.vocab

    def _tokenize(self, text: str) -> List[str]:
        return self._tokenizer.tokenize(text.lower() if self.do_lower_case else text)

    def _convert_token_to_id(self, token: str) -> int:
        return self._tokenizer.id(token)

    def _convert_id_to_token(self, index: int) -> str:
        return self._tokenizer.token(index)

    def convert_tokens_to_string(self, tokens: List[str]) -> str:
        return self._tokenizer.decode(list(tokens))

    # -- [CLS] A [SEP]   |   [CLS] A [SEP] B [SEP]
    def build_inputs_with_special_tokens(self, token_ids_0: List[int], token_ids_1: Optional[List[int]] = None) -> List[int]:
        cls, sep = [self.cls_token_id], [self.sep_token_id]
        out = cls + list(token_ids_0) + sep
        return out if token_ids_1 is None else out + list(token_ids_1) + sep

    def create_token_type_ids_from_sequences(self, token_ids_0: List[int], token_ids_1: Optional[List[int]] = None) -> List[int]:
        first = [0] * (len(token_ids_0) + 2)
        return first if token_ids_1 is None else first + [1] * (len(token_ids_1) + 1)

    def get_special_tokens_mask(self, token_ids_0, token_ids_1=None, already_has_special_tokens: bool = False) -> List[int]:
        if already_has_special_tokens:
            return super().get_special_tokens_mask(token_ids_0, token_ids_1, True)
        mask = [1] + [0] * len(token_ids_0) + [1]
        return mask if token_ids_1 is None else mask + [0] * len(token_ids_1) + [1]

    def num_special_tokens_to_add(self, pair: bool = False) -> int:
        return 3 if pair else 2

    def save_vocabulary(self, save_directory: str, filename_prefix: Optional[str] = None) -> Tuple[str]:
        out = os.path.join(save_directory, ((filename_prefix + "-") if filename_prefix else "") + self.vocab_files_names["vocab_file"])
        if os.path.abspath(out) != os.path.abspath(self.vocab_file):
            shutil.copyfile(self.vocab_file, out)
        return (out,)


def get_debertav2_tokenizer(name: str) -> DebertaV2Tokenizer:
    return DebertaV2Tokenizer.from_pretrained(name)


def debertav2_tokenize(texts: List[str], tokenizer: DebertaV2Tokenizer, max_length: int = MAX_LENGTH):
    enc = tokenizer.batch_encode_plus(list(texts), return_tensors="pt", padding="longest", max_length=max_length, truncation=True)
    return enc.input_ids, enc.attention_mask
