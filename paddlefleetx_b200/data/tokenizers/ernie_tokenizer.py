"""ERNIE WordPiece tokenizer (reference ppfleetx/data/tokenizers/ernie_tokenizer.py:14-25 just re-exports paddlenlp's
``ErnieTokenizer``; paddlenlp is not a dependency here, so the tokenizer itself lives in this file).

Pipeline: clean-up → CJK characters isolated as single tokens → whitespace split → optional lower-casing + accent stripping →
punctuation split → greedy longest-match WordPiece (continuations prefixed ``##``).  Template ``[CLS] A [SEP]`` /
``[CLS] A [SEP] B [SEP]`` with token-type ids 0 / 1, the layout ``ErnieDataset`` and the GLUE-style fine-tuning sets produce.
``from_pretrained`` takes a local directory with ``vocab.txt``.
"""
from __future__ import annotations

import os
import unicodedata
from typing import Dict, List, Optional, Tuple

from .tokenization_utils_base import PreTrainedTokenizer


def _is_whitespace(ch: str) -> bool:
    return ch in " \t\n\r" or unicodedata.category(ch) == "Zs"


def _is_control(ch: str) -> bool:
    return ch not in "\t\n\r" and unicodedata.category(ch).startswith("C")


def _is_punctuation(ch: str) -> bool:
    cp = ord(ch)
    return (33 <= cp <= 47) or (58 <= cp <= 64) or (91 <= cp <= 96) or (123 <= cp <= 126) or unicodedata.category(ch).startswith("P")


_CJK_RANGES = ((0x4E00, 0x9FFF), (0x3400, 0x4DBF), (0x20000, 0x2A6DF), (0x2A700, 0x2B73F), (0x2B740, 0x2B81F), (0x2B820, 0x2CEAF),
               (0xF900, 0xFAFF), (0x2F800, 0x2FA1F))


def _is_cjk(ch: str) -> bool:
    cp = ord(ch)
    return any(lo <= cp <= hi for lo, hi in _CJK_RANGES)


class BasicTokenizer:
    def __init__(self, do_lower_case: bool = True, never_split: Optional[List[str]] = None, tokenize_chinese_chars: bool = True,
                 strip_accents: Optional[bool] = None):
        self.do_lower_case, self.never_split = do_lower_case, set(never_split or ())
        self.tokenize_chinese_chars, self.strip_accents = tokenize_chinese_chars, strip_accents

    def tokenize(self, text: str, never_split: Optional[List[str]] = None) -> List[str]:
        keep = self.never_split | set(never_split or ())
        chars = []
        for ch in text:
            if ord(ch) in (0, 0xFFFD) or _is_control(ch):
                continue
            if _is_whitespace(ch):
                chars.append(" ")
            elif self.tokenize_chinese_chars and _is_cjk(ch):
                chars.extend((" ", ch, " "))
            else:
                chars.append(ch)
        out: List[str] = []
        for word in unicodedata.normalize("NFC", "".join(chars)).split():
            if word in keep:
                out.append(word)
                continue
            if self.do_lower_case:
                word = word.lower()
                if self.strip_accents is not False:
                    word = self._strip_accents(word)
            elif self.strip_accents:
                word = self._strip_accents(word)
            out.extend(self._split_on_punct(word))
        return out

    @staticmethod
    def _strip_accents(word: str) -> str:
        return "".join(ch for ch in unicodedata.normalize("NFD", word) if unicodedata.category(ch) != "Mn")

    @staticmethod
    def _split_on_punct(word: str) -> List[str]:
        out, cur = [], ""
        for ch in word:
            if _is_punctuation(ch):
                if cur:
                    out.append(cur)
                out.append(ch)
                cur = ""
            else:
                cur += ch
        if cur:
            out.append(cur)
        return out


class WordpieceTokenizer:
    def __init__(self, vocab: Dict[str, int], unk_token: str, max_input_chars_per_word: int = 100):
        self.vocab, self.unk_token, self.max_chars = vocab, unk_token, max_input_chars_per_word

    def tokenize(self, text: str) -> List[str]:
        out: List[str] = []
        for word in text.split():
            if len(word) > self.max_chars:
                out.append(self.unk_token)
                continue
            pieces, start, bad = [], 0, False
            while start < len(word):
                end, cur = len(word), None
                while start < end:
                    cand = ("##" if start else "") + word[start:end]
                    if cand in self.vocab:
                        cur = cand
                        break
                    end -= 1
                if cur is None:
                    bad = True
                    break
                pieces.append(cur)
                start = end
            out.extend([self.unk_token] if bad else pieces)
        return out


def load_vocab(vocab_file: str) -> Dict[str, int]:
    vocab: Dict[str, int] = {}
    with open(vocab_file, encoding="utf-8") as f:
        for i, line in enumerate(f):
            vocab.setdefault(line.rstrip("\n").split("\t")[0], i)
    return vocab


class ErnieTokenizer(PreTrainedTokenizer):
    vocab_files_names = {"vocab_file": "vocab.txt"}
    model_input_names = ["input_ids", "token_type_ids", "attention_mask"]

    def __init__(self, vocab_file: str, do_lower_case: bool = True, unk_token: str = "[UNK]", sep_token: str = "[SEP]", pad_token: str = "[PAD]",
                 cls_token: str = "[CLS]", mask_token: str = "[MASK]", tokenize_chinese_chars: bool = True, strip_accents: Optional[bool] = None,
                 model_max_length: Optional[int] = 512, **kwargs):
        if not os.path.isfile(vocab_file):
            raise FileNotFoundError(f"vocabulary file {vocab_file!r} not found")
        super().__init__(unk_token=unk_token, sep_token=sep_token, pad_token=pad_token, cls_token=cls_token, mask_token=mask_token,
                         model_max_length=model_max_length, **kwargs)
        self.vocab_file, self.do_lower_case = vocab_file, do_lower_case
        self.vocab = load_vocab(vocab_file)
        self.ids_to_tokens = {i: t for t, i in self.vocab.items()}
        self.basic_tokenizer = BasicTokenizer(do_lower_case, None, tokenize_chinese_chars, strip_accents)
        self.wordpiece_tokenizer = WordpieceTokenizer(self.vocab, unk_token)
        self._init_kwargs = {"do_lower_case": do_lower_case}

    @property
    def vocab_size(self) -> int:
        return len(self.vocab)

    def get_vocab(self) -> Dict[str, int]:
        return dict(self.vocab, **self.added_tokens_encoder)

    def _tokenize(self, text: str) -> List[str]:
        keep, out = set(self.all_special_tokens), []
        for word in self.basic_tokenizer.tokenize(text, never_split=keep):
            out.extend([word] if word in keep else self.wordpiece_tokenizer.tokenize(word))
        return out

    def _convert_token_to_id(self, token: str) -> int:
        return self.vocab.get(token, self.vocab.get(self.unk_token, 0))

    def _convert_id_to_token(self, index: int) -> str:
        return self.ids_to_tokens.get(index, self.unk_token)

    def convert_tokens_to_string(self, tokens: List[str]) -> str:
        return " ".join(tokens).replace(" ##", "").strip()

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
        with open(out, "w", encoding="utf-8") as f:
            for tok, _ in sorted(self.vocab.items(), key=lambda kv: kv[1]):
                f.write(tok + "\n")
        return (out,)

    # data tools call ``encode(text)`` for raw ids without the [CLS]/[SEP] frame when building a pre-training corpus
    def encode_plain(self, text: str) -> List[int]:
        return self.convert_tokens_to_ids(self.tokenize(text))


_TOKENIZERS: Dict[str, ErnieTokenizer] = {}


def get_ernie_tokenizer(tokenizer_type: str) -> ErnieTokenizer:
    """Process-wide cached tokenizer, keyed by name / directory (the reference caches a single global instance)."""
    if tokenizer_type not in _TOKENIZERS:
        _TOKENIZERS[tokenizer_type] = ErnieTokenizer.from_pretrained(tokenizer_type)
    return _TOKENIZERS[tokenizer_type]
