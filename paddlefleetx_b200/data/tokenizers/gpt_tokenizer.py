"""Byte-level BPE tokenizer for GPT (GPT-2 vocabulary format: ``vocab.json`` + ``merges.txt``).

The reference tokenizer (ppfleetx/data/tokenizers/gpt_tokenizer.py:97-819) downloads ``gpt2-vocab.json`` /
``gpt2-merges.txt`` on first use; this box is offline, so ``from_pretrained`` takes a *local directory* (or the
``PFX_GPT_VOCAB_DIR`` environment variable, or ``~/.cache/ppfleetx/<name>``) and raises a clear error otherwise.
``GPTTokenizer.byte_fallback()`` builds a 256-symbol byte vocabulary (+ ``<|endoftext|>``) that needs no files — handy
for tests and smoke runs.  Pre-training itself never needs a tokenizer (``eos_id`` is a dataset option).
"""
from __future__ import annotations

import json
import os
from functools import lru_cache
from typing import Dict, List, Optional, Tuple

import regex as re


@lru_cache()
def bytes_to_unicode() -> Dict[int, str]:
    keep = list(range(ord("!"), ord("~") + 1)) + list(range(0xA1, 0xAC + 1)) + list(range(0xAE, 0xFF + 1))
    chars, extra = keep[:], 0
    for b in range(256):
        if b not in keep:
            keep.append(b)
            chars.append(256 + extra)
            extra += 1
    return dict(zip(keep, (chr(c) for c in chars)))


_SPLIT = re.compile(r"""'s|'t|'re|'ve|'m|'ll|'d| ?\p{L}+| ?\p{N}+| ?[^\s\p{L}\p{N}]+|\s+(?!\S)|\s+""")


def get_pairs(word):
    """Set of adjacent symbol pairs of a word given as a tuple of symbols (the candidates of one BPE merge step)."""
    return set(zip(word, word[1:]))


class GPTTokenizer:
    eos_token = "<|endoftext|>"

    def __init__(self, vocab: Dict[str, int], merges: List[Tuple[str, str]], errors: str = "replace", max_len: Optional[int] = None,
                 pad_token: Optional[str] = None):
        self.encoder = dict(vocab)
        self.decoder = {v: k for k, v in self.encoder.items()}
        self.bpe_ranks = {tuple(m): i for i, m in enumerate(merges)}
        self.byte_encoder = bytes_to_unicode()
        self.byte_decoder = {v: k for k, v in self.byte_encoder.items()}
        self.errors = errors
        self.max_len = max_len or int(1e12)
        self._cache: Dict[str, List[str]] = {}
        self.eos_token_id = self.encoder.get(self.eos_token, len(self.encoder) - 1)
        self.bos_token_id = self.eos_token_id
        self.pad_token_id = self.encoder.get(pad_token, self.eos_token_id) if pad_token else self.eos_token_id
        self.padding_side = "left"

    # ---------------------------------------------------------------- construction
    @classmethod
    def from_pretrained(cls, name_or_dir: str = "gpt2", **kw) -> "GPTTokenizer":
        cands = [name_or_dir, os.environ.get("PFX_GPT_VOCAB_DIR", ""), os.path.expanduser(os.path.join("~/.cache/ppfleetx", name_or_dir))]
        for d in cands:
            if d and os.path.isdir(d):
                for vname, mname in (("vocab.json", "merges.txt"), ("gpt2-vocab.json", "gpt2-merges.txt")):
                    vp, mp = os.path.join(d, vname), os.path.join(d, mname)
                    if os.path.isfile(vp) and os.path.isfile(mp):
                        return cls.from_files(vp, mp, **kw)
        raise FileNotFoundError(
            f"GPT vocabulary for {name_or_dir!r} not found locally (looked in {[c for c in cands if c]}). This machine is offline: "
            "place vocab.json + merges.txt in a directory and pass it (or set PFX_GPT_VOCAB_DIR), or use GPTTokenizer.byte_fallback().")

    @classmethod
    def from_files(cls, vocab_file: str, merges_file: str, **kw) -> "GPTTokenizer":
        with open(vocab_file, encoding="utf-8") as f:
            vocab = json.load(f)
        with open(merges_file, encoding="utf-8") as f:
            lines = f.read().split("\n")
        merges = [tuple(l.split()) for l in lines[1:] if l and not l.startswith("#") and len(l.split()) == 2]
        return cls(vocab, merges, **kw)

    @classmethod
    def byte_fallback(cls, **kw) -> "GPTTokenizer":
        b2u = bytes_to_unicode()
        vocab = {b2u[i]: i for i in range(256)}
        vocab[cls.eos_token] = 256
        return cls(vocab, [], **kw)

    # ---------------------------------------------------------------- BPE
    def _bpe(self, token: str) -> List[str]:
        if token in self._cache:
            return self._cache[token]
        word = list(token)
        while len(word) > 1:
            pairs = [(self.bpe_ranks.get((a, b), 1 << 60), i) for i, (a, b) in enumerate(zip(word, word[1:]))]
            rank, _ = min(pairs)
            if rank == 1 << 60:
                break
            first, second = self._pair_of_rank(word, rank)
            merged, i = [], 0
            while i < len(word):
                if i < len(word) - 1 and word[i] == first and word[i + 1] == second:
                    merged.append(first + second)
                    i += 2
                else:
                    merged.append(word[i])
                    i += 1
            word = merged
        self._cache[token] = word
        return word

    def _pair_of_rank(self, word: List[str], rank: int) -> Tuple[str, str]:
        for a, b in zip(word, word[1:]):
            if self.bpe_ranks.get((a, b)) == rank:
                return a, b
        raise KeyError(rank)

    def tokenize(self, text: str) -> List[str]:
        out: List[str] = []
        for piece in _SPLIT.findall(text):
            mapped = "".join(self.byte_encoder[b] for b in piece.encode("utf-8"))
            out.extend(self._bpe(mapped))
        return out

    def convert_tokens_to_ids(self, tokens) -> List[int]:
        if isinstance(tokens, str):
            return self.encoder[tokens]
        return [self.encoder[t] for t in tokens]

    def convert_ids_to_tokens(self, ids, skip_special_tokens: bool = False) -> List[str]:
        # ids outside the vocabulary (a model whose embedding is padded past the tokenizer, or the byte-level fallback vocabulary
        # next to a 50304-row model) decode to nothing instead of raising
        return [self.decoder.get(int(i), "") for i in ids if not (skip_special_tokens and int(i) == self.eos_token_id)]

    def encode(self, text: str) -> List[int]:
        return self.convert_tokens_to_ids(self.tokenize(text))

    def decode(self, ids, skip_special_tokens: bool = False) -> str:
        text = "".join(self.convert_ids_to_tokens(ids, skip_special_tokens))
        return bytearray(self.byte_decoder[c] for c in text if c in self.byte_decoder).decode("utf-8", errors=self.errors)

    convert_ids_to_string = decode

    def __len__(self) -> int:
        return len(self.encoder)

    @property
    def vocab_size(self) -> int:
        return len(self.encoder)

    def __call__(self, text, padding: bool = False, max_length: Optional[int] = None, return_attention_mask: bool = True, **unused):
        texts = [text] if isinstance(text, str) else list(text)
        ids = [self.encode(t)[: max_length or self.max_len] for t in texts]
        if padding:
            mx = max(len(i) for i in ids)
            if self.padding_side == "left":
                masks = [[0] * (mx - len(i)) + [1] * len(i) for i in ids]
                ids = [[self.pad_token_id] * (mx - len(i)) + i for i in ids]
            else:
                masks = [[1] * len(i) + [0] * (mx - len(i)) for i in ids]
                ids = [i + [self.pad_token_id] * (mx - len(i)) for i in ids]
        else:
            masks = [[1] * len(i) for i in ids]
        out = {"input_ids": ids if not isinstance(text, str) else ids[0]}
        if return_attention_mask:
            out["attention_mask"] = masks if not isinstance(text, str) else masks[0]
        return out


class GPTChineseTokenizer:
    """Sentencepiece tokenizer of the Chinese GPT (CPM) checkpoints (the reference imports paddlenlp's ``GPTChineseTokenizer``,
    language_module.py:37-44 / gpt_dataset.py:30-39).  Words are segmented first (jieba when importable, otherwise the text is
    passed through unsegmented), spaces / newlines are carried through sentencepiece as ``\u2582`` / ``\u2583`` and restored by ``decode``."""

    vocab_file_name = "sentencepiece.model"
    _FWD = str.maketrans(" \n", "\u2582\u2583")

    def __init__(self, model_file: str, max_len: int = 512, unk_token: str = "<unk>", bos_token: str = "<bod>", eos_token: str = "<eod>",
                 eol_token: str = "\u2583"):
        import sentencepiece as spm

        if not os.path.isfile(model_file):
            raise FileNotFoundError(model_file)
        self.model_file, self.max_len = model_file, max_len or int(1e12)
        self.sp = spm.SentencePieceProcessor()
        self.sp.Load(model_file)
        self.unk_token, self.bos_token, self.eos_token, self.eol_token = unk_token, bos_token, eos_token, eol_token
        self.padding_side = "left"

    @classmethod
    def from_pretrained(cls, name_or_dir: str = "gpt-cpm-large-cn", **kw) -> "GPTChineseTokenizer":
        cands = [name_or_dir, os.environ.get("PFX_GPT_VOCAB_DIR", ""), os.path.expanduser(os.path.join("~/.cache/ppfleetx", name_or_dir))]
        for d in cands:
            if d and os.path.isfile(d):
                return cls(d, **kw)
            if d and os.path.isdir(d):
                for name in (cls.vocab_file_name, "gpt-cpm-cn-sentencepiece.model", "spiece.model"):
                    if os.path.isfile(os.path.join(d, name)):
                        return cls(os.path.join(d, name), **kw)
        raise FileNotFoundError(f"sentencepiece model for {name_or_dir!r} not found locally (looked in {[c for c in cands if c]}); this machine is offline")

    def _id_or_unk(self, piece: str) -> int:
        i = self.sp.piece_to_id(piece)
        return i

    @property
    def eos_token_id(self) -> int:
        return self._id_or_unk(self.eos_token)

    @property
    def bos_token_id(self) -> int:
        return self._id_or_unk(self.bos_token)

    @property
    def eol_token_id(self) -> int:
        return self._id_or_unk(self.eol_token)

    pad_token_id = eos_token_id

    @property
    def vocab_size(self) -> int:
        return self.sp.get_piece_size()

    def __len__(self) -> int:
        return self.sp.get_piece_size()

    def _segment(self, text: str) -> List[str]:
        try:
            import jieba

            return list(jieba.cut(text, cut_all=False))
        except ImportError:
            return [text]

    def tokenize(self, text: str) -> List[str]:
        joined = " ".join(w.translate(self._FWD) for w in self._segment(text))
        return self.sp.encode(joined, out_type=str)

    def convert_tokens_to_ids(self, tokens):
        if isinstance(tokens, str):
            return self.sp.piece_to_id(tokens)
        return [self.sp.piece_to_id(t) for t in tokens]

    def convert_ids_to_tokens(self, ids, skip_special_tokens: bool = False) -> List[str]:
        n = self.sp.get_piece_size()
        return [self.sp.IdToPiece(int(i)) if 0 <= int(i) < n else "" for i in ids if not (skip_special_tokens and int(i) == self.eos_token_id)]

    def encode(self, text: str) -> List[int]:
        return self.convert_tokens_to_ids(self.tokenize(text))

    def decode(self, ids, skip_special_tokens: bool = False) -> str:
        n = self.sp.get_piece_size()
        keep = [int(i) for i in ids if 0 <= int(i) < n and not (skip_special_tokens and int(i) == self.eos_token_id)]
        text = self.sp.decode(keep)
        return text.replace(" ", "").replace("\u2582", " ").replace("\u2583", "\n")

    convert_ids_to_string = decode

    def __call__(self, text, padding: bool = False, max_length: Optional[int] = None, return_attention_mask: bool = True, **unused):
        return GPTTokenizer.__call__(self, text, padding=padding, max_length=max_length, return_attention_mask=return_attention_mask)
