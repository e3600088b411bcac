"""ERNIE pre-training dataset: sentence-pair samples with MLM masking (whole-word / n-gram) and sentence-order labels.

Reference: ppfleetx/data/dataset/ernie/{ernie_dataset.py:46-479, dataset_utils.py:78-675}.  Corpus format: ``<prefix>_ids.npy``
(flat token ids) + ``<prefix>_idx.npz`` with ``lens`` (tokens per *sentence*) and ``docs`` (sentence index where each document
starts, length n_docs + 1).  Sample spans come from the C++ ``build_mapping`` (cached as
``<prefix>_<name>_indexmap_<E>ep_<N>mns_<L>msl_<P>ssp_<S>s.npy``).  Special token ids are options (no tokenizer download).
"""
from __future__ import annotations

import os
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from ..gpt_dataset import find_corpus_prefix, train_valid_test_split
from .dataset_utils import (MMapIndexedDataset, create_masked_lm_predictions, create_tokens_and_tokentypes, get_a_and_b_segments,  # noqa: F401
                            get_samples_mapping, pad_and_convert_to_numpy, truncate_segments)


def get_local_rank() -> int:
    """Rank inside the node: decides who builds the index caches (the launcher's ``LOCAL_RANK``; Paddle's ``PADDLE_RANK_IN_NODE`` is honoured too)."""
    return int(os.environ.get("LOCAL_RANK", os.environ.get("PADDLE_RANK_IN_NODE", 0)))


def get_train_data_file(input_dir: str):
    """Corpus prefixes found in ``input_dir`` (same lookup as the GPT dataset; reference ernie_dataset.py)."""
    from ..gpt_dataset import get_train_data_file as _find

    return _find(input_dir)


def get_train_valid_test_split_(splits, size: int):
    from .dataset_utils import get_train_valid_test_split_ as _split

    return _split(splits, size)


def build_training_sample(sample, target_seq_length, max_seq_length, vocab_id_list=None, vocab_id_to_token_dict=None, vocab_token_to_id_dict=None,
                          cls_id=1, sep_id=2, mask_id=3, pad_id=0, masked_lm_prob=0.15, np_rng=None, binary_head=True, favor_longer_ngram=False,
                          max_ngrams=3, vocab_size=None):
    """One pre-training example from a list of sentences (each a sequence of token ids) — reference ernie_dataset.py:156-260: split into
    segments A / B (swapped half of the time when ``binary_head``), truncate to ``target_seq_length``, add [CLS] / [SEP] and token types,
    mask n-grams for the MLM objective, pad to ``max_seq_length``.  Random replacement ids are drawn from ``vocab_id_list`` (or
    ``range(vocab_size)``).  Returns the reference's dict: ``text``, ``types``, ``labels`` (-1 where nothing is predicted), ``is_random``,
    ``loss_mask``, ``padding_mask``, ``truncated`` — plus ``masked_positions`` / ``masked_labels``, the compact form our collate uses."""
    np_rng = np_rng if np_rng is not None else np.random.RandomState(1234)
    if binary_head:
        assert len(sample) > 1, "The sentence num should be large than 1."
    assert target_seq_length <= max_seq_length
    if binary_head:
        a, b, swapped = get_a_and_b_segments(sample, np_rng)
    else:
        a, b, swapped = [int(t) for sent in sample for t in sent], [], False
    truncated = truncate_segments(a, b, target_seq_length, np_rng)
    tokens, types = create_tokens_and_tokentypes(a, b, cls_id, sep_id)
    if vocab_size is None:
        vocab_size = (max(vocab_id_list) + 1) if vocab_id_list is not None else 40000
    max_pred = int(masked_lm_prob * max_seq_length) + 1
    masked, positions, labels = create_masked_lm_predictions(tokens, vocab_size, cls_id, sep_id, mask_id, masked_lm_prob, max_pred, np_rng, max_ngrams,
                                                             favor_longer_ngram=favor_longer_ngram)
    tok, typ, pad_mask, pos, lab = pad_and_convert_to_numpy(masked, types, positions, labels, pad_id, max_seq_length)
    dense_labels = np.full(max_seq_length, -1, dtype=np.int64)
    loss_mask = np.zeros(max_seq_length, dtype=np.int64)
    dense_labels[pos], loss_mask[pos] = lab, 1
    return {"text": tok, "types": typ, "labels": dense_labels, "is_random": int(swapped), "loss_mask": loss_mask, "padding_mask": pad_mask,
            "truncated": int(truncated), "masked_positions": pos, "masked_labels": lab}


class ErnieDataset(torch.utils.data.Dataset):
    def __init__(self, input_dir: str, split: Sequence[float] = (949, 50, 1), max_seq_len: int = 512, max_seq_length: Optional[int] = None,
                 num_samples: Optional[int] = None, mode: str = "Train", masked_lm_prob: float = 0.15, short_seq_prob: float = 0.1, seed: int = 1234,
                 binary_head: bool = True, vocab_size: Optional[int] = None, cls_id: int = 1, sep_id: int = 2, mask_id: int = 3, pad_id: int = 0,
                 max_ngrams: int = 3, favor_longer_ngram: bool = False, share_folder: bool = False, tokenizer_type=None, **unused):
        # special ids and the id range random replacements are drawn from: the vocabulary named by ``tokenizer_type`` when it is on this
        # machine (reference: get_ernie_tokenizer, ernie_dataset.py:57), otherwise the conventional ERNIE layout ([PAD] [CLS] [SEP] [MASK] = 0..3)
        if tokenizer_type:
            try:
                from ...tokenizers import get_ernie_tokenizer

                tok = get_ernie_tokenizer(tokenizer_type)
                cls_id, sep_id, mask_id, pad_id = tok.cls_token_id, tok.sep_token_id, tok.mask_token_id, tok.pad_token_id
                vocab_size = vocab_size or tok.vocab_size
            except FileNotFoundError:
                pass
        vocab_size = vocab_size or 40000
        self.max_seq_length = max_seq_length or max_seq_len
        self.mode, self.seed, self.binary_head = mode, seed, binary_head
        self.masked_lm_prob, self.vocab_size = masked_lm_prob, vocab_size
        self.cls_id, self.sep_id, self.mask_id, self.pad_id = cls_id, sep_id, mask_id, pad_id
        self.max_ngrams, self.favor_longer_ngram = max_ngrams, favor_longer_ngram
        prefix = find_corpus_prefix(input_dir)
        self.indexed = MMapIndexedDataset(prefix)
        n_docs = len(self.indexed.doc_idx) - 1
        bounds = train_valid_test_split(split, n_docs)
        k = {"Train": 0, "Eval": 1, "Test": 2}[mode]
        doc_idx = self.indexed.doc_idx[bounds[k]:bounds[k + 1] + 1]
        local_rank = get_local_rank()
        self.samples_mapping = get_samples_mapping(self.indexed, doc_idx, prefix, None, num_samples, self.max_seq_length - 3, short_seq_prob, seed,
                                                   "ernie_" + mode, binary_head, build=(local_rank == 0))

    def __len__(self):
        return self.samples_mapping.shape[0]

    def __getitem__(self, idx: int):
        start, end, target_len = (int(v) for v in self.samples_mapping[idx])
        sample = [self.indexed[i] for i in range(start, end)]
        rng = np.random.RandomState(seed=((self.seed + idx) % 2 ** 32))
        ex = build_training_sample(sample, min(target_len, self.max_seq_length - 3), self.max_seq_length, cls_id=self.cls_id, sep_id=self.sep_id,
                                   mask_id=self.mask_id, pad_id=self.pad_id, masked_lm_prob=self.masked_lm_prob, np_rng=rng,
                                   binary_head=self.binary_head, favor_longer_ngram=self.favor_longer_ngram, max_ngrams=self.max_ngrams,
                                   vocab_size=self.vocab_size)
        return [ex["text"], ex["types"], ex["padding_mask"], ex["masked_positions"], ex["masked_labels"],
                np.asarray([ex["is_random"]], dtype=np.int64)]


class SyntheticErnieDataset(torch.utils.data.Dataset):
    """Random MLM + SOP samples in the exact sample format of :class:`ErnieDataset` (no corpus needed): smoke runs, benchmarks."""

    def __init__(self, max_seq_length: int = 512, vocab_size: int = 40000, masked_lm_prob: float = 0.15, num_samples: int = 1 << 16, seed: int = 1234,
                 **unused):
        self.seq, self.vocab, self.n, self.seed = int(max_seq_length), int(vocab_size), int(num_samples), int(seed)
        self.n_mask = max(1, int(round(self.seq * masked_lm_prob)))

    def __len__(self):
        return self.n

    def __getitem__(self, idx: int):
        rng = np.random.default_rng(self.seed * 1000003 + idx)
        tok = rng.integers(4, self.vocab, size=self.seq, dtype=np.int64)
        split = int(rng.integers(self.seq // 4, 3 * self.seq // 4))
        typ = (np.arange(self.seq) >= split).astype(np.int64)
        pos = np.sort(rng.choice(self.seq, size=self.n_mask, replace=False)).astype(np.int64)
        lab = rng.integers(4, self.vocab, size=self.n_mask, dtype=np.int64)
        return [tok, typ, np.ones(self.seq, dtype=np.float32), pos, lab, np.asarray([int(rng.integers(0, 2))], dtype=np.int64)]


class ErnieSeqClsDataset(torch.utils.data.Dataset):
    """TSV ``text[\\ttext_b]\\tlabel`` classification data.  ``tokenizer_type`` (a vocabulary directory or cached name, reference
    ernie_dataset.py:328-334) selects the WordPiece tokenizer — rows then go through its pair encoding (``[CLS] a [SEP] b [SEP]``, longest-first
    truncation); without it a byte-level fallback keeps the recipe runnable on a box with no vocabulary."""

    def __init__(self, dataset_type: str = "chnsenticorp_v2", input_dir: Optional[str] = None, path: Optional[str] = None, mode: str = "Train",
                 max_seq_len: int = 128, tokenizer=None, tokenizer_type: Optional[str] = None, cls_id: int = 1, sep_id: int = 2, pad_id: int = 0, **unused):
        path = path or os.path.join(input_dir, {"Train": "train.tsv", "Eval": "dev.tsv", "Test": "test.tsv"}[mode])
        self.max_seq_len, self.cls_id, self.sep_id, self.pad_id = max_seq_len, cls_id, sep_id, pad_id
        self.wordpiece = None
        if tokenizer is None and tokenizer_type:
            from ...tokenizers import get_ernie_tokenizer

            self.wordpiece = tokenizer = get_ernie_tokenizer(tokenizer_type)
        if tokenizer is None:
            from ...tokenizers import GPTTokenizer

            tokenizer = GPTTokenizer.byte_fallback()
        self.tok = tokenizer
        self.rows = []
        with open(path, encoding="utf-8") as f:
            for i, line in enumerate(f):
                parts = line.rstrip("\n").split("\t")
                if i == 0 and not parts[-1].strip().lstrip("-").isdigit():
                    continue
                self.rows.append((parts[0], parts[1] if len(parts) > 2 else None, int(parts[-1])))

    def __len__(self):
        return len(self.rows)

    def __getitem__(self, i):
        a, b, y = self.rows[i]
        if self.wordpiece is not None:
            enc = self.wordpiece(a, text_pair=b, max_length=self.max_seq_len, truncation=True, return_token_type_ids=True)
            return {"input_ids": np.asarray(enc["input_ids"], np.int64), "token_type_ids": np.asarray(enc["token_type_ids"], np.int64), "labels": np.asarray(y, np.int64)}
        ia = [t + 4 for t in self.tok.encode(a)]
        ib = [t + 4 for t in self.tok.encode(b)] if b else []
        budget = self.max_seq_len - (3 if ib else 2)
        ia, ib = ia[: budget - min(len(ib), budget // 2)], ib[: budget // 2]
        tokens, types = create_tokens_and_tokentypes(ia, ib, self.cls_id, self.sep_id)
        return {"input_ids": np.asarray(tokens, np.int64), "token_type_ids": np.asarray(types, np.int64), "labels": np.asarray(y, np.int64)}
