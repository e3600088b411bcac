"""ERNIE dataset helpers (reference ppfleetx/data/dataset/ernie/dataset_utils.py:32-700): memory-mapped corpus reader, the sample-span
mapping built by the C++ helper, sentence-pair sampling, truncation, span / whole-word masking and padding.  ``ernie_dataset.py`` builds
the datasets on top of these.

Corpus format: ``<prefix>_ids.npy`` (flat token ids) + ``<prefix>_idx.npz`` with ``lens`` (tokens per *sentence*) and ``docs`` (sentence
index where each document starts, length n_docs + 1).  Sample spans come from the C++ ``build_mapping`` (cached as
``<prefix>_<name>_indexmap_<E>ep_<N>mns_<L>msl_<P>ssp_<S>s.npy``)."""
from __future__ import annotations

import os
import time
from typing import List, Optional, Sequence, Tuple

import numpy as np

from ....utils.log import logger
from ..gpt_dataset import _helpers


def get_local_rank() -> int:
    return int(os.environ.get("LOCAL_RANK", os.environ.get("PADDLE_RANK_IN_NODE", 0)))


def get_datasets_weights_and_num_samples(data_prefix, train_valid_test_num_samples):
    """``[w1, prefix1, w2, prefix2, ...]`` -> ``(prefixes, normalised weights, per-dataset [train, valid, test] sample counts)``; each count is
    the weighted share plus 0.5 % head-room so that a blend that does not draw perfectly evenly never runs dry (reference dataset_utils.py:46-75)."""
    import math

    assert len(data_prefix) % 2 == 0, "data_prefix must alternate weight, prefix"
    weights = [float(w) for w in data_prefix[0::2]]
    prefixes = [str(p).strip() for p in data_prefix[1::2]]
    total = sum(weights)
    assert total > 0.0
    weights = [w / total for w in weights]
    counts = [[int(math.ceil(n * w * 1.005)) for n in train_valid_test_num_samples] for w in weights]
    return prefixes, weights, counts


class MMapIndexedDataset:
    """Sentence-addressable view of the flat token stream."""

    def __init__(self, prefix: str):
        self.ids = np.load(prefix + "_ids.npy", mmap_mode="r", allow_pickle=True)
        idx = np.load(prefix + "_idx.npz")
        self.sizes = idx["lens"].astype(np.int32)
        self.doc_idx = idx["docs"].astype(np.int64) if "docs" in idx.files else np.arange(len(self.sizes) + 1, dtype=np.int64)
        self.starts = np.concatenate([[0], np.cumsum(self.sizes, dtype=np.int64)])

    def __len__(self):
        return len(self.sizes)

    def __getitem__(self, i: int) -> np.ndarray:
        return np.asarray(self.ids[self.starts[i]:self.starts[i + 1]])


def get_samples_mapping(indexed: MMapIndexedDataset, doc_idx: np.ndarray, prefix: str, num_epochs: Optional[int], max_num_samples: Optional[int],
                        max_seq_length: int, short_seq_prob: float, seed: int, name: str, binary_head: bool, build: bool = True) -> np.ndarray:
    if not num_epochs:
        assert max_num_samples, "Need to specify either max_num_samples or num_epochs"
        num_epochs = np.iinfo(np.int32).max - 1
    if not max_num_samples:
        max_num_samples = np.iinfo(np.int64).max - 1
    fname = f"{prefix}_{name}_indexmap"
    if num_epochs != np.iinfo(np.int32).max - 1:
        fname += f"_{num_epochs}ep"
    if max_num_samples != np.iinfo(np.int64).max - 1:
        fname += f"_{max_num_samples}mns"
    fname += f"_{max_seq_length}msl_{short_seq_prob:0.2f}ssp_{seed}s.npy"
    if build and not os.path.isfile(fname):
        t0 = time.time()
        mapping = _helpers().build_mapping(doc_idx, indexed.sizes, int(num_epochs), int(max_num_samples), int(max_seq_length), float(short_seq_prob),
                                           int(seed), False, 2 if binary_head else 1)
        np.save(fname, mapping, allow_pickle=True)
        logger.info(f"built ERNIE samples mapping {os.path.basename(fname)} ({mapping.shape[0]} samples) in {time.time() - t0:.2f}s")
    else:
        while not os.path.isfile(fname):
            time.sleep(1)
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
    return np.load(fname, allow_pickle=True, mmap_mode="r")


# ------------------------------------------------------------------------------------------- sample construction
def get_a_and_b_segments(sample: List[np.ndarray], rng: np.random.RandomState) -> Tuple[List[int], List[int], bool]:
    n = len(sample)
    assert n > 1, "make sure each sample has at least two sentences."
    a_end = 1 if n < 3 else rng.randint(1, n)
    a = [int(t) for s in sample[:a_end] for t in s]
    b = [int(t) for s in sample[a_end:] for t in s]
    swapped = rng.random() < 0.5
    if swapped:
        a, b = b, a
    return a, b, swapped


def truncate_segments(a: List[int], b: List[int], max_tokens: int, rng: np.random.RandomState) -> bool:
    if len(a) + len(b) <= max_tokens:
        return False
    while len(a) + len(b) > max_tokens:
        seg = a if len(a) > len(b) else b
        if rng.random() < 0.5:
            del seg[0]
        else:
            seg.pop()
    return True


def create_tokens_and_tokentypes(a: List[int], b: List[int], cls_id: int, sep_id: int) -> Tuple[List[int], List[int]]:
    tokens = [cls_id] + a + [sep_id]
    types = [0] * len(tokens)
    if b:
        tokens += b + [sep_id]
        types += [1] * (len(b) + 1)
    return tokens, types


def create_masked_lm_predictions(tokens: List[int], vocab_size: int, cls_id: int, sep_id: int, mask_id: int, masked_lm_prob: float,
                                 max_predictions: int, rng: np.random.RandomState, max_ngrams: int = 3, is_word_start=None,
                                 favor_longer_ngram: bool = False):
    """BERT/ERNIE masking: n-gram spans over (whole-)word units, 80 % [MASK] / 10 % random / 10 % keep."""
    units: List[List[int]] = []
    for i, t in enumerate(tokens):
        if t in (cls_id, sep_id):
            continue
        starts_word = True if is_word_start is None else bool(is_word_start(t))
        if units and not starts_word:
            units[-1].append(i)
        else:
            units.append([i])
    out = list(tokens)
    num_to_predict = min(max_predictions, max(1, int(round(len(tokens) * masked_lm_prob))))
    ngrams = np.arange(1, max_ngrams + 1)
    pvals = 1.0 / ngrams
    pvals = pvals / pvals.sum()
    if favor_longer_ngram:
        pvals = pvals[::-1]
    order = list(range(len(units)))
    rng.shuffle(order)
    covered, picked = set(), []
    for u in order:
        if len(picked) >= num_to_predict:
            break
        n = int(rng.choice(ngrams, p=pvals))
        span = [i for unit in units[u:u + n] for i in unit]
        while span and len(picked) + len(span) > num_to_predict:
            n -= 1
            span = [i for unit in units[u:u + n] for i in unit] if n > 0 else []
        if not span or any(i in covered for i in span):
            continue
        for i in span:
            covered.add(i)
            r = rng.random()
            if r < 0.8:
                out[i] = mask_id
            elif r < 0.9:
                out[i] = int(rng.randint(0, vocab_size))
            picked.append((i, tokens[i]))
    picked.sort()
    return out, [p for p, _ in picked], [l for _, l in picked]


def pad_and_convert_to_numpy(tokens, tokentypes, positions, labels, pad_id: int, max_seq_length: int):
    n = len(tokens)
    pad = max_seq_length - n
    assert pad >= 0
    tok = np.asarray(tokens + [pad_id] * pad, dtype=np.int64)
    typ = np.asarray(tokentypes + [pad_id] * pad, dtype=np.int64)
    mask = np.asarray([1] * n + [0] * pad, dtype=np.float32)
    return tok, typ, mask, np.asarray(positions, dtype=np.int64), np.asarray(labels, dtype=np.int64)


def make_indexed_dataset(data_prefix, data_impl=None, skip_warmup=False):
    """``<prefix>_ids.npy`` + ``<prefix>_idx.npz`` -> ``MMapIndexedDataset`` (``data_impl`` / ``skip_warmup`` kept for call-site parity)."""
    return MMapIndexedDataset(data_prefix)


def get_indexed_dataset_(data_prefix, data_impl=None, skip_warmup=False):
    """``make_indexed_dataset`` plus the consistency check and the statistics the reference logs (dataset_utils.py:529-545)."""
    import time

    t0 = time.time()
    ds = make_indexed_dataset(data_prefix, data_impl, skip_warmup)
    assert ds.sizes.shape[0] == ds.doc_idx[-1], "the last document boundary must equal the number of sentences"
    logger.info(f" > indexed dataset {data_prefix}: {ds.doc_idx.shape[0] - 1} documents, {ds.sizes.shape[0]} sentences "
                f"({time.time() - t0:.4f}s)")
    return ds


def is_start_piece(piece: str) -> bool:
    """WordPiece continuation pieces start with ``##``; everything else starts a word (whole-word masking groups on this)."""
    return not piece.startswith("##")


def get_train_valid_test_split_(splits, size: int):
    """``"949,50,1"`` or ``[949, 50, 1]`` -> four cumulative document boundaries ``[0, a, b, size]``."""
    parts = [float(s) for s in (splits.split(",") if isinstance(splits, str) else splits)]
    parts = (parts + [0.0, 0.0, 0.0])[:3]
    total = sum(parts)
    assert total > 0
    bounds, acc = [0], 0.0
    for p in parts:
        acc += p / total
        bounds.append(int(round(acc * size)))
    bounds[-1] = size
    return bounds
