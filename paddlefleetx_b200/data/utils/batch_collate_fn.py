"""Task-level collate functions (reference ppfleetx/data/utils/batch_collate_fn.py:31-211): GPT pre-training / evaluation batches, the ERNIE
masked-LM collator with its flattened mask positions, the padding collator of the fine-tuning tasks and Imagen's (image, text embedding,
mask) batches.  Built on the ``Stack`` / ``Pad`` primitives of ``data/sampler/collate.py``."""
from __future__ import annotations

from typing import List

import numpy as np
import torch

from ..sampler.collate import Dict, Pad, Stack, Tuple  # noqa: F401  (the primitives are part of this module's public surface too)

def _to_tensor(x):
    return torch.from_numpy(np.ascontiguousarray(x)) if isinstance(x, np.ndarray) else x


def gpt_collate_fn(batch):
    """Stack every field of the GPT sample tuple."""
    return [_to_tensor(np.stack([np.asarray(s[i]) for s in batch])) for i in range(len(batch[0]))]


def default_collate_fn(batch):
    return torch.utils.data.default_collate(batch)


def gpt_eval_collate_fn(batch):
    return gpt_collate_fn(batch)


class ErnieCollateData:
    """Flatten masked positions to 1-D (padded to a multiple of 8) and split into micro-batches
    (reference data/utils/batch_collate_fn.py:99-147)."""

    def __init__(self, micro_batch_size: int = 1):
        self.micro_batch_size = micro_batch_size

    def _one(self, data):
        num_fields = len(data[0])
        out = [[] for _ in range(num_fields)]
        batch_size, seq_len = len(data), len(data[0][0])
        for i in (0, 1, 2, 5):
            out[i] = np.stack([np.asarray(x[i]) for x in data])
        mlm_pos, mlm_lab = [], []
        for b, x in enumerate(data):
            for p, l in zip(np.asarray(x[3]).reshape(-1), np.asarray(x[4]).reshape(-1)):
                mlm_pos.append(b * seq_len + int(p)); mlm_lab.append(int(l))
        pad = (-len(mlm_pos)) % 8
        mlm_pos += [0] * pad; mlm_lab += [-1] * pad
        out[3] = np.asarray(mlm_pos, dtype=np.int32).reshape(-1)
        out[4] = np.asarray(mlm_lab, dtype=np.int64).reshape(-1, 1)
        return [_to_tensor(o) for o in out]

    def __call__(self, data):
        n = len(data)
        if n % self.micro_batch_size:
            return self._one(data)
        chunks = [data[i:i + self.micro_batch_size] for i in range(0, n, self.micro_batch_size)]
        if len(chunks) == 1:
            return self._one(data)
        return [self._one(c) for c in chunks]


class DataCollatorWithPadding:
    def __init__(self, tokenizer=None, pad_token_id: int = 0, padding: bool = True, max_length=None, return_attention_mask=None, **unused):
        self.pad_token_id = tokenizer.pad_token_id if tokenizer is not None and getattr(tokenizer, "pad_token_id", None) is not None else pad_token_id

    def __call__(self, features: List[dict]):
        keys = features[0].keys()
        out = {}
        for k in keys:
            vals = [np.asarray(f[k]) for f in features]
            if vals[0].ndim == 0:
                out[k] = torch.from_numpy(np.stack(vals))
            else:
                out[k] = torch.from_numpy(Pad(self.pad_token_id if "ids" in k else 0)(vals))
        return out


def imagen_collate_fn(batch):
    """Dict batch for Imagen: images stacked, text embeds/masks padded to the longest caption."""
    out = {}
    for k in batch[0]:
        vals = [np.asarray(b[k]) for b in batch]
        same = all(v.shape == vals[0].shape for v in vals)
        out[k] = torch.from_numpy(np.stack(vals) if same else Pad(0)(vals))
    return out

collate_fn = default_collate_fn
