"""GLUE fine-tuning datasets: CoLA, SST-2, MNLI, QNLI, RTE, WNLI, MRPC, QQP, STS-B (reference
data/dataset/glue_dataset.py:48-841).  Each reads the standard ``<root>/<TASK>/{train,dev}.tsv`` files and yields
``{"input_ids": ..., "labels": ...}`` (GPT-style: single sequence, the pair joined by the eos/sep token); the
collator pads to the longest sample.  Data must be present locally (this box is offline)."""
from __future__ import annotations

import csv
import os
from typing import Optional

import numpy as np
import torch

_TASKS = {
    # name: (dir, train file, dev file, text_a col, text_b col, label col, labels or None for regression, skip_header)
    "CoLA": ("CoLA", "train.tsv", "dev.tsv", 3, None, 1, ["0", "1"], False),
    "SST2": ("SST-2", "train.tsv", "dev.tsv", 0, None, 1, ["0", "1"], True),
    "MNLI": ("MNLI", "train.tsv", "dev_matched.tsv", 8, 9, -1, ["contradiction", "entailment", "neutral"], True),
    "QNLI": ("QNLI", "train.tsv", "dev.tsv", 1, 2, -1, ["entailment", "not_entailment"], True),
    "RTE": ("RTE", "train.tsv", "dev.tsv", 1, 2, -1, ["entailment", "not_entailment"], True),
    "WNLI": ("WNLI", "train.tsv", "dev.tsv", 1, 2, -1, ["0", "1"], True),
    "MRPC": ("MRPC", "train.tsv", "dev.tsv", 3, 4, 0, ["0", "1"], True),
    "QQP": ("QQP", "train.tsv", "dev.tsv", 3, 4, 5, ["0", "1"], True),
    "STSB": ("STS-B", "train.tsv", "dev.tsv", 7, 8, -1, None, True),
}


class GlueDataset(torch.utils.data.Dataset):
    task = None

    def __init__(self, root: str, split: str = "train", max_length: int = 128, tokenizer=None, tokenizer_type: str = "GPTTokenizer",
                 tokenizer_name: str = "gpt2", vocab_dir: Optional[str] = None, dev_file: Optional[str] = None, **unused):
        d, train_f, dev_f, ca, cb, cl, labels, header = _TASKS[self.task]
        self.max_length = max_length
        self.labels = labels
        if tokenizer is None:
            from ..tokenizers import GPTTokenizer

            try:
                tokenizer = GPTTokenizer.from_pretrained(vocab_dir or tokenizer_name)
            except FileNotFoundError:
                tokenizer = GPTTokenizer.byte_fallback()
        self.tokenizer = tokenizer
        base = os.path.join(root, d) if os.path.isdir(os.path.join(root, d)) else root
        if split == "train":
            names = [train_f, os.path.join("raw", "in_domain_train.tsv")]
        elif split in ("dev", "eval", "validation"):
            names = [dev_file or dev_f, "dev.tsv", os.path.join("raw", "in_domain_dev.tsv")]
        else:                         # test, dev_matched, dev_mismatched, ...: the split names its own file
            names = [dev_file or f"{split}.tsv", dev_f]
        path = next((os.path.join(base, n) for n in names if n and os.path.isfile(os.path.join(base, n))), os.path.join(base, names[0]))
        self.samples = []
        with open(path, encoding="utf-8") as f:
            rows = list(csv.reader(f, delimiter="\t", quoting=csv.QUOTE_NONE))
        for r in rows[1 if header else 0:]:
            try:
                a = r[ca]
                b = r[cb] if cb is not None else None
                y = r[cl]
            except IndexError:
                continue
            label = float(y) if labels is None else labels.index(y)
            self.samples.append((a, b, label))

    def __len__(self):
        return len(self.samples)

    def __getitem__(self, i):
        a, b, y = self.samples[i]
        ids = self.tokenizer.encode(a)
        if b is not None:
            ids = ids + [self.tokenizer.eos_token_id] + self.tokenizer.encode(b)
        ids = ids[: self.max_length - 1] + [self.tokenizer.eos_token_id]
        return {"input_ids": np.asarray(ids, dtype=np.int64),
                "labels": np.asarray(y, dtype=np.float32 if self.labels is None else np.int64)}


def _make(name):
    return type(name, (GlueDataset,), {"task": name})


CoLA, SST2, MNLI, QNLI, RTE, WNLI, MRPC, QQP, STSB = (_make(n) for n in ("CoLA", "SST2", "MNLI", "QNLI", "RTE", "WNLI", "MRPC", "QQP", "STSB"))
