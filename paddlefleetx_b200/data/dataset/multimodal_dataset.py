"""``ImagenDataset`` (reference data/dataset/multimodal_dataset.py:40-202): text-image shards listed in ``input_path`` (one
shard file per line), each shard a TSV whose columns include a caption and a base64-encoded image; shards are split across the
data-parallel ranks.  Captions are tokenised to ``input_ids`` / ``attention_mask`` for the frozen text tower, or pre-computed
text embeddings are passed through.  ``SyntheticImagenDataset`` yields random images + embeddings for benchmarking."""
from __future__ import annotations

import base64
import io
import os

import numpy as np
import torch

from ...distributed.apis import env


def get_keys(data_path: str, gpu_num: int, rank: int = None, seed: int = 0):
    """Shard files of ``data_path`` (one per line) for this rank: the list is padded with randomly re-drawn shards to a multiple of
    ``gpu_num`` so that every rank gets the same number, then dealt out round-robin (reference multimodal_dataset.py:40-62).  The padding draw
    is seeded, so all ranks agree on it without communicating."""
    import random

    with open(data_path) as f:
        files = [line.strip() for line in f if line.strip()]
    rank = env.get_data_world_rank() if rank is None else rank
    extended = list(files)
    if files and len(files) % gpu_num:
        need = gpu_num - len(files) % gpu_num
        rng = random.Random(seed)
        extended += rng.sample(files, need) if need <= len(files) else [rng.choice(files) for _ in range(need)]
    return extended[rank::gpu_num]


class ImagenDataset(torch.utils.data.Dataset):
    def __init__(self, input_path: str, image_size: int = 64, text_max_len: int = 128, filter_image_resolution: int = 128, image_format: str = "base64",
                 caption_col: int = 2, image_col: int = 5, tokenizer=None, split: bool = True, **unused):
        self.image_size, self.text_max_len, self.min_res = image_size, text_max_len, filter_image_resolution
        self.caption_col, self.image_col = caption_col, image_col
        if split and env.get_data_world_size() > 1:
            shards = get_keys(input_path, env.get_data_world_size())      # padded: every rank reads the same number of shards
        else:
            with open(input_path) as f:
                shards = [l.strip() for l in f if l.strip()]
        self.rows = []
        for shard in shards:
            path = shard if os.path.isabs(shard) else os.path.join(os.path.dirname(input_path), shard)
            with open(path, encoding="utf-8") as f:
                for line in f:
                    parts = line.rstrip("\n").split("\t")
                    if len(parts) > max(caption_col, image_col):
                        self.rows.append((parts[caption_col], parts[image_col]))
        if isinstance(tokenizer, str):           # a text-tower name or a vocabulary directory (t5-* / *deberta*)
            from ..tokenizers import get_text_tokenizer

            tokenizer = get_text_tokenizer(tokenizer, strict=True)
        if tokenizer is None:
            from ..tokenizers import GPTTokenizer

            tokenizer = GPTTokenizer.byte_fallback()
        self.tok = tokenizer

    def __len__(self):
        return len(self.rows)

    def _image(self, b64: str) -> np.ndarray:
        from PIL import Image

        img = Image.open(io.BytesIO(base64.b64decode(b64))).convert("RGB")
        s = self.image_size
        img = img.resize((s, s), Image.BICUBIC)
        return np.asarray(img, np.float32).transpose(2, 0, 1) / 255.0

    def __getitem__(self, i):
        caption, b64 = self.rows[i]
        ids = self.tok.encode(caption)[: self.text_max_len]
        return {"images": self._image(b64), "input_ids": np.asarray(ids, np.int64), "attention_mask": np.ones(len(ids), np.int64)}


class SyntheticImagenDataset(torch.utils.data.Dataset):
    def __init__(self, image_size: int = 64, text_embed_dim: int = 1024, text_len: int = 32, num_samples: int = 1 << 14, seed: int = 0, **unused):
        self.s, self.d, self.t, self.n, self.seed = image_size, text_embed_dim, text_len, num_samples, seed

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        rng = np.random.default_rng(self.seed * 9973 + i)
        return {"images": rng.random((3, self.s, self.s), dtype=np.float32), "text_embeds": rng.standard_normal((self.t, self.d), dtype=np.float32),
                "text_masks": np.ones(self.t, np.int64)}
