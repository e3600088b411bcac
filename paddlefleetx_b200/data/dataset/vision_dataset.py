"""Vision datasets (reference data/dataset/vision_dataset.py:33-426): ``GeneralClsDataset`` (list file ``path label``),
``ImageFolder``, ``CIFAR10`` (python pickle batches), ``ContrativeLearningDataset`` (two augmented views, MoCo) and a
``SyntheticImageDataset`` for benchmarking."""
from __future__ import annotations

import os
import pickle
from typing import Optional

import numpy as np
import torch

from ..transforms.preprocess import build_transforms, transform


class GeneralClsDataset(torch.utils.data.Dataset):
    def __init__(self, image_root: str, cls_label_path: str, transform_ops=None, delimiter: Optional[str] = None, class_num: Optional[int] = None,
                 multi_label: bool = False, **unused):
        self.root, self.ops = image_root, build_transforms(transform_ops)
        self.images, self.labels = [], []
        with open(cls_label_path) as f:
            for line in f:
                parts = line.strip().split(delimiter or " ")
                if len(parts) < 2:
                    continue
                self.images.append(os.path.join(image_root, parts[0]))
                self.labels.append(np.int64(parts[1]))

    def __len__(self):
        return len(self.images)

    def __getitem__(self, i):
        with open(self.images[i], "rb") as f:
            img = f.read()
        return transform(img, self.ops), self.labels[i]


class ImageFolder(torch.utils.data.Dataset):
    EXT = (".jpg", ".jpeg", ".png", ".ppm", ".bmp", ".pgm", ".tif", ".tiff", ".webp")

    def __init__(self, root: str, transform_ops=None, **unused):
        self.ops = build_transforms(transform_ops)
        classes = sorted(d for d in os.listdir(root) if os.path.isdir(os.path.join(root, d)))
        self.class_to_idx = {c: i for i, c in enumerate(classes)}
        self.samples = []
        for c in classes:
            for dp, _, files in sorted(os.walk(os.path.join(root, c))):
                for fn in sorted(files):
                    if fn.lower().endswith(self.EXT):
                        self.samples.append((os.path.join(dp, fn), self.class_to_idx[c]))

    def __len__(self):
        return len(self.samples)

    def __getitem__(self, i):
        path, y = self.samples[i]
        with open(path, "rb") as f:
            return transform(f.read(), self.ops), np.int64(y)


class CIFAR10(torch.utils.data.Dataset):
    def __init__(self, root: str, mode: str = "train", transform_ops=None, **unused):
        self.ops = build_transforms(transform_ops)
        files = [f"data_batch_{i}" for i in range(1, 6)] if mode == "train" else ["test_batch"]
        xs, ys = [], []
        for fn in files:
            with open(os.path.join(root, fn), "rb") as f:
                d = pickle.load(f, encoding="latin1")
            xs.append(np.asarray(d["data"]).reshape(-1, 3, 32, 32).transpose(0, 2, 3, 1))
            ys.extend(d.get("labels", d.get("fine_labels")))
        self.data, self.targets = np.concatenate(xs), np.asarray(ys, np.int64)

    def __len__(self):
        return len(self.data)

    def __getitem__(self, i):
        return transform(self.data[i], self.ops), self.targets[i]


class ContrativeLearningDataset(torch.utils.data.Dataset):
    """Two independently augmented views of every image (MoCo)."""

    def __init__(self, root: str, transform_ops=None, **unused):
        self.base = ImageFolder(root, None)
        self.ops = build_transforms(transform_ops)

    def __len__(self):
        return len(self.base)

    def __getitem__(self, i):
        path, _ = self.base.samples[i]
        with open(path, "rb") as f:
            raw = f.read()
        return transform(raw, self.ops), transform(raw, self.ops)


class SyntheticImageDataset(torch.utils.data.Dataset):
    def __init__(self, image_size: int = 224, class_num: int = 1000, num_samples: int = 1 << 16, seed: int = 0, two_views: bool = False, **unused):
        self.size, self.classes, self.n, self.seed, self.two = image_size, class_num, num_samples, seed, two_views

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        rng = np.random.default_rng(self.seed * 7919 + i)
        img = rng.standard_normal((3, self.size, self.size), dtype=np.float32)
        if self.two:
            return img, rng.standard_normal((3, self.size, self.size), dtype=np.float32)
        return img, np.int64(rng.integers(0, self.classes))
