"""Deterministic distributed batch samplers.

``GPTBatchSampler`` (reference ppfleetx/data/sampler/batch_sampler.py:31-192): walks the dataset in index order
(shuffling lives in the dataset's cached ``shuffle_idx``; the ``shuffle`` flag is accepted and ignored exactly
like the reference), groups ``batch_size * num_replicas`` consecutive indices into a global window and hands
replica ``rank`` its contiguous ``batch_size`` slice.  ``consumed_samples`` lets a resumed run continue from a
global sample offset.  Data-parallel rank = ``dp_rank * sharding_size + sharding_rank`` (env.py:158-178).
"""
from __future__ import annotations

import math
from typing import Iterator, List, Optional

import torch

from ...distributed.apis import env


class GPTBatchSampler(torch.utils.data.Sampler):
    def __init__(self, dataset, batch_size: int = 1, num_replicas: Optional[int] = None, rank: Optional[int] = None, shuffle: bool = False,
                 drop_last: bool = True, consumed_samples: int = 0, **unused):
        self.dataset = dataset
        assert isinstance(batch_size, int) and batch_size > 0, "batch_size should be a positive integer"
        self.batch_size = batch_size
        self.nranks = env.get_data_world_size() if num_replicas is None else int(num_replicas)
        self.local_rank = env.get_data_world_rank() if rank is None else int(rank)
        assert 0 <= self.local_rank < self.nranks
        self.drop_last = drop_last
        self.shuffle = shuffle          # intentionally unused
        self.epoch = 0
        self.consumed_samples = int(consumed_samples)
        self.skip_batches = 0           # one-shot: the next ``__iter__`` starts this many local batches in (resume without loading consumed data)
        self.total_size = len(dataset)
        per = self.total_size / self.nranks
        self.num_samples = int(math.floor(per) if drop_last else math.ceil(per))

    def set_epoch(self, epoch: int = 0, consumed_samples: int = 0) -> None:
        self.epoch = epoch
        self.consumed_samples = int(consumed_samples)

    def __iter__(self) -> Iterator[List[int]]:
        window = self.batch_size * self.nranks
        lo = self.local_rank * self.batch_size
        skip, self.skip_batches = self.skip_batches, 0
        pos = self.consumed_samples + skip * window
        while pos + window <= self.total_size:
            yield list(range(pos + lo, pos + lo + self.batch_size))
            pos += window
        if not self.drop_last and pos < self.total_size:
            rest = list(range(pos, self.total_size))
            mine = rest[lo:lo + self.batch_size]
            if mine:
                yield mine

    def __len__(self) -> int:
        remaining = max(self.total_size - self.consumed_samples, 0)
        n = remaining // (self.batch_size * self.nranks)
        if not self.drop_last and remaining % (self.batch_size * self.nranks):
            n += 1
        return n


class DistributedBatchSampler(torch.utils.data.Sampler):
    """Epoch-shuffling sampler for the vision / fine-tune datasets (Paddle's ``DistributedBatchSampler``)."""

    def __init__(self, dataset, batch_size: int = 1, num_replicas: Optional[int] = None, rank: Optional[int] = None, shuffle: bool = False,
                 drop_last: bool = False, seed: int = 0, **unused):
        self.dataset, self.batch_size, self.shuffle, self.drop_last, self.seed = dataset, batch_size, shuffle, drop_last, seed
        self.nranks = env.get_data_world_size() if num_replicas is None else int(num_replicas)
        self.local_rank = env.get_data_world_rank() if rank is None else int(rank)
        self.epoch = 0
        self.skip_batches = 0           # one-shot, as in GPTBatchSampler
        self.num_samples = int(math.ceil(len(dataset) / self.nranks))
        self.total_size = self.num_samples * self.nranks

    def set_epoch(self, epoch: int) -> None:
        self.epoch = epoch

    def __iter__(self):
        n = len(self.dataset)
        if self.shuffle:
            g = torch.Generator()
            g.manual_seed(self.seed + self.epoch)
            idx = torch.randperm(n, generator=g).tolist()
        else:
            idx = list(range(n))
        idx += idx[: self.total_size - n]
        idx = idx[self.local_rank:self.total_size:self.nranks]
        skip, self.skip_batches = self.skip_batches, 0
        idx = idx[skip * self.batch_size:]
        batch = []
        for i in idx:
            batch.append(i)
            if len(batch) == self.batch_size:
                yield batch
                batch = []
        if batch and not self.drop_last:
            yield batch

    def __len__(self) -> int:
        return self.num_samples // self.batch_size if self.drop_last else (self.num_samples + self.batch_size - 1) // self.batch_size
