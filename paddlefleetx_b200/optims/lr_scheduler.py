"""Learning-rate schedules.

Behaviour follows ppfleetx/optims/lr_scheduler.py:31-192 with its latent bugs fixed (SURVEY F8:
``LinearDecayWithWarmup`` used an undefined ``total_steps``; ``CosineDecay`` read ``lr`` before assignment):

  * ``CosineAnnealingWithWarmupDecay`` — linear warm-up over ``warmup_rate * decay_steps``, cosine to
    ``min_lr``, constant afterwards.  ``step(epoch=N)`` advances the counter by N: GPT configs run it with
    ``use_increments`` so the counter is in *samples* (``decay_steps *= global_batch_size``).
  * ``LinearDecayWithWarmup``, ``ViTLRScheduler`` (cosine | linear + warm-up), ``MultiStepDecay``, ``CosineDecay``.
"""
from __future__ import annotations

import math
from typing import List, Optional


class LRScheduler:
    def __init__(self, learning_rate: float = 0.1, last_epoch: int = 0):
        self.base_lr = float(learning_rate)
        self.last_epoch = last_epoch
        self.last_lr = self.get_lr()

    def get_lr(self) -> float:
        raise NotImplementedError

    def __call__(self) -> float:
        return self.last_lr

    def step(self, epoch: Optional[int] = None) -> None:
        self.last_epoch += 1 if epoch is None else epoch
        self.last_lr = self.get_lr()

    def state_dict(self) -> dict:
        return {"last_epoch": self.last_epoch, "last_lr": self.last_lr}

    def set_state_dict(self, sd: dict) -> None:
        self.last_epoch = sd["last_epoch"]
        self.last_lr = sd.get("last_lr", self.get_lr())

    load_state_dict = set_state_dict


class CosineAnnealingWithWarmupDecay(LRScheduler):
    def __init__(self, max_lr: float, min_lr: float, warmup_rate: float, decay_steps: int, last_epoch: int = 0,
                 use_increments: bool = False, **unused):
        self.max_lr, self.min_lr = float(max_lr), float(min_lr)
        self.decay_steps = int(decay_steps)
        self.warmup_step = warmup_rate * decay_steps
        self.use_increments = use_increments
        super().__init__(max_lr, last_epoch)

    def get_lr(self) -> float:
        t = self.last_epoch
        if self.warmup_step > 0 and t <= self.warmup_step:
            return self.max_lr * t / self.warmup_step
        if t > self.decay_steps:
            return self.min_lr
        ratio = (t - self.warmup_step) / max(self.decay_steps - self.warmup_step, 1e-12)
        return self.min_lr + 0.5 * (math.cos(math.pi * ratio) + 1.0) * (self.max_lr - self.min_lr)


class LinearDecayWithWarmup(LRScheduler):
    """Linear warm-up to ``learning_rate`` over ``warmup`` (fraction if <1, else steps), then ``lr * (1 - t / total_steps)``: linear to 0 at
    ``total_steps`` (= ``step_each_epoch * epochs`` injected by tools/train.py)."""

    def __init__(self, learning_rate: float, step_each_epoch: int = 1, epochs: int = 1, warmup: float = 0.0, total_steps: Optional[int] = None,
                 last_epoch: int = 0, **unused):
        self.total = int(total_steps) if total_steps else int(step_each_epoch * epochs)
        self.warmup_steps = int(warmup * self.total) if warmup < 1 else int(warmup)
        super().__init__(learning_rate, last_epoch)

    def get_lr(self) -> float:
        t = self.last_epoch
        if self.warmup_steps > 0 and t < self.warmup_steps:
            return self.base_lr * t / self.warmup_steps
        return self.base_lr * max(0.0, 1.0 - t / max(1, self.total))      # reference formula (optims/lr_scheduler.py:96-100): 1 - t/T_max


class ViTLRScheduler(LRScheduler):
    def __init__(self, learning_rate: float, step_each_epoch: int = 1, epochs: int = 1, decay_type: str = "cosine", linear_end: float = 1e-5,
                 warmup_steps: int = 0, last_epoch: int = 0, total_steps: Optional[int] = None, **unused):
        self.T_max = int(total_steps) if total_steps else int(epochs * step_each_epoch)
        self.warmup_steps = int(warmup_steps)
        self.decay_type = decay_type
        self.linear_end = linear_end
        super().__init__(learning_rate, last_epoch)

    def get_lr(self) -> float:
        t = self.last_epoch
        progress = (t - self.warmup_steps) / float(max(1, self.T_max - self.warmup_steps))
        progress = min(1.0, max(0.0, progress))
        if self.decay_type == "linear":
            lr = self.linear_end + (self.base_lr - self.linear_end) * (1.0 - progress)
        elif self.decay_type == "cosine":
            lr = 0.5 * self.base_lr * (1.0 + math.cos(math.pi * progress))
        else:
            raise ValueError(f"unknown decay_type {self.decay_type}")
        if self.warmup_steps:
            lr = lr * min(1.0, t / self.warmup_steps)
        return lr


class MultiStepDecay(LRScheduler):
    def __init__(self, learning_rate: float, milestones: List[int], gamma: float = 0.1, last_epoch: int = 0, **unused):
        self.milestones = sorted(int(m) for m in milestones)
        self.gamma = gamma
        super().__init__(learning_rate, last_epoch)

    def get_lr(self) -> float:
        passed = sum(1 for m in self.milestones if self.last_epoch >= m)
        return self.base_lr * (self.gamma ** passed)


class CosineDecay(LRScheduler):
    def __init__(self, learning_rate: float, step_each_epoch: int = 1, epochs: int = 1, update_unit: str = "epoch", warmups: int = 0,
                 last_epoch: int = 0, **unused):
        self.T_max = epochs if update_unit == "epoch" else step_each_epoch * epochs
        self.warmups = warmups if update_unit == "epoch" else step_each_epoch * warmups
        assert self.warmups < self.T_max
        super().__init__(learning_rate, last_epoch)

    def get_lr(self) -> float:
        t = self.last_epoch
        if self.warmups > 0 and t < self.warmups:
            return self.base_lr * (t + 1) / self.warmups
        progress = (t - self.warmups) / float(max(1, self.T_max - self.warmups))
        return 0.5 * self.base_lr * (1.0 + math.cos(math.pi * min(1.0, progress)))
