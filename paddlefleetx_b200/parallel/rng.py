"""Named RNG streams (``global_seed`` / ``local_seed``) for dropout under tensor parallelism.

Contract from the reference call sites (SURVEY §2.5 ``get_rng_state_tracker``;
ppfleetx/distributed/apis/env.py:86-88, gpt/dygraph/hybrid_model.py:328,651,732):
``tracker.add(name, seed)`` registers a stream, ``with tracker.rng_state(name):`` runs a block on
that stream and restores the default stream afterwards.

B200-first twist: besides swapping the torch generator state (needed for library dropout and for
the CPU path) every stream owns a *Philox (seed, offset)* pair.  Our fused dropout kernels are
stateless counter-based generators: they take ``(seed, offset)``, and the backward pass regenerates
the mask from the same pair instead of storing it — no mask tensor ever hits HBM.
"""
from __future__ import annotations

import contextlib
from typing import Dict, Tuple

import torch

_DEFAULT = "__default__"


class RNGStatesTracker:
    def __init__(self) -> None:
        self.reset()

    def reset(self) -> None:
        self._cpu_states: Dict[str, torch.Tensor] = {}
        self._cuda_states: Dict[str, torch.Tensor] = {}
        self._seeds: Dict[str, int] = {}
        self._philox_offsets: Dict[str, int] = {}
        self._active = _DEFAULT

    def add(self, name: str, seed: int) -> None:
        if name in self._seeds:
            raise ValueError(f"rng state {name} already exists")
        if seed in self._seeds.values():
            raise ValueError(f"seed {seed} already exists")
        self._seeds[name] = int(seed)
        self._philox_offsets[name] = 0
        cpu_prev = torch.get_rng_state()
        torch.manual_seed(seed)  # seeds CPU and (lazily) every CUDA device
        self._cpu_states[name] = torch.get_rng_state()
        torch.set_rng_state(cpu_prev)
        if torch.cuda.is_available():
            cuda_prev = torch.cuda.get_rng_state()
            torch.cuda.manual_seed(seed)
            self._cuda_states[name] = torch.cuda.get_rng_state()
            torch.cuda.set_rng_state(cuda_prev)

    def has(self, name: str) -> bool:
        return name in self._seeds

    def get_states_tracker(self) -> dict:
        return {"cpu": dict(self._cpu_states), "cuda": dict(self._cuda_states),
                "seeds": dict(self._seeds), "offsets": dict(self._philox_offsets)}

    def set_states_tracker(self, st: dict) -> None:
        self._cpu_states = dict(st["cpu"])
        self._cuda_states = dict(st["cuda"])
        self._seeds = dict(st["seeds"])
        self._philox_offsets = dict(st["offsets"])

    @contextlib.contextmanager
    def rng_state(self, name: str = "global_seed"):
        if name not in self._seeds:
            # untracked name (e.g. single-card run before set_seed): use the default stream
            yield
            return
        prev_active = self._active
        cpu_prev = torch.get_rng_state()
        torch.set_rng_state(self._cpu_states[name])
        cuda_prev = None
        if name in self._cuda_states:
            cuda_prev = torch.cuda.get_rng_state()
            torch.cuda.set_rng_state(self._cuda_states[name])
        self._active = name
        try:
            yield
        finally:
            self._cpu_states[name] = torch.get_rng_state()
            torch.set_rng_state(cpu_prev)
            if cuda_prev is not None:
                self._cuda_states[name] = torch.cuda.get_rng_state()
                torch.cuda.set_rng_state(cuda_prev)
            self._active = prev_active

    # -- counter-based stream for fused kernels ---------------------------------------
    def philox(self, numel: int, name: str | None = None) -> Tuple[int, int]:
        """Reserve ``numel`` random numbers; returns (seed, offset) for a stateless kernel."""
        name = name or self._active
        if name not in self._seeds:
            if _DEFAULT not in self._seeds:
                self._seeds[_DEFAULT] = int(torch.initial_seed()) & 0x7FFFFFFFFFFFFFFF
                self._philox_offsets[_DEFAULT] = 0
            name = _DEFAULT
        off = self._philox_offsets[name]
        # each Philox call yields 4 x 32 bit; keep offsets in units of 4 numbers
        self._philox_offsets[name] = off + (numel + 3) // 4
        return self._seeds[name], off


_TRACKER = RNGStatesTracker()


def get_rng_state_tracker() -> RNGStatesTracker:
    return _TRACKER
