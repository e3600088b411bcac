"""Activation recomputation that preserves every RNG stream of the tracker.

Contract of the reference's ``fleet.utils.recompute`` / ``recompute_hybrid`` (call sites
hybrid_model.py:378,457,638; moe_layer.py:213-217): run ``fn`` without saving activations, re-run it during
backward under the *same* dropout randomness and autocast state.

Besides torch's default generator states we snapshot the tracker's named streams (global/local seed) and
its Philox offsets, so both library dropout and our counter-based dropout kernels replay identically.
``recompute_hybrid`` additionally supports offloading the saved inputs to pinned host memory and
partitioning them across the mp group (``recompute_ctx{mp_group, offload, partition}``).
"""
from __future__ import annotations

from typing import Any, Callable, Optional

import torch
import torch.distributed as dist

from .rng import get_rng_state_tracker


class _RecomputeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, fn, ctx_opts, n_tensor, *args):
        ctx.fn = fn
        ctx.opts = ctx_opts or {}
        tracker = get_rng_state_tracker()
        ctx.cpu_rng = torch.get_rng_state()
        ctx.cuda_rng = torch.cuda.get_rng_state() if torch.cuda.is_available() and torch.cuda.is_initialized() else None
        ctx.tracker_state = tracker.get_states_tracker()
        ctx.autocast = (torch.is_autocast_enabled(), torch.get_autocast_gpu_dtype()) if torch.cuda.is_available() else (False, None)
        tensors, meta = [], []
        for a in args:
            if isinstance(a, torch.Tensor):
                meta.append(("t", len(tensors), a.requires_grad))
                tensors.append(_stash(a, ctx.opts))
            else:
                meta.append(("o", a, False))
        ctx.meta = meta
        ctx.stash = tensors
        with torch.no_grad():
            out = fn(*args)
        return out

    @staticmethod
    def backward(ctx, *grads):
        args = []
        for kind, val, rg in ctx.meta:
            if kind == "t":
                t = _unstash(ctx.stash[val], ctx.opts).detach()
                t.requires_grad_(rg)
                args.append(t)
            else:
                args.append(val)
        tracker = get_rng_state_tracker()
        cur_cpu = torch.get_rng_state()
        cur_cuda = torch.cuda.get_rng_state() if ctx.cuda_rng is not None else None
        cur_tracker = tracker.get_states_tracker()
        torch.set_rng_state(ctx.cpu_rng)
        if ctx.cuda_rng is not None:
            torch.cuda.set_rng_state(ctx.cuda_rng)
        tracker.set_states_tracker(ctx.tracker_state)
        try:
            with torch.enable_grad():
                if ctx.autocast[0]:
                    with torch.autocast("cuda", dtype=ctx.autocast[1]):
                        out = ctx.fn(*args)
                else:
                    out = ctx.fn(*args)
        finally:
            torch.set_rng_state(cur_cpu)
            if cur_cuda is not None:
                torch.cuda.set_rng_state(cur_cuda)
            tracker.set_states_tracker(cur_tracker)
        outs = out if isinstance(out, (tuple, list)) else (out,)
        pairs = [(o, g) for o, g in zip(outs, grads) if isinstance(o, torch.Tensor) and o.requires_grad and g is not None]
        if pairs:
            torch.autograd.backward([o for o, _ in pairs], [g for _, g in pairs])
        in_grads = [a.grad if isinstance(a, torch.Tensor) and a.requires_grad else None for a in args]
        return (None, None, None, *in_grads)


def _stash(t: torch.Tensor, opts: dict):
    group = opts.get("mp_group")
    if opts.get("partition") and group is not None and group.nranks > 1 and t.numel() % group.nranks == 0 and t.is_floating_point():
        shard = t.detach().reshape(-1).chunk(group.nranks)[group.rank].clone()
        t_small, shape = shard, t.shape
    else:
        t_small, shape = t.detach(), None
    if opts.get("offload") and t_small.is_cuda:
        host = torch.empty(t_small.shape, dtype=t_small.dtype, device="cpu", pin_memory=True)
        host.copy_(t_small, non_blocking=True)
        return ("host", host, shape, t.device)
    return ("dev", t_small, shape, t.device)


def _unstash(entry, opts: dict) -> torch.Tensor:
    kind, t, shape, device = entry
    if kind == "host":
        t = t.to(device, non_blocking=True)
    if shape is not None:
        group = opts["mp_group"]
        full = torch.empty(group.nranks * t.numel(), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(full, t.contiguous(), group=group.process_group)
        t = full.view(shape)
    return t


def recompute(fn: Callable[..., Any], *args, **kwargs):
    """``fleet.utils.recompute(fn, *args)``.  kwargs are bound into the callable."""
    if kwargs:
        inner = fn
        fn = lambda *a: inner(*a, **kwargs)  # noqa: E731
    if not torch.is_grad_enabled():
        return fn(*args)
    return _RecomputeFn.apply(fn, None, 0, *args)


def recompute_hybrid(ctx_opts: Optional[dict], fn: Callable[..., Any], *args):
    """``recompute_hybrid(ctx, fn, *args)`` with ``ctx = {mp_group, offload, partition}``."""
    if not torch.is_grad_enabled():
        return fn(*args)
    return _RecomputeFn.apply(fn, ctx_opts or {}, 0, *args)
