"""MoE routing bookkeeping on the device (no custom-op dependency, no per-layer host sync except the one that
sizes the variable all-to-all in the NCCL path; ``assign_pos`` — NCCL path only — reads the drop count).

Equivalents of Paddle's MoE helper ops used by the reference (moe/utils.py:23,93-126):
``_number_count`` -> ``number_count``; ``_assign_pos`` -> ``assign_pos``; ``_limit_by_capacity`` ->
``limit_by_capacity_counts``; ``_prune_gate_by_capacity`` -> ``prune_gate_by_capacity``; ``_random_routing``.
An expert index of -1 means "dropped".
"""
from __future__ import annotations


import torch
import torch.distributed as dist


def _pg(group):
    return None if group is None else group.process_group


def _world(group) -> int:
    return 1 if group is None else group.nranks


def number_count(gate_idx: torch.Tensor, tot_expert: int) -> torch.Tensor:
    """Slots per expert, dropped slots (-1) ignored.  No boolean indexing and no ``bincount``: both make the host wait for the device
    (8 GPUs, 24 MoE layers: the gates' three waits per layer left the GPUs at 330 W of 1000)."""
    flat = gate_idx.reshape(-1)
    counts = torch.zeros(tot_expert + 1, dtype=torch.int64, device=flat.device)
    counts.scatter_add_(0, flat.to(torch.int64) + 1, torch.ones_like(flat, dtype=torch.int64))      # -1 lands in slot 0
    return counts[1:]


def assign_pos(gate_idx: torch.Tensor) -> torch.Tensor:
    """Indices of the (token, k) slots sorted by expert id, dropped slots removed: ``x[pos // k]`` gives the rows in
    expert-major order (what ``_assign_pos`` + ``index_select`` produce)."""
    flat = gate_idx.reshape(-1)
    order = torch.argsort(flat, stable=True)
    n_drop = int((flat < 0).sum()) if flat.numel() else 0
    return order[n_drop:]


def alltoall_counts(local_counts: torch.Tensor, group) -> torch.Tensor:
    """[world * E_local] counts: entry (r, e) on rank s = how many of s's tokens go to expert e of rank r; after the
    exchange rank r holds for every source s the count for its own experts."""
    if _world(group) == 1 or _pg(group) is None:
        return local_counts.clone()
    out = torch.empty_like(local_counts)
    if local_counts.is_cuda:
        dist.all_to_all_single(out, local_counts.contiguous(), group=_pg(group))
    else:   # gloo lacks all_to_all_single for some dtypes/sizes: emulate with all_gather
        w = _world(group)
        gathered = [torch.empty_like(local_counts) for _ in range(w)]
        dist.all_gather(gathered, local_counts.contiguous(), group=_pg(group))
        e = local_counts.numel() // w
        out = torch.cat([g[group.rank * e:(group.rank + 1) * e] for g in gathered])
    return out


def count_by_gate(gate_idx: torch.Tensor, num_expert: int, world_size: int, require_pos: bool = True, group=None):
    with torch.no_grad():
        lec = number_count(gate_idx, num_expert * world_size)
        gec = alltoall_counts(lec, group) if world_size > 1 else lec
        pos = assign_pos(gate_idx) if require_pos else None
    return pos, lec, gec


def prepare_forward(gate, num_expert: int, world_size: int, moe_group=None):
    """``(pos, local_expert_count, global_expert_count, fwd_expert_count, fwd_batch_size)`` for a top-k gate index tensor (reference
    moe/utils.py:26-38).  ``fwd_batch_size`` is a Python int, i.e. this call synchronises with the device — the layer's default path does not
    use it (routing tables stay in device memory, moe/fused_dispatch.py); it serves the NCCL fallback and external callers."""
    pos, lec, gec = count_by_gate(gate, num_expert, world_size, group=moe_group)
    with torch.no_grad():
        fwd_expert_count = gec.view(world_size, num_expert).sum(0)
        fwd_batch_size = int(fwd_expert_count.sum().item())
    return pos, lec, gec, fwd_expert_count, fwd_batch_size


def limit_by_capacity_counts(gec: torch.Tensor, capacity: torch.Tensor, world_size: int) -> torch.Tensor:
    """gec: [world * E_local] incoming counts ordered (source rank, expert).  Each expert accepts at most ``capacity[e]``
    tokens, granted to source ranks in rank order."""
    e = capacity.numel()
    g = gec.view(world_size, e)
    before = torch.cumsum(g, 0) - g
    allowed = (capacity.unsqueeze(0) - before).clamp(min=0)
    return torch.minimum(g, allowed).reshape(-1)


def prune_gate_by_capacity(gate_idx: torch.Tensor, new_lec: torch.Tensor, num_expert: int, world_size: int) -> torch.Tensor:
    """Keep, per expert, only the first ``new_lec[e]`` slots (token order); the rest become -1."""
    flat = gate_idx.reshape(-1)
    tot = num_expert * world_size
    valid = flat >= 0
    onehot = torch.zeros(flat.numel(), tot, dtype=torch.int64, device=flat.device)
    onehot.scatter_(1, flat.clamp(min=0).unsqueeze(1), valid.to(torch.int64).unsqueeze(1))       # (no boolean indexing: it syncs)
    rank_in_expert = (torch.cumsum(onehot, 0) - onehot).gather(1, flat.clamp(min=0).unsqueeze(1)).squeeze(1)
    keep = valid & (rank_in_expert < new_lec[flat.clamp(min=0)])
    return torch.where(keep, flat, torch.full_like(flat, -1)).view_as(gate_idx)


_CAP_CACHE = {}


def _capacity_tensor(num_expert: int, capacity: int, device) -> torch.Tensor:
    key = (num_expert, int(capacity), str(device))
    if key not in _CAP_CACHE:
        if len(_CAP_CACHE) > 64:
            _CAP_CACHE.clear()
        _CAP_CACHE[key] = torch.full((num_expert,), int(capacity), dtype=torch.int64, device=device)
    return _CAP_CACHE[key]


def limit_by_capacity(topk_idx: torch.Tensor, num_expert: int, world_size: int, capacity: int, group=None, lec: torch.Tensor = None):
    """``lec``: the local per-expert slot counts when the caller has them already (``number_count(topk_idx, num_expert * world_size)``)."""
    with torch.no_grad():
        cap = _capacity_tensor(num_expert, capacity, topk_idx.device)
        if lec is None:
            _, lec, gec = count_by_gate(topk_idx, num_expert, world_size, require_pos=False, group=group)
        else:
            gec = alltoall_counts(lec, group) if world_size > 1 else lec
        new_gec = limit_by_capacity_counts(gec, cap, world_size)
        new_lec = alltoall_counts(new_gec, group) if world_size > 1 else new_gec
        topk_idx = prune_gate_by_capacity(topk_idx, new_lec, num_expert, world_size)
    return new_lec, new_gec, topk_idx


def random_routing(topk_idx: torch.Tensor, topk_val: torch.Tensor, prob: torch.Tensor) -> torch.Tensor:
    """GShard second-expert policy: drop expert #2 of a token when ``2 * gate_2 < u``."""
    out = topk_idx.clone()
    drop = (2.0 * topk_val[:, 1].float()) < prob
    out[:, 1] = torch.where(drop, torch.full_like(out[:, 1], -1), out[:, 1])
    return out
