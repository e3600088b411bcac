"""Capacity-based top-1 / top-2 gating and the einsum dispatch/combine layer (GShard, Lepikhin et al. 2020; Switch, Fedus et
al. 2021) — the algorithm behind the reference's ``moe_exp/sharded_moe.py:119-470``.  The reference ships this path with its
all-to-all commented out ("HACK disable AllToAll"); here the exchange is live: ``[E, C, M]`` buckets are exchanged with an equal-
split all-to-all over the expert-parallel group, so ``ep_size > 1`` really shards experts.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

from ....parallel import comm_ops as C


def einsum(rule: str, a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """The handful of contractions of the einsum-formulated MoE layer (dispatch / combine), each lowered to ONE GEMM or broadcast multiply —
    ``se,sc->sec`` outer product, ``sec,sm->ecm`` dispatch, ``sec,ecm->sm`` combine, ``se,se->s`` row dot, ``s,se->se`` row scale,
    ``ks,ksm->sm`` top-k merge — and ``torch.einsum`` for anything else (reference moe_exp/sharded_moe.py:87-117)."""
    if rule == "s,se->se":
        return a.reshape(a.shape[0], -1) * b
    if rule == "se,sc->sec":
        return a.unsqueeze(2) * b.unsqueeze(1)
    if rule == "se,se->s":
        return (a * b).sum(-1)
    if rule == "sec,sm->ecm":
        s_, e, c = a.shape
        return torch.matmul(a.reshape(s_, e * c).t(), b).reshape(e, c, b.shape[1])
    if rule == "sec,ecm->sm":
        return torch.matmul(a.reshape(a.shape[0], -1), b.reshape(-1, b.shape[-1]))
    if rule == "ks,ksm->sm":
        return torch.bmm(a.t().unsqueeze(1), b.permute(1, 0, 2)).squeeze(1)
    return torch.einsum(rule, a, b)


def multiplicative_jitter(x: torch.Tensor, epsilon: float = 1e-2) -> torch.Tensor:
    if epsilon == 0:
        return x
    return x * torch.empty_like(x).uniform_(1.0 - epsilon, 1.0 + epsilon)


def gumbel_rsample(shape, device=None, dtype=torch.float32) -> torch.Tensor:
    u = torch.rand(shape, device=device, dtype=dtype).clamp_(1e-9, 1.0 - 1e-9)
    return -torch.log(-torch.log(u))


class _AllToAll(torch.autograd.Function):
    """Equal-split all-to-all along dim 0; its own transpose in backward."""

    @staticmethod
    def forward(ctx, group, x):
        ctx.group = group
        return _all_to_all(x.contiguous(), group)

    @staticmethod
    def backward(ctx, g):
        return None, _all_to_all(g.contiguous(), ctx.group)


def _all_to_all(x: torch.Tensor, group) -> torch.Tensor:
    world = C.group_size(group)
    if world == 1:
        return x
    out = torch.empty_like(x)
    if x.is_cuda:
        dist.all_to_all_single(out, x, group=group.process_group)
        return out
    ins = list(x.chunk(world, 0))                 # gloo: gather everybody's chunks, keep the ones addressed to me
    gathered = [torch.empty_like(x) for _ in range(world)]
    dist.all_gather(gathered, x, group=group.process_group)
    return torch.cat([g.chunk(world, 0)[group.rank] for g in gathered], 0) if ins else out


def _capacity(num_tokens: int, num_experts: int, capacity_factor: float, min_capacity: int) -> int:
    return max(int(math.ceil(num_tokens / num_experts * capacity_factor)), int(min_capacity))


def _one_hot(idx: torch.Tensor, n: int) -> torch.Tensor:
    return F.one_hot(idx, n).to(torch.float32)


def top1gating(logits: torch.Tensor, capacity_factor: float, min_capacity: int, used_token: Optional[torch.Tensor] = None,
               noisy_gate_policy: Optional[str] = None, drop_tokens: bool = True, use_rts: bool = True):
    """returns (l_aux, combine_weights [S,E,C], dispatch_mask [S,E,C] bool, exp_counts [E])."""
    if noisy_gate_policy == "RSample":
        noisy = logits + gumbel_rsample(logits.shape, logits.device, logits.dtype)
    gates = F.softmax(logits, dim=1)
    S, E = gates.shape
    capacity = _capacity(S, E, capacity_factor, min_capacity)
    idx = torch.argmax(noisy if noisy_gate_policy == "RSample" else gates, dim=1)
    mask1 = _one_hot(idx, E)
    if used_token is not None:
        mask1 = mask1 * used_token.reshape(-1, 1).to(mask1.dtype)
    exp_counts = mask1.sum(0).detach().to(torch.int64)
    if not drop_tokens:                                    # capacity grows to the busiest expert (synchronised over the world)
        cap_t = exp_counts.max()
        if dist.is_initialized():
            dist.all_reduce(cap_t, op=dist.ReduceOp.MAX)
        capacity = int(cap_t)
    me, ce = gates.mean(0), mask1.mean(0)
    l_aux = (me * ce).sum() * E
    if use_rts:                                            # random token selection: which tokens keep their slot is randomised
        prio = mask1 * torch.rand_like(mask1)
    else:
        prio = mask1
    keep_idx = torch.topk(prio, k=min(capacity, S), dim=0).indices          # [C, E] tokens kept per expert
    keep = torch.zeros_like(mask1).scatter_(0, keep_idx, 1.0)
    mask1 = mask1 * keep
    loc = torch.cumsum(mask1, 0) - 1                                        # slot of each kept token inside its expert
    loc_s = (loc * mask1).sum(1).to(torch.int64)
    gates1 = (gates * mask1).sum(1)
    loc_sc = _one_hot(loc_s.clamp(min=0), capacity)
    combine = (gates1.unsqueeze(1) * mask1).unsqueeze(2) * loc_sc.unsqueeze(1)
    return l_aux, combine, combine.bool(), exp_counts


def top2gating(logits: torch.Tensor, capacity_factor: float, min_capacity: int):
    gates = F.softmax(logits, dim=1)
    S, E = gates.shape
    capacity = _capacity(S, E, 2.0 * capacity_factor, min_capacity)
    idx1 = torch.argmax(gates, dim=1)
    mask1 = _one_hot(idx1, E)
    noisy = logits + gumbel_rsample(logits.shape, logits.device, logits.dtype)      # second expert sampled ~ gumbel-max
    idx2 = torch.argmax(noisy.masked_fill(mask1.bool(), float("-inf")), dim=1)
    mask2 = _one_hot(idx2, E)
    loc1 = torch.cumsum(mask1, 0) - 1
    loc2 = torch.cumsum(mask2, 0) - 1 + mask1.sum(0, keepdim=True)                  # second choices queue behind first choices
    exp_counts = mask1.sum(0).detach().to(torch.int64)
    me, ce = gates.mean(0), mask1.mean(0)
    l_aux = (me * ce).mean() * E * E
    mask1 = mask1 * (loc1 < capacity)
    mask2 = mask2 * (loc2 < capacity)
    loc1_s = (loc1 * mask1).sum(1).to(torch.int64)
    loc2_s = (loc2 * mask2).sum(1).to(torch.int64)
    g1 = (gates * mask1).sum(1)
    g2 = (gates * mask2).sum(1)
    denom = (g1 + g2).clamp(min=torch.finfo(gates.dtype).eps)
    g1, g2 = g1 / denom, g2 / denom
    c1 = (g1.unsqueeze(1) * mask1).unsqueeze(2) * _one_hot(loc1_s.clamp(min=0), capacity).unsqueeze(1)
    c2 = (g2.unsqueeze(1) * mask2).unsqueeze(2) * _one_hot(loc2_s.clamp(min=0), capacity).unsqueeze(1)
    combine = c1 + c2
    return l_aux, combine, combine.bool(), exp_counts


class TopKGate(nn.Module):
    """fp32 router ``wg`` + top-1 / top-2 capacity gating (reference sharded_moe.py:300-376)."""

    def __init__(self, model_dim: int, num_experts: int, k: int = 1, capacity_factor: float = 1.0, eval_capacity_factor: float = 1.0,
                 min_capacity: int = 8, noisy_gate_policy: Optional[str] = None, drop_tokens: bool = True, use_rts: bool = True):
        super().__init__()
        if k not in (1, 2):
            raise ValueError("Only top-1 and top-2 gatings are supported.")
        self.wg = nn.Linear(model_dim, num_experts, bias=False).float()
        self.k = k
        self.capacity_factor, self.eval_capacity_factor, self.min_capacity = capacity_factor, eval_capacity_factor, min_capacity
        self.noisy_gate_policy, self.drop_tokens, self.use_rts = noisy_gate_policy, drop_tokens, use_rts

    def forward(self, x: torch.Tensor, used_token: Optional[torch.Tensor] = None):
        xf = x.float()
        if self.noisy_gate_policy == "Jitter" and self.training:
            xf = multiplicative_jitter(xf)
        logits = F.linear(xf, self.wg.weight.float())
        cf = self.capacity_factor if self.training else self.eval_capacity_factor
        if self.k == 1:
            return top1gating(logits, cf, self.min_capacity, used_token, self.noisy_gate_policy if self.training else None,
                              self.drop_tokens, self.use_rts)
        return top2gating(logits, cf, self.min_capacity)


class MOELayer(nn.Module):
    """dispatch (einsum) -> all-to-all -> local experts -> all-to-all -> combine (einsum)."""

    def __init__(self, gate: nn.Module, experts: nn.Module, ep_group_name, ep_size: int, num_local_experts: int):
        super().__init__()
        self.gate, self.experts = gate, experts
        self.ep_group = None
        self.ep_size, self.ep_group_name, self.num_local_experts = ep_size, ep_group_name, num_local_experts
        self.l_aux = torch.zeros(())
        self.exp_counts = None

    def _set_ep_group(self, ep_group) -> None:
        self.ep_group = ep_group

    def get_loss(self) -> torch.Tensor:
        return self.l_aux

    def forward(self, x: torch.Tensor, used_token: Optional[torch.Tensor] = None) -> torch.Tensor:
        d_model = x.shape[-1]
        tokens = x.reshape(-1, d_model)
        self.l_aux, combine, dispatch, self.exp_counts = self.gate(tokens, used_token)
        dispatched = torch.einsum("sec,sm->ecm", dispatch.to(tokens.dtype), tokens)            # [E, C, M]
        if self.ep_size > 1:
            dispatched = _AllToAll.apply(self.ep_group, dispatched)
        dispatched = dispatched.reshape(self.ep_size, self.num_local_experts, -1, d_model)      # [ep, E_local, C, M]
        out = self.experts(dispatched)
        out = out.reshape(self.ep_size * self.num_local_experts, -1, d_model)
        if self.ep_size > 1:
            out = _AllToAll.apply(self.ep_group, out)
        combined = torch.einsum("sec,ecm->sm", combine.to(tokens.dtype), out)
        return combined.reshape(x.shape)
