"""Expert-parallel data movement with autograd: ``MoEScatter`` / ``MoEGather`` (variable-size all-to-all, the
reference's ``global_scatter`` / ``global_gather``), ``AllGather`` / ``Slice`` for the mp axis
(reference moe/comm_ops.py:28-160).  The fixed-capacity, host-sync-free P2P path lives in ``fused_dispatch.py``."""
from __future__ import annotations

import torch
import torch.distributed as dist

from ....parallel import comm_ops as C


def _split_sizes(local_count: torch.Tensor, global_count: torch.Tensor, world: int):
    e = local_count.numel() // world
    send = local_count.view(world, e).sum(1).tolist()       # host sync: sizes of the variable all-to-all
    recv = global_count.view(world, e).sum(1).tolist()
    return [int(s) for s in send], [int(r) for r in recv]


def _reorder_recv(x: torch.Tensor, global_count: torch.Tensor, world: int, inverse: bool = False) -> torch.Tensor:
    """all_to_all delivers rows grouped (source rank, expert); experts want (expert, source rank)."""
    e = global_count.numel() // world
    if world == 1 or e == 1 or x.shape[0] == 0:
        return x
    cnt = global_count.view(world, e)
    src_major = torch.arange(world * e, device=x.device).view(world, e)          # block id in (source, expert) order
    exp_major = src_major.t().reshape(-1)                                          # same blocks listed expert-major
    sizes = cnt.reshape(-1)
    starts = torch.cumsum(sizes, 0) - sizes
    sel_sizes = sizes[exp_major]
    total = int(x.shape[0])
    blk = torch.repeat_interleave(torch.arange(exp_major.numel(), device=x.device), sel_sizes, output_size=total)
    within = torch.arange(total, device=x.device) - torch.repeat_interleave(torch.cumsum(sel_sizes, 0) - sel_sizes, sel_sizes, output_size=total)
    idx = starts[exp_major][blk] + within                    # idx[j] = row (in source-major order) of expert-major row j
    if inverse:
        out = torch.empty_like(x)
        out[idx] = x
        return out
    return x[idx]


def global_scatter(x: torch.Tensor, local_count: torch.Tensor, global_count: torch.Tensor, group=None) -> torch.Tensor:
    world = C.group_size(group)
    if world == 1 or group.process_group is None:
        return x
    send, recv = _split_sizes(local_count, global_count, world)
    out = torch.empty((sum(recv), x.shape[1]), dtype=x.dtype, device=x.device)
    _all_to_all(out, x.contiguous(), recv, send, group)
    return _reorder_recv(out, global_count, world)


def global_gather(x: torch.Tensor, local_count: torch.Tensor, global_count: torch.Tensor, group=None) -> torch.Tensor:
    world = C.group_size(group)
    if world == 1 or group.process_group is None:
        return x
    send, recv = _split_sizes(local_count, global_count, world)
    x = _reorder_recv(x, global_count, world, inverse=True)
    out = torch.empty((sum(send), x.shape[1]), dtype=x.dtype, device=x.device)
    _all_to_all(out, x.contiguous(), send, recv, group)
    return out


def _all_to_all(out, inp, out_splits, in_splits, group):
    if inp.is_cuda:
        dist.all_to_all_single(out, inp, out_splits, in_splits, group=group.process_group)
        return
    # gloo: emulate with point-to-point
    world, me = group.nranks, group.rank
    ins = list(inp.split(in_splits)) if inp.shape[0] else [inp[:0]] * world
    outs = list(out.split(out_splits)) if out.shape[0] else [out[:0]] * world
    reqs = []
    for r in range(world):
        if r == me:
            outs[r].copy_(ins[r])
            continue
        if in_splits[r]:
            reqs.append(dist.isend(ins[r].contiguous(), group.ranks[r], group=group.process_group))
    for r in range(world):
        if r != me and out_splits[r]:
            buf = torch.empty_like(outs[r])
            dist.recv(buf, group.ranks[r], group=group.process_group)
            outs[r].copy_(buf)
    for q in reqs:
        q.wait()


class MoEScatter(torch.autograd.Function):
    """local gather by ``pos`` + global scatter to the expert owners."""

    @staticmethod
    def forward(ctx, inp, pos, local_count, global_count, fwd_batch_size, world_size, group, topk):
        buf = inp.index_select(0, pos // topk)
        out = global_scatter(buf, local_count, global_count, group) if world_size > 1 else buf
        ctx.moe_args = (inp.shape[0], world_size, group, topk)
        ctx.save_for_backward(pos, local_count, global_count)
        return out

    @staticmethod
    def backward(ctx, g):
        pos, lc, gc = ctx.saved_tensors
        n_in, world, group, topk = ctx.moe_args
        buf = global_gather(g.contiguous(), lc, gc, group) if world > 1 else g
        gi = torch.zeros(n_in, g.shape[1], dtype=torch.float32, device=g.device)
        gi.index_add_(0, pos // topk, buf.float())
        return gi.to(g.dtype), None, None, None, None, None, None, None


class MoEGather(torch.autograd.Function):
    """global gather back to the token owners + scatter into [tokens * k, h] slot order (fp32 scatter as in the
    reference, moe/utils.py:61-75)."""

    @staticmethod
    def forward(ctx, x, pos, local_count, global_count, out_batch_size, world_size, group):
        buf = global_gather(x.contiguous(), local_count, global_count, group) if world_size > 1 else x
        out = torch.zeros(out_batch_size, x.shape[1], dtype=torch.float32, device=x.device)
        out[pos] = buf.float()
        ctx.moe_args = (world_size, group)
        ctx.save_for_backward(pos, local_count, global_count)
        return out.to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        pos, lc, gc = ctx.saved_tensors
        world, group = ctx.moe_args
        buf = g.index_select(0, pos)
        out = global_scatter(buf, lc, gc, group) if world > 1 else buf
        return out, None, None, None, None, None, None


class AllGather(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, rank, world_size, group):
        ctx.args = (rank, world_size, x.shape[0])
        return C.all_gather_dim0(x, group)

    @staticmethod
    def backward(ctx, g):
        rank, world, n = ctx.args
        return g[rank * n:(rank + 1) * n], None, None, None


class Slice(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, rank, world_size, group):
        n = x.shape[0] // world_size
        ctx.group = group
        return x[rank * n:(rank + 1) * n].contiguous()

    @staticmethod
    def backward(ctx, g):
        return C.all_gather_dim0(g.contiguous(), ctx.group), None, None, None
