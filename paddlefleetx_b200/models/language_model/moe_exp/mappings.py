"""Token gather / drop across the tensor-parallel group for expert tensor parallelism (reference moe_exp/mappings.py:27-92):
``gather_tokens`` all-gathers along ``axis`` in forward and keeps the local slice of the gradient; ``drop_tokens`` is its
transpose."""
import torch

from ....parallel import comm_ops as C


def _gather(x, group, axis):
    if C.group_size(group) == 1:
        return x
    x = x.transpose(0, axis).contiguous()
    return C.all_gather_dim0(x, group).transpose(0, axis).contiguous()


def _drop(x, group, axis):
    world = C.group_size(group)
    if world == 1:
        return x
    n = x.shape[axis] // world
    return x.narrow(axis, group.rank * n, n).contiguous()


class _GatherTokens(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, group, axis):
        ctx.group, ctx.axis = group, axis
        return _gather(x, group, axis)

    @staticmethod
    def backward(ctx, g):
        return _drop(g, ctx.group, ctx.axis), None, None


class _DropTokens(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, group, axis):
        ctx.group, ctx.axis = group, axis
        return _drop(x, group, axis)

    @staticmethod
    def backward(ctx, g):
        return _gather(g, ctx.group, ctx.axis), None, None


def gather_tokens(x, group=None, axis=0):
    return x if C.group_size(group) == 1 else _GatherTokens.apply(x, group, axis)


def drop_tokens(x, group=None, axis=0):
    return x if C.group_size(group) == 1 else _DropTokens.apply(x, group, axis)
