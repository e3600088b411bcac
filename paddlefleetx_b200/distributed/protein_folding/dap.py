"""Dynamic Axial Parallelism (FastFold-style) primitives — reference distributed/protein_folding/dap.py:37-426.

MSA / pair activations are sharded along one of their two "sequence" axes across the ``dap`` group; ``row_to_col`` /
``col_to_row`` transpose which axis is sharded with an **all-to-all** (the Ulysses-like operation of this code base);
``scatter`` / ``gather`` move between sharded and replicated layouts; gradients of replicated parameters are all-reduced over the
group.  Every op is an autograd function with the conjugate collective in backward.

Two call styles:

* one-shot: ``gather_full(x, axis)`` / ``exchange(x, in_axis, out_axis)`` / ``row_to_col`` / ``col_to_row`` return the finished tensor — what
  the model code here uses;
* the reference's split-phase ("duality") pairs: ``y = all_gather(x, axis)`` STARTS the collective and returns the rank-major stack
  (``[n * d0, ...]``), independent work may follow, ``z = all_gather_opp(y, axis)`` finishes it (waits, moves the stack to ``axis``, and
  carries the backward collective); likewise ``all_to_all`` / ``all_to_all_opp``.  The reference declares the asynchronous mode but leaves
  the wait unimplemented (dap.py:81-82,181-183); here ``set_dap_sync_op(False)`` really launches the collective asynchronously on the
  communicator's stream and the ``*_opp`` call is where the compute stream waits for it.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from ...parallel import comm_ops as C
from .scg import scg

_SYNC_OP = True


def set_dap_sync_op(sync_op: bool) -> None:
    """``False``: ``all_gather`` / ``all_to_all`` return immediately and their ``*_opp`` partner waits (communication overlaps whatever runs in
    between).  ``True`` (default): the first call of the pair already blocks."""
    global _SYNC_OP
    _SYNC_OP = bool(sync_op)


def get_dap_sync_op() -> bool:
    return _SYNC_OP


def get_world_size() -> int:
    return scg.get_dap_world_size()


def get_rank_in_group() -> int:
    return scg.get_dap_rank()


def ensure_divisibility(numerator: int, denominator: int) -> None:
    assert numerator % denominator == 0, f"{numerator} is not divisible by {denominator}"


def divide(numerator: int, denominator: int) -> int:
    ensure_divisibility(numerator, denominator)
    return numerator // denominator


def _grp(group=None):
    return group if group is not None else scg.get_dap_group()


def _all_to_all(x: torch.Tensor, group, scatter_dim: int, gather_dim: int) -> torch.Tensor:
    """Split ``scatter_dim`` over the group, concatenate received pieces along ``gather_dim``."""
    w = C.group_size(group)
    if w == 1 or group.process_group is None:
        return x
    parts = [p.contiguous() for p in x.chunk(w, dim=scatter_dim)]
    out = [torch.empty_like(parts[0]) for _ in range(w)]
    if x.is_cuda:
        dist.all_to_all(out, parts, group=group.process_group)
    else:   # gloo: emulate
        reqs = [dist.isend(parts[r], group.ranks[r], group=group.process_group) for r in range(w) if r != group.rank]
        for r in range(w):
            if r == group.rank:
                out[r].copy_(parts[r])
            else:
                dist.recv(out[r], group.ranks[r], group=group.process_group)
        for q in reqs:
            q.wait()
    return torch.cat(out, dim=gather_dim)


class _Scatter(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, dim, group):
        ctx.dim, ctx.group = dim, group
        return C.split_dim(x, group, dim)

    @staticmethod
    def backward(ctx, g):
        return C.all_gather_dim(g.contiguous(), ctx.group, ctx.dim), None, None


class _Gather(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, dim, group):
        ctx.dim, ctx.group = dim, group
        return C.all_gather_dim(x.contiguous(), group, dim)

    @staticmethod
    def backward(ctx, g):
        return C.split_dim(g, ctx.group, ctx.dim), None, None


class _AllGatherSum(torch.autograd.Function):
    """all-gather whose backward is reduce-scatter (used when the gathered tensor feeds a computation every rank repeats)."""

    @staticmethod
    def forward(ctx, x, dim, group):
        ctx.dim, ctx.group = dim, group
        return C.all_gather_dim(x.contiguous(), group, dim)

    @staticmethod
    def backward(ctx, g):
        grp = ctx.group
        if C.group_size(grp) == 1 or grp.process_group is None:
            return g, None, None
        g = g.contiguous().clone()
        dist.all_reduce(g, group=grp.process_group)
        return C.split_dim(g, grp, ctx.dim), None, None


class _AllToAll(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, scatter_dim, gather_dim, group):
        ctx.args = (scatter_dim, gather_dim, group)
        return _all_to_all(x, group, scatter_dim, gather_dim)

    @staticmethod
    def backward(ctx, g):
        s, ga, grp = ctx.args
        return _all_to_all(g.contiguous(), grp, ga, s), None, None, None


def scatter(x, axis: int = 0, group=None):
    g = _grp(group)
    return x if C.group_size(g) == 1 else _Scatter.apply(x, axis, g)


def gather(x, axis: int = 0, group=None):
    g = _grp(group)
    return x if C.group_size(g) == 1 else _Gather.apply(x, axis, g)


def gather_full(x, axis: int = 0, group=None):
    """One-shot all-gather along ``axis`` whose backward is a reduce-scatter (the gathered tensor feeds work every rank repeats)."""
    g = _grp(group)
    return x if C.group_size(g) == 1 else _AllGatherSum.apply(x, axis, g)


def exchange(x, in_axis: int, out_axis: int, group=None):
    """One-shot all-to-all: split ``in_axis`` over the group, concatenate what arrives along ``out_axis``."""
    g = _grp(group)
    return x if C.group_size(g) == 1 else _AllToAll.apply(x, in_axis, out_axis, g)


def row_to_col(x, group=None):
    """[.., R/n, C, ..] (rows sharded, dims 1/2) -> [.., R, C/n, ..] (columns sharded)."""
    return exchange(x, 2, 1, group)


def col_to_row(x, group=None):
    return exchange(x, 1, 2, group)


# ---------------------------------------------------------------- split-phase pairs (reference call style)
_PENDING = []          # (stacked output tensor, work handle) of collectives started and not yet finished by their *_opp call


def _start(out: torch.Tensor, work):
    if work is not None:
        if _SYNC_OP:
            work.wait()
        else:
            _PENDING.append((out, work))
    return out


def _finish(t: torch.Tensor) -> None:
    """Wait for the collective that produces ``t`` (matched by storage); a tensor we cannot match waits for everything outstanding."""
    if not _PENDING:
        return
    ptr = t.untyped_storage().data_ptr()
    mine = [i for i, (o, _) in enumerate(_PENDING) if o.untyped_storage().data_ptr() == ptr]
    for i in (mine or range(len(_PENDING))):
        _PENDING[i][1].wait()
    keep = [] if not mine else [e for i, e in enumerate(_PENDING) if i not in mine]
    _PENDING[:] = keep


def _stack_all_gather(x: torch.Tensor, group):
    x = x.contiguous()
    out = torch.empty((group.nranks * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    return _start(out, dist.all_gather_into_tensor(out, x, group=group.process_group, async_op=True))


def _stack_all_to_all(x: torch.Tensor, group):
    """``x`` is a rank-major stack along dim 0: chunk r goes to rank r, chunk r of the result came from rank r."""
    x = x.contiguous()
    out = torch.empty_like(x)
    if x.is_cuda:
        return _start(out, dist.all_to_all_single(out, x, group=group.process_group, async_op=True))
    w = group.nranks                      # gloo has no all-to-all: pairwise exchange, complete on return
    src, dst = x.chunk(w, 0), out.chunk(w, 0)
    reqs = [dist.isend(src[r].contiguous(), group.ranks[r], group=group.process_group) for r in range(w) if r != group.rank]
    for r in range(w):
        if r == group.rank:
            dst[r].copy_(src[r])
        else:
            buf = torch.empty_like(src[r])
            dist.recv(buf, group.ranks[r], group=group.process_group)
            dst[r].copy_(buf)
    for q in reqs:
        q.wait()
    return out


def _to_stack(x, axis, n):
    return x if axis == 0 else torch.cat(x.chunk(n, dim=axis), dim=0)


def _from_stack(x, axis, n):
    return x if axis == 0 else torch.cat(x.chunk(n, dim=0), dim=axis)


def _reduce_scatter_stack(g: torch.Tensor, group):
    g = g.contiguous()
    out = torch.empty((g.shape[0] // group.nranks,) + tuple(g.shape[1:]), dtype=g.dtype, device=g.device)
    if g.is_cuda:
        dist.reduce_scatter_tensor(out, g, group=group.process_group)
    else:
        g = g.clone()
        dist.all_reduce(g, group=group.process_group)
        out.copy_(g.chunk(group.nranks, 0)[group.rank])
    return out


class Scatter(torch.autograd.Function):
    """Keep this rank's slice of ``axis``; backward all-gathers (reference dap.py:106-116)."""

    @staticmethod
    def forward(ctx, input, axis=-1):
        ctx.axis = axis
        return C.split_dim(input, _grp(), axis)

    @staticmethod
    def backward(ctx, grad_output):
        return C.all_gather_dim(grad_output.contiguous(), _grp(), ctx.axis), None


class Gather(torch.autograd.Function):
    """All-gather along ``axis``; backward keeps this rank's slice (reference dap.py:131-141)."""

    @staticmethod
    def forward(ctx, input, axis=-1):
        ctx.axis = axis
        return C.all_gather_dim(input.contiguous(), _grp(), axis)

    @staticmethod
    def backward(ctx, grad_output):
        return C.split_dim(grad_output, _grp(), ctx.axis), None


class AllGather(torch.autograd.Function):
    """First half of the all-gather pair: starts the collective and returns the rank-major stack.  Its backward is the reduce-scatter of the
    stack's gradient (the reference puts that in ``AllGather_Opp.backward`` and passes a wrongly shaped tensor through here, which Paddle's
    PyLayer tolerates; autograd checks shapes, and the composite of the pair is the same)."""

    @staticmethod
    def forward(ctx, input, axis=-1, sync_op=True):
        return _stack_all_gather(input, _grp())

    @staticmethod
    def backward(ctx, grad_output):
        return _reduce_scatter_stack(grad_output, _grp()), None, None


class AllGather_Opp(torch.autograd.Function):
    """Second half: marks the point where the gathered stack is consumed (identity in both directions; the wait happens before it)."""

    @staticmethod
    def forward(ctx, input, axis=-1, sync_op=True):
        return input.view_as(input)

    @staticmethod
    def backward(ctx, grad_output):
        return grad_output, None, None


class All_to_All(torch.autograd.Function):
    """First half of the all-to-all pair.  In backward the roles swap: ``All_to_All_Opp.backward`` starts the conjugate all-to-all and this
    function is where it is waited for."""

    @staticmethod
    def forward(ctx, input, in_axis=-1, out_axis=-1, sync_op=True):
        return _stack_all_to_all(input, _grp())

    @staticmethod
    def backward(ctx, grad_output):
        _finish(grad_output)
        return grad_output, None, None, None


class All_to_All_Opp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, output, in_axis=-1, out_axis=-1, sync_op=True):
        return output.view_as(output)

    @staticmethod
    def backward(ctx, grad_output):
        return _stack_all_to_all(grad_output, _grp()), None, None, None


class All2All(torch.autograd.Function):
    """Synchronous all-to-all on a rank-major stack, conjugate all-to-all in backward (reference dap.py:345-355)."""

    @staticmethod
    def forward(ctx, input, in_axis=-1, out_axis=-1):
        out = _stack_all_to_all(input, _grp())
        _finish(out)
        return out

    @staticmethod
    def backward(ctx, grad_output):
        out = _stack_all_to_all(grad_output, _grp())
        _finish(out)
        return out, None, None


def all_gather(input, axis: int = -1):
    """START gathering ``input`` from every rank of the dap group; returns the rank-major stack ``[n * d0, ...]`` whatever ``axis`` is.  Finish
    with ``all_gather_opp(result, axis)``; in asynchronous mode do not overwrite ``input`` before that (reference dap.py:201-217)."""
    if get_world_size() == 1:
        return input
    if input.requires_grad:
        return AllGather.apply(input, axis, _SYNC_OP)
    return _stack_all_gather(input, _grp())


def all_gather_opp(output, axis: int = -1):
    """FINISH an ``all_gather``: wait for it, attach the reduce-scatter backward, move the stack to ``axis`` (reference dap.py:220-241)."""
    n = get_world_size()
    if n == 1:
        return output
    _finish(output)
    if output.requires_grad:
        output = AllGather_Opp.apply(output, axis, _SYNC_OP)
    return _from_stack(output, axis, n)


def all_to_all(input, in_axis: int, out_axis: int):
    """START an all-to-all that splits ``in_axis``; returns the received chunks as a rank-major stack.  Finish with ``all_to_all_opp``."""
    n = get_world_size()
    if n == 1:
        return input
    ensure_divisibility(input.shape[in_axis], n)
    input = _to_stack(input, in_axis, n)
    if input.requires_grad:
        return All_to_All.apply(input, in_axis, out_axis, _SYNC_OP)
    return _stack_all_to_all(input, _grp())


def all_to_all_opp(output, in_axis: int, out_axis: int):
    """FINISH an ``all_to_all``: wait, attach the conjugate all-to-all as backward, concatenate the stack along ``out_axis``."""
    n = get_world_size()
    if n == 1:
        return output
    _finish(output)
    if output.requires_grad:
        output = All_to_All_Opp.apply(output, in_axis, out_axis, _SYNC_OP)
    ensure_divisibility(output.shape[0], n)
    return _from_stack(output, out_axis, n)


def _is_sharded(prm) -> bool:
    """A tensor-parallel shard (every rank holds a different slice): its gradient must not be summed.  Paddle marks those with a boolean
    ``is_distributed`` attribute; ``torch.Tensor.is_distributed`` is an unrelated built-in method, so only a plain bool counts."""
    flag = getattr(prm, "is_distributed", False)
    return (flag is True) or bool(getattr(prm, "tp_sharded", False))


def grad_sync(params, group=None) -> None:
    """All-reduce (mean-free SUM) the gradients of parameters replicated across the dap group (dap.py tail).  Accepts a parameter list or the
    reference's optimizer ``param_groups`` (only groups flagged ``dap: True`` take part, tensors marked ``is_distributed`` are skipped)."""
    params = list(params)
    if params and isinstance(params[0], dict):
        params = [q for grp in params if grp.get("dap", False) for q in grp["params"] if not _is_sharded(q)]
    C.fused_allreduce_gradients(params, _grp(group), scale=1.0)
