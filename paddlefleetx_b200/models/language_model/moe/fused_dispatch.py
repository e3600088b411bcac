"""Peer-memory MoE dispatch / combine (``ops/csrc/moe_kernels.cu``) with autograd.

Replaces the reference's ``MoEScatter`` / ``MoEGather`` pair (moe/comm_ops.py:28-160: index_select + variable-size NCCL
all-to-all + host-side split sizes) by two kernels that store / load token rows directly in the expert owners' memory over
NVLink.  One :class:`MoEDispatcher` per expert-parallel group owns the symmetric buffers:

* ``cnt``      int32 [world, world*E_local] — count matrix, row s written by rank s
* ``flags``    uint32 [3, 16] — arrive / done / ready epoch flags
* ``stage_in`` [cap_rows, H]  — rows pushed TO me when nobody needs to keep them (dy in the backward pass, x in eval)
* ``stage_out``[cap_rows, H]  — rows I produced for peers to pull (expert outputs y in forward, dx in backward)
* per-layer receive buffers for x in training: they double as the activation the expert backward needs, so nothing is copied

``cap_rows`` bounds the rows one rank can receive (``capacity_factor`` x the balanced share, rounded up to the alignment); an
overflow sets a device flag that is checked whenever the segment table is read.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from ....ops import _native
from ....parallel.symmetric_memory import get_allocator

ALIGN = 256      # expert blocks start on 256-row boundaries: a 2-CTA (256-row) GEMM tile never straddles two experts


class _LocalMemory:
    """The symmetric-allocator surface for an expert group of ONE rank: the dispatch / combine kernels then "exchange" with themselves, and
    the layer still gets the device-side routing, the expert-major layout and the grouped expert GEMMs (no per-expert loop, no host sync)."""

    def __init__(self, device):
        self.device = device
        self._keep = []

    def alloc_tensor(self, numel: int, dtype: torch.dtype, device=None) -> torch.Tensor:
        t = torch.zeros(numel, dtype=dtype, device=self.device)
        self._keep.append(t)
        return t

    def empty(self, shape, dtype: torch.dtype) -> torch.Tensor:
        n = 1
        for d in shape:
            n *= int(d)
        return self.alloc_tensor(n, dtype).view(*shape)

    def peer_ptrs(self, t: torch.Tensor):
        return [t.data_ptr()]

    def barrier(self, channel: int = 0) -> None:
        return None


class MoEDispatcher:
    def __init__(self, group, hidden: int, e_local: int, dtype: torch.dtype, capacity_factor: float = 2.0, num_ctas: int = 0):
        self.group = group
        self.world, self.rank = (group.nranks, group.rank) if group is not None else (1, 0)
        self.hidden, self.e_local, self.dtype = hidden, e_local, dtype
        self.capacity_factor = capacity_factor
        self.lib = _native.require()
        self.alloc = get_allocator(group) if self.world > 1 else _LocalMemory(torch.device("cuda", torch.cuda.current_device()))
        dev = self.alloc.device
        e_total = self.world * e_local
        self.cnt = self.alloc.alloc_tensor(self.world * e_total, torch.int32)
        self.flags = self.alloc.alloc_tensor(64, torch.int32)
        self.cnt.zero_(); self.flags.zero_()
        self.block_counter = torch.zeros(1, dtype=torch.int32, device=dev)
        self.cap_rows = 0
        self.stage_in: Optional[torch.Tensor] = None
        self.stage_out: Optional[torch.Tensor] = None
        self.layer_bufs: Dict[int, torch.Tensor] = {}
        self.epoch = 0
        # receive-buffer overflow of the sync-free path: a sticky device flag, mirrored to pinned memory asynchronously and examined one
        # step late (the host never waits for the routing result)
        self.overflow = torch.zeros(1, dtype=torch.int32, device=dev)
        self._ovf_host = torch.zeros(1, dtype=torch.int32).pin_memory()
        self._ovf_event = None
        props = torch.cuda.get_device_properties(dev)
        self.num_ctas = num_ctas or 2 * props.multi_processor_count
        torch.cuda.synchronize()
        self.alloc.barrier()

    # -- buffers (collective: every rank calls with the same arguments) ----------------------------
    def ensure_capacity(self, num_slots: int):
        want = int(self.capacity_factor * num_slots) + ALIGN * self.e_local
        want = min(max(want, ALIGN), self.world * num_slots + ALIGN * self.e_local)
        want = (want + ALIGN - 1) // ALIGN * ALIGN
        if want > self.cap_rows:
            assert not self.layer_bufs, "MoE receive capacity must not grow after layers were bound (batch size changed?)"
            self.cap_rows = want
            self.stage_in = self.alloc.empty((want, self.hidden), self.dtype)
            self.stage_out = self.alloc.empty((want, self.hidden), self.dtype)

    def recv_buffer(self, layer_key: Optional[int]) -> torch.Tensor:
        if layer_key is None:
            return self.stage_in
        if layer_key not in self.layer_bufs:
            self.layer_bufs[layer_key] = self.alloc.empty((self.cap_rows, self.hidden), self.dtype)
        return self.layer_bufs[layer_key]

    def poll_overflow(self):
        ev = self._ovf_event
        if ev is not None:
            if not ev.query():
                return
            self._ovf_event = None
            if int(self._ovf_host[0]):
                raise RuntimeError(f"MoE receive buffer overflow (capacity {self.cap_rows} rows per rank): raise moe p2p_capacity_factor")
        self._ovf_host.copy_(self.overflow, non_blocking=True)
        self._ovf_event = torch.cuda.Event()
        self._ovf_event.record()

    def tile_table(self, seg: torch.Tensor):
        return self.lib.moe_tile_table(seg, self.e_local, ALIGN, self.cap_rows, self.overflow)

    # -- kernels ----------------------------------------------------------------------------------
    def route(self, gate_idx: torch.Tensor):
        return self.lib.moe_route(gate_idx.reshape(-1).contiguous(), self.world * self.e_local)

    def dispatch(self, src, scale, gate_idx_flat, slot_rank, counts, recv, src_div):
        self.epoch += 1
        return self.lib.moe_dispatch(src, scale, gate_idx_flat, slot_rank, counts, self.alloc.peer_ptrs(recv), self.alloc.peer_ptrs(self.cnt),
                                     self.alloc.peer_ptrs(self.flags), self.block_counter, src_div, self.e_local, self.rank, ALIGN,
                                     self.cap_rows, self.epoch, self.num_ctas)

    def combine(self, slot_loc, weights, tokens: int, topk: int, keep_rows: bool):
        self.epoch += 1
        out = torch.empty(tokens, self.hidden, dtype=self.dtype, device=slot_loc.device)
        rows = torch.empty(tokens * topk, self.hidden, dtype=self.dtype, device=slot_loc.device) if keep_rows else None
        self.lib.moe_combine(self.alloc.peer_ptrs(self.stage_out), slot_loc, weights, out, rows, self.alloc.peer_ptrs(self.flags), topk,
                             self.rank, self.epoch, self.num_ctas)
        return out, rows


class _Plan:
    """Routing state shared by the dispatch and combine halves of one MoE layer invocation."""
    __slots__ = ("disp", "gate_flat", "slot_rank", "counts", "slot_loc", "seg", "topk", "tokens", "rows_total", "tile_group", "seg2", "sync_free")


class FusedDispatch(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, plan: _Plan, layer_key, sync_free=False):
        d = plan.disp
        recv = d.recv_buffer(layer_key)
        plan.slot_loc, plan.seg = d.dispatch(x.contiguous(), None, plan.gate_flat, plan.slot_rank, plan.counts, recv, plan.topk)
        plan.sync_free = sync_free
        ctx.plan = plan
        if sync_free:
            # grouped expert GEMMs: the whole fixed-capacity buffer goes on, the per-block expert table stays on the device
            plan.tile_group, plan.seg2 = d.tile_table(plan.seg)
            plan.rows_total = d.cap_rows
            return recv[:d.cap_rows], None
        seg = plan.seg.tolist()                     # host sync: sizes the expert loop
        if seg[-1]:
            raise RuntimeError(f"MoE receive buffer overflow: need {seg[-2]} rows, capacity {d.cap_rows}; raise moe capacity_factor")
        plan.rows_total = seg[-2]
        ctx.plan = plan
        return recv[:plan.rows_total], seg

    @staticmethod
    def backward(ctx, g_rows, _):
        plan = ctx.plan
        d = plan.disp
        if g_rows.data_ptr() != d.stage_out.data_ptr():          # the grouped backward writes dX straight into the staging buffer
            d.stage_out[:plan.rows_total].copy_(g_rows)
        dx, _ = d.combine(plan.slot_loc, None, plan.tokens, plan.topk, keep_rows=False)
        return dx, None, None, None


class FusedCombine(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ys, weights, plan: _Plan):
        d = plan.disp
        if ys.data_ptr() != d.stage_out.data_ptr():               # the grouped FFN2 writes its output straight into the staging buffer
            d.stage_out[:plan.rows_total].copy_(ys)
        w32 = weights.reshape(-1).float().contiguous()
        out, rows = d.combine(plan.slot_loc, w32, plan.tokens, plan.topk, keep_rows=weights.requires_grad)
        ctx.plan = plan
        ctx.save_for_backward(w32, rows if rows is not None else w32.new_empty(0))
        ctx.has_rows = rows is not None
        ctx.w_shape, ctx.w_dtype = weights.shape, weights.dtype
        return out

    @staticmethod
    def backward(ctx, g_out):
        plan = ctx.plan
        d = plan.disp
        w32, rows = ctx.saved_tensors
        g_out = g_out.contiguous()
        # combine-bwd: rows of w[t,k] * g_out[t] travel to the expert owners (shared staging: consumed right away)
        d.dispatch(g_out, w32, plan.gate_flat, plan.slot_rank, plan.counts, d.stage_in, plan.topk)
        # sync-free: the grouped expert backward runs next in stream order and is the only reader before the next dispatch into stage_in
        g_ys = d.stage_in[:plan.rows_total] if plan.sync_free else d.stage_in[:plan.rows_total].clone()
        g_w = None
        if ctx.has_rows:
            valid = (plan.slot_loc >= 0).view(plan.tokens, plan.topk, 1)
            g_w = (rows.view(plan.tokens, plan.topk, -1).float() * g_out.float().unsqueeze(1) * valid).sum(-1)
            g_w = g_w.reshape(ctx.w_shape).to(ctx.w_dtype)
        return g_ys, g_w, None


def make_plan(disp: MoEDispatcher, gate_idx: torch.Tensor, tokens: int) -> _Plan:
    p = _Plan()
    p.disp = disp
    p.topk = gate_idx.shape[1] if gate_idx.dim() > 1 else 1
    p.tokens = tokens
    p.gate_flat = gate_idx.reshape(-1).contiguous()
    disp.ensure_capacity(p.gate_flat.numel())
    p.slot_rank, p.counts = disp.route(p.gate_flat)
    return p


_DISPATCHERS: Dict[tuple, MoEDispatcher] = {}


def get_dispatcher(group, hidden: int, e_local: int, dtype: torch.dtype, capacity_factor: float = 2.0) -> MoEDispatcher:
    key = (id(group), hidden, e_local, dtype)
    if key not in _DISPATCHERS:
        _DISPATCHERS[key] = MoEDispatcher(group, hidden, e_local, dtype, capacity_factor)
    return _DISPATCHERS[key]
