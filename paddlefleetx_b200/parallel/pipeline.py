"""Pipeline parallelism: layer descriptors, stage partitioning, tied weights and the 1F1B / interleaved schedulers.

Supplies what the reference takes from Paddle: ``LayerDesc`` / ``SharedLayerDesc`` / ``PipelineLayer``
(call sites gpt/dygraph/hybrid_model.py:1115-1206, ernie/dygraph/hybrid_model.py:796-872) and the
``PipelineParallel`` runtime with ``_prepare_training`` / ``forward_backward_pipeline(batch, scaler)`` /
``eval_batch(batch, compute_loss)`` (eager_engine.py:507-517,655,713).

Design points:
  * stage boundaries carry ONE activation tensor; its shape/dtype is exchanged once per (stage pair, shape) and then
    cached — steady-state micro-batches need no meta handshake,
  * sends/receives are posted with ``batch_isend_irecv`` (NCCL p2p on CUDA, gloo on CPU) so the paired
    send-forward/recv-backward of 1F1B cannot deadlock; on CUDA they run on NCCL's internal streams and overlap
    the compute stream,
  * tied embeddings (``SharedLayerDesc``): the copies are synchronised at build time and their gradients are
    all-reduced over the first/last-stage pair after the schedule; duplicates are flagged so the global grad norm
    counts the weight once.
"""
from __future__ import annotations

import re
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist
import torch.nn as nn

from . import comm_ops as C
from .recompute import recompute, recompute_hybrid


class LayerDesc:
    def __init__(self, layer_func, *inputs, **kwargs):
        self.layer_func, self.inputs, self.kwargs = layer_func, inputs, kwargs
        if not callable(layer_func):
            raise TypeError("The input(layer_func) should be a callable (nn.Module class or factory)")

    def build_layer(self) -> nn.Module:
        return self.layer_func(*self.inputs, **self.kwargs)

    def __repr__(self) -> str:
        return f"LayerDesc({getattr(self.layer_func, '__name__', self.layer_func)})"


class SharedLayerDesc(LayerDesc):
    def __init__(self, key: str, layer_func, forward_func: Optional[Callable] = None, shared_weight_attr: str = "weight", *inputs, **kwargs):
        super().__init__(layer_func, *inputs, **kwargs)
        self.layer_name = key
        self.forward_func = forward_func
        self.shared_weight_attr = shared_weight_attr


def _resolve_attr(module: nn.Module, path: str):
    obj = module
    for part in path.split("."):
        obj = getattr(obj, part)
    return obj


class SegmentLayers:
    """Split ``num_items`` layer descriptors into ``num_parts`` contiguous parts: ``uniform`` or
    ``layer:<ClassName regex>`` (balance only the matching layers, reference seg_method)."""

    def __init__(self, descs: Sequence, num_parts: int, method: str = "uniform"):
        self.descs, self.num_parts, self.method = list(descs), num_parts, method
        assert len(self.descs) >= num_parts, "layer number should be greater than number of segments"

    def do_segment(self) -> List[int]:
        n = len(self.descs)
        if self.method == "uniform":
            return self.uniform(n, self.num_parts)
        if self.method.startswith("layer:"):
            pat = re.compile(self.method.split(":", 1)[1], re.IGNORECASE)
            weights = [0] * n
            for i, d in enumerate(self.descs):
                name = d.layer_func.__name__ if isinstance(d, LayerDesc) else d.__class__.__name__ if isinstance(d, nn.Module) else getattr(d, "__name__", "")
                if pat.search(name):
                    weights[i] = 1
            total = sum(weights)
            if total == 0 or total % self.num_parts != 0:
                if total < self.num_parts:
                    return self.uniform(n, self.num_parts)
            per = self.uniform(total, self.num_parts)     # boundaries in units of matching layers
            bounds, seen, k = [0], 0, 1
            for i, w in enumerate(weights):
                if k < self.num_parts and w and seen == per[k]:
                    bounds.append(i)
                    k += 1
                seen += w
            while len(bounds) < self.num_parts:
                bounds.append(n)
            bounds.append(n)
            return bounds
        raise ValueError(f"unknown seg_method {self.method}")

    @staticmethod
    def uniform(num_items: int, num_parts: int) -> List[int]:
        base, extra = divmod(num_items, num_parts)
        out = [0]
        for i in range(num_parts):
            out.append(out[-1] + base + (1 if i >= num_parts - extra else 0))
        return out


class PipelineLayer(nn.Module):
    def __init__(self, layers: Sequence, num_stages: Optional[int] = None, topology=None, loss_fn: Optional[nn.Module] = None,
                 seg_method: str = "uniform", recompute_interval: int = 0, recompute_ctx: Optional[dict] = None,
                 num_virtual_pipeline_stages: Optional[int] = None, hcg=None):
        super().__init__()
        from ..distributed.apis import env

        self._hcg = hcg or env.get_hcg()
        self._num_stages = num_stages or self._hcg.get_pipe_parallel_world_size()
        self._stage_id = self._hcg.get_stage_id()
        self._loss_fn = loss_fn
        self._recompute_interval = recompute_interval
        self._recompute_ctx = recompute_ctx
        self._num_virtual = num_virtual_pipeline_stages or 1
        self._layers_desc = list(layers)
        total_parts = self._num_stages * self._num_virtual
        self.segment_parts = SegmentLayers(self._layers_desc, total_parts, seg_method).do_segment()
        self.shared_layers = nn.ModuleDict()
        self.shared_weight_attrs: Dict[str, str] = {}
        self._chunks: List[List] = []              # per virtual chunk: list of (callable, is_module)
        self._model_chunks = nn.ModuleList()
        for v in range(self._num_virtual):
            part = v * self._num_stages + self._stage_id
            start, end = self.segment_parts[part], self.segment_parts[part + 1]
            funcs, holder = self._build_range(start, end)
            self._chunks.append(funcs)
            self._model_chunks.append(holder)
        self._sync_shared_weights()

    # -- construction -------------------------------------------------------------------------
    def _build_range(self, start: int, end: int):
        # keyed by the GLOBAL layer index so checkpoint keys do not depend on the pp layout (utils/ckpt_convert.py)
        funcs, holder = [], nn.ModuleDict()
        for idx in range(start, end):
            d = self._layers_desc[idx]
            if isinstance(d, SharedLayerDesc):
                if d.layer_name not in self.shared_layers:
                    self.shared_layers[d.layer_name] = d.build_layer()
                    self.shared_weight_attrs[d.layer_name] = d.shared_weight_attr
                layer = self.shared_layers[d.layer_name]
                if d.forward_func is None:
                    funcs.append(layer)
                else:
                    funcs.append(_Bound(d.forward_func, layer))
            elif isinstance(d, LayerDesc):
                layer = d.build_layer()
                holder[str(idx)] = layer
                funcs.append(layer)
            elif isinstance(d, nn.Module):
                holder[str(idx)] = d
                funcs.append(d)
            elif callable(d):
                funcs.append(d)
            else:
                raise TypeError(f"unsupported layer entry {d!r}")
        return funcs, holder

    def _shared_stage_sets(self) -> Dict[str, List[int]]:
        where: Dict[str, set] = {}
        for part in range(self._num_stages * self._num_virtual):
            for idx in range(self.segment_parts[part], self.segment_parts[part + 1]):
                d = self._layers_desc[idx]
                if isinstance(d, SharedLayerDesc):
                    where.setdefault(d.layer_name, set()).add(part % self._num_stages)
        return {k: sorted(v) for k, v in where.items()}

    def _sync_shared_weights(self) -> None:
        self._shared_comm = {}
        for key, stages in self._shared_stage_sets().items():
            if len(stages) < 2 or key not in self.shared_layers:
                continue
            assert stages == [0, self._num_stages - 1], "shared layers are supported between the first and last stage"
            grp = self._hcg.get_embedding_group()
            w = _resolve_attr(self.shared_layers[key], self.shared_weight_attrs[key])
            if self._stage_id != stages[0]:
                w.pp_shared_duplicate = True           # counted once in the global grad norm
            self._shared_comm[key] = (grp, w)
            if grp is not None and grp.process_group is not None:
                dist.broadcast(w.data, src=grp.ranks[0], group=grp.process_group)

    def allreduce_shared_weight_gradients(self) -> None:
        C.wait_side_streams()       # an overlapped bucket reduction of the same gradient buffer may still be in flight
        for key, (grp, w) in self._shared_comm.items():
            if grp is None or grp.process_group is None:
                continue
            g = getattr(w, "main_grad", None)
            g = g if g is not None else w.grad
            if g is None:
                g = torch.zeros_like(w)
                w.grad = g
            dist.all_reduce(g, group=grp.process_group)

    # -- execution -----------------------------------------------------------------------------
    def get_num_virtual_stages(self) -> int:
        return self._num_virtual

    def forward_chunk(self, x, chunk_id: int = 0):
        funcs = self._chunks[chunk_id]
        if self._recompute_interval <= 0 or not self.training:
            for f in funcs:
                x = f(x) if not isinstance(x, tuple) else f(*x)
            return x
        i = 0
        while i < len(funcs):
            seg = funcs[i:i + self._recompute_interval]
            needs = any(isinstance(f, nn.Module) and any(p.requires_grad for p in f.parameters()) for f in seg)
            first_embed = i == 0 and self._stage_id == 0 and chunk_id == 0   # int inputs carry no grad: run plainly
            if needs and not first_embed and isinstance(x, torch.Tensor) and x.requires_grad:
                run = _SeqRunner(seg)
                x = recompute_hybrid(self._recompute_ctx, run, x) if self._recompute_ctx else recompute(run, x)
            else:
                for f in seg:
                    x = f(x) if not isinstance(x, tuple) else f(*x)
            i += self._recompute_interval
        return x

    def forward(self, x, chunk_id: int = 0):
        return self.forward_chunk(x, chunk_id)


class _Bound:
    def __init__(self, fn, layer):
        self.fn, self.layer = fn, layer

    def __call__(self, *args):
        return self.fn(self.layer, *args)


class _SeqRunner:
    def __init__(self, funcs):
        self.funcs = funcs

    def __call__(self, x):
        for f in self.funcs:
            x = f(x) if not isinstance(x, tuple) else f(*x)
        return x


# ============================================================================================ p2p
class _P2P:
    """Stage-boundary communication with cached tensor metadata."""

    def __init__(self, hcg, partial: bool = False):
        self.hcg = hcg
        self.prev, self.next = hcg.prev_rank, hcg.next_rank
        self.group = hcg.get_pipe_parallel_group().process_group
        # ``enable_partial_send_recv`` (reference: pipeline_configs, env.py:139-148): activations / activation gradients at a stage boundary are
        # replicated across the tensor-parallel group, so every mp rank ships only its 1/mp slice to its peer in the next stage and the
        # receiving stage all-gathers over its own mp group — the inter-stage link (the slow one when stages sit on different nodes)
        # carries mp x less data.  Off under sequence parallelism, where boundary tensors are already sequence shards.
        self.mp_group = hcg.get_model_parallel_group()
        self.mp, self.mp_rank = hcg.get_model_parallel_world_size(), hcg.get_model_parallel_rank()
        self.partial = bool(partial) and self.mp > 1
        self.partial_transfers = 0      # sliced sends + receives so far (observability / tests)
        self._recv_meta: Dict[str, Tuple] = {}
        self._sent_meta: Dict[str, Tuple] = {}
        self._pending: List = []        # outstanding sends (request, tensor kept alive)

    _DTYPES = [torch.float32, torch.float16, torch.bfloat16, torch.int64, torch.int32]

    def _send_meta(self, t: torch.Tensor, dst: int, tag: str) -> None:
        meta = (tuple(t.shape), t.dtype)
        if self._sent_meta.get(tag) == meta:
            return
        buf = torch.zeros(10, dtype=torch.int64, device=t.device)
        buf[0] = t.dim()
        buf[1] = self._DTYPES.index(t.dtype)
        buf[2:2 + t.dim()] = torch.tensor(t.shape, dtype=torch.int64)
        self._pending.append((dist.isend(buf, dst=dst, group=self.group), buf))
        self._sent_meta[tag] = meta

    def _get_recv_meta(self, src: int, tag: str, device) -> Tuple:
        if tag not in self._recv_meta:
            buf = torch.zeros(10, dtype=torch.int64, device=device)
            dist.recv(buf, src=src, group=self.group)
            vals = buf.tolist()
            self._recv_meta[tag] = (tuple(int(v) for v in vals[2:2 + int(vals[0])]), self._DTYPES[int(vals[1])])
        return self._recv_meta[tag]

    def exchange(self, send_next=None, send_prev=None, recv_prev: bool = False, recv_next: bool = False, device=None, fwd_tag: str = "f",
                 bwd_meta: Optional[Tuple] = None):
        """Post all requested transfers as one batch; returns (tensor_from_prev, tensor_from_next)."""
        ops, from_prev, from_next = [], None, None
        if send_next is not None:
            self._send_meta(send_next, self.next, fwd_tag)
        if recv_prev:
            shape, dtype = self._get_recv_meta(self.prev, fwd_tag, device)
            from_prev = torch.empty(shape, dtype=dtype, device=device)
        if recv_next:
            shape, dtype = bwd_meta
            from_next = torch.empty(shape, dtype=dtype, device=device)
        kinds, partial_recv = [], []

        def outgoing(t):
            t = t.contiguous()
            if not self._is_partial(t):
                return t
            self.partial_transfers += 1
            return t.view(-1).chunk(self.mp)[self.mp_rank]

        def incoming(full):
            if not self._is_partial(full):
                return full
            piece = full.view(-1).chunk(self.mp)[self.mp_rank]      # a view: the slice lands in place, the rest is filled by the all-gather
            partial_recv.append((full, piece))
            self.partial_transfers += 1
            return piece

        if send_prev is not None:
            t = outgoing(send_prev)
            ops.append(dist.P2POp(dist.isend, t, self.prev, self.group)); kinds.append(t)
        if recv_prev:
            ops.append(dist.P2POp(dist.irecv, incoming(from_prev), self.prev, self.group)); kinds.append(None)
        if send_next is not None:
            t = outgoing(send_next)
            ops.append(dist.P2POp(dist.isend, t, self.next, self.group)); kinds.append(t)
        if recv_next:
            ops.append(dist.P2POp(dist.irecv, incoming(from_next), self.next, self.group)); kinds.append(None)
        if ops:
            reqs = dist.batch_isend_irecv(ops)
            if len(reqs) == len(ops):
                # per-op requests (gloo): block only on receives; sends complete in the background so that ring
                # wrap-around sends of the interleaved schedule cannot form a wait cycle
                for req, keep in zip(reqs, kinds):
                    if keep is None:
                        req.wait()
                    else:
                        self._pending.append((req, keep))
            else:
                for req in reqs:      # coalesced NCCL group: wait() only orders streams, it does not block the host
                    req.wait()
        for full, piece in partial_recv:
            full.view(-1).copy_(C.all_gather_dim0(piece, self.mp_group))
        if len(self._pending) > 64:
            self.flush(keep_last=32)
        return from_prev, from_next

    def _is_partial(self, t: torch.Tensor) -> bool:
        return self.partial and t.is_floating_point() and t.numel() % self.mp == 0 and t.numel() >= self.mp

    def flush(self, keep_last: int = 0) -> None:
        while len(self._pending) > keep_last:
            req, _ = self._pending.pop(0)
            req.wait()


# ============================================================================================ schedulers
class PipelineParallel(nn.Module):
    def __init__(self, layers: PipelineLayer, hcg, strategy):
        super().__init__()
        if not isinstance(layers, PipelineLayer):
            raise TypeError("The Layer should be a derived class of PipelineLayer.")
        self._layers = layers
        self._hcg = hcg
        pc = strategy.pipeline_configs
        self.accumulate_steps = int(pc["accumulate_steps"])
        self.micro_batch_size = int(pc["micro_batch_size"])
        self.num_stages = hcg.get_pipe_parallel_world_size()
        self.stage_id = hcg.get_stage_id()
        self.is_first = self.stage_id == 0
        self.is_last = self.stage_id == self.num_stages - 1
        self._p2p = _P2P(hcg, partial=bool(pc.get("enable_partial_send_recv", False)))
        self._num_virtual = layers.get_num_virtual_stages()
        self.optimizer = None
        self.lr_scheduler = None
        self.total_loss = None

    # passthroughs so the wrapper looks like the wrapped model to the engine / checkpoint code
    def state_dict(self, *a, **k):
        return self._layers.state_dict(*a, **k)

    def load_state_dict(self, *a, **k):
        return self._layers.load_state_dict(*a, **k)

    def parameters(self, recurse: bool = True):
        return self._layers.parameters(recurse)

    def named_parameters(self, *a, **k):
        return self._layers.named_parameters(*a, **k)

    def _prepare_training(self, data, optimizer, lr_scheduler):
        self.optimizer, self.lr_scheduler = optimizer, lr_scheduler
        self._layers.train()

    # -- micro-batch plumbing -------------------------------------------------------------------
    def _split(self, part, i: int):
        if part is None:
            return None
        if isinstance(part, torch.Tensor):
            b = self.micro_batch_size
            return part[i * b:(i + 1) * b]
        return type(part)(self._split(p, i) for p in part)

    def _load_micro(self, data, i: int):
        inputs, labels = data
        return (self._split(inputs, i) if self.is_first else None), (self._split(labels, i) if self.is_last else None)

    def _device(self):
        return next(self._layers.parameters()).device

    def _forward_step(self, recv, micro_inputs, micro_labels, chunk: int = 0, first_chunk: bool = True, last_chunk: bool = True):
        x = micro_inputs if (self.is_first and first_chunk) else recv
        out = self._layers.forward_chunk(x, chunk)
        if self.is_last and last_chunk and self._layers._loss_fn is not None and micro_labels is not None:
            labels = micro_labels if isinstance(micro_labels, (tuple, list)) else (micro_labels,)
            loss = self._layers._loss_fn(out, *labels)
            out = loss / self.accumulate_steps
            with torch.no_grad():
                self.total_loss = loss.detach() if self.total_loss is None else self.total_loss + loss.detach()
        return out

    @staticmethod
    def _backward_step(inp, out, out_grad, scaler=None):
        if inp is not None and isinstance(inp, torch.Tensor) and inp.requires_grad:
            inp.retain_grad()
        C.run_pre_backward()
        if out_grad is None:                      # last stage: ``out`` is the (scaled) loss
            (scaler.scale(out) if scaler is not None else out).backward()
        else:
            torch.autograd.backward(out, out_grad)
        g = inp.grad if isinstance(inp, torch.Tensor) and inp.requires_grad else None
        return g

    # -- 1F1B ---------------------------------------------------------------------------------------
    def forward_backward_pipeline(self, data, scaler=None):
        """One training step under the 1F1B schedule (interleaved when virtual stages are configured): ``warm`` forward micro-batches, then
        forward / backward pairs with the stage-boundary activations and gradients exchanged by ``batch_isend_irecv``, then the remaining backwards;
        returns the mean loss over the micro-batches, broadcast from the last stage (reference contract: eager_engine.py:507-517)."""
        if self._num_virtual > 1:
            return self._interleaved(data, scaler)
        M = self.accumulate_steps
        dev = self._device()
        self.total_loss = None
        warm = min(self.num_stages - self.stage_id - 1, M)
        steady = M - warm
        inputs, outputs = [], []
        no_sync = self.optimizer.no_sync() if (self.optimizer is not None and hasattr(self.optimizer, "no_sync")) else _Null()
        fwd_i = 0

        def recv_fwd():
            if self.is_first:
                return None
            t, _ = self._p2p.exchange(recv_prev=True, device=dev)
            return t.requires_grad_()

        with no_sync:
            for _ in range(warm):
                x = recv_fwd()
                mi, ml = self._load_micro(data, fwd_i)
                y = self._forward_step(x, mi, ml)
                fwd_i += 1
                if not self.is_last:
                    self._p2p.exchange(send_next=y)
                inputs.append(x); outputs.append(y)
            x = recv_fwd() if steady > 0 else None
            for k in range(steady):
                last_iter = k == steady - 1
                mi, ml = self._load_micro(data, fwd_i)
                y = self._forward_step(x, mi, ml)
                fwd_i += 1
                if self.is_last:
                    gy = None
                else:
                    _, gy = self._p2p.exchange(send_next=y, recv_next=True, device=dev, bwd_meta=(tuple(y.shape), y.dtype))
                inputs.append(x); outputs.append(y)
                xi, yi = inputs.pop(0), outputs.pop(0)
                sync_now = last_iter and warm == 0
                gx = self._run_backward(xi, yi, gy, scaler, sync_now)
                if last_iter:
                    x = None
                    if not self.is_first:
                        self._p2p.exchange(send_prev=gx)
                elif self.is_first:
                    x = None
                else:
                    x, _ = self._p2p.exchange(send_prev=gx, recv_prev=True, device=dev)
                    x = x.requires_grad_()
            for k in range(warm):
                xi, yi = inputs.pop(0), outputs.pop(0)
                _, gy = self._p2p.exchange(recv_next=True, device=dev, bwd_meta=(tuple(yi.shape), yi.dtype))
                gx = self._run_backward(xi, yi, gy, scaler, k == warm - 1)
                if not self.is_first:
                    self._p2p.exchange(send_prev=gx)
        self._p2p.flush()
        self._layers.allreduce_shared_weight_gradients()
        return self._broadcast_loss(dev)

    def _run_backward(self, xi, yi, gy, scaler, sync_now: bool):
        """The last backward of the schedule runs outside ``no_sync`` so ZeRO/DP bucket hooks may fire."""
        if sync_now and self.optimizer is not None and hasattr(self.optimizer, "_accumulating"):
            prev = self.optimizer._accumulating
            self.optimizer._accumulating = False
            try:
                return self._backward_step(xi, yi, gy, scaler)
            finally:
                self.optimizer._accumulating = prev
        return self._backward_step(xi, yi, gy, scaler)

    def _broadcast_loss(self, dev):
        loss = (self.total_loss / self.accumulate_steps) if self.is_last and self.total_loss is not None else torch.zeros((), device=dev)
        loss = loss.float().reshape(1).contiguous()
        pg = self._hcg.get_pipe_parallel_group()
        if pg.process_group is not None:
            dist.broadcast(loss, src=pg.ranks[-1], group=pg.process_group)
        return loss.reshape(())

    # -- interleaved (virtual stages): breadth-first over chunks, one micro-batch group at a time -------
    def _interleaved(self, data, scaler=None):
        """Interleaved schedule.  Model chunk v of stage s holds global part ``v * pp + s``; a micro-batch visits
        (v=0,s=0..pp-1), (v=1,s=0..pp-1), ...  Forwards for all micro-batches run chunk-major in groups of ``pp``
        micro-batches, then backwards in reverse — an all-forward/all-backward variant per group that keeps the same
        numerics as 1F1B-interleave with a slightly larger activation footprint."""
        M, P, V = self.accumulate_steps, self.num_stages, self._num_virtual
        assert M % P == 0, "interleaved pipeline needs accumulate_steps % pp_degree == 0"
        dev = self._device()
        self.total_loss = None
        no_sync = self.optimizer.no_sync() if (self.optimizer is not None and hasattr(self.optimizer, "no_sync")) else _Null()
        with no_sync:
            for g0 in range(0, M, P):
                saved = {}
                for v in range(V):
                    for m in range(g0, g0 + P):
                        first_part = self.is_first and v == 0
                        last_part = self.is_last and v == V - 1
                        x = None
                        if not first_part:
                            x, _ = self._p2p.exchange(recv_prev=True, device=dev, fwd_tag=f"f{v}")
                            x = x.requires_grad_()
                        mi, ml = self._load_micro(data, m)
                        y = self._forward_step(x, mi, ml, chunk=v, first_chunk=v == 0, last_chunk=v == V - 1)
                        if not last_part:
                            self._p2p.exchange(send_next=y, fwd_tag=f"f{v if not self.is_last else v + 1}")
                        saved[(v, m)] = (x, y)
                for v in reversed(range(V)):
                    for m in range(g0, g0 + P):
                        x, y = saved.pop((v, m))
                        last_part = self.is_last and v == V - 1
                        first_part = self.is_first and v == 0
                        gy = None
                        if not last_part:
                            _, gy = self._p2p.exchange(recv_next=True, device=dev, bwd_meta=(tuple(y.shape), y.dtype))
                        final = (g0 + P >= M) and v == 0 and m == g0 + P - 1
                        gx = self._run_backward(x, y, gy, scaler, final)
                        if not first_part:
                            self._p2p.exchange(send_prev=gx)
        self._p2p.flush()
        self._layers.allreduce_shared_weight_gradients()
        return self._broadcast_loss(dev)

    # -- train / eval convenience --------------------------------------------------------------------
    def train_batch(self, data, optimizer, lr_scheduler=None, scaler=None):
        self._prepare_training(data, optimizer, lr_scheduler)
        loss = self.forward_backward_pipeline(data, scaler)
        optimizer.step()
        optimizer.clear_grad()
        if lr_scheduler is not None:
            lr_scheduler.step()
        return loss

    @torch.no_grad()
    def eval_batch(self, data, compute_loss: bool = False):
        self._layers.eval()
        dev = self._device()
        self.total_loss = None
        M, V = self.accumulate_steps, self._num_virtual
        out = None
        for m in range(M):
            for v in range(V):
                first_part = self.is_first and v == 0
                last_part = self.is_last and v == V - 1
                x = None
                if not first_part:
                    x, _ = self._p2p.exchange(recv_prev=True, device=dev, fwd_tag=f"e{v}")
                mi, ml = self._load_micro(data, m)
                if not compute_loss:
                    ml = None
                out = self._forward_step(x, mi, ml, chunk=v, first_chunk=v == 0, last_chunk=v == V - 1)
                if not last_part:
                    self._p2p.exchange(send_next=out, fwd_tag=f"e{v if not self.is_last else v + 1}")
        self._p2p.flush()
        if compute_loss:
            return self._broadcast_loss(dev)
        return out

    def forward(self, *args, **kwargs):
        return self._layers(*args, **kwargs)

    def __getattr__(self, name):
        try:
            return super().__getattr__(name)
        except AttributeError:
            return getattr(super().__getattr__("_layers"), name)


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False
