"""Mixed-precision helpers: ``GradScaler`` (dynamic fp16 / static bf16) and the ``MixPrecision*`` wrappers.

Reference: eager_engine.py:185-224 (fp16 -> dynamic ``init_loss_scaling = scale_loss``; bf16 -> static 1.0 and
``scaler.step`` is never called) and distributed/apis/amp.py:30-234 (main-grad layer/optimizer/scaler).  In this
framework the unscale + inf/nan check + clip are folded into the optimizer's device-side ``clip_coef`` kernel
(csrc/loss_optim.cu), so the scaler only owns the loss-scale value and its growth/backoff policy; the
main-grad accumulation lives in the flat optimizer (``use_main_grad``) and in the wgrad GEMM epilogue.
"""
from __future__ import annotations

import contextlib

import torch
import torch.distributed as dist


class GradScaler:
    def __init__(self, enable: bool = True, init_loss_scaling: float = 32768.0, use_dynamic_loss_scaling: bool = True,
                 incr_ratio: float = 2.0, decr_ratio: float = 0.5, incr_every_n_steps: int = 2000, decr_every_n_nan_or_inf: int = 1,
                 hcg=None):
        self._enable = enable
        self._scale = float(init_loss_scaling) if enable else 1.0
        self._dynamic = use_dynamic_loss_scaling and enable
        self._incr_ratio, self._decr_ratio = incr_ratio, decr_ratio
        self._incr_every, self._decr_every = incr_every_n_steps, decr_every_n_nan_or_inf
        self._good, self._bad = 0, 0
        self._found_inf = False
        self.hcg = hcg

    def scale(self, loss: torch.Tensor) -> torch.Tensor:
        return loss * self._scale if self._enable and self._scale != 1.0 else loss

    def get_scale(self) -> float:
        return self._scale

    def step(self, optimizer) -> None:
        optimizer.loss_scale = self._scale
        optimizer.step()
        self._found_inf = optimizer.found_inf() if self._dynamic else False
        if self._dynamic and self.hcg is not None and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            flag = torch.tensor([1.0 if self._found_inf else 0.0], device=optimizer._dev if hasattr(optimizer, "_dev") else "cpu")
            dist.all_reduce(flag, op=dist.ReduceOp.MAX)
            self._found_inf = bool(flag.item() > 0)

    def update(self) -> None:
        if not self._dynamic:
            return
        if self._found_inf:
            self._good, self._bad = 0, self._bad + 1
            if self._bad >= self._decr_every:
                self._scale = max(self._scale * self._decr_ratio, 1.0)
                self._bad = 0
        else:
            self._bad, self._good = 0, self._good + 1
            if self._good >= self._incr_every:
                self._scale *= self._incr_ratio
                self._good = 0

    def minimize(self, optimizer, loss=None) -> None:
        self.step(optimizer)
        self.update()

    @property
    def found_inf(self) -> bool:
        return self._found_inf

    def state_dict(self) -> dict:
        return dict(scale=self._scale, good=self._good, bad=self._bad)

    def load_state_dict(self, sd: dict) -> None:
        self._scale, self._good, self._bad = sd["scale"], sd.get("good", 0), sd.get("bad", 0)


def unscale_method(self, optimizer) -> None:
    """Explicit unscale for code that drives the scaler by hand (reference amp.py:193-225, bound onto the scaler as ``_unscale``): divides
    every fp32 ``main_grad`` (or plain ``grad``) of the optimizer's parameters by the loss scale in place, records whether any value was
    non-finite in ``self._found_inf`` and, when the job has model / pipeline / sharding ranks, agrees on that flag across all ranks.
    The engine does not come through here: the flat optimizer folds unscale + check + clip into its ``clip_coef`` kernel."""
    if not self._enable:
        return
    params = []
    groups = getattr(optimizer, "param_groups", None) or getattr(optimizer, "_param_groups", None)
    if groups and isinstance(groups[0], dict):
        for g in groups:
            params.extend(g["params"])
    else:
        params = list(getattr(optimizer, "_parameter_list", None) or [])
    grads = []
    for prm in params:
        g = getattr(prm, "main_grad", None)
        if g is not None:
            assert g.dtype == torch.float32, "main_grad accumulates in fp32"
        else:
            g = prm.grad
        if g is not None:
            grads.append(g)
    found = False
    if grads:
        dev = grads[0].device
        found_inf = torch.zeros(1, dtype=torch.float32, device=dev)
        inv = torch.full((1,), 1.0 / self._scale, dtype=torch.float32, device=dev)
        by_dtype = {}
        for g in grads:
            by_dtype.setdefault((g.dtype, g.device), []).append(g)
        for (_, gdev), gl in by_dtype.items():
            torch._amp_foreach_non_finite_check_and_unscale_(gl, found_inf.to(gdev) if gdev != dev else found_inf, inv.to(gdev))
        found = bool(found_inf.item() > 0)
    hcg = self.hcg
    if hcg is None:
        from . import env

        hcg = env._hcg if getattr(env, "_hcg", None) is not None else None
    if hcg is not None and dist.is_available() and dist.is_initialized() and hcg.nranks > hcg.get_data_parallel_world_size():
        flag = torch.tensor([1.0 if found else 0.0], device=grads[0].device if grads else "cpu")
        dist.all_reduce(flag, op=dist.ReduceOp.MAX)
        found = bool(flag.item() > 0)
    self._found_inf = found


GradScaler._unscale = unscale_method
MixPrecisionScaler = GradScaler


class MixPrecisionLayer(torch.nn.Module):
    """Marks a model for fp32 main-grad accumulation (reference amp.py:30-118).  The actual buffers are created by
    the flat optimizer (``use_main_grad=True``); this wrapper only forwards calls and records the request."""

    def __init__(self, layers: torch.nn.Module, dtype: str = "float16"):
        super().__init__()
        self._layers = layers
        self._dtype = dtype
        self.use_main_grad = True

    def forward(self, *args, **kwargs):
        return self._layers(*args, **kwargs)

    def state_dict(self, *a, **k):
        return self._layers.state_dict(*a, **k)

    def load_state_dict(self, *a, **k):
        return self._layers.load_state_dict(*a, **k)


class MixPrecisionOptimizer:
    """API-compat shell over an optimizer that already handles main grads."""

    def __init__(self, optimizer):
        self._inner_opt = optimizer

    def __getattr__(self, item):
        return getattr(self._inner_opt, item)


def autocast_context(enable: bool, dtype: str = "bfloat16", level: str = "O2", device_type: str = "cuda"):
    """O2 models already hold low-precision parameters, so no autocast is needed; O1 uses torch.autocast."""
    if not enable or str(level).upper() == "O2":
        return contextlib.nullcontext()
    td = torch.bfloat16 if dtype == "bfloat16" else torch.float16
    return torch.autocast(device_type=device_type, dtype=td)
