"""Component builders for explicit training loops (reference examples/transformer/utils/components.py:32-210)."""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..", "..")))

from paddlefleetx_b200.data import build_batch_sampler, build_dataloader, build_dataset  # noqa: E402,F401
from paddlefleetx_b200.distributed.apis import env  # noqa: E402
from paddlefleetx_b200.optims import build_grad_clip, build_lr_scheduler  # noqa: E402,F401
from paddlefleetx_b200.optims import build_optimizer as _build_optimizer  # noqa: E402
from paddlefleetx_b200.utils.profiler import StepProfiler  # noqa: E402


def build_optimizer(config, model, lr_scheduler=None, multi_precision=False, dist_config=None, amp_config=None):
    cfg = dict(config)
    cfg.setdefault("multi_precision", multi_precision)
    hcg = env.get_hcg() if env.world_size() > 1 else None
    return _build_optimizer(cfg, model, lr_scheduler, hcg=hcg, dist_config=dist_config, amp_config=amp_config)


def build_profiler(profiler_config):
    if not profiler_config or not profiler_config.get("enable", False):
        return None
    return StepProfiler(profiler_config)


def profiler_done(profiler, profiler_config=None):
    if profiler is not None:
        profiler.finish()
