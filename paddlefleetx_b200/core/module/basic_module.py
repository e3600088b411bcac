"""``BasicModule`` — the task-glue contract between the engine and a model family.

Hook names and call order follow ppfleetx/core/module/basic_module.py:29-86: the engine calls
``pretreating_batch`` -> ``training_step`` -> ``backward`` per micro-batch, ``training_step_end`` at every logging
boundary, and the ``validation_*`` / ``test_*`` twins for eval / predict.  Sub-classes implement ``get_model``
(and usually ``get_loss_fn``, ``training_step``, ``*_step_end``).
"""
from __future__ import annotations

import torch.nn as nn


class BasicModule(nn.Module):
    def __init__(self, configs, *args, **kwargs):
        super().__init__()
        self.configs = configs
        self.nranks = _world_size()
        self.process_configs(configs)
        self.model = self.get_model()
        self.loss_fn = self.get_loss_fn() if hasattr(self, "get_loss_fn") else None

    # -- construction hooks
    def process_configs(self, configs):
        return configs

    def get_model(self):
        raise NotImplementedError

    def get_loss_fn(self):
        return None

    # -- execution hooks
    def forward(self, *args, **kwargs):
        return self.model(*args, **kwargs)

    def pretreating_batch(self, batch):
        return batch

    def training_step(self, *args, **kwargs):
        raise NotImplementedError

    def training_step_end(self, *args, **kwargs):
        pass

    def validation_step(self, *args, **kwargs):
        raise NotImplementedError

    def validation_step_end(self, *args, **kwargs):
        pass

    def test_step(self, *args, **kwargs):
        raise NotImplementedError

    def test_step_end(self, *args, **kwargs):
        pass

    def backward(self, loss):
        loss.backward()

    def input_spec(self):
        raise NotImplementedError

    def inference_end(self, outputs):
        return outputs

    def training_epoch_end(self, *args, **kwargs):
        pass

    def validation_epoch_end(self, *args, **kwargs):
        pass


def _world_size() -> int:
    import torch.distributed as dist

    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
