"""``ImagenModule`` / ``MultiModalModule`` (reference multimodal_model/multimodal_module.py:24-111): builds
``imagen.<name>(**Model)``, trains with ``ImagenCriterion`` on dict batches, logs ``ips`` in images/sec."""
from __future__ import annotations

import copy

import torch

from ...core.module.basic_module import BasicModule
from ...distributed.apis import env
from ...utils.log import logger
from .imagen import modeling as imagen


class MultiModalModule(BasicModule):
    def __init__(self, configs):
        self.nranks = env.world_size()
        super().__init__(configs)
        self.loss_fn = self.get_loss_fn()

    def process_configs(self, configs):
        from .utils import process_configs

        return process_configs(configs)

    def training_step_end(self, log_dict):
        ips = self.configs.Global.global_batch_size / log_dict["train_cost"]
        logger.train("[train] epoch: %d, batch: %d, loss: %.9f, avg_batch_cost: %.5f sec, speed: %.2f step/s, ips: %.2f images/sec, learning rate: %.5e"
                     % (log_dict["epoch"], log_dict["batch"], log_dict["loss"], log_dict["train_cost"], 1.0 / log_dict["train_cost"], ips, log_dict["lr"]))

    def training_epoch_end(self, log_dict):
        logger.info("[Training] epoch: %d, total time: %.5f sec" % (log_dict["epoch"], log_dict["train_cost"]))


class ImagenModule(MultiModalModule):
    def get_model(self):
        from ..language_model.language_module import _device

        cfg = copy.deepcopy(dict(self.configs.Model))
        cfg.pop("module", None)
        name = cfg.pop("name")
        self._loss_cfg = dict(self.configs.get("Loss", {"name": "mse_loss", "p2_loss_weight_k": 1.0}))
        cfg.pop("fused_linear", None)
        return getattr(imagen, name)(**cfg).to(_device(self.configs))      # unknown Model keys raise in the constructor: nothing is dropped silently

    def get_loss_fn(self):
        c = dict(self._loss_cfg)
        return imagen.ImagenCriterion(c.get("name", "mse_loss"), c.get("p2_loss_weight_k", 1.0))

    def pretreating_batch(self, batch):
        return batch

    def training_step(self, batch):
        if isinstance(batch, dict):
            out = self.model(batch["images"], text_embeds=batch.get("text_embeds"), text_masks=batch.get("text_masks"),
                             input_ids=batch.get("input_ids"), attention_mask=batch.get("attention_mask"))
        else:
            out = self.model(*batch)
        return self.loss_fn(*out)

    def validation_step(self, batch):
        return self.training_step(batch)

    def input_spec(self):
        return [dict(shape=[None, 3, 64, 64], name="images", dtype="float32")]
