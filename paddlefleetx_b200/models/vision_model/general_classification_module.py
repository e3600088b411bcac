"""``GeneralClsModule`` — image classification task glue (reference vision_model/general_classification_module.py:31-187):
model / loss / metric built by name through the vision factory, ``ips: N images/sec`` log line, eval all-gathers
logits + labels across the world before computing the metric."""
from __future__ import annotations

import copy

import torch
import torch.distributed as dist

from ...core.module.basic_module import BasicModule
from ...distributed.apis import env
from ...utils.log import logger
from .factory import build


class GeneralClsModule(BasicModule):
    def __init__(self, configs):
        self.nranks = env.world_size()
        super().__init__(configs)
        self.loss_config = copy.deepcopy(dict(configs.Model.get("loss", {"train": {"name": "CELoss"}, "eval": {"name": "CELoss"}})))
        self.train_loss = build(self.loss_config["train"])
        self.eval_loss = build(self.loss_config.get("eval", self.loss_config["train"]))
        metric_cfg = configs.Model.get("metric")
        self.train_metric = build(metric_cfg["train"]) if metric_cfg and metric_cfg.get("train") else None
        self.eval_metric = build(metric_cfg["eval"]) if metric_cfg and metric_cfg.get("eval") else None
        self.best_metric, self.acc_list = 0.0, []

    def process_configs(self, configs):
        g = configs.Global
        for mode in ("Train", "Eval", "Test"):
            if mode in configs.get("Data", {}) and "sampler" in configs.Data[mode]:
                configs.Data[mode].sampler.setdefault("batch_size", g.local_batch_size)
        return configs

    def get_model(self):
        from ..language_model.language_module import _device, _param_dtype

        cfg = copy.deepcopy(dict(self.configs.Model.model))
        model = build(cfg)
        return model.to(device=_device(self.configs), dtype=_param_dtype(self.configs))

    def forward(self, inputs):
        return self.model(inputs.to(next(self.model.parameters()).dtype))

    def training_step(self, batch):
        inputs, labels = batch
        logits = self(inputs)
        return self.train_loss(logits, labels)

    def training_step_end(self, log_dict):
        ips = self.configs.Global.global_batch_size / log_dict["train_cost"]
        logger.train("[train] epoch: %d, step: [%d/%d], learning rate: %.7f, loss: %.9f, batch_cost: %.5f sec, ips: %.2f images/sec"
                     % (log_dict["epoch"], log_dict["batch"], log_dict["total_batch"], log_dict["lr"], log_dict["loss"], log_dict["train_cost"], ips))

    def validation_step(self, batch):
        inputs, labels = batch
        logits = self(inputs)
        loss = self.eval_loss(logits, labels)
        if self.eval_metric is not None:
            if self.nranks > 1:
                gl = [torch.empty_like(logits) for _ in range(self.nranks)]
                gy = [torch.empty_like(labels) for _ in range(self.nranks)]
                dist.all_gather(gl, logits.contiguous()); dist.all_gather(gy, labels.contiguous())
                logits, labels = torch.cat(gl), torch.cat(gy)
            self.acc_list.append({k: float(v) for k, v in self.eval_metric(logits, labels).items()})
        return loss

    def validation_step_end(self, log_dict):
        logger.eval("[eval] epoch: %d, batch: %d, loss: %.9f, avg_eval_cost: %.5f sec" % (log_dict["epoch"], log_dict["batch"], log_dict["loss"], log_dict["eval_cost"]))

    def validation_epoch_end(self, log_dict):
        msg = ""
        if self.acc_list:
            keys = self.acc_list[0].keys()
            avg = {k: sum(a[k] for a in self.acc_list) / len(self.acc_list) for k in keys}
            first = next(iter(avg.values()))
            self.best_metric = max(self.best_metric, first)
            msg = ", ".join(f"{k}: {v:.5f}" for k, v in avg.items()) + f", best: {self.best_metric:.5f}"
            self.acc_list = []
        logger.eval("[Eval] epoch: %d, total time: %.5f sec, %s" % (log_dict["epoch"], log_dict["eval_cost"], msg))

    def test_step(self, batch):
        return self.validation_step(batch)

    def training_epoch_end(self, log_dict):
        logger.info("[Training] epoch: %d, total time: %.5f sec" % (log_dict["epoch"], log_dict["train_cost"]))

    def input_spec(self):
        s = self.configs.Model.model.get("img_size", 224)
        return [dict(shape=[None, 3, s, s], name="images", dtype="float32")]


GeneralClsModuleAuto = GeneralClsModule
