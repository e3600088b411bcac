"""``MOCOModule`` (pre-training, InfoNCE) and ``MOCOClsModule`` (linear probe with a frozen backbone) — reference
vision_model/moco_module.py:32-292."""
from __future__ import annotations

import copy

import torch
import torch.nn.functional as F

from ...core.module.basic_module import BasicModule
from ...distributed.apis import env
from ...utils.log import logger
from .moco import MoCo, MoCoClassifier
from . import resnet as R
from .metrics import TopkAcc


class MOCOModule(BasicModule):
    def __init__(self, configs):
        self.nranks = env.world_size()
        super().__init__(configs)

    def get_model(self):
        from ..language_model.language_module import _device

        cfg = copy.deepcopy(dict(self.configs.Model.get("model", {})))
        cfg.pop("name", None)
        return MoCo(**cfg).to(_device(self.configs))

    def training_step(self, batch):
        im_q, im_k = batch
        logits, labels = self.model(im_q.float(), im_k.float())
        return F.cross_entropy(logits.float(), labels)

    def training_step_end(self, log_dict):
        ips = self.configs.Global.global_batch_size / log_dict["train_cost"]
        logger.train("[train] epoch: %d, step: [%d/%d], learning rate: %.7f, loss: %.9f, batch_cost: %.5f sec, ips: %.2f images/sec"
                     % (log_dict["epoch"], log_dict["batch"], log_dict["total_batch"], log_dict["lr"], log_dict["loss"], log_dict["train_cost"], ips))

    def training_epoch_end(self, log_dict):
        logger.info("[Training] epoch: %d, total time: %.5f sec" % (log_dict["epoch"], log_dict["train_cost"]))

    def input_spec(self):
        return [dict(shape=[None, 3, 224, 224], name="images", dtype="float32")]


class MOCOClsModule(BasicModule):
    def __init__(self, configs):
        self.nranks = env.world_size()
        super().__init__(configs)
        self.metric = TopkAcc((1, 5))
        self.acc_list, self.best_metric = [], 0.0

    def get_model(self):
        from ..language_model.language_module import _device

        cfg = copy.deepcopy(dict(self.configs.Model.get("model", {})))
        backbone = getattr(R, cfg.get("backbone", "resnet50"))(num_classes=0, with_pool=False)
        head = MoCoClassifier(True, backbone.out_features, int(cfg.get("class_num", 1000)))
        pretrained = cfg.get("pretrained")
        if pretrained:
            state = torch.load(pretrained, map_location="cpu", weights_only=False)
            enc = {k[len("base.0."):]: v for k, v in state.items() if k.startswith("base.0.")}
            backbone.load_state_dict(enc, strict=False)
        for p in backbone.parameters():
            p.requires_grad = False
        model = torch.nn.Sequential(backbone, head).to(_device(self.configs))
        return model

    def training_step(self, batch):
        x, y = batch
        self.model[0].eval()                  # frozen backbone keeps BN statistics fixed
        with torch.no_grad():
            feat = self.model[0](x.float())
        return F.cross_entropy(self.model[1](feat).float(), y)

    def training_step_end(self, log_dict):
        ips = self.configs.Global.global_batch_size / log_dict["train_cost"]
        logger.train("[train] epoch: %d, step: [%d/%d], learning rate: %.7f, loss: %.9f, batch_cost: %.5f sec, ips: %.2f images/sec"
                     % (log_dict["epoch"], log_dict["batch"], log_dict["total_batch"], log_dict["lr"], log_dict["loss"], log_dict["train_cost"], ips))

    def validation_step(self, batch):
        x, y = batch
        logits = self.model(x.float())
        self.acc_list.append({k: float(v) for k, v in self.metric(logits, y).items()})
        return F.cross_entropy(logits.float(), y)

    def validation_epoch_end(self, log_dict):
        if self.acc_list:
            avg = {k: sum(a[k] for a in self.acc_list) / len(self.acc_list) for k in self.acc_list[0]}
            self.best_metric = max(self.best_metric, avg["top1"])
            logger.eval("[Eval] epoch: %d, %s, best top1: %.5f" % (log_dict["epoch"], avg, self.best_metric))
            self.acc_list = []
