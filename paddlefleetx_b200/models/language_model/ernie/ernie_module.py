"""ERNIE task modules (reference ernie/ernie_module.py:44-382): ``ErnieModule`` (MLM + SOP pre-training on the
6-field batch ``input_ids, segment_ids, input_mask, masked_lm_positions, masked_lm_labels, next_sentence_labels``) and
``ErnieSeqClsModule`` (sequence classification fine-tune).  ``*Auto`` names map to the same eager classes."""
from __future__ import annotations

import os


import torch

from ....core.module.basic_module import BasicModule
from ....distributed.apis import env
from ....utils.log import logger
from ..language_module import _device, _param_dtype, process_optim_configs
from . import model as E

_KEYS = ("vocab_size", "hidden_size", "num_hidden_layers", "num_layers", "num_attention_heads", "ffn_hidden_size", "intermediate_size",
         "hidden_act", "hidden_dropout_prob", "attention_probs_dropout_prob", "max_position_embeddings", "type_vocab_size",
         "task_type_vocab_size", "task_id", "use_task_id", "initializer_range", "pad_token_id", "use_recompute", "use_flash_attn",
         "sequence_parallel")


def get_model_size(l, h, v, s) -> float:
    p = 12 * l * h * h * (1 + 13 / (12 * h) + (v + s) / (12 * l * h))
    logger.info("Model Size: {:.2f} B".format(p / 1e9))
    return p


def process_data_configs(config) -> None:
    g, eng = config.Global, config.Engine
    eval_freq = eng.eval_freq if eng.eval_freq and eng.eval_freq > 0 else max(eng.max_steps, 1)
    n = {"Train": g.global_batch_size * eng.max_steps, "Eval": g.global_batch_size * (eng.max_steps // eval_freq + 1) * eng.eval_iters,
         "Test": g.global_batch_size * eng.test_iters}
    for mode in ("Train", "Eval", "Test"):
        if mode in config.get("Data", {}):
            ds = config.Data[mode].dataset
            ds.setdefault("num_samples", n[mode])
            ds["mode"] = mode
            ds.setdefault("seed", g.seed)
            ds.setdefault("binary_head", g.get("binary_head", True))
            if ds.get("vocab_size") is None and not ds.get("tokenizer_type") and config.get("Model", {}).get("vocab_size"):
                ds["vocab_size"] = config.Model.vocab_size          # random-replacement ids of the MLM masking stay inside the embedding table
            if "sampler" in config.Data[mode]:
                config.Data[mode].sampler["batch_size"] = g.local_batch_size
            col = config.Data[mode].get("loader", {}).get("collate_fn")
            if isinstance(col, dict) and col.get("name") == "ErnieCollateData":
                col["micro_batch_size"] = g.micro_batch_size


def process_model_configs(config) -> None:
    """``Model.intermediate_size`` defaults to four times the hidden size (reference ernie_module.py:75-78)."""
    m = config["Model"]
    m.setdefault("intermediate_size", m["hidden_size"] * 4)


def process_auto_model_configs(config) -> None:
    """Auto variant (reference ernie/auto/auto_module.py:63-68): also records the process mesh under ``Model.mesh``."""
    from ..auto_module import process_mesh_config

    config["Model"].update({"mesh": process_mesh_config(config["Distributed"])})
    process_model_configs(config)


FINETUNE_CONFIGS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "finetune_configs.yaml")


def process_finetune_configs(task: str, config, path: str = FINETUNE_CONFIGS) -> None:
    """Overlay the per-dataset fine-tuning hyper-parameters of ``finetune_configs.yaml`` (``{task: {dataset_type: {num_train_epochs,
    learning_rate, max_seq_length, batch_size}}}``) on a recipe (reference ernie_module.py:81-117): epochs go to the engine, the learning rate to
    ``Optimizer.lr.max_lr``, the sequence length to every dataset section, the batch size to ``Global`` (global = local x dp x pp, as
    the reference computes it)."""
    import yaml

    with open(path, encoding="utf-8") as f:
        table = yaml.safe_load(f)
    dataset_type = config.Data.Train.dataset.dataset_type
    assert task in table and dataset_type in table[task], (f"{dataset_type} is an invalid dataset type ! Only support the types of dataset shown "
                                                           f"in {path}")
    hp = table[task][dataset_type]
    if hp.get("num_train_epochs") is not None:
        config.Engine["num_train_epochs"] = hp["num_train_epochs"]
    if hp.get("learning_rate") is not None:
        config.Optimizer["lr"]["max_lr"] = hp["learning_rate"]
    if hp.get("max_seq_length") is not None:
        for mode in ("Train", "Eval", "Test"):
            if mode in config.Data:
                config.Data[mode]["dataset"]["max_seq_len"] = hp["max_seq_length"]
    if hp.get("batch_size") is not None:
        assert hp["batch_size"] % config.Global["micro_batch_size"] == 0
        config.Global["local_batch_size"] = hp["batch_size"]
        config.Global["global_batch_size"] = hp["batch_size"] * config.Distributed["dp_degree"] * config.Distributed["pp_degree"]


class ErnieModule(BasicModule):
    def __init__(self, configs):
        from ....parallel import tp_layers as _tp

        _tp.configure(configs.get("Fused", {}))
        self.nranks = env.world_size()
        self.binary_head = bool(configs.Global.get("binary_head", True))
        super().__init__(configs)
        self.loss_fn = self.get_loss_fn()

    def process_configs(self, configs):
        process_data_configs(configs)
        m = configs.Model
        if m.get("ffn_hidden_size") is None and m.get("intermediate_size") is None:
            m["ffn_hidden_size"] = 4 * m["hidden_size"]
        process_optim_configs(configs)
        return configs

    def _kwargs(self):
        return {k: self.configs.Model[k] for k in _KEYS if k in self.configs.Model and self.configs.Model[k] is not None}

    def get_model(self):
        cfg, d = self.configs, self.configs.Distributed
        hcg = env.get_hcg()
        mp_group = hcg.get_model_parallel_group() if d.mp_degree > 1 else None
        kw = self._kwargs()
        unit = 128 * d.mp_degree
        kw["vocab_size"] = (kw.get("vocab_size", 40000) + unit - 1) // unit * unit
        cfg.Model["vocab_size"] = kw["vocab_size"]
        l = kw.get("num_layers") or kw.get("num_hidden_layers", 12)
        get_model_size(l, kw["hidden_size"], kw["vocab_size"], kw.get("max_position_embeddings", 512))
        dtype, device = _param_dtype(cfg), _device(cfg)
        if d.pp_degree > 1:
            from .pipe import ErnieForPretrainingPipe

            return ErnieForPretrainingPipe(hcg=hcg, mp_group=mp_group, binary_head=self.binary_head, dtype=dtype, device=device, **kw)
        ernie = E.ErnieModel(mp_group=mp_group, dtype=dtype, device=device, fused_tp_comm=bool(cfg.get("Fused", {}).get("tp_comm", False)), **kw)
        model = E.ErnieForPretraining(ernie, kw["vocab_size"], kw.get("hidden_act", "gelu"), self.binary_head)
        if ernie.sequence_parallel:
            from ....parallel.tp_layers import register_sequence_parallel_allreduce_hooks

            register_sequence_parallel_allreduce_hooks(model, cfg.Engine.accumulate_steps, d.get("fuse_sequence_parallel_allreduce", False), mp_group)
        return model

    def get_loss_fn(self):
        d = self.configs.Distributed
        if d.pp_degree > 1:
            return None
        hcg = env.get_hcg()
        return E.ErniePretrainingCriterion(self.binary_head, hcg.get_model_parallel_group() if d.mp_degree > 1 else None)

    def pretreating_batch(self, batch):
        if self.configs.Distributed.pp_degree > 1:
            ids, seg, mask, pos, labels, nsp = batch
            return [(ids, seg, mask), (pos, labels, nsp)]
        return batch

    def forward(self, tokens):
        return self.model(tokens)

    def training_step(self, batch):
        input_ids, segment_ids, input_mask, masked_lm_positions, masked_lm_labels, next_sentence_labels = batch
        scores, rel = self.model(input_ids, segment_ids, None, input_mask, masked_lm_positions)
        out = self.loss_fn(scores, rel, masked_lm_labels, next_sentence_labels if self.binary_head else None)
        if isinstance(out, tuple):
            self._last_parts = (out[1].detach(), out[2].detach())
            return out[0]
        return out

    def training_step_end(self, log_dict):
        speed = 1.0 / log_dict["train_cost"]
        gbs = self.configs.Global.global_batch_size
        seq = self.configs.Data.Train.dataset.get("max_seq_len", self.configs.Data.Train.dataset.get("max_seq_length", 512))
        logger.train("[train] epoch: %d, batch: %d, loss: %.9f, avg_batch_cost: %.5f sec, speed: %.2f step/s, ips_total: %.0f tokens/s, "
                     "ips: %.0f tokens/s, learning rate: %.5e"
                     % (log_dict["epoch"], log_dict["batch"], log_dict["loss"], log_dict["train_cost"], speed, speed * gbs * seq,
                        speed * gbs * seq / max(env.get_data_world_size(), 1), log_dict["lr"]))

    def validation_step(self, batch):
        return self.training_step(batch)

    def validation_step_end(self, log_dict):
        logger.eval("[eval] epoch: %d, batch: %d, loss: %.9f, avg_eval_cost: %.5f sec, speed: %.2f step/s"
                    % (log_dict["epoch"], log_dict["batch"], log_dict["loss"], log_dict["eval_cost"], 1.0 / log_dict["eval_cost"]))

    def test_step(self, batch):
        return self.training_step(batch)

    def test_step_end(self, log_dict):
        logger.eval("[test] epoch: %d, batch: %d, loss: %.9f, avg_test_cost: %.5f sec, speed: %.2f step/s"
                    % (log_dict["epoch"], log_dict["batch"], log_dict["loss"], log_dict["test_cost"], 1.0 / log_dict["test_cost"]))

    def training_epoch_end(self, log_dict):
        logger.info("[Training] epoch: %d, total time: %.5f sec" % (log_dict["epoch"], log_dict["train_cost"]))

    def input_spec(self):
        return [dict(shape=[None, None], name="input_ids", dtype="int64"), dict(shape=[None, None], name="token_type_ids", dtype="int64")]


class ErnieSeqClsModule(ErnieModule):
    def process_configs(self, configs):
        process_optim_configs(configs)
        return configs

    def get_model(self):
        cfg = self.configs
        kw = self._kwargs()
        ernie = E.ErnieModel(dtype=_param_dtype(cfg), device=_device(cfg), **kw)
        return E.ErnieForSequenceClassification(ernie, int(cfg.Model.get("num_classes", 2)), cfg.Model.get("classifier_dropout"))

    def get_loss_fn(self):
        return lambda logits, y: torch.nn.functional.cross_entropy(logits.float(), y.long().reshape(-1))

    def training_step(self, batch):
        if isinstance(batch, dict):
            ids, seg, labels = batch["input_ids"], batch.get("token_type_ids"), batch["labels"]
        else:
            ids, seg, labels = batch[0], batch[1] if len(batch) > 2 else None, batch[-1]
        return self.loss_fn(self.model(ids, seg), labels)

    def validation_step(self, batch):
        """Loss like training, plus running accuracy reported once per evaluation pass (the reference module only trains,
        ernie_module.py:344; a fine-tune without a dev-set score is hard to use)."""
        if isinstance(batch, dict):
            ids, seg, labels = batch["input_ids"], batch.get("token_type_ids"), batch["labels"]
        else:
            ids, seg, labels = batch[0], batch[1] if len(batch) > 2 else None, batch[-1]
        logits = self.model(ids, seg)
        hit = (logits.argmax(-1) == labels.reshape(-1)).sum()
        self._hits = getattr(self, "_hits", 0) + hit
        self._seen = getattr(self, "_seen", 0) + labels.numel()
        return self.loss_fn(logits, labels)

    def validation_epoch_end(self, log_dict):
        seen = getattr(self, "_seen", 0)
        if not seen:
            return
        stat = torch.stack([self._hits.double(), torch.tensor(float(seen), dtype=torch.float64, device=self._hits.device)])
        if env.world_size() > 1 and env.get_data_world_size() > 1:      # every data replica scored its own slice of the dev set
            from ....parallel import collective as C

            C.all_reduce(stat, group=env.get_hcg().get_dp_sharding_group())
        acc = float(stat[0] / stat[1])
        self.best_metric = max(getattr(self, "best_metric", 0.0), acc)
        logger.eval("[Eval] epoch: %d, total time: %.5f sec, accuracy: %.5f, best: %.5f" % (log_dict["epoch"], log_dict["eval_cost"], acc, self.best_metric))
        self._hits, self._seen = 0, 0


ErnieModuleAuto = ErnieModule
ErnieSeqClsModuleAuto = ErnieSeqClsModule
