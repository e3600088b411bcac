"""``GPTEvalModule`` — offline evaluation: WikiText perplexity (overlapping windows) and LAMBADA cloze accuracy
(reference language_module.py:600-733; formulas in SURVEY §2.7): ppl = exp(min(20, loss_sum / (N_tok - 1))),
adjusted ppl rescales by (N_tok - 1) / (N_orig - 1); LAMBADA counts a sample correct iff every target piece is arg-max."""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from ...utils.log import logger
from .gpt import model as gpt
from .language_module import LanguageModule, _device, _param_dtype, model_kwargs


class GPTEvalModule(LanguageModule):
    def __init__(self, configs):
        self.eval_cfgs = configs.Offline_Eval
        super().__init__(configs)
        self.post_process_configs()
        self.first_step = True
        self.total_score = 0.0
        self.score_name = "loss" if not self.eval_cfgs.get("cloze_eval", False) else "number correct"

    def process_configs(self, configs):
        return configs

    def post_process_configs(self):
        ev = self.eval_cfgs
        ds = self.configs.Data.setdefault("Eval", {}).setdefault("dataset", {}) if "Data" in self.configs else {}
        if isinstance(ds, dict):
            ds.update(name="Lambada_Eval_Dataset" if ev.get("cloze_eval", False) else "LM_Eval_Dataset", input_dir=ev.get("eval_path"),
                      max_seq_len=ev.get("max_seq_len", 1024))
            if not ev.get("cloze_eval", False):
                ds["overlapping_eval"] = ev.get("overlapping_eval", 32)

    def get_model(self):
        cfg = self.configs
        kw = model_kwargs(cfg)
        kw["use_flash_attn"] = False        # explicit mask path, like the reference eval
        core = gpt.GPTModel(dtype=_param_dtype(cfg), device=_device(cfg), **kw)
        return gpt.GPTForPretraining(core)

    def get_loss_fn(self):
        return None

    def forward(self, tokens, ids, mask):
        return self.model(tokens, ids, mask)

    def validation_step(self, batch):
        tokens, loss_mask, attention_mask, position_ids, labels, info = batch
        add_mask = (1.0 - attention_mask.float()) * -1e4
        preds = self(tokens, position_ids, add_mask).float()
        if not self.eval_cfgs.get("cloze_eval", False):
            if self.first_step:
                self.num_original_tokens, self.num_tokenized_tokens = int(info[0][0]), int(info[0][1])
            ce = F.cross_entropy(preds.reshape(-1, preds.shape[-1]), labels.reshape(-1), reduction="none").view_as(labels)
            loss = (ce * loss_mask).sum()
            return loss
        if self.first_step:
            self.num_examples = int(info[0][0])
        outputs = preds.argmax(-1)
        acc = ((outputs == labels) | (loss_mask == 0)).all(-1).float().sum()
        return acc

    def validation_step_end(self, log_dict):
        self.first_step = False
        self.total_score += float(log_dict["loss"]) * self.configs.Engine.get("logging_freq", 1)
        logger.eval("[eval] epoch: %d, batch: %d, %s: %.9f, speed: %.2f step/s"
                    % (log_dict["epoch"], log_dict["batch"], self.score_name, self.total_score, 1.0 / max(log_dict["eval_cost"], 1e-9)))

    def validation_epoch_end(self, log_dict):
        if not self.eval_cfgs.get("cloze_eval", False):
            total_loss = self.total_score
            ppl = math.exp(min(20, total_loss / max(self.num_tokenized_tokens - 1, 1)))
            ratio = (self.num_tokenized_tokens - 1) / max(self.num_original_tokens - 1, 1)
            adj = math.exp(min(20, total_loss / max(self.num_tokenized_tokens - 1, 1) * ratio))
            s = "validation results on {} | avg loss: {:.4E} | ppl: {:.4E} | adjusted ppl: {:.4E} | token ratio: {} |".format(
                self.eval_cfgs.get("eval_path"), total_loss / max(self.num_tokenized_tokens - 1, 1), ppl, adj, ratio)
        else:
            n = getattr(self, "num_examples", 1)
            s = "validation results on {} | number correct: {:.4E} | total examples: {:.4E} | avg accuracy: {:.4E}".format(
                self.eval_cfgs.get("eval_path"), self.total_score, n, self.total_score / max(n, 1))
        logger.eval(s)
        self.last_summary = s

    def input_spec(self):
        s = self.eval_cfgs.get("max_seq_len", 1024)
        return [dict(shape=[None, s], name="tokens", dtype="int64"), dict(shape=[None, s], name="ids", dtype="int64")]
