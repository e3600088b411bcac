"""``GPTGenerationModule`` (reference language_module.py:490-597): builds ``GPTForGeneration`` from the ``Generation:``
YAML block, left-pads prompts, ``generate(text)`` returns decoded strings.  Hybrid generation = single card or pure
data parallel, as in the reference (language_module.py:525-526)."""
from __future__ import annotations

from typing import List, Union

import torch

from ...data.tokenizers import GPTTokenizer
from ...distributed.apis import env
from ...utils.log import logger
from .gpt import model as gpt
from .gpt.generation import GPTForGeneration
from .language_module import LanguageModule, _device, _param_dtype, model_kwargs


class GPTGenerationModule(LanguageModule):
    def __init__(self, configs):
        self.generation_cfgs = configs.Generation
        super().__init__(configs)

    def process_configs(self, configs):
        return configs

    def _tokenizer(self):
        vocab_dir = self.configs.Generation.get("vocab_dir") or self.configs.Model.get("vocab_dir") or "gpt2"
        try:
            return GPTTokenizer.from_pretrained(vocab_dir)
        except FileNotFoundError as e:
            logger.warning(f"{e}  -> falling back to the byte-level vocabulary")
            return GPTTokenizer.byte_fallback()

    def get_model(self):
        cfg = self.configs
        d = cfg.Distributed
        assert d.mp_degree == 1 and d.pp_degree == 1 and d.sharding.sharding_degree == 1, \
            "generation runs on a single card or with pure data parallelism"
        self.tokenizer = self._tokenizer()
        self.tokenizer.padding_side = "left"
        kw = model_kwargs(cfg)
        kw["vocab_size"] = gpt.vocab_size_with_padding(cfg.Model.get("vocab_size", len(self.tokenizer)), cfg.Model.get("vocab_size_divisible_unit", 128), 1)
        kw.pop("use_recompute", None)
        core = gpt.GPTModel(dtype=_param_dtype(cfg), device=_device(cfg), use_recompute=False, **kw)
        gen = dict(self.generation_cfgs)
        gen.setdefault("eos_token_id", self.tokenizer.eos_token_id)
        gen.setdefault("pad_token_id", self.tokenizer.eos_token_id)
        gen["max_dec_len"] = min(int(gen.get("max_dec_len", 20)), 512)       # reference clamps to 512
        return GPTForGeneration(core, gen)

    def get_loss_fn(self):
        return None

    def adjust_length_to_model(self, length: int, max_sequence_length: int) -> int:
        if length < 0 or length > max_sequence_length:
            return max_sequence_length
        return length

    def left_padding(self, inputs: List[List[int]], pad_id: int):
        mx = max(len(x) for x in inputs)
        ids = [[pad_id] * (mx - len(x)) + x for x in inputs]
        mask = [[0] * (mx - len(x)) + [1] * len(x) for x in inputs]
        return ids, mask

    @torch.no_grad()
    def generate(self, input_text: Union[str, List[str]], seed=None) -> List[str]:
        return self(input_text, seed=seed)

    @torch.no_grad()
    def forward(self, input_text, seed=None):
        texts = [input_text] if isinstance(input_text, str) else list(input_text)
        enc = [self.tokenizer.encode(t) for t in texts]
        ids, mask = self.left_padding(enc, self.tokenizer.eos_token_id)
        dev = next(self.model.parameters()).device
        ids_t = torch.tensor(ids, dtype=torch.long, device=dev)
        mask_t = torch.tensor(mask, dtype=torch.long, device=dev)
        out_ids, _ = self.model.generate(ids_t, attention_mask=mask_t, seed=seed)
        results = []
        for i, row in enumerate(out_ids.tolist()):
            text = self.tokenizer.decode(row, skip_special_tokens=True)
            results.append(texts[i % len(texts)] + text)
        return results

    def input_spec(self):
        return [dict(shape=[None, None], name="input_ids", dtype="int64")]
