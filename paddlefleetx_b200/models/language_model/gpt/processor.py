"""Logits processors for generation (reference gpt/dygraph/processor.py:22-200): ``MinLengthLogitsProcessor``,
``RepetitionPenaltyLogitsProcessor``, ``HammingDiversityLogitsProcessor``, ``ForcedBOSTokenLogitsProcessor``,
``ForcedEOSTokenLogitsProcessor`` and the ``LogitsProcessorList`` container.  All operate in place-free style on
``[batch, vocab]`` logits and keep everything on the device."""
from __future__ import annotations


import torch


class LogitsProcessor:
    def __call__(self, input_ids: torch.Tensor, logits: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError


class LogitsProcessorList(list):
    def __call__(self, input_ids, logits, **kwargs):
        for proc in self:
            logits = proc(input_ids, logits, **kwargs) if kwargs else proc(input_ids, logits)
        return logits


class MinLengthLogitsProcessor(LogitsProcessor):
    def __init__(self, min_length: int, eos_token_id: int):
        if min_length < 0:
            raise ValueError("`min_length` should be a non-negative integer")
        self.min_length, self.eos_token_id = int(min_length), int(eos_token_id)

    def __call__(self, input_ids, logits):
        if input_ids.shape[-1] < self.min_length:
            logits = logits.clone()
            logits[:, self.eos_token_id] = float("-inf")
        return logits


class RepetitionPenaltyLogitsProcessor(LogitsProcessor):
    def __init__(self, penalty: float):
        if not penalty > 0:
            raise ValueError("`penalty` has to be a strictly positive float")
        self.penalty = float(penalty)

    def __call__(self, input_ids, logits):
        score = logits.gather(1, input_ids)
        score = torch.where(score < 0, score * self.penalty, score / self.penalty)
        return logits.scatter(1, input_ids, score)


class HammingDiversityLogitsProcessor(LogitsProcessor):
    def __init__(self, diversity_rate: float, num_beams: int, num_beam_groups: int):
        self.diversity_rate, self.num_beams, self.num_sub_beams = float(diversity_rate), num_beams, num_beams // num_beam_groups

    def __call__(self, input_ids, logits, current_tokens=None, beam_group_idx: int = 0):
        if current_tokens is None or beam_group_idx == 0:
            return logits
        batch = current_tokens.shape[0] // self.num_beams
        start = beam_group_idx * self.num_sub_beams
        end = min(start + self.num_sub_beams, self.num_beams)
        size = end - start
        vocab = logits.shape[-1]
        logits = logits.clone()
        for b in range(batch):
            prev = current_tokens[b * self.num_beams:b * self.num_beams + start]
            freq = torch.bincount(prev, minlength=vocab).to(logits.dtype)
            logits[b * size:(b + 1) * size] -= self.diversity_rate * freq
        return logits


class ForcedBOSTokenLogitsProcessor(LogitsProcessor):
    def __init__(self, forced_bos_token_id: int):
        self.token = int(forced_bos_token_id)

    def __call__(self, input_ids, logits):
        if input_ids.shape[-1] == 1:
            out = torch.full_like(logits, float("-inf"))
            out[:, self.token] = 0
            return out
        return logits


class ForcedEOSTokenLogitsProcessor(LogitsProcessor):
    def __init__(self, max_length: int, forced_eos_token_id: int):
        self.max_length, self.token = int(max_length), int(forced_eos_token_id)

    def __call__(self, input_ids, logits):
        if input_ids.shape[-1] == self.max_length - 1:
            out = torch.full_like(logits, -1e9)
            out[:, self.token] = 0
            return out
        return logits
