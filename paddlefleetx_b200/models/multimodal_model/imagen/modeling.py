"""``ImagenModel`` — cascaded continuous-time Gaussian diffusion (reference multimodal_model/imagen/modeling.py:36-1026,
utils.py:26-489): cosine / linear log-SNR schedules, ``q_sample``, noise / x0 / v objectives, conditioning dropout for
classifier-free guidance, low-resolution noise augmentation for the SR stages, dynamic thresholding and ancestral sampling;
``ImagenCriterion`` = per-sample mean loss weighted by ``(k + exp(log_snr))^-gamma`` (p2 weighting)."""
from __future__ import annotations

import math
from typing import Optional, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import unet as U


def log_snr_cosine(t, s: float = 0.008):
    return -torch.log(torch.clamp(torch.cos((t + s) / (1 + s) * math.pi * 0.5) ** -2 - 1, min=1e-5))


def log_snr_linear(t):
    return -torch.log(torch.expm1(1e-4 + 10 * t ** 2))


class GaussianDiffusionContinuousTimes(nn.Module):
    def __init__(self, noise_schedule: str = "cosine", timesteps: int = 1000):
        super().__init__()
        self.log_snr = log_snr_cosine if noise_schedule == "cosine" else log_snr_linear
        self.num_timesteps = timesteps

    def get_times(self, batch, noise_level, device):
        return torch.full((batch,), noise_level, device=device, dtype=torch.float32)

    def sample_random_times(self, batch, device):
        return torch.rand(batch, device=device)

    def get_sampling_timesteps(self, batch, device):
        times = torch.linspace(1.0, 0.0, self.num_timesteps + 1, device=device)
        times = times[None].expand(batch, -1)
        return list(zip(times[:, :-1].unbind(1), times[:, 1:].unbind(1)))

    @staticmethod
    def alpha_sigma(log_snr):
        return torch.sqrt(torch.sigmoid(log_snr)), torch.sqrt(torch.sigmoid(-log_snr))

    def q_sample(self, x0, t, noise=None):
        noise = torch.randn_like(x0) if noise is None else noise
        ls = self.log_snr(t).view(-1, 1, 1, 1)
        a, s = self.alpha_sigma(ls)
        return a * x0 + s * noise, ls.view(-1), a, s

    def q_posterior(self, x0, xt, t, t_next):
        ls, ln = self.log_snr(t).view(-1, 1, 1, 1), self.log_snr(t_next).view(-1, 1, 1, 1)
        c = -torch.expm1(ls - ln)
        a, s = self.alpha_sigma(ls)
        an, sn = self.alpha_sigma(ln)
        mean = an * (xt * (1 - c) / a + c * x0)
        var = sn ** 2 * c
        return mean, var, torch.log(var.clamp(min=1e-20))

    def predict_start_from_noise(self, xt, t, noise):
        a, s = self.alpha_sigma(self.log_snr(t).view(-1, 1, 1, 1))
        return (xt - s * noise) / a.clamp(min=1e-8)

    def predict_start_from_v(self, xt, t, v):
        a, s = self.alpha_sigma(self.log_snr(t).view(-1, 1, 1, 1))
        return a * xt - s * v

    def calculate_v(self, x0, t, noise):
        a, s = self.alpha_sigma(self.log_snr(t).view(-1, 1, 1, 1))
        return a * noise - s * x0


def resize_image_to(img, size):
    return img if img.shape[-1] == size else F.interpolate(img, size=(size, size), mode="nearest" if size < img.shape[-1] else "bilinear",
                                                           **({} if size < img.shape[-1] else {"align_corners": False}))


class ImagenModel(nn.Module):
    """``unets``: one U-Net per cascade stage (modules, or ``{"name": <preset>, ...}`` dicts).  ``image_sizes[k]`` is stage k's
    resolution; a stage trained on its own with low-resolution conditioning lists the conditioning resolution after its own —
    ``image_sizes=(256, 64)`` with one ``lowres_cond`` U-Net, the form of the reference's super-resolution recipes
    (modeling.py:1000-1026).  Every keyword of the reference constructor is accepted; unknown ones raise."""

    def __init__(self, unets: Sequence, image_sizes: Sequence[int] = (64,), text_encoder_name: Optional[str] = None, text_embed_dim: Optional[int] = 1024,
                 in_chans: Optional[int] = None, channels: int = 3, timesteps: int = 1000, cond_drop_prob: float = 0.1, noise_schedules="cosine",
                 pred_objectives="noise", lowres_noise_schedule: str = "linear", lowres_sample_noise_level: float = 0.2,
                 per_sample_random_aug_noise_level: bool = False, condition_on_text: bool = True, auto_normalize_img: bool = True,
                 dynamic_thresholding: bool = True, dynamic_thresholding_percentile: float = 0.95, p2_loss_weight_gamma: float = 0.5,
                 p2_loss_weight_k: float = 1.0, unet_number: Optional[int] = None, only_train_unet_number: Optional[int] = None, is_sr: bool = False,
                 text_encoder=None):
        super().__init__()
        if isinstance(unets, nn.Module):
            unets = [unets]
        text_embed_dim = text_embed_dim or 1024
        built = [u if isinstance(u, nn.Module) else getattr(U, u["name"])(**{k: v for k, v in u.items() if k != "name"}, text_embed_dim=text_embed_dim)
                 for u in unets]
        in_chans = channels if in_chans is None else in_chans
        for k, u in enumerate(built):
            # every stage agrees with the cascade on channels / text conditioning; stages after the first condition on the previous stage's image
            want_lowres = u.lowres_cond or k > 0
            if (u.channels != in_chans or u.cond_on_text != condition_on_text or u._locals["text_embed_dim"] != text_embed_dim
                    or u.lowres_cond != want_lowres):
                built[k] = u.__class__(**{**u._locals, "channels": in_chans, "channels_out": u._locals["channels_out"], "cond_on_text": condition_on_text,
                                          "text_embed_dim": text_embed_dim, "lowres_cond": want_lowres})
        self.unets = nn.ModuleList(built)
        n = len(self.unets)
        self.image_sizes = list(image_sizes)
        assert len(self.image_sizes) >= n, "one image size per U-Net (plus the conditioning size of a stand-alone super-resolution stage)"
        ns = [noise_schedules] * n if isinstance(noise_schedules, str) else list(noise_schedules)
        self.noise_schedulers = nn.ModuleList([GaussianDiffusionContinuousTimes(s, timesteps) for s in ns])
        self.lowres_noise_schedule = GaussianDiffusionContinuousTimes(lowres_noise_schedule)
        self.pred_objectives = [pred_objectives] * n if isinstance(pred_objectives, str) else list(pred_objectives)
        self.cond_drop_prob, self.lowres_sample_noise_level = cond_drop_prob, lowres_sample_noise_level
        self.per_sample_random_aug_noise_level = per_sample_random_aug_noise_level
        self.condition_on_text, self.auto_normalize_img, self.is_sr = condition_on_text, auto_normalize_img, is_sr
        self.dynamic_thresholding, self.dt_pct = dynamic_thresholding, dynamic_thresholding_percentile
        self.p2_gamma, self.p2_k = p2_loss_weight_gamma, p2_loss_weight_k
        self.unet_number = unet_number or only_train_unet_number or 1
        self.in_chans = in_chans
        self.text_encoder = text_encoder           # frozen T5 / DeBERTa, optional (pre-computed embeddings are accepted too)
        self.text_encoder_name = None if text_encoder_name in (None, "None", "") else text_encoder_name
        from ....data.tokenizers import get_text_tokenizer

        self.tokenizer = get_text_tokenizer(self.text_encoder_name)      # None when the vocabulary is not on this (offline) machine
        if self.text_encoder is not None:
            for p in self.text_encoder.parameters():
                p.requires_grad = False

    def _prev_size(self, k: int) -> Optional[int]:
        """Resolution of the image stage ``k`` is conditioned on (None for an unconditioned base stage)."""
        if k > 0:
            return self.image_sizes[k - 1]
        if self.unets[k].lowres_cond and len(self.image_sizes) > len(self.unets):
            return self.image_sizes[len(self.unets)]
        return None

    def encode_text(self, input_ids, attention_mask):
        with torch.no_grad():
            self.text_encoder.eval()
            return self.text_encoder(input_ids, attention_mask).detach()

    def tokenize_captions(self, texts, max_length: int = 256):
        """captions -> (input_ids, attention_mask) on the model's device with the text tower's own tokenizer."""
        assert self.tokenizer is not None, f"no local vocabulary for text encoder {self.text_encoder_name!r} (set PFX_TOKENIZER_DIR)"
        enc = self.tokenizer.batch_encode_plus(list(texts), return_tensors="pt", padding="longest", max_length=max_length, truncation=True)
        dev = next(self.parameters()).device
        return enc.input_ids.to(dev), enc.attention_mask.to(dev)

    def p_losses(self, unet, x0, times, scheduler, objective, text_embeds=None, text_mask=None, lowres_cond_img=None, lowres_aug_times=None, noise=None):
        norm = (lambda im: im * 2 - 1) if self.auto_normalize_img else (lambda im: im)
        x0 = norm(x0)
        noise = torch.randn_like(x0) if noise is None else noise
        xt, log_snr, _, _ = scheduler.q_sample(x0, times, noise)
        lowres_noisy = None
        if lowres_cond_img is not None:
            lowres_noisy, _, _, _ = self.lowres_noise_schedule.q_sample(norm(lowres_cond_img), lowres_aug_times)
        pred = unet(xt, scheduler.log_snr(times), lowres_cond_img=lowres_noisy,
                    lowres_noise_times=None if lowres_aug_times is None else self.lowres_noise_schedule.log_snr(lowres_aug_times),
                    text_embeds=text_embeds, text_mask=text_mask, cond_drop_prob=self.cond_drop_prob)
        target = {"noise": noise, "x_start": x0, "v": scheduler.calculate_v(x0, times, noise)}[objective]
        return pred, target, log_snr, self.p2_gamma

    def forward(self, images, text_embeds=None, text_masks=None, input_ids=None, attention_mask=None, unet_number: Optional[int] = None):
        k = (unet_number or self.unet_number) - 1
        unet, sched, obj, size = self.unets[k], self.noise_schedulers[k], self.pred_objectives[k], self.image_sizes[k]
        if text_embeds is None and input_ids is not None and self.text_encoder is not None:
            text_embeds, text_masks = self.encode_text(input_ids, attention_mask), attention_mask
        b = images.shape[0]
        times = sched.sample_random_times(b, images.device)
        lowres, aug_t = None, None
        prev = self._prev_size(k)
        if prev is not None:
            lowres = resize_image_to(resize_image_to(images, prev), size)
            aug_t = self.lowres_noise_schedule.sample_random_times(b if self.per_sample_random_aug_noise_level else 1, images.device).expand(b)
        if not self.condition_on_text:
            text_embeds = text_masks = None
        x0 = resize_image_to(images, size)
        return self.p_losses(unet, x0, times, sched, obj, text_embeds, text_masks, lowres, aug_t)

    # ------------------------------------------------------------------ sampling
    def _threshold(self, x0):
        if not self.dynamic_thresholding:
            return x0.clamp(-1, 1)
        s = torch.quantile(x0.flatten(1).abs().float(), self.dt_pct, dim=-1).clamp(min=1.0).view(-1, 1, 1, 1)
        return x0.clamp(-s, s) / s

    @torch.no_grad()
    def sample(self, text_embeds=None, text_masks=None, batch_size: int = 1, cond_scale: float = 1.0, stop_at_unet_number: Optional[int] = None):
        dev = next(self.parameters()).device
        img = None
        outputs = []
        for k, (unet, sched, obj, size) in enumerate(zip(self.unets, self.noise_schedulers, self.pred_objectives, self.image_sizes)):
            lowres, lowres_t = None, None
            if k > 0:
                lowres = resize_image_to(img, size) * 2 - 1
                lt = self.lowres_noise_schedule.get_times(batch_size, self.lowres_sample_noise_level, dev)
                lowres, _, _, _ = self.lowres_noise_schedule.q_sample(lowres, lt)
                lowres_t = self.lowres_noise_schedule.log_snr(lt)
            x = torch.randn(batch_size, self.in_chans, size, size, device=dev)
            for t, t_next in sched.get_sampling_timesteps(batch_size, dev):
                pred = unet.forward_with_cond_scale(x, sched.log_snr(t), lowres_cond_img=lowres, lowres_noise_times=lowres_t, text_embeds=text_embeds,
                                                    text_mask=text_masks, cond_scale=cond_scale)
                x0 = {"noise": sched.predict_start_from_noise, "v": sched.predict_start_from_v}.get(obj, lambda a, b_, c: c)(x, t, pred)
                x0 = self._threshold(x0)
                mean, var, _ = sched.q_posterior(x0, x, t, t_next)
                noise = torch.randn_like(x) * (t_next > 0).float().view(-1, 1, 1, 1)
                x = mean + var.sqrt() * noise
            img = (x.clamp(-1, 1) + 1) * 0.5
            outputs.append(img)
            if stop_at_unet_number is not None and stop_at_unet_number == k + 1:
                break
        return outputs[-1]


class ImagenCriterion(nn.Module):
    def __init__(self, name: str = "mse_loss", p2_loss_weight_k: float = 1.0):
        super().__init__()
        self.fn = {"mse_loss": F.mse_loss, "l1_loss": F.l1_loss, "smooth_l1_loss": F.smooth_l1_loss}[name]
        self.k = p2_loss_weight_k

    def forward(self, pred, target, log_snr, p2_loss_weight_gamma):
        losses = self.fn(pred.float(), target.float(), reduction="none").flatten(1).mean(1)
        if p2_loss_weight_gamma > 0:
            losses = losses * (self.k + log_snr.exp()) ** -p2_loss_weight_gamma
        return losses.mean()


def _cascade(unet_builders, sizes):
    """Preset factory with the keyword surface of the reference constructors (modeling.py:952-1026): ``use_recompute`` goes to the
    U-Nets, ``lowres_cond`` to a stand-alone stage, the rest to ``ImagenModel``."""
    def build(**kw):
        use_recompute = bool(kw.pop("use_recompute", False))
        kw.pop("recompute_granularity", None)
        lowres_cond = bool(kw.pop("lowres_cond", False))
        ted = kw.get("text_embed_dim") or 1024
        unets = [b(text_embed_dim=ted, use_recompute=use_recompute, **({"lowres_cond": True} if (lowres_cond and len(unet_builders) == 1) else {}))
                 for b in unet_builders]
        return ImagenModel(unets, image_sizes=sizes if (lowres_cond or len(sizes) == len(unets)) else sizes[:len(unets)], **kw)
    return build


imagen_397M_text2im_64 = _cascade([U.Unet64_397M], (64,))
imagen_text2im_64 = _cascade([U.BaseUnet64], (64,))
imagen_2B_text2im_64 = imagen_text2im_64
imagen_text2im_64_debertav2 = _cascade([lambda **k: U.BaseUnet64(dim=360, **k)], (64,))
imagen_text2im_64_SR256 = _cascade([U.BaseUnet64, U.SRUnet256], (64, 256))
imagen_SR256 = _cascade([U.SRUnet256], (256, 64))
imagen_SR1024 = _cascade([lambda **k: U.SRUnet1024(dim=128, **k)], (1024, 256))
