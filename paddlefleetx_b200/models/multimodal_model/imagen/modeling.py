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

from ....utils.log import logger
from . import unet as U
from .unet import BaseUnet64, SRUnet256, SRUnet1024, Unet64_397M  # noqa: F401  (the reference defines the presets in this module)


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

    def q_sample_from_to(self, x_from, from_t, to_t, noise=None):
        """Re-noise a sample from time ``from_t`` to the noisier time ``to_t`` (the resampling step of RePaint-style inpainting): scale by
        ``alpha_to / alpha_from`` and add exactly the noise that brings the variance up to ``sigma_to^2``, so that the result is distributed like
        ``q_sample(x0, to_t)``.  The reference (imagen/utils.py:448-469) adds ``(sigma_to * alpha - sigma * alpha_to) / alpha`` instead, which
        under-shoots that variance; the exact form is used here."""
        b = x_from.shape[0]
        from_t = torch.full((b,), from_t, device=x_from.device) if isinstance(from_t, float) else from_t
        to_t = torch.full((b,), to_t, device=x_from.device) if isinstance(to_t, float) else to_t
        noise = torch.randn_like(x_from) if noise is None else noise
        a, s = self.alpha_sigma(self.log_snr(from_t).view(-1, 1, 1, 1))
        a_to, s_to = self.alpha_sigma(self.log_snr(to_t).view(-1, 1, 1, 1))
        ratio = a_to / a
        return x_from * ratio + noise * torch.sqrt((s_to ** 2 - (ratio * s) ** 2).clamp(min=0.0))

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
        if text_embed_dim is None and text_encoder_name not in (None, "None", ""):
            text_embed_dim = self._text_tower_width(text_encoder_name)      # reference: default(text_embed_dim, get_encoded_dim(name))
        text_embed_dim = text_embed_dim or 1024
        self.text_embed_dim = text_embed_dim
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
        if self.text_encoder is None and self.text_encoder_name is not None:
            self.text_encoder = self._load_text_encoder(self.text_encoder_name)
        if self.text_encoder is not None:
            for p in self.text_encoder.parameters():
                p.requires_grad = False

    @staticmethod
    def _text_tower_width(name: str) -> Optional[int]:
        try:
            if "deberta" in str(name).lower():
                from ..debertav2.modeling import get_debertav2_encoded_dim

                return get_debertav2_encoded_dim(name)
            from ..t5.modeling import get_encoded_dim

            return get_encoded_dim(name)
        except (FileNotFoundError, KeyError, ImportError):
            return None

    @staticmethod
    def _load_text_encoder(name: str):
        """The frozen text tower stored under the directory ``name`` (``t5/t5-11b``, ``.../deberta-v2-xxlarge``: ``config.json`` + weights), as
        the reference builds it at construction time (imagen/modeling.py:221-241).  A name whose directory is not on this machine gives
        ``None``: training then runs on pre-computed embeddings, and ``sample(texts=...)`` explains what is missing."""
        import os

        if not os.path.isdir(str(name)):
            return None
        low = str(name).lower()
        try:
            if "t5" in low:
                from ..t5.modeling import get_t5_model

                return get_t5_model(name, pretrained=True)
            if "deberta" in low:
                from ..debertav2.modeling import get_debertav2_model

                return get_debertav2_model(name, pretrained=True)
        except FileNotFoundError as exc:
            logger.warning(f"text encoder {name!r} not loaded: {exc}")
            return None
        raise NotImplementedError("Please implement the text encoder.")

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
            out = self.text_encoder(input_ids, attention_mask)
            return getattr(out, "last_hidden_state", out).detach()

    def encode_captions(self, texts, max_length: int = 256):
        """captions -> ``(text_embeds, text_masks)`` through the model's own tokenizer and frozen text tower (fp32, no autocast)."""
        assert self.text_encoder is not None, (f"no text encoder weights for {self.text_encoder_name!r} on this machine: pass text_embeds, or place the "
                                               "encoder directory (config.json + weights) at that path")
        input_ids, attention_mask = self.tokenize_captions(texts, max_length)
        with torch.autocast(device_type=input_ids.device.type, enabled=False):
            return self.encode_text(input_ids, attention_mask), attention_mask

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

    def _p_sample_loop(self, unet, shape, sched, obj, *, text_embeds, text_mask, cond_images, cond_scale, lowres_cond_img, lowres_noise_times,
                       inpaint_images=None, inpaint_masks=None, inpaint_resample_times=5, init_images=None, skip_steps=None):
        """Ancestral sampling of one cascade stage (reference imagen/modeling.py:436-538): optional start from ``init_images`` + noise, optional
        skipped leading steps, self-conditioning on the previous x0 estimate, and RePaint-style inpainting (known pixels are re-imposed at every
        step, each step is repeated ``inpaint_resample_times`` times with re-noising in between).  Returns images in [-1, 1]."""
        b, dev = shape[0], next(unet.parameters()).device
        x = torch.randn(shape, device=dev)
        if init_images is not None:
            x = x + init_images
        x0 = None
        inpaint = inpaint_images is not None and inpaint_masks is not None
        resample = inpaint_resample_times if inpaint else 1
        if inpaint:
            known = resize_image_to(inpaint_images * 2 - 1 if self.auto_normalize_img else inpaint_images, shape[-1])
            m = inpaint_masks if inpaint_masks.dim() == 4 else inpaint_masks[:, None]
            keep = resize_image_to(m.float(), shape[-1]) > 0.5                      # True where the given pixels are kept
        steps = sched.get_sampling_timesteps(b, dev)[int(skip_steps or 0):]
        for t, t_next in steps:
            last = bool((t_next == 0).all())
            for r in reversed(range(resample)):
                if inpaint:
                    noised_known, _, _, _ = sched.q_sample(known, t)
                    x = torch.where(keep, noised_known, x)
                kw = {"self_cond": x0} if getattr(unet, "self_cond", False) else {}
                pred = unet.forward_with_cond_scale(x, sched.log_snr(t), lowres_cond_img=lowres_cond_img, lowres_noise_times=lowres_noise_times,
                                                    text_embeds=text_embeds, text_mask=text_mask, cond_scale=cond_scale,
                                                    **({"cond_images": cond_images} if cond_images is not None else {}), **kw)
                x0 = {"noise": sched.predict_start_from_noise, "v": sched.predict_start_from_v}.get(obj, lambda a, b_, c: c)(x, t, pred)
                x0 = self._threshold(x0)
                mean, var, _ = sched.q_posterior(x0, x, t, t_next)
                noise = torch.randn_like(x) * (t_next > 0).float().view(-1, 1, 1, 1)
                x = mean + var.sqrt() * noise
                if inpaint and r > 0 and not last:
                    x = sched.q_sample_from_to(x, t_next, t)
        x = x.clamp(-1, 1)
        if inpaint:
            x = torch.where(keep, known, x)
        return x

    @torch.no_grad()
    def sample(self, texts=None, text_masks=None, text_embeds=None, cond_images=None, inpaint_images=None, inpaint_masks=None,
               inpaint_resample_times: int = 5, init_images=None, skip_steps=None, batch_size: int = 1, cond_scale=1.0,
               lowres_sample_noise_level: Optional[float] = None, start_at_unet_number: int = 1, start_image_or_video=None,
               stop_at_unet_number: Optional[int] = None, return_all_unet_outputs: bool = False, return_pil_images: bool = False):
        """Run the cascade (reference imagen/modeling.py:541-676).  Conditioning: ``texts`` (encoded by the frozen text tower) or
        ``text_embeds`` (+ ``text_masks``, default: non-zero rows); ``cond_images`` for U-Nets built with ``cond_images_channels``.  Editing:
        ``inpaint_images`` + ``inpaint_masks`` (True / 1 = keep the given pixel), ``init_images`` and ``skip_steps`` (one value or one per stage).
        ``cond_scale`` may be one value or one per stage; ``start_at_unet_number`` > 1 upsamples ``start_image_or_video``.  Returns the last
        stage's images in [0, 1] — all stages as a list with ``return_all_unet_outputs`` (the reference's default there is True), PIL images
        with ``return_pil_images``."""
        was_training = self.training
        self.eval()
        try:
            dev = next(self.parameters()).device
            if cond_images is not None and cond_images.dtype == torch.uint8:
                cond_images = cond_images.float() / 255.0
            if texts is not None and text_embeds is None and self.condition_on_text:
                text_embeds, text_masks = self.encode_captions(texts)
            if self.condition_on_text:
                assert text_embeds is not None, "text or text encodings must be passed into imagen if specified"
                assert text_embeds.shape[-1] == self.text_embed_dim, f"invalid text embedding dimension being passed in (should be {self.text_embed_dim})"
                text_embeds = text_embeds.to(dev)
                text_masks = (text_embeds != 0).any(-1) if text_masks is None else text_masks.to(dev)
                batch_size = text_embeds.shape[0]
            else:
                assert text_embeds is None, "imagen specified not to be conditioned on text, yet it is presented"
            assert (inpaint_images is None) == (inpaint_masks is None), "inpaint images and masks must be both passed in to do inpainting"
            if inpaint_images is not None:
                if not self.condition_on_text and batch_size == 1:
                    batch_size = inpaint_images.shape[0]
                assert inpaint_images.shape[0] == batch_size, "number of inpainting images must be equal to the specified batch size on sample"
            n = len(self.unets)
            per_stage = lambda v: tuple(v) if isinstance(v, (tuple, list)) else (v,) * n      # noqa: E731
            cond_scales, inits, skips = per_stage(cond_scale), per_stage(init_images), per_stage(skip_steps)
            noise_level = self.lowres_sample_noise_level if lowres_sample_noise_level is None else lowres_sample_noise_level
            img = None
            if start_at_unet_number > 1:
                assert stop_at_unet_number is None or start_at_unet_number <= stop_at_unet_number
                assert start_image_or_video is not None, "starting image or video must be supplied if only doing upscaling"
            if start_image_or_video is not None:          # also feeds a stand-alone super-resolution stage (one low-res-conditioned U-Net)
                prev = self._prev_size(start_at_unet_number - 1)
                img = start_image_or_video.to(dev) if prev is None else resize_image_to(start_image_or_video.to(dev), prev)
            outputs = []
            for k, (unet, sched, obj, size) in enumerate(zip(self.unets, self.noise_schedulers, self.pred_objectives, self.image_sizes)):
                if k + 1 < start_at_unet_number:
                    continue
                lowres, lowres_t = None, None
                if unet.lowres_cond:
                    assert img is not None, "a low-resolution-conditioned stage needs the previous stage's output (or start_image_or_video)"
                    lowres = resize_image_to(img, size)
                    lowres = lowres * 2 - 1 if self.auto_normalize_img else lowres
                    lt = self.lowres_noise_schedule.get_times(batch_size, noise_level, dev)
                    lowres, _, _, _ = self.lowres_noise_schedule.q_sample(lowres, lt)
                    lowres_t = self.lowres_noise_schedule.log_snr(lt)
                init = inits[k]
                if init is not None:
                    init = resize_image_to(init.to(dev) * 2 - 1 if self.auto_normalize_img else init.to(dev), size)
                x = self._p_sample_loop(unet, (batch_size, self.in_chans, size, size), sched, obj, text_embeds=text_embeds, text_mask=text_masks,
                                        cond_images=None if cond_images is None else cond_images.to(dev), cond_scale=cond_scales[k],
                                        lowres_cond_img=lowres, lowres_noise_times=lowres_t,
                                        inpaint_images=None if inpaint_images is None else inpaint_images.to(dev),
                                        inpaint_masks=None if inpaint_masks is None else inpaint_masks.to(dev),
                                        inpaint_resample_times=inpaint_resample_times, init_images=init, skip_steps=skips[k])
                img = (x + 1) * 0.5 if self.auto_normalize_img else x
                outputs.append(img)
                if stop_at_unet_number is not None and stop_at_unet_number == k + 1:
                    break
        finally:
            self.train(was_training)
        if return_pil_images:
            from PIL import Image

            to_pil = lambda t: Image.fromarray((t.clamp(0, 1).permute(1, 2, 0).float().cpu().numpy() * 255 + 0.5).astype("uint8").squeeze())   # noqa: E731
            pil = [[to_pil(im) for im in stage] for stage in outputs]
            return pil if return_all_unet_outputs else pil[-1]
        return outputs if return_all_unet_outputs else outputs[-1]


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
