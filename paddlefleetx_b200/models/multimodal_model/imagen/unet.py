"""Imagen U-Net — the full efficient-U-Net family of the reference (multimodal_model/imagen/unet.py:33-1562, presets in
modeling.py:36-92), written for PyTorch.

Conditioning: learned-sinusoidal log-SNR embedding -> time hiddens (+ low-resolution noise-level hiddens for the super-resolution
stages) feed every ResNet block as scale/shift; time tokens, Perceiver-resampled text tokens and a mean-pooled text hidden feed
the cross-attention layers; classifier-free guidance through learned null text embeddings.  Structure per resolution:
[pre-downsample (memory-efficient variant)] -> ResNet block with (linear) cross-attention -> n ResNet blocks with global-context
gating -> multi-query self-attention transformer block (or linear attention, or nothing) -> [post-downsample]; mirrored on the way up
with scaled skip connections, pixel-shuffle (or nearest + conv) upsampling, an optional combiner over all upsampling feature maps,
an optional residual from the initial convolution, final ResNet block and a zero-initialised output convolution.

Every constructor option of the reference is honoured (an unknown keyword raises — nothing is silently dropped), so the four
presets build the same networks: ``Unet64_397M`` (397 M parameters), ``BaseUnet64`` (2 B), ``SRUnet256``, ``SRUnet1024``.
"""
from __future__ import annotations

import math
from functools import partial
from typing import Optional, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.utils.checkpoint import checkpoint


def _exists(v) -> bool:
    return v is not None


def _default(v, d):
    return v if v is not None else (d() if callable(d) else d)


def _cast_tuple(v, n: int) -> tuple:
    if isinstance(v, (list, tuple)):
        return tuple(v)
    return (v,) * n


def _l2norm(t):
    return F.normalize(t, dim=-1)


def prob_mask_like(shape, prob: float, device) -> torch.Tensor:
    if prob == 1:
        return torch.ones(shape, dtype=torch.bool, device=device)
    if prob == 0:
        return torch.zeros(shape, dtype=torch.bool, device=device)
    return torch.rand(shape, device=device) < prob


def resize_image_to(img: torch.Tensor, size: int) -> torch.Tensor:
    return img if img.shape[-1] == size else F.interpolate(img, size=(size, size), mode="nearest")


# ------------------------------------------------------------------------------------------ small pieces
class GainLayerNorm(nn.Module):
    """Layer norm with a gain and no bias, over ``dim`` (-1 = features of a token, -3 = channels of a feature map); the ``stable`` variant
    divides by the running maximum first (reference LayerNorm / ChanLayerNorm, unet.py:33-56)."""

    def __init__(self, feats: int, stable: bool = False, dim: int = -1):
        super().__init__()
        self.stable, self.dim = stable, dim
        self.g = nn.Parameter(torch.ones(feats, *((1,) * (-dim - 1))))

    def forward(self, x):
        if self.stable:
            x = x / x.amax(dim=self.dim, keepdim=True).detach()
        eps = 1e-5 if x.dtype == torch.float32 else 1e-3
        var = x.var(dim=self.dim, unbiased=False, keepdim=True)
        mean = x.mean(dim=self.dim, keepdim=True)
        return (x - mean) * (var + eps).rsqrt() * self.g.to(x.dtype)


ChanLayerNorm = partial(GainLayerNorm, dim=-3)


LayerNorm = GainLayerNorm          # the reference's name (imagen/unet.py:33-54)


class Residual(nn.Module):
    """``fn(x, **kwargs) + x`` (reference imagen/unet.py:60-66)."""

    def __init__(self, fn):
        super().__init__()
        self.fn = fn

    def forward(self, x, **kwargs):
        return self.fn(x, **kwargs) + x


class Always(nn.Module):
    def __init__(self, value):
        super().__init__()
        self.value = value

    def forward(self, *args, **kwargs):
        return self.value


class PassThrough(nn.Module):
    """Identity that tolerates the (x, context) call of the attention slots."""

    def forward(self, x, *args, **kwargs):
        return x


class Parallel(nn.Module):
    def __init__(self, *fns):
        super().__init__()
        self.fns = nn.ModuleList(fns)

    def forward(self, x):
        return sum(fn(x) for fn in self.fns)


class SinusoidalPosEmb(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.dim = dim

    def forward(self, t):
        half = self.dim // 2
        freqs = torch.exp(torch.arange(half, device=t.device, dtype=torch.float32) * -(math.log(10000) / max(half - 1, 1)))
        args = t.float()[:, None] * freqs[None]
        return torch.cat([args.sin(), args.cos()], -1)


class LearnedSinusoidalPosEmb(nn.Module):
    """[x, sin(2 pi w x), cos(2 pi w x)] with learned frequencies w: ``dim + 1`` features (unet.py:363-380)."""

    def __init__(self, dim: int):
        super().__init__()
        assert dim % 2 == 0
        self.weights = nn.Parameter(torch.randn(dim // 2))

    def forward(self, x):
        x = x[:, None].float()
        freqs = x * self.weights[None, :].float() * 2 * math.pi
        return torch.cat((x, freqs.sin(), freqs.cos()), dim=-1)


def FeedForward(dim: int, mult: float = 2.0) -> nn.Sequential:
    hidden = int(dim * mult)
    return nn.Sequential(GainLayerNorm(dim), nn.Linear(dim, hidden, bias=False), nn.GELU(), GainLayerNorm(hidden), nn.Linear(hidden, dim, bias=False))


def ChanFeedForward(dim: int, mult: float = 2.0) -> nn.Sequential:
    hidden = int(dim * mult)
    return nn.Sequential(ChanLayerNorm(dim), nn.Conv2d(dim, hidden, 1, bias=False), nn.GELU(), ChanLayerNorm(hidden), nn.Conv2d(hidden, dim, 1, bias=False))


def Upsample(dim: int, dim_out: Optional[int] = None) -> nn.Sequential:
    return nn.Sequential(nn.Upsample(scale_factor=2, mode="nearest"), nn.Conv2d(dim, _default(dim_out, dim), 3, padding=1))


class PixelShuffleUpsample(nn.Module):
    """1x1 conv to 4x channels -> SiLU -> pixel shuffle; the four sub-pixel filters start identical, which removes the checkerboard
    pattern at initialisation (unet.py:314-340)."""

    def __init__(self, dim: int, dim_out: Optional[int] = None):
        super().__init__()
        dim_out = _default(dim_out, dim)
        conv = nn.Conv2d(dim, dim_out * 4, 1)
        self.net = nn.Sequential(conv, nn.SiLU(), nn.PixelShuffle(2))
        with torch.no_grad():
            w = torch.empty(dim_out, dim, 1, 1)
            nn.init.kaiming_uniform_(w)
            conv.weight.copy_(w.repeat_interleave(4, dim=0))
            conv.bias.zero_()

    def forward(self, x):
        return self.net(x)


class _SpaceToDepth(nn.Module):
    def forward(self, x):
        b, c, h, w = x.shape
        x = x.view(b, c, h // 2, 2, w // 2, 2).permute(0, 1, 3, 5, 2, 4)
        return x.reshape(b, c * 4, h // 2, w // 2)


def Downsample(dim: int, dim_out: Optional[int] = None) -> nn.Sequential:
    return nn.Sequential(_SpaceToDepth(), nn.Conv2d(dim * 4, _default(dim_out, dim), 1))


class CrossEmbedLayer(nn.Module):
    """Parallel convolutions of several kernel sizes whose outputs are concatenated (half of the channels to the smallest kernel, a
    quarter to the next, ...; unet.py:794-820)."""

    def __init__(self, dim_in: int, dim_out: Optional[int] = None, *, kernel_sizes: Sequence[int], stride: int = 2):
        # (dim_in, dim_out) positional like every other down-sampler: the reference's ``partial(CrossEmbedLayer, kernel_sizes=...)(dim_in, dim_out)``
        # collides with its own positional ``kernel_sizes`` (unet.py:1086-1090,794), i.e. ``cross_embed_downsample=True`` cannot be built there
        super().__init__()
        assert all(k % 2 == stride % 2 for k in kernel_sizes)
        dim_out = _default(dim_out, dim_in)
        kernel_sizes = sorted(kernel_sizes)
        scales = [int(dim_out / (2 ** i)) for i in range(1, len(kernel_sizes))]
        scales = [*scales, dim_out - sum(scales)]
        self.convs = nn.ModuleList(nn.Conv2d(dim_in, d, k, stride=stride, padding=(k - stride) // 2) for k, d in zip(kernel_sizes, scales))

    def forward(self, x):
        return torch.cat([conv(x) for conv in self.convs], dim=1)


class GlobalContext(nn.Module):
    """Attention-pooled squeeze-excitation gate (unet.py:677-694)."""

    def __init__(self, *, dim_in: int, dim_out: int):
        super().__init__()
        self.to_k = nn.Conv2d(dim_in, 1, 1)
        hidden = max(3, dim_out // 2)
        self.net = nn.Sequential(nn.Conv2d(dim_in, hidden, 1), nn.SiLU(), nn.Conv2d(hidden, dim_out, 1), nn.Sigmoid())

    def forward(self, x):
        ctx = self.to_k(x).flatten(2).softmax(dim=-1)                    # [b, 1, n]
        pooled = torch.einsum("bin,bcn->bci", ctx, x.flatten(2))         # [b, c, 1]
        return self.net(pooled.unsqueeze(-1))


# ------------------------------------------------------------------------------------------ attention
def _masked_softmax(sim, mask, pad_left: int):
    if _exists(mask):
        mask = F.pad(mask, (pad_left, 0), value=True)
        sim = sim.masked_fill(~mask[:, None, None, :].bool(), -torch.finfo(sim.dtype).max)
    return sim.float().softmax(dim=-1).to(sim.dtype)


class PerceiverAttention(nn.Module):
    def __init__(self, *, dim: int, dim_head: int = 64, heads: int = 8, cosine_sim_attn: bool = False):
        super().__init__()
        self.scale = dim_head ** -0.5 if not cosine_sim_attn else 1.0
        self.cosine_sim_attn, self.cosine_sim_scale = cosine_sim_attn, (16 if cosine_sim_attn else 1)
        self.heads = heads
        inner = dim_head * heads
        self.norm, self.norm_latents = nn.LayerNorm(dim), nn.LayerNorm(dim)
        self.to_q, self.to_kv = nn.Linear(dim, inner, bias=False), nn.Linear(dim, inner * 2, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner, dim, bias=False), nn.LayerNorm(dim))

    def forward(self, x, latents, mask=None):
        x, latents = self.norm(x), self.norm_latents(latents)
        b, h = x.shape[0], self.heads
        q = self.to_q(latents)
        k, v = self.to_kv(torch.cat((x, latents), dim=-2)).chunk(2, dim=-1)     # keys / values also cover the latents themselves
        q, k, v = (t.view(b, t.shape[1], h, -1).transpose(1, 2) for t in (q, k, v))
        q = q * self.scale
        if self.cosine_sim_attn:
            q, k = _l2norm(q), _l2norm(k)
        sim = torch.einsum("bhid,bhjd->bhij", q, k) * self.cosine_sim_scale
        if _exists(mask):
            mask = F.pad(mask, (0, latents.shape[-2]), value=True)
            sim = sim.masked_fill(~mask[:, None, None, :].bool(), -torch.finfo(sim.dtype).max)
        attn = sim.float().softmax(dim=-1).to(sim.dtype)
        out = torch.einsum("bhij,bhjd->bhid", attn, v).transpose(1, 2).reshape(b, latents.shape[1], -1)
        return self.to_out(out)


class PerceiverResampler(nn.Module):
    """Text tokens -> a fixed number of latents (+ latents derived from the mean-pooled sequence); unet.py:126-186."""

    def __init__(self, *, dim: int, depth: int, dim_head: int = 64, heads: int = 8, num_latents: int = 64, num_latents_mean_pooled: int = 4,
                 max_seq_len: int = 512, ff_mult: float = 4, cosine_sim_attn: bool = False):
        super().__init__()
        self.pos_emb = nn.Embedding(max_seq_len, dim)
        self.latents = nn.Parameter(torch.randn(num_latents, dim))
        self.num_latents_mean_pooled = num_latents_mean_pooled
        self.to_latents_from_mean_pooled_seq = None
        if num_latents_mean_pooled > 0:
            self.to_latents_from_mean_pooled_seq = nn.Sequential(GainLayerNorm(dim), nn.Linear(dim, dim * num_latents_mean_pooled))
        self.layers = nn.ModuleList(nn.ModuleList([PerceiverAttention(dim=dim, dim_head=dim_head, heads=heads, cosine_sim_attn=cosine_sim_attn),
                                                   FeedForward(dim=dim, mult=ff_mult)]) for _ in range(depth))

    def forward(self, x, mask=None):
        n = x.shape[1]
        x_pos = x + self.pos_emb(torch.arange(n, device=x.device))
        latents = self.latents.to(x.dtype).unsqueeze(0).expand(x.shape[0], -1, -1)
        if _exists(self.to_latents_from_mean_pooled_seq):
            pooled = self.to_latents_from_mean_pooled_seq(x.mean(dim=1))
            latents = torch.cat((pooled.view(x.shape[0], self.num_latents_mean_pooled, -1), latents), dim=-2)
        for attn, ff in self.layers:
            latents = attn(x_pos, latents, mask=mask) + latents
            latents = ff(latents) + latents
        return latents


class Attention(nn.Module):
    """Multi-query self-attention over image tokens (one shared key / value head), with a learned null key / value and optional
    extra keys / values projected from the conditioning tokens (unet.py:189-283)."""

    def __init__(self, dim: int, *, dim_head: int = 64, heads: int = 8, context_dim: Optional[int] = None, cosine_sim_attn: bool = False,
                 use_recompute: bool = False):
        super().__init__()
        self.use_recompute = use_recompute
        self.scale = dim_head ** -0.5 if not cosine_sim_attn else 1.0
        self.cosine_sim_attn, self.cosine_sim_scale = cosine_sim_attn, (16 if cosine_sim_attn else 1)
        self.heads = heads
        inner = dim_head * heads
        self.norm = GainLayerNorm(dim)
        self.null_kv = nn.Parameter(torch.randn(2, dim_head))
        self.to_q = nn.Linear(dim, inner, bias=False)
        self.to_kv = nn.Linear(dim, dim_head * 2, bias=False)
        self.to_context = nn.Sequential(nn.LayerNorm(context_dim), nn.Linear(context_dim, dim_head * 2)) if _exists(context_dim) else None
        self.to_out = nn.Sequential(nn.Linear(inner, dim, bias=False), GainLayerNorm(dim))

    def forward(self, x, context=None, mask=None, attn_bias=None):
        if self.use_recompute and self.training and torch.is_grad_enabled():
            return checkpoint(self._forward, x, context, mask, attn_bias, use_reentrant=False)
        return self._forward(x, context, mask, attn_bias)

    def _forward(self, x, context=None, mask=None, attn_bias=None):
        b, n = x.shape[:2]
        x = self.norm(x)
        q = self.to_q(x).view(b, n, self.heads, -1).transpose(1, 2) * self.scale
        k, v = self.to_kv(x).chunk(2, dim=-1)
        nk, nv = (t.to(x.dtype).view(1, 1, -1).expand(b, 1, -1) for t in self.null_kv.unbind(dim=-2))
        k, v = torch.cat((nk, k), dim=-2), torch.cat((nv, v), dim=-2)
        if _exists(context):
            assert _exists(self.to_context), "this attention layer was built without a conditioning dimension"
            ck, cv = self.to_context(context).chunk(2, dim=-1)
            k, v = torch.cat((ck, k), dim=-2), torch.cat((cv, v), dim=-2)
        if self.cosine_sim_attn:
            q, k = _l2norm(q), _l2norm(k)
        sim = torch.einsum("bhid,bjd->bhij", q, k) * self.cosine_sim_scale
        if _exists(attn_bias):
            sim = sim + attn_bias
        attn = _masked_softmax(sim, mask, 1)
        out = torch.einsum("bhij,bjd->bhid", attn, v).transpose(1, 2).reshape(b, n, -1)
        return self.to_out(out)


class CrossAttention(nn.Module):
    def __init__(self, dim: int, *, context_dim: Optional[int] = None, dim_head: int = 64, heads: int = 8, norm_context: bool = False,
                 cosine_sim_attn: bool = False):
        super().__init__()
        self.scale = dim_head ** -0.5 if not cosine_sim_attn else 1.0
        self.cosine_sim_attn, self.cosine_sim_scale = cosine_sim_attn, (16 if cosine_sim_attn else 1)
        self.heads = heads
        inner = dim_head * heads
        context_dim = _default(context_dim, dim)
        self.norm = GainLayerNorm(dim)
        self.norm_context = GainLayerNorm(context_dim) if norm_context else nn.Identity()
        self.null_kv = nn.Parameter(torch.randn(2, dim_head))
        self.to_q = nn.Linear(dim, inner, bias=False)
        self.to_kv = nn.Linear(context_dim, inner * 2, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner, dim, bias=False), GainLayerNorm(dim))

    def _qkv(self, x, context):
        b = x.shape[0]
        x, context = self.norm(x), self.norm_context(context)
        q = self.to_q(x)
        k, v = self.to_kv(context).chunk(2, dim=-1)
        q, k, v = (t.view(b, t.shape[1], self.heads, -1).transpose(1, 2) for t in (q, k, v))
        nk, nv = (t.to(x.dtype).view(1, 1, 1, -1).expand(b, self.heads, 1, -1) for t in self.null_kv.unbind(dim=-2))
        return q, torch.cat((nk, k), dim=-2), torch.cat((nv, v), dim=-2)

    def forward(self, x, context, mask=None):
        b, n = x.shape[:2]
        q, k, v = self._qkv(x, context)
        q = q * self.scale
        if self.cosine_sim_attn:
            q, k = _l2norm(q), _l2norm(k)
        sim = torch.einsum("bhid,bhjd->bhij", q, k) * self.cosine_sim_scale
        attn = _masked_softmax(sim, mask, 1)
        out = torch.einsum("bhij,bhjd->bhid", attn, v).transpose(1, 2).reshape(b, n, -1)
        return self.to_out(out)


class LinearCrossAttention(CrossAttention):
    """Same parameters, O(n) evaluation: softmax over features for the queries and over positions for the keys (unet.py:544-583)."""

    def forward(self, x, context, mask=None):
        b, n = x.shape[:2]
        q, k, v = self._qkv(x, context)
        if _exists(mask):
            m = F.pad(mask, (1, 0), value=True)[:, None, :, None].bool()
            k = k.masked_fill(~m, -torch.finfo(k.dtype).max)
            v = v.masked_fill(~m, 0.0)
        q = q.softmax(dim=-1) * self.scale
        k = k.softmax(dim=-2)
        ctx = torch.einsum("bhnd,bhne->bhde", k, v)
        out = torch.einsum("bhnd,bhde->bhne", q, ctx).transpose(1, 2).reshape(b, n, -1)
        return self.to_out(out)


class LinearAttention(nn.Module):
    """Linear attention on feature maps with depthwise-conv projections (unet.py:586-674)."""

    def __init__(self, dim: int, dim_head: int = 32, heads: int = 8, dropout: float = 0.05, context_dim: Optional[int] = None, **unused_attn_kwargs):
        super().__init__()
        self.scale, self.heads = dim_head ** -0.5, heads
        inner = dim_head * heads
        self.norm = ChanLayerNorm(dim)
        self.nonlin = nn.SiLU()

        def proj():
            return nn.Sequential(nn.Dropout(dropout), nn.Conv2d(dim, inner, 1, bias=False), nn.Conv2d(inner, inner, 3, bias=False, padding=1, groups=inner))

        self.to_q, self.to_k, self.to_v = proj(), proj(), proj()
        self.to_context = nn.Sequential(nn.LayerNorm(context_dim), nn.Linear(context_dim, inner * 2, bias=False)) if _exists(context_dim) else None
        self.to_out = nn.Sequential(nn.Conv2d(inner, dim, 1, bias=False), ChanLayerNorm(dim))

    def forward(self, fmap, context=None):
        h, (x, y) = self.heads, fmap.shape[-2:]
        b = fmap.shape[0]
        fmap = self.norm(fmap)
        q, k, v = (fn(fmap).view(b, h, -1, x * y).transpose(-1, -2) for fn in (self.to_q, self.to_k, self.to_v))       # [b, h, xy, c]
        if _exists(context):
            assert _exists(self.to_context)
            ck, cv = self.to_context(context).chunk(2, dim=-1)
            ck, cv = (t.view(b, t.shape[1], h, -1).transpose(1, 2) for t in (ck, cv))
            k, v = torch.cat((k, ck), dim=-2), torch.cat((v, cv), dim=-2)
        q = q.softmax(dim=-1) * self.scale
        k = k.softmax(dim=-2)
        ctx = torch.einsum("bhnd,bhne->bhde", k, v)
        out = torch.einsum("bhnd,bhde->bhne", q, ctx)
        out = out.transpose(-1, -2).reshape(b, -1, x, y)
        return self.to_out(self.nonlin(out))


class TransformerBlock(nn.Module):
    def __init__(self, dim: int, *, depth: int = 1, heads: int = 8, dim_head: int = 32, ff_mult: float = 2, context_dim: Optional[int] = None,
                 cosine_sim_attn: bool = False, use_recompute: bool = False):
        super().__init__()
        self.layers = nn.ModuleList(nn.ModuleList([
            Attention(dim=dim, heads=heads, dim_head=dim_head, context_dim=context_dim, cosine_sim_attn=cosine_sim_attn, use_recompute=use_recompute),
            FeedForward(dim=dim, mult=ff_mult)]) for _ in range(depth))

    def forward(self, x, context=None):
        b, c, h, w = x.shape
        x = x.flatten(2).transpose(1, 2)
        for attn, ff in self.layers:
            x = attn(x, context=context) + x
            x = ff(x) + x
        return x.transpose(1, 2).reshape(b, c, h, w)


class LinearAttentionTransformerBlock(nn.Module):
    def __init__(self, dim: int, *, depth: int = 1, heads: int = 8, dim_head: int = 32, ff_mult: float = 2, context_dim: Optional[int] = None, **unused_attn_kwargs):
        super().__init__()
        self.layers = nn.ModuleList(nn.ModuleList([LinearAttention(dim=dim, heads=heads, dim_head=dim_head, context_dim=context_dim),
                                                   ChanFeedForward(dim=dim, mult=ff_mult)]) for _ in range(depth))

    def forward(self, x, context=None):
        for attn, ff in self.layers:
            x = attn(x, context=context) + x
            x = ff(x) + x
        return x


# ------------------------------------------------------------------------------------------ residual blocks
class Block(nn.Module):
    def __init__(self, dim: int, dim_out: int, groups: int = 8, norm: bool = True):
        super().__init__()
        self.groupnorm = nn.GroupNorm(groups, dim) if norm else nn.Identity()
        self.activation = nn.SiLU()
        self.project = nn.Conv2d(dim, dim_out, 3, padding=1)

    def forward(self, x, scale_shift=None):
        x = self.groupnorm(x)
        if _exists(scale_shift):
            scale, shift = scale_shift
            x = x * (scale + 1) + shift
        return self.project(self.activation(x))


class ResnetBlock(nn.Module):
    def __init__(self, dim: int, dim_out: int, *, cond_dim: Optional[int] = None, time_cond_dim: Optional[int] = None, groups: int = 8,
                 linear_attn: bool = False, use_gca: bool = False, squeeze_excite: bool = False, use_recompute: bool = False, **attn_kwargs):
        super().__init__()
        self.use_recompute = use_recompute
        self.time_mlp = nn.Sequential(nn.SiLU(), nn.Linear(time_cond_dim, dim_out * 2)) if _exists(time_cond_dim) else None
        self.cross_attn = None
        if _exists(cond_dim):
            klass = LinearCrossAttention if linear_attn else CrossAttention
            self.cross_attn = klass(dim=dim_out, context_dim=cond_dim, **attn_kwargs)
        self.block1 = Block(dim, dim_out, groups=groups)
        self.block2 = Block(dim_out, dim_out, groups=groups)
        self.gca = GlobalContext(dim_in=dim_out, dim_out=dim_out) if use_gca else Always(1)
        self.res_conv = nn.Conv2d(dim, dim_out, 1) if dim != dim_out else nn.Identity()

    def forward(self, x, time_emb=None, cond=None):
        if self.use_recompute and self.training and torch.is_grad_enabled():
            return checkpoint(self._forward, x, time_emb, cond, use_reentrant=False)
        return self._forward(x, time_emb, cond)

    def _forward(self, x, time_emb=None, cond=None):
        scale_shift = None
        if _exists(self.time_mlp) and _exists(time_emb):
            scale_shift = self.time_mlp(time_emb)[:, :, None, None].chunk(2, dim=1)
        h = self.block1(x)
        if _exists(self.cross_attn):
            assert _exists(cond), "this block cross-attends: conditioning tokens are required"
            b, c, hh, ww = h.shape
            t = h.flatten(2).transpose(1, 2)
            t = self.cross_attn(t, context=cond) + t
            h = t.transpose(1, 2).reshape(b, c, hh, ww)
        h = self.block2(h, scale_shift=scale_shift)
        h = h * self.gca(h)
        return h + self.res_conv(x)


class UpsampleCombiner(nn.Module):
    def __init__(self, dim: int, *, enabled: bool = False, dim_ins: Sequence[int] = (), dim_outs=()):
        super().__init__()
        dim_outs = _cast_tuple(dim_outs, len(dim_ins))
        assert len(dim_ins) == len(dim_outs)
        self.enabled = enabled
        if not enabled:
            self.dim_out = dim
            return
        self.fmap_convs = nn.ModuleList(Block(i, o) for i, o in zip(dim_ins, dim_outs))
        self.dim_out = dim + (sum(dim_outs) if len(dim_outs) > 0 else 0)

    def forward(self, x, fmaps=None):
        fmaps = _default(fmaps, ())
        if not self.enabled or len(fmaps) == 0 or len(self.fmap_convs) == 0:
            return x
        size = x.shape[-1]
        outs = [conv(resize_image_to(f, size)) for f, conv in zip(fmaps, self.fmap_convs)]
        return torch.cat((x, *outs), dim=1)


# ------------------------------------------------------------------------------------------ the U-Net
class Unet(nn.Module):
    def __init__(self, *, dim, image_embed_dim=1024, text_embed_dim=1024, num_resnet_blocks=1, cond_dim=None, num_image_tokens=4, num_time_tokens=2,
                 learned_sinu_pos_emb_dim=16, out_dim=None, dim_mults=(1, 2, 4, 8), cond_images_channels=0, channels=3, channels_out=None,
                 attn_dim_head=64, attn_heads=8, ff_mult=2.0, lowres_cond=False, layer_attns=True, layer_attns_depth=1, layer_mid_attns_depth=1,
                 layer_attns_add_text_cond=True, attend_at_middle=True, layer_cross_attns=True, use_linear_attn=False, use_linear_cross_attn=False,
                 cond_on_text=True, max_text_len=256, init_dim=None, resnet_groups=8, init_conv_kernel_size=7, init_cross_embed=True,
                 init_cross_embed_kernel_sizes=(3, 7, 15), cross_embed_downsample=False, cross_embed_downsample_kernel_sizes=(2, 4),
                 attn_pool_text=True, attn_pool_num_latents=32, dropout=0.0, memory_efficient=False, init_conv_to_final_conv_residual=False,
                 use_global_context_attn=True, scale_skip_connection=True, final_resnet_block=True, final_conv_kernel_size=3,
                 cosine_sim_attn=False, self_cond=False, combine_upsample_fmaps=False, pixel_shuffle_upsample=True, use_recompute=False):
        super().__init__()
        assert attn_heads > 1, "you need to have more than 1 attention head, ideally at least 4 or 8"
        self._locals = {k: v for k, v in locals().items() if k not in ("self", "__class__")}
        self.use_recompute = use_recompute
        self.channels = channels
        self.channels_out = _default(channels_out, channels)
        self.self_cond = self_cond
        self.has_cond_image = cond_images_channels > 0
        self.cond_images_channels = cond_images_channels
        init_channels = channels * (1 + int(lowres_cond) + int(self_cond)) + cond_images_channels
        init_dim = _default(init_dim, dim)

        self.init_conv = CrossEmbedLayer(init_channels, dim_out=init_dim, kernel_sizes=init_cross_embed_kernel_sizes, stride=1) if init_cross_embed \
            else nn.Conv2d(init_channels, init_dim, init_conv_kernel_size, padding=init_conv_kernel_size // 2)
        dims = [init_dim, *[dim * m for m in dim_mults]]
        in_out = list(zip(dims[:-1], dims[1:]))

        # ---- time (log-SNR) conditioning
        cond_dim = _default(cond_dim, dim)
        time_cond_dim = dim * 4 * (2 if lowres_cond else 1)
        self.num_time_tokens = num_time_tokens
        self.to_time_hiddens = nn.Sequential(LearnedSinusoidalPosEmb(learned_sinu_pos_emb_dim), nn.Linear(learned_sinu_pos_emb_dim + 1, time_cond_dim), nn.SiLU())
        self.to_time_cond = nn.Sequential(nn.Linear(time_cond_dim, time_cond_dim))
        self.to_time_tokens = nn.Sequential(nn.Linear(time_cond_dim, cond_dim * num_time_tokens))
        self.lowres_cond = lowres_cond
        if lowres_cond:
            self.to_lowres_time_hiddens = nn.Sequential(LearnedSinusoidalPosEmb(learned_sinu_pos_emb_dim),
                                                        nn.Linear(learned_sinu_pos_emb_dim + 1, time_cond_dim), nn.SiLU())
            self.to_lowres_time_cond = nn.Sequential(nn.Linear(time_cond_dim, time_cond_dim))
            self.to_lowres_time_tokens = nn.Sequential(nn.Linear(time_cond_dim, cond_dim * num_time_tokens))
        self.norm_cond = nn.LayerNorm(cond_dim)

        # ---- text conditioning
        self.text_to_cond = None
        if cond_on_text:
            assert _exists(text_embed_dim), "text_embed_dim must be given to the unet if cond_on_text is True"
            self.text_to_cond = nn.Linear(text_embed_dim, cond_dim)
        self.cond_on_text = cond_on_text
        self.attn_pool = PerceiverResampler(dim=cond_dim, depth=2, dim_head=attn_dim_head, heads=attn_heads, num_latents=attn_pool_num_latents,
                                            cosine_sim_attn=cosine_sim_attn) if attn_pool_text else None
        self.max_text_len = max_text_len
        self.null_text_embed = nn.Parameter(torch.randn(1, max_text_len, cond_dim))
        self.null_text_hidden = nn.Parameter(torch.randn(1, time_cond_dim))
        self.to_text_non_attn_cond = None
        if cond_on_text:
            self.to_text_non_attn_cond = nn.Sequential(nn.LayerNorm(cond_dim), nn.Linear(cond_dim, time_cond_dim), nn.SiLU(),
                                                       nn.Linear(time_cond_dim, time_cond_dim))

        # ---- per-resolution settings
        attn_kwargs = dict(heads=attn_heads, dim_head=attn_dim_head, cosine_sim_attn=cosine_sim_attn)
        n_layers = len(in_out)
        num_resnet_blocks = _cast_tuple(num_resnet_blocks, n_layers)
        resnet_groups = _cast_tuple(resnet_groups, n_layers)
        layer_attns = _cast_tuple(layer_attns, n_layers)
        layer_attns_depth = _cast_tuple(layer_attns_depth, n_layers)
        layer_cross_attns = _cast_tuple(layer_cross_attns, n_layers)
        use_linear_attn = _cast_tuple(use_linear_attn, n_layers)
        use_linear_cross_attn = _cast_tuple(use_linear_cross_attn, n_layers)
        assert all(len(t) == n_layers for t in (num_resnet_blocks, resnet_groups, layer_attns, layer_attns_depth, layer_cross_attns,
                                                use_linear_attn, use_linear_cross_attn)), "per-layer options must have one entry per resolution"
        resnet_klass = partial(ResnetBlock, use_recompute=use_recompute, **attn_kwargs)
        downsample_klass = partial(CrossEmbedLayer, kernel_sizes=cross_embed_downsample_kernel_sizes) if cross_embed_downsample else Downsample
        self.init_resnet_block = resnet_klass(init_dim, init_dim, time_cond_dim=time_cond_dim, groups=resnet_groups[0],
                                              use_gca=use_global_context_attn) if memory_efficient else None
        self.skip_connect_scale = 1.0 if not scale_skip_connection else 2 ** -0.5

        def attn_block(layer_attn, layer_linear, d, depth):
            if layer_attn:
                return TransformerBlock(dim=d, depth=depth, ff_mult=ff_mult, context_dim=cond_dim, use_recompute=use_recompute, **attn_kwargs)
            if layer_linear:
                return LinearAttentionTransformerBlock(dim=d, depth=depth, ff_mult=ff_mult, context_dim=cond_dim, **attn_kwargs)
            return PassThrough()

        layer_params = [num_resnet_blocks, resnet_groups, layer_attns, layer_attns_depth, layer_cross_attns, use_linear_attn, use_linear_cross_attn]
        self.downs = nn.ModuleList()
        skip_dims = []
        for ind, ((d_in, d_out), n_blocks, groups, l_attn, l_depth, l_cross, l_lin, l_lin_cross) in enumerate(zip(in_out, *layer_params)):
            is_last = ind >= n_layers - 1
            layer_cond_dim = cond_dim if (l_cross or l_lin_cross) else None
            cur = d_in
            pre = None
            if memory_efficient:
                pre = downsample_klass(d_in, d_out)
                cur = d_out
            skip_dims.append(cur)
            post = None
            if not memory_efficient:
                post = downsample_klass(cur, d_out) if not is_last else Parallel(nn.Conv2d(d_in, d_out, 3, padding=1), nn.Conv2d(d_in, d_out, 1))
            self.downs.append(nn.ModuleList([
                pre if pre is not None else nn.Identity(),
                resnet_klass(cur, cur, cond_dim=layer_cond_dim, linear_attn=l_lin_cross, time_cond_dim=time_cond_dim, groups=groups),
                nn.ModuleList(ResnetBlock(cur, cur, time_cond_dim=time_cond_dim, groups=groups, use_gca=use_global_context_attn,
                                          use_recompute=use_recompute) for _ in range(n_blocks)),
                attn_block(l_attn, l_lin, cur, l_depth),
                post if post is not None else nn.Identity()]))

        mid = dims[-1]
        self.mid_block1 = ResnetBlock(mid, mid, cond_dim=cond_dim, time_cond_dim=time_cond_dim, groups=resnet_groups[-1], use_recompute=use_recompute)
        self.mid_attn = TransformerBlock(mid, depth=layer_mid_attns_depth, use_recompute=use_recompute, **attn_kwargs) if attend_at_middle else None
        self.mid_block2 = ResnetBlock(mid, mid, cond_dim=cond_dim, time_cond_dim=time_cond_dim, groups=resnet_groups[-1], use_recompute=use_recompute)

        upsample_klass = PixelShuffleUpsample if pixel_shuffle_upsample else Upsample
        self.ups = nn.ModuleList()
        up_fmap_dims = []
        for ind, ((d_in, d_out), n_blocks, groups, l_attn, l_depth, l_cross, l_lin, l_lin_cross) in enumerate(
                zip(reversed(in_out), *[tuple(reversed(p)) for p in layer_params])):
            is_last = ind == n_layers - 1
            layer_cond_dim = cond_dim if (l_cross or l_lin_cross) else None
            skip = skip_dims.pop()
            up_fmap_dims.append(d_out)
            self.ups.append(nn.ModuleList([
                resnet_klass(d_out + skip, d_out, cond_dim=layer_cond_dim, linear_attn=l_lin_cross, time_cond_dim=time_cond_dim, groups=groups),
                nn.ModuleList(ResnetBlock(d_out + skip, d_out, time_cond_dim=time_cond_dim, groups=groups, use_gca=use_global_context_attn,
                                          use_recompute=use_recompute) for _ in range(n_blocks)),
                attn_block(l_attn, l_lin, d_out, l_depth),
                upsample_klass(d_out, d_in) if (not is_last or memory_efficient) else nn.Identity()]))

        self.upsample_combiner = UpsampleCombiner(dim=dim, enabled=combine_upsample_fmaps, dim_ins=up_fmap_dims, dim_outs=dim)
        self.init_conv_to_final_conv_residual = init_conv_to_final_conv_residual
        final_dim = self.upsample_combiner.dim_out + (dim if init_conv_to_final_conv_residual else 0)
        self.final_res_block = ResnetBlock(final_dim, dim, time_cond_dim=time_cond_dim, groups=resnet_groups[0], use_gca=True,
                                           use_recompute=use_recompute) if final_resnet_block else None
        final_in = (dim if final_resnet_block else final_dim) + (channels if lowres_cond else 0)
        self.final_conv = nn.Conv2d(final_in, self.channels_out, final_conv_kernel_size, padding=final_conv_kernel_size // 2)
        nn.init.zeros_(self.final_conv.weight)
        nn.init.zeros_(self.final_conv.bias)

    # -------------------------------------------------------------------------------------- cascade plumbing
    def cast_model_parameters(self, *, text_embed_dim, channels, channels_out, cond_on_text):
        """Rebuild with the settings a cascading DDPM needs at this position, if they differ (unet.py:1299-1313)."""
        if (channels == self.channels and cond_on_text == self.cond_on_text and text_embed_dim == self._locals["text_embed_dim"]
                and channels_out == self.channels_out):
            return self
        return self.__class__(**{**self._locals, "text_embed_dim": text_embed_dim, "channels": channels, "channels_out": channels_out,
                                 "cond_on_text": cond_on_text})

    def to_config_and_state_dict(self):
        return self._locals, self.state_dict()

    @classmethod
    def from_config_and_state_dict(cls, config, state_dict):
        unet = cls(**config)
        unet.load_state_dict(state_dict)
        return unet

    def persist_to_file(self, path):
        import os

        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        config, state = self.to_config_and_state_dict()
        torch.save(dict(config=config, state_dict=state), path)

    @classmethod
    def hydrate_from_file(cls, path):
        pkg = torch.load(path, map_location="cpu", weights_only=False)
        return Unet.from_config_and_state_dict(pkg["config"], pkg["state_dict"])

    # -------------------------------------------------------------------------------------- forward
    def forward_with_cond_scale(self, *args, cond_scale=1.0, **kwargs):
        logits = self.forward(*args, **kwargs)
        if cond_scale == 1:
            return logits
        null_logits = self.forward(*args, cond_drop_prob=1.0, **kwargs)
        return null_logits + (logits - null_logits) * cond_scale

    def forward(self, x, time, *, lowres_cond_img=None, lowres_noise_times=None, text_embeds=None, text_mask=None, self_cond=None,
                cond_images=None, cond_drop_prob=0.0, use_recompute=False):
        """Predict the noise (or ``x0`` / ``v``) for noisy images ``x`` at continuous ``time``: time (+ low-resolution noise level) conditioning
        tokens, text conditioning through the Perceiver / attention pooling with classifier-free dropout of probability ``cond_drop_prob``, then the
        down path, the middle blocks and the up path with skip connections (reference unet.py:1371-1562)."""
        b = x.shape[0]
        if self.self_cond:
            x = torch.cat((x, _default(self_cond, lambda: torch.zeros_like(x))), dim=1)
        assert not (self.lowres_cond and not _exists(lowres_cond_img)), "low resolution conditioning image must be present"
        assert not (self.lowres_cond and not _exists(lowres_noise_times)), "low resolution conditioning noise time must be present"
        if _exists(lowres_cond_img):
            x = torch.cat((x, lowres_cond_img), dim=1)
        assert not (self.has_cond_image ^ _exists(cond_images)), "conditioning image requested but not supplied (or the reverse)"
        if _exists(cond_images):
            assert cond_images.shape[1] == self.cond_images_channels, "conditioning image has the wrong number of channels"
            x = torch.cat((resize_image_to(cond_images, x.shape[-1]), x), dim=1)

        x = self.init_conv(x)
        init_residual = x.clone() if self.init_conv_to_final_conv_residual else None

        time_hiddens = self.to_time_hiddens(time).to(x.dtype)
        time_tokens = self.to_time_tokens(time_hiddens).view(b, self.num_time_tokens, -1)
        t = self.to_time_cond(time_hiddens)
        if self.lowres_cond:
            lr_hiddens = self.to_lowres_time_hiddens(lowres_noise_times).to(x.dtype)
            t = t + self.to_lowres_time_cond(lr_hiddens)
            time_tokens = torch.cat((time_tokens, self.to_lowres_time_tokens(lr_hiddens).view(b, self.num_time_tokens, -1)), dim=-2)

        text_tokens = None
        if _exists(text_embeds) and self.cond_on_text:
            keep = prob_mask_like((b,), 1 - cond_drop_prob, x.device)
            keep_embed, keep_hidden = keep[:, None, None], keep[:, None]
            text_tokens = self.text_to_cond(text_embeds.to(x.dtype))[:, :self.max_text_len]
            if _exists(text_mask):
                text_mask = text_mask[:, :self.max_text_len]
            remainder = self.max_text_len - text_tokens.shape[1]
            if remainder > 0:
                text_tokens = F.pad(text_tokens, (0, 0, 0, remainder))
            if _exists(text_mask):
                tm = text_mask.bool()
                if remainder > 0:
                    tm = F.pad(tm, (0, remainder), value=False)
                keep_embed = tm[:, :, None] & keep_embed
            text_tokens = torch.where(keep_embed, text_tokens, self.null_text_embed.to(text_tokens.dtype))
            if _exists(self.attn_pool):
                text_tokens = self.attn_pool(text_tokens)
            text_hiddens = self.to_text_non_attn_cond(text_tokens.mean(dim=-2))
            t = t + torch.where(keep_hidden, text_hiddens, self.null_text_hidden.to(t.dtype))

        c = time_tokens if not _exists(text_tokens) else torch.cat((time_tokens, text_tokens), dim=-2)
        c = self.norm_cond(c)

        if _exists(self.init_resnet_block):
            x = self.init_resnet_block(x, t)
        hiddens = []
        for pre, init_block, blocks, attn, post in self.downs:
            x = pre(x)
            x = init_block(x, t, c)
            for blk in blocks:
                x = blk(x, t)
                hiddens.append(x)
            x = attn(x, c)
            hiddens.append(x)
            x = post(x)

        x = self.mid_block1(x, t, c)
        if _exists(self.mid_attn):
            x = self.mid_attn(x)
        x = self.mid_block2(x, t, c)

        def with_skip(y):
            return torch.cat((y, hiddens.pop() * self.skip_connect_scale), dim=1)

        up_hiddens = []
        for init_block, blocks, attn, upsample in self.ups:
            x = init_block(with_skip(x), t, c)
            for blk in blocks:
                x = blk(with_skip(x), t)
            x = attn(x, c)
            up_hiddens.append(x)
            x = upsample(x)

        x = self.upsample_combiner(x, up_hiddens)
        if self.init_conv_to_final_conv_residual:
            x = torch.cat((x, init_residual), dim=1)
        if _exists(self.final_res_block):
            x = self.final_res_block(x, t)
        if _exists(lowres_cond_img):
            x = torch.cat((x, lowres_cond_img), dim=1)
        return self.final_conv(x)


# ------------------------------------------------------------------------------------------ presets (reference modeling.py:36-92)
class Unet64_397M(Unet):
    def __init__(self, **kwargs):
        super().__init__(**{**dict(dim=256, dim_mults=(1, 2, 3, 4), num_resnet_blocks=3, layer_attns=(False, True, True, True),
                                   layer_cross_attns=(False, True, True, True), attn_heads=8, ff_mult=2.0, memory_efficient=False), **kwargs})


class BaseUnet64(Unet):
    def __init__(self, **kwargs):
        super().__init__(**{**dict(dim=512, cond_dim=512, dim_mults=(1, 2, 3, 4), num_resnet_blocks=3, layer_attns=(False, True, True, True),
                                   layer_cross_attns=(False, True, True, True), attn_heads=8, ff_mult=2.0, memory_efficient=False), **kwargs})


class SRUnet256(Unet):
    def __init__(self, **kwargs):
        super().__init__(**{**dict(dim=128, dim_mults=(1, 2, 4, 8), num_resnet_blocks=(2, 4, 8, 8), layer_attns=(False, False, False, True),
                                   layer_cross_attns=(False, False, False, True), attn_heads=8, ff_mult=2.0, memory_efficient=True), **kwargs})


class SRUnet1024(Unet):
    def __init__(self, **kwargs):
        super().__init__(**{**dict(dim=128, dim_mults=(1, 2, 4, 8), num_resnet_blocks=(2, 4, 8, 8), layer_attns=False,
                                   layer_cross_attns=(False, False, False, True), attn_heads=8, ff_mult=2.0, memory_efficient=True), **kwargs})
