"""Imagen U-Net (reference multimodal_model/imagen/unet.py:33-1562): time + text conditioning (mean-pooled text hidden
state + Perceiver-resampled attention tokens), ResNet blocks with scale-shift conditioning and optional cross-attention,
per-resolution transformer blocks (full or linear attention), low-resolution conditioning for the super-resolution stages,
and the four presets ``Unet64_397M`` / ``BaseUnet64`` / ``SRUnet256`` / ``SRUnet1024``."""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F


def _groups(ch, g=8):
    while ch % g:
        g //= 2
    return max(g, 1)


class SinusoidalPosEmb(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.dim = dim

    def forward(self, t):
        half = self.dim // 2
        freqs = torch.exp(torch.arange(half, device=t.device, dtype=torch.float32) * -(math.log(10000) / max(half - 1, 1)))
        args = t.float()[:, None] * freqs[None]
        return torch.cat([args.sin(), args.cos()], -1)


class Block(nn.Module):
    def __init__(self, dim, dim_out, groups=8):
        super().__init__()
        self.norm = nn.GroupNorm(_groups(dim, groups), dim)
        self.proj = nn.Conv2d(dim, dim_out, 3, padding=1)

    def forward(self, x, scale_shift=None):
        x = self.norm(x)
        if scale_shift is not None:
            s, b = scale_shift
            x = x * (s + 1) + b
        return self.proj(F.silu(x))


class CrossAttention(nn.Module):
    def __init__(self, dim, context_dim, heads=8, dim_head=64):
        super().__init__()
        inner = heads * dim_head
        self.heads, self.dh = heads, dim_head
        self.norm, self.norm_ctx = nn.LayerNorm(dim), nn.LayerNorm(context_dim)
        self.to_q, self.to_kv = nn.Linear(dim, inner, bias=False), nn.Linear(context_dim, inner * 2, bias=False)
        self.null_kv = nn.Parameter(torch.randn(2, dim_head))
        self.to_out = nn.Sequential(nn.Linear(inner, dim, bias=False), nn.LayerNorm(dim))

    def forward(self, x, context, mask=None):
        b, n, _ = x.shape
        q = self.to_q(self.norm(x)).view(b, n, self.heads, self.dh).transpose(1, 2)
        k, v = self.to_kv(self.norm_ctx(context)).chunk(2, -1)
        k, v = (t.view(b, -1, self.heads, self.dh).transpose(1, 2) for t in (k, v))
        nk, nv = (t.view(1, 1, 1, self.dh).expand(b, self.heads, 1, self.dh) for t in self.null_kv.unbind(0))
        k, v = torch.cat([nk, k], 2), torch.cat([nv, v], 2)
        am = None
        if mask is not None:
            am = F.pad(mask.bool(), (1, 0), value=True)[:, None, None, :]
        o = F.scaled_dot_product_attention(q, k, v, attn_mask=am)
        return self.to_out(o.transpose(1, 2).reshape(b, n, -1))


class ResnetBlock(nn.Module):
    def __init__(self, dim, dim_out, cond_dim=None, time_cond_dim=None, groups=8, use_cross_attn=False):
        super().__init__()
        self.time_mlp = nn.Sequential(nn.SiLU(), nn.Linear(time_cond_dim, dim_out * 2)) if time_cond_dim else None
        self.cross_attn = CrossAttention(dim_out, cond_dim) if (use_cross_attn and cond_dim) else None
        self.block1, self.block2 = Block(dim, dim_out, groups), Block(dim_out, dim_out, groups)
        self.res_conv = nn.Conv2d(dim, dim_out, 1) if dim != dim_out else nn.Identity()

    def forward(self, x, time_emb=None, cond=None, cond_mask=None):
        ss = None
        if self.time_mlp is not None and time_emb is not None:
            ss = self.time_mlp(time_emb)[:, :, None, None].chunk(2, 1)
        h = self.block1(x)
        if self.cross_attn is not None and cond is not None:
            b, c, hh, ww = h.shape
            t = h.flatten(2).transpose(1, 2)
            h = (self.cross_attn(t, cond, cond_mask) + t).transpose(1, 2).reshape(b, c, hh, ww)
        h = self.block2(h, ss)
        return h + self.res_conv(x)


class SelfAttention2d(nn.Module):
    """Full attention over pixels (+ optional text context as extra keys)."""

    def __init__(self, dim, heads=8, dim_head=64, context_dim=None):
        super().__init__()
        inner = heads * dim_head
        self.heads, self.dh = heads, dim_head
        self.norm = nn.LayerNorm(dim)
        self.to_qkv = nn.Linear(dim, inner * 3, bias=False)
        self.to_ctx = nn.Sequential(nn.LayerNorm(context_dim), nn.Linear(context_dim, inner * 2)) if context_dim else None
        self.to_out = nn.Sequential(nn.Linear(inner, dim, bias=False), nn.LayerNorm(dim))

    def forward(self, x, context=None):
        b, c, hh, ww = x.shape
        t = self.norm(x.flatten(2).transpose(1, 2))
        q, k, v = (u.view(b, -1, self.heads, self.dh).transpose(1, 2) for u in self.to_qkv(t).chunk(3, -1))
        if self.to_ctx is not None and context is not None:
            ck, cv = (u.view(b, -1, self.heads, self.dh).transpose(1, 2) for u in self.to_ctx(context).chunk(2, -1))
            k, v = torch.cat([ck, k], 2), torch.cat([cv, v], 2)
        o = F.scaled_dot_product_attention(q, k, v)
        o = self.to_out(o.transpose(1, 2).reshape(b, hh * ww, -1))
        return x + o.transpose(1, 2).reshape(b, c, hh, ww)


class LinearAttention2d(nn.Module):
    def __init__(self, dim, heads=8, dim_head=32):
        super().__init__()
        inner = heads * dim_head
        self.heads, self.dh = heads, dim_head
        self.norm = nn.GroupNorm(1, dim)
        self.to_qkv = nn.Conv2d(dim, inner * 3, 1, bias=False)
        self.to_out = nn.Sequential(nn.Conv2d(inner, dim, 1), nn.GroupNorm(1, dim))

    def forward(self, x):
        b, c, h, w = x.shape
        q, k, v = (t.view(b, self.heads, self.dh, h * w) for t in self.to_qkv(self.norm(x)).chunk(3, 1))
        q, k = q.softmax(-2) * self.dh ** -0.5, k.softmax(-1)
        ctx = torch.einsum("bhdn,bhen->bhde", k, v)
        out = torch.einsum("bhde,bhdn->bhen", ctx, q).reshape(b, -1, h, w)
        return x + self.to_out(out)


class PerceiverResampler(nn.Module):
    def __init__(self, dim, depth=2, dim_head=64, heads=8, num_latents=32, max_seq_len=512):
        super().__init__()
        self.pos_emb = nn.Embedding(max_seq_len, dim)
        self.latents = nn.Parameter(torch.randn(num_latents, dim))
        self.layers = nn.ModuleList([nn.ModuleList([CrossAttention(dim, dim, heads, dim_head),
                                                    nn.Sequential(nn.LayerNorm(dim), nn.Linear(dim, dim * 2), nn.GELU(), nn.Linear(dim * 2, dim))])
                                     for _ in range(depth)])

    def forward(self, x, mask=None):
        n = x.shape[1]
        x = x + self.pos_emb(torch.arange(n, device=x.device))[None]
        lat = self.latents[None].expand(x.shape[0], -1, -1)
        for attn, ff in self.layers:
            ctx = torch.cat([x, lat], 1)
            m = None if mask is None else F.pad(mask.bool(), (0, lat.shape[1]), value=True)
            lat = attn(lat, ctx, m) + lat
            lat = ff(lat) + lat
        return lat


class Unet(nn.Module):
    def __init__(self, dim=128, text_embed_dim=1024, cond_dim=None, channels=3, channels_out=None, dim_mults=(1, 2, 4, 8), num_resnet_blocks=1,
                 layer_attns=(False, False, False, True), layer_cross_attns=(False, True, True, True), attn_heads=8, attn_dim_head=64,
                 lowres_cond=False, cond_on_text=True, max_text_len=256, num_latents=32, memory_efficient=False, use_linear_attn=False,
                 resnet_groups=8, learned_sinu_pos_emb_dim=16, **unused):
        super().__init__()
        self.channels, self.lowres_cond, self.cond_on_text = channels, lowres_cond, cond_on_text
        self.channels_out = channels_out or channels
        cond_dim = cond_dim or dim
        time_cond_dim = dim * 4
        init_ch = channels * (2 if lowres_cond else 1)
        self.init_conv = nn.Conv2d(init_ch, dim, 7, padding=3)
        self.to_time_hiddens = nn.Sequential(SinusoidalPosEmb(dim), nn.Linear(dim, time_cond_dim), nn.SiLU())
        self.to_time_cond = nn.Linear(time_cond_dim, time_cond_dim)
        self.to_time_tokens = nn.Linear(time_cond_dim, cond_dim * 2)
        self.num_time_tokens, self.cond_dim = 2, cond_dim
        if lowres_cond:
            self.to_lowres_time_hiddens = nn.Sequential(SinusoidalPosEmb(dim), nn.Linear(dim, time_cond_dim), nn.SiLU())
            self.to_lowres_time_cond = nn.Linear(time_cond_dim, time_cond_dim)
        if cond_on_text:
            self.text_to_cond = nn.Linear(text_embed_dim, cond_dim)
            self.attn_pool = PerceiverResampler(cond_dim, 2, attn_dim_head, attn_heads, num_latents, max_text_len)
            self.to_text_non_attn_cond = nn.Sequential(nn.LayerNorm(cond_dim), nn.Linear(cond_dim, time_cond_dim), nn.SiLU(),
                                                       nn.Linear(time_cond_dim, time_cond_dim))
            self.null_text_embed = nn.Parameter(torch.randn(1, max_text_len, cond_dim))
            self.null_text_hidden = nn.Parameter(torch.randn(1, time_cond_dim))
            self.max_text_len = max_text_len
        self.norm_cond = nn.LayerNorm(cond_dim)
        dims = [dim] + [dim * m for m in dim_mults]
        io = list(zip(dims[:-1], dims[1:]))
        n = len(io)
        nrb = [num_resnet_blocks] * n if isinstance(num_resnet_blocks, int) else list(num_resnet_blocks)
        la = list(layer_attns) if isinstance(layer_attns, (tuple, list)) else [layer_attns] * n
        lca = list(layer_cross_attns) if isinstance(layer_cross_attns, (tuple, list)) else [layer_cross_attns] * n
        self.downs, self.ups = nn.ModuleList(), nn.ModuleList()
        skip_dims = []
        for i, (di, do) in enumerate(io):
            last = i == n - 1
            pre = nn.Conv2d(di, do, 4, 2, 1) if (memory_efficient and not last) else None
            cur = do if pre is not None else di
            blocks = nn.ModuleList([ResnetBlock(cur, cur, cond_dim, time_cond_dim, resnet_groups, lca[i])] +
                                   [ResnetBlock(cur, cur, None, time_cond_dim, resnet_groups) for _ in range(nrb[i])])
            attn = SelfAttention2d(cur, attn_heads, attn_dim_head, cond_dim) if la[i] else (LinearAttention2d(cur) if use_linear_attn else nn.Identity())
            post = None if (memory_efficient or last) else nn.Conv2d(cur, do, 4, 2, 1)
            if last and pre is None:
                post = nn.Conv2d(cur, do, 3, padding=1)
            skip_dims.append(cur)
            self.downs.append(nn.ModuleList([pre if pre is not None else nn.Identity(), blocks, attn, post if post is not None else nn.Identity()]))
        mid = dims[-1]
        self.mid_block1 = ResnetBlock(mid, mid, cond_dim, time_cond_dim, resnet_groups, True)
        self.mid_attn = SelfAttention2d(mid, attn_heads, attn_dim_head)
        self.mid_block2 = ResnetBlock(mid, mid, cond_dim, time_cond_dim, resnet_groups, True)
        for i, (di, do) in enumerate(reversed(io)):
            j = n - 1 - i
            last = i == n - 1
            skip = skip_dims[j]
            blocks = nn.ModuleList([ResnetBlock(do + skip, do, cond_dim, time_cond_dim, resnet_groups, lca[j])] +
                                   [ResnetBlock(do + skip, do, None, time_cond_dim, resnet_groups) for _ in range(nrb[j])])
            attn = SelfAttention2d(do, attn_heads, attn_dim_head, cond_dim) if la[j] else (LinearAttention2d(do) if use_linear_attn else nn.Identity())
            if j == n - 1:          # innermost level: the down path did not halve here, so only the channel projection is mirrored
                up = nn.Sequential(nn.Conv2d(do, di, 3, padding=1))
            elif last:
                up = nn.Identity()
            else:
                up = nn.Sequential(nn.Upsample(scale_factor=2, mode="nearest"), nn.Conv2d(do, di, 3, padding=1))
            self.ups.append(nn.ModuleList([blocks, attn, up, nn.Identity()]))
        self._io = io
        self.final_res = ResnetBlock(dim * 2, dim, None, time_cond_dim, resnet_groups)
        self.final_conv = nn.Conv2d(dim, self.channels_out, 3, padding=1)
        nn.init.zeros_(self.final_conv.weight); nn.init.zeros_(self.final_conv.bias)

    def forward_with_cond_scale(self, *args, cond_scale=1.0, **kwargs):
        logits = self.forward(*args, **kwargs)
        if cond_scale == 1:
            return logits
        null = self.forward(*args, cond_drop_prob=1.0, **kwargs)
        return null + (logits - null) * cond_scale

    def forward(self, x, time, lowres_cond_img=None, lowres_noise_times=None, text_embeds=None, text_mask=None, cond_drop_prob=0.0):
        b = x.shape[0]
        if self.lowres_cond:
            assert lowres_cond_img is not None, "low resolution conditioning image must be present"
            x = torch.cat([x, lowres_cond_img], 1)
        x = self.init_conv(x)
        r = x
        th = self.to_time_hiddens(time)
        t = self.to_time_cond(th)
        tokens = self.to_time_tokens(th).view(b, self.num_time_tokens, self.cond_dim)
        if self.lowres_cond and lowres_noise_times is not None:
            lh = self.to_lowres_time_hiddens(lowres_noise_times)
            t = t + self.to_lowres_time_cond(lh)
        c, cmask = tokens, None
        if self.cond_on_text and text_embeds is not None:
            keep = torch.rand(b, device=x.device) >= cond_drop_prob
            te = self.text_to_cond(text_embeds)[:, : self.max_text_len]
            tm = text_mask[:, : self.max_text_len].bool() if text_mask is not None else torch.ones(te.shape[:2], dtype=torch.bool, device=x.device)
            pad = self.max_text_len - te.shape[1]
            if pad > 0:
                te, tm = F.pad(te, (0, 0, 0, pad)), F.pad(tm, (0, pad), value=False)
            km = tm & keep[:, None]
            te = torch.where(km[..., None], te, self.null_text_embed.to(te.dtype).expand(b, -1, -1))
            mean = (te * km[..., None]).sum(1) / km.sum(1, keepdim=True).clamp(min=1)
            hid = self.to_text_non_attn_cond(mean)
            hid = torch.where(keep[:, None], hid, self.null_text_hidden.to(hid.dtype).expand(b, -1))
            t = t + hid
            c = torch.cat([tokens, self.attn_pool(te, None)], 1)
        c = self.norm_cond(c)
        skips = []
        for pre, blocks, attn, post in self.downs:
            x = pre(x)
            x = blocks[0](x, t, c, cmask)
            for blk in blocks[1:]:
                x = blk(x, t)
            x = attn(x, c) if isinstance(attn, SelfAttention2d) else attn(x)
            skips.append(x)
            x = post(x)
        x = self.mid_block2(self.mid_attn(self.mid_block1(x, t, c, cmask)), t, c, cmask)
        for (blocks, attn, up, _), skip in zip(self.ups, reversed(skips)):
            if x.shape[-2:] != skip.shape[-2:]:
                x = F.interpolate(x, size=skip.shape[-2:], mode="nearest")
            x = blocks[0](torch.cat([x, skip], 1), t, c, cmask)
            for blk in blocks[1:]:
                x = blk(torch.cat([x, skip], 1), t)
            x = attn(x, c) if isinstance(attn, SelfAttention2d) else attn(x)
            x = up(x)
        if x.shape[-2:] != r.shape[-2:]:
            x = F.interpolate(x, size=r.shape[-2:], mode="nearest")
        x = self.final_res(torch.cat([x, r], 1), t)
        return self.final_conv(x)


def Unet64_397M(**kw):
    return Unet(**{**dict(dim=256, dim_mults=(1, 2, 3, 4), num_resnet_blocks=3, layer_attns=(False, True, True, True),
                          layer_cross_attns=(False, True, True, True), attn_heads=8, memory_efficient=False), **kw})


def BaseUnet64(**kw):
    return Unet(**{**dict(dim=512, dim_mults=(1, 2, 3, 4), num_resnet_blocks=3, layer_attns=(False, True, True, True),
                          layer_cross_attns=(False, True, True, True), attn_heads=8, memory_efficient=False), **kw})


def SRUnet256(**kw):
    return Unet(**{**dict(dim=128, dim_mults=(1, 2, 4, 8), num_resnet_blocks=(2, 4, 8, 8), layer_attns=(False, False, False, True),
                          layer_cross_attns=(False, False, False, True), attn_heads=8, memory_efficient=True, lowres_cond=True), **kw})


def SRUnet1024(**kw):
    return Unet(**{**dict(dim=128, dim_mults=(1, 2, 4, 8), num_resnet_blocks=(2, 4, 8, 8), layer_attns=False,
                          layer_cross_attns=(False, False, False, True), attn_heads=8, memory_efficient=True, lowres_cond=True), **kw})
