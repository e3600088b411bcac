"""Vision Transformer family (14 presets, tiny -> 6B) — reference ppfleetx/models/vision_model/vit/vit.py:54-609.

One ``Block`` implementation serves both of the reference's variants: its ``FusedBlock`` (Paddle ``FusedMultiHeadAttention`` +
``FusedFeedForward``) exists only to reach fused kernels, whereas here every block already runs the fused path (tcgen05 GEMMs
with bias epilogue, fused LayerNorm, flash attention, fused bias+GELU, Philox dropout+residual).  ``use_fused_attn`` is
accepted for config parity; checkpoints always use the un-fused key names (``blocks.N.attn.qkv.weight`` ...), which is also
what the reference writes (``replaced_dict``, vit.py:283-420).
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from ....ops import attention as ATT
from ....ops import functional as OF
from ..layers import DropPath, ViTPatchEmbed, trunc_normal_


class ViTAttention(nn.Module):
    def __init__(self, dim, num_heads=8, qkv_bias=False, qk_scale=None, attn_drop=0.0, proj_drop=0.0, dtype=None, device=None):
        super().__init__()
        self.num_heads, self.head_dim = num_heads, dim // num_heads
        self.scale = qk_scale or self.head_dim ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias, dtype=dtype, device=device)
        self.proj = nn.Linear(dim, dim, dtype=dtype, device=device)
        self.attn_drop, self.proj_drop = attn_drop, proj_drop

    def forward(self, x):
        b, n, c = x.shape
        qkv = OF.linear(x, self.qkv.weight, self.qkv.bias).view(b, n, 3, self.num_heads, self.head_dim)
        q, k, v = qkv.unbind(2)
        o = ATT.attention(q, k, v, causal=False, dropout_p=self.attn_drop if self.training else 0.0, scale=self.scale)
        return OF.linear(o.reshape(b, n, c), self.proj.weight, None), self.proj.bias


class ViTMLP(nn.Module):
    def __init__(self, dim, hidden, drop=0.0, act="gelu", dtype=None, device=None):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden, dtype=dtype, device=device)
        self.fc2 = nn.Linear(hidden, dim, dtype=dtype, device=device)
        self.drop, self.act = drop, act

    def forward(self, x):
        h = OF.linear(x, self.fc1.weight, None)
        if self.act == "gelu":
            h = OF.bias_gelu(h, self.fc1.bias, exact=True)        # nn.GELU (erf form), as in the reference ViT MLP
        else:
            h = F.relu(h + self.fc1.bias)
        h = OF.dropout(h, self.drop, self.training)
        return OF.linear(h, self.fc2.weight, None), self.fc2.bias


class Block(nn.Module):
    def __init__(self, dim, num_heads, mlp_ratio=4.0, qkv_bias=False, qk_scale=None, drop=0.0, attn_drop=0.0, drop_path=0.0, act="gelu",
                 epsilon=1e-5, dtype=None, device=None):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=epsilon, dtype=dtype, device=device)
        self.attn = ViTAttention(dim, num_heads, qkv_bias, qk_scale, attn_drop, drop, dtype, device)
        self.drop_path = DropPath(drop_path) if drop_path > 0.0 else None
        self.norm2 = nn.LayerNorm(dim, eps=epsilon, dtype=dtype, device=device)
        self.mlp = ViTMLP(dim, int(dim * mlp_ratio), drop, act, dtype, device)
        self.drop = drop

    def forward(self, x):
        y, b = self.attn(OF.layer_norm(x, self.norm1.weight, self.norm1.bias, self.norm1.eps))
        if self.drop_path is None:
            x = OF.bias_dropout_add(y, b, x, self.drop, self.training)
        else:
            x = x + self.drop_path(OF.dropout(y + b, self.drop, self.training))
        y, b = self.mlp(OF.layer_norm(x, self.norm2.weight, self.norm2.bias, self.norm2.eps))
        if self.drop_path is None:
            return OF.bias_dropout_add(y, b, x, self.drop, self.training)
        return x + self.drop_path(OF.dropout(y + b, self.drop, self.training))


FusedBlock = Block


class ViT(nn.Module):
    def __init__(self, img_size=224, patch_size=16, in_chans=3, class_num=1000, embed_dim=768, depth=12, num_heads=12, mlp_ratio=4,
                 qkv_bias=False, qk_scale=None, drop_rate=0.0, attn_drop_rate=0.0, drop_path_rate=0.0, epsilon=1e-5, representation_size=None,
                 use_fused_attn=False, use_recompute=False, dtype=None, device=None, **unused):
        super().__init__()
        self.class_num, self.num_features, self.embed_dim = class_num, embed_dim, embed_dim
        self.use_recompute = use_recompute
        self.patch_embed = ViTPatchEmbed(img_size, patch_size, in_chans, embed_dim, dtype=dtype, device=device)
        n = self.patch_embed.num_patches
        self.pos_embed = nn.Parameter(torch.zeros(1, n + 1, embed_dim, dtype=dtype, device=device))
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim, dtype=dtype, device=device))
        self.pos_drop = drop_rate
        dpr = [float(v) for v in torch.linspace(0, drop_path_rate, depth)]
        self.blocks = nn.ModuleList([Block(embed_dim, num_heads, mlp_ratio, qkv_bias, qk_scale, drop_rate, attn_drop_rate, dpr[i], "gelu", epsilon,
                                           dtype, device) for i in range(depth)])
        self.norm = nn.LayerNorm(embed_dim, eps=epsilon, dtype=dtype, device=device)
        if representation_size is not None:
            self.head0 = nn.Linear(embed_dim, representation_size, dtype=dtype, device=device)
            self.num_features = representation_size
        else:
            self.head0 = None
        self.head = nn.Linear(self.num_features, class_num, dtype=dtype, device=device) if class_num > 0 else nn.Identity()
        trunc_normal_(self.pos_embed, std=0.02)
        self.apply(self._init_weights)
        if isinstance(self.head, nn.Linear):
            nn.init.zeros_(self.head.weight); nn.init.zeros_(self.head.bias)

    @staticmethod
    def _init_weights(m):
        if isinstance(m, nn.Linear):
            nn.init.xavier_uniform_(m.weight)
            if m.bias is not None:
                nn.init.normal_(m.bias, std=1e-6)
        elif isinstance(m, nn.LayerNorm):
            nn.init.zeros_(m.bias); nn.init.ones_(m.weight)

    def forward_features(self, x):
        b = x.shape[0]
        x = self.patch_embed(x)
        x = torch.cat([self.cls_token.expand(b, -1, -1).to(x.dtype), x], 1) + self.pos_embed.to(x.dtype)
        x = OF.dropout(x, self.pos_drop, self.training)
        for blk in self.blocks:
            if self.use_recompute and self.training:
                from ....parallel.recompute import recompute

                x = recompute(blk, x)
            else:
                x = blk(x)
        x = OF.layer_norm(x, self.norm.weight, self.norm.bias, self.norm.eps)
        return x[:, 0]

    def forward(self, x):
        x = self.forward_features(x)
        if self.head0 is not None:
            x = torch.tanh(self.head0(x))
        return self.head(x)

    def load_pretrained(self, prefix_path: str, finetune: bool = False):
        """Load ``<prefix>.pdparams``; on fine-tune the classifier is re-initialised and the position embedding is
        bilinearly interpolated to the new grid (reference vit.py:210-282)."""
        state = torch.load(prefix_path + ".pdparams" if not prefix_path.endswith(".pdparams") else prefix_path, map_location="cpu", weights_only=False)
        own = self.state_dict()
        if finetune:
            for k in ("head.weight", "head.bias", "head0.weight", "head0.bias"):
                state.pop(k, None)
            pe = state.get("pos_embed")
            if pe is not None and pe.shape != own["pos_embed"].shape:
                cls, grid = pe[:, :1], pe[:, 1:]
                g0 = int(math.sqrt(grid.shape[1])); g1 = int(math.sqrt(own["pos_embed"].shape[1] - 1))
                grid = F.interpolate(grid.reshape(1, g0, g0, -1).permute(0, 3, 1, 2).float(), size=(g1, g1), mode="bilinear", align_corners=False)
                state["pos_embed"] = torch.cat([cls, grid.permute(0, 2, 3, 1).reshape(1, g1 * g1, -1).to(pe.dtype)], 1)
        load = {k: v.to(own[k].dtype) for k, v in state.items() if k in own and own[k].shape == v.shape}
        return self.load_state_dict(load, strict=False)


_PRESETS = {
    "ViT_tiny_patch16_224": dict(patch_size=16, embed_dim=192, depth=12, num_heads=3, mlp_ratio=4, representation_size=192),
    "ViT_base_patch16_224": dict(patch_size=16, embed_dim=768, depth=12, num_heads=12, mlp_ratio=4, representation_size=768),
    "ViT_base_patch16_384": dict(img_size=384, patch_size=16, embed_dim=768, depth=12, num_heads=12, mlp_ratio=4, representation_size=None),
    "ViT_base_patch32_224": dict(patch_size=32, embed_dim=768, depth=12, num_heads=12, mlp_ratio=4, representation_size=768),
    "ViT_base_patch32_384": dict(img_size=384, patch_size=32, embed_dim=768, depth=12, num_heads=12, mlp_ratio=4, representation_size=None),
    "ViT_large_patch16_224": dict(patch_size=16, embed_dim=1024, depth=24, num_heads=16, mlp_ratio=4, representation_size=1024),
    "ViT_large_patch16_384": dict(img_size=384, patch_size=16, embed_dim=1024, depth=24, num_heads=16, mlp_ratio=4, representation_size=None),
    "ViT_large_patch32_224": dict(patch_size=32, embed_dim=1024, depth=24, num_heads=16, mlp_ratio=4, representation_size=1024),
    "ViT_large_patch32_384": dict(img_size=384, patch_size=32, embed_dim=1024, depth=24, num_heads=16, mlp_ratio=4, representation_size=None),
    "ViT_huge_patch14_224": dict(patch_size=14, embed_dim=1280, depth=32, num_heads=16, mlp_ratio=4, representation_size=1280),
    "ViT_huge_patch14_384": dict(img_size=384, patch_size=14, embed_dim=1280, depth=32, num_heads=16, mlp_ratio=4, representation_size=None),
    "ViT_g_patch14_224": dict(img_size=224, patch_size=14, embed_dim=1408, depth=40, num_heads=16, mlp_ratio=4.364, representation_size=1408),
    "ViT_G_patch14_224": dict(img_size=224, patch_size=14, embed_dim=1664, depth=48, num_heads=16, mlp_ratio=4.9231, representation_size=1664),
    "ViT_6B_patch14_224": dict(img_size=224, patch_size=14, embed_dim=2320, depth=80, num_heads=16, mlp_ratio=4.955, representation_size=2320),
}


def _factory(name):
    def build(**kwargs):
        cfg = dict(_PRESETS[name], qkv_bias=True, epsilon=1e-6)
        cfg.update(kwargs)
        pretrained = cfg.pop("pretrained", None)
        model = ViT(**cfg)
        if pretrained and pretrained.get("prefix_path"):
            import os

            if os.path.exists(pretrained["prefix_path"] + ".pdparams"):
                model.load_pretrained(pretrained["prefix_path"], bool(pretrained.get("finetune", False)))
        return model
    build.__name__ = name
    return build


for _n in _PRESETS:
    globals()[_n] = _factory(_n)
__all__ = ["ViT", "Block", "FusedBlock"] + list(_PRESETS)
