"""Vision building blocks (reference vision_model/layers/*.py): patch embedding, stochastic depth, initialisers."""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


def trunc_normal_(t: torch.Tensor, mean=0.0, std=1.0, a=-2.0, b=2.0):
    with torch.no_grad():
        return nn.init.trunc_normal_(t, mean, std, a, b)


def to_2tuple(x):
    return tuple(x) if isinstance(x, (tuple, list)) else (x, x)


class DropPath(nn.Module):
    def __init__(self, drop_prob: float = 0.0):
        super().__init__()
        self.drop_prob = float(drop_prob)

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1.0 - self.drop_prob
        mask = x.new_empty((x.shape[0],) + (1,) * (x.dim() - 1)).bernoulli_(keep)
        return x / keep * mask


class ViTPatchEmbed(nn.Module):
    """Patchify with a strided conv expressed as unfold + GEMM so that it runs on the tcgen05 GEMM."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768, dtype=None, device=None):
        super().__init__()
        self.img_size, self.patch_size = to_2tuple(img_size), to_2tuple(patch_size)
        self.num_patches = (self.img_size[0] // self.patch_size[0]) * (self.img_size[1] // self.patch_size[1])
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=self.patch_size, stride=self.patch_size, dtype=dtype, device=device)

    def forward(self, x):
        b, c, h, w = x.shape
        assert (h, w) == self.img_size, f"Input image size ({h}*{w}) doesn't match model ({self.img_size[0]}*{self.img_size[1]})."
        ph, pw = self.patch_size
        x = x.to(self.proj.weight.dtype)
        patches = x.reshape(b, c, h // ph, ph, w // pw, pw).permute(0, 2, 4, 1, 3, 5).reshape(b, -1, c * ph * pw)
        from ....ops import functional as OF

        return OF.linear(patches.contiguous(), self.proj.weight.reshape(self.proj.weight.shape[0], -1), self.proj.bias)


def drop_path(x: torch.Tensor, drop_prob: float = 0.0, training: bool = False) -> torch.Tensor:
    """Stochastic depth, functional form: drop whole samples of a residual branch and rescale the survivors."""
    if drop_prob == 0.0 or not training:
        return x
    keep = 1.0 - drop_prob
    mask = torch.bernoulli(torch.full((x.shape[0],) + (1,) * (x.dim() - 1), keep, device=x.device, dtype=x.dtype))
    return x * mask / keep


class Identity(nn.Module):
    def forward(self, x):
        return x


def xavier_uniform_2d_(weight: torch.Tensor) -> torch.Tensor:
    """Xavier-uniform for a conv / patch-embedding kernel viewed as a 2-D matrix ``[out, in * kh * kw]`` (the ViT reference initialiser)."""
    w2 = weight.view(weight.shape[0], -1)
    bound = (6.0 / (w2.shape[0] + w2.shape[1])) ** 0.5
    with torch.no_grad():
        return weight.uniform_(-bound, bound)
