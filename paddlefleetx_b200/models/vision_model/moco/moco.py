"""MoCo v1 / v2 (reference vision_model/moco/moco.py:36-246): query + momentum key encoders, dictionary queue of K keys,
batch-shuffle BN across the world (all-gather + broadcast of the permutation), InfoNCE logits; ``MoCoV2Projector`` (MLP
head) and ``MoCoClassifier`` (linear probe on a frozen backbone)."""
from __future__ import annotations

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

from .. import resnet as R


@torch.no_grad()
def concat_all_gather(t: torch.Tensor) -> torch.Tensor:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return t
    out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t.contiguous())
    return torch.cat(out, 0)


class MoCoV2Projector(nn.Module):
    def __init__(self, with_pool: bool, in_dim: int, out_dim: int):
        super().__init__()
        self.with_pool = with_pool
        self.avgpool = nn.AdaptiveAvgPool2d(1) if with_pool else None
        self.mlp = nn.Sequential(nn.Linear(in_dim, out_dim), nn.ReLU())

    def forward(self, x):
        if self.with_pool:
            x = self.avgpool(x)
        return self.mlp(torch.flatten(x, 1))


class MoCoClassifier(nn.Module):
    def __init__(self, with_pool: bool, num_features: int, class_num: int):
        super().__init__()
        self.avgpool = nn.AdaptiveAvgPool2d(1) if with_pool else None
        self.fc = nn.Linear(num_features, class_num)
        nn.init.normal_(self.fc.weight, 0.0, 0.01); nn.init.zeros_(self.fc.bias)

    def forward(self, x):
        if self.avgpool is not None:
            x = self.avgpool(x)
        return self.fc(torch.flatten(x, 1))


class MoCo(nn.Module):
    def __init__(self, base_encoder=None, base_projector=None, base_classifier=None, momentum_encoder=None, momentum_projector=None,
                 momentum_classifier=None, dim: int = 128, K: int = 65536, m: float = 0.999, T: float = 0.07, backbone: str = "resnet50",
                 mlp: bool = False, **unused):
        super().__init__()
        self.K, self.m, self.T = K, m, T

        def make():
            enc = getattr(R, backbone)(num_classes=0, with_pool=False)
            feat = enc.out_features
            proj = MoCoV2Projector(True, feat, feat) if mlp else nn.Sequential(nn.AdaptiveAvgPool2d(1), nn.Flatten(1))
            return nn.Sequential(enc, proj, nn.Linear(feat, dim))

        self.base = make()
        self.momentum = make()
        for pq, pk in zip(self.base.parameters(), self.momentum.parameters()):
            pk.data.copy_(pq.data)
            pk.requires_grad = False
        self.register_buffer("queue", F.normalize(torch.randn(dim, K), dim=0))
        self.register_buffer("queue_ptr", torch.zeros(1, dtype=torch.long))

    @torch.no_grad()
    def _momentum_update(self):
        for pq, pk in zip(self.base.parameters(), self.momentum.parameters()):
            pk.data.mul_(self.m).add_(pq.data, alpha=1.0 - self.m)

    @torch.no_grad()
    def _dequeue_and_enqueue(self, keys):
        keys = concat_all_gather(keys)
        b = keys.shape[0]
        ptr = int(self.queue_ptr)
        assert self.K % b == 0, "queue size must be divisible by the global batch"
        self.queue[:, ptr:ptr + b] = keys.t()
        self.queue_ptr[0] = (ptr + b) % self.K

    @torch.no_grad()
    def _batch_shuffle(self, x):
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        if world == 1:
            perm = torch.randperm(x.shape[0], device=x.device)
            return x[perm], torch.argsort(perm)
        bs = x.shape[0]
        xg = concat_all_gather(x)
        perm = torch.randperm(xg.shape[0], device=x.device)
        dist.broadcast(perm, src=0)
        r = dist.get_rank()
        return xg[perm.view(world, -1)[r]], torch.argsort(perm)

    @torch.no_grad()
    def _batch_unshuffle(self, x, unshuffle):
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        if world == 1:
            return x[unshuffle]
        xg = concat_all_gather(x)
        r = dist.get_rank()
        return xg[unshuffle.view(world, -1)[r]]

    def forward(self, im_q, im_k=None):
        if im_k is None:                       # feature extraction
            return self.base(im_q)
        q = F.normalize(self.base(im_q), dim=1)
        with torch.no_grad():
            self._momentum_update()
            ks, unshuffle = self._batch_shuffle(im_k)
            k = F.normalize(self.momentum(ks), dim=1)
            k = self._batch_unshuffle(k, unshuffle)
        l_pos = torch.einsum("nc,nc->n", q, k).unsqueeze(-1)
        l_neg = torch.einsum("nc,ck->nk", q, self.queue.clone().detach().to(q.dtype))
        logits = torch.cat([l_pos, l_neg], 1) / self.T
        labels = torch.zeros(logits.shape[0], dtype=torch.long, device=logits.device)
        self._dequeue_and_enqueue(k)
        return logits, labels
