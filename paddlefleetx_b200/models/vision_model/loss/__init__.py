"""Classification losses (reference vision_model/loss/cross_entropy.py:25,64): ``CELoss`` (optional label smoothing,
soft labels) and ``ViTCELoss`` (sigmoid / BCE-style pre-training loss with eps)."""
import torch
import torch.nn as nn
import torch.nn.functional as F


class CELoss(nn.Module):
    def __init__(self, epsilon=None):
        super().__init__()
        self.epsilon = epsilon if epsilon is not None and 0 < epsilon < 1 else None

    def forward(self, x, label):
        if isinstance(x, dict):
            x = x["logits"]
        x = x.float()
        if self.epsilon is not None:
            n = x.shape[-1]
            onehot = F.one_hot(label.reshape(-1).long(), n).float()
            soft = onehot * (1 - self.epsilon) + self.epsilon / n
            return (-(soft * F.log_softmax(x, -1)).sum(-1)).mean()
        if label.dim() == x.dim() and label.shape[-1] == x.shape[-1] and label.dtype.is_floating_point:
            return (-(label * F.log_softmax(x, -1)).sum(-1)).mean()
        return F.cross_entropy(x, label.reshape(-1).long())


class ViTCELoss(nn.Module):
    def __init__(self, epsilon=None):
        super().__init__()
        self.epsilon = epsilon

    def forward(self, x, label):
        if isinstance(x, dict):
            x = x["logits"]
        x = x.float()
        n = x.shape[-1]
        if label.dim() == 1 or label.shape[-1] != n:
            label = F.one_hot(label.reshape(-1).long(), n).float()
        if self.epsilon is not None:
            label = label * (1 - self.epsilon) + self.epsilon / n
        return F.binary_cross_entropy_with_logits(x, label, reduction="none").sum(-1).mean()
