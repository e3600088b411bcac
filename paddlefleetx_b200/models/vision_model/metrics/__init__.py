"""``TopkAcc`` (reference vision_model/metrics/accuracy.py:19)."""
import torch
import torch.nn as nn


class TopkAcc(nn.Module):
    def __init__(self, topk=(1, 5)):
        super().__init__()
        self.topk = [topk] if isinstance(topk, int) else list(topk)

    def forward(self, x, label):
        if isinstance(x, dict):
            x = x["logits"]
        k = min(max(self.topk), x.shape[-1])
        pred = x.float().topk(k, -1).indices
        hit = pred == label.reshape(-1, 1)
        return {f"top{t}": hit[:, :min(t, k)].any(-1).float().mean() for t in self.topk}
