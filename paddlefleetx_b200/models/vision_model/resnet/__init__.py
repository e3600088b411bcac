"""ResNet-18/34/50/101/152 (the reference re-exports ``paddle.vision.models.resnet*`` as the MoCo backbone,
vision_model/resnet/__init__.py:14-23 — supplied natively here).  ``with_pool`` / ``num_classes <= 0`` return features."""
from __future__ import annotations

import torch
import torch.nn as nn


class BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, inp, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = nn.Conv2d(inp, planes, 3, stride, 1, bias=False); self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, 1, 1, bias=False); self.bn2 = nn.BatchNorm2d(planes)
        self.downsample = downsample

    def forward(self, x):
        idt = x if self.downsample is None else self.downsample(x)
        out = torch.relu(self.bn1(self.conv1(x)))
        return torch.relu(self.bn2(self.conv2(out)) + idt)


class BottleneckBlock(nn.Module):
    expansion = 4

    def __init__(self, inp, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = nn.Conv2d(inp, planes, 1, bias=False); self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, stride, 1, bias=False); self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = nn.Conv2d(planes, planes * 4, 1, bias=False); self.bn3 = nn.BatchNorm2d(planes * 4)
        self.downsample = downsample

    def forward(self, x):
        idt = x if self.downsample is None else self.downsample(x)
        out = torch.relu(self.bn1(self.conv1(x)))
        out = torch.relu(self.bn2(self.conv2(out)))
        return torch.relu(self.bn3(self.conv3(out)) + idt)


class ResNet(nn.Module):
    def __init__(self, block, depth_cfg, num_classes=1000, with_pool=True, zero_init_residual=False):
        super().__init__()
        self.inplanes = 64
        self.conv1 = nn.Conv2d(3, 64, 7, 2, 3, bias=False); self.bn1 = nn.BatchNorm2d(64)
        self.maxpool = nn.MaxPool2d(3, 2, 1)
        self.layer1 = self._make(block, 64, depth_cfg[0]); self.layer2 = self._make(block, 128, depth_cfg[1], 2)
        self.layer3 = self._make(block, 256, depth_cfg[2], 2); self.layer4 = self._make(block, 512, depth_cfg[3], 2)
        self.with_pool, self.num_classes = with_pool, num_classes
        self.avgpool = nn.AdaptiveAvgPool2d(1) if with_pool else None
        self.fc = nn.Linear(512 * block.expansion, num_classes) if num_classes > 0 else None
        self.out_features = 512 * block.expansion
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
        if zero_init_residual:
            for m in self.modules():
                if isinstance(m, BottleneckBlock):
                    nn.init.zeros_(m.bn3.weight)
                elif isinstance(m, BasicBlock):
                    nn.init.zeros_(m.bn2.weight)

    def _make(self, block, planes, n, stride=1):
        down = None
        if stride != 1 or self.inplanes != planes * block.expansion:
            down = nn.Sequential(nn.Conv2d(self.inplanes, planes * block.expansion, 1, stride, bias=False), nn.BatchNorm2d(planes * block.expansion))
        layers = [block(self.inplanes, planes, stride, down)]
        self.inplanes = planes * block.expansion
        layers += [block(self.inplanes, planes) for _ in range(1, n)]
        return nn.Sequential(*layers)

    def forward(self, x):
        x = self.maxpool(torch.relu(self.bn1(self.conv1(x))))
        x = self.layer4(self.layer3(self.layer2(self.layer1(x))))
        if self.with_pool:
            x = self.avgpool(x)
        if self.fc is not None:
            x = self.fc(torch.flatten(x, 1))
        return x


def resnet18(**kw): return ResNet(BasicBlock, [2, 2, 2, 2], **kw)
def resnet34(**kw): return ResNet(BasicBlock, [3, 4, 6, 3], **kw)
def resnet50(**kw): return ResNet(BottleneckBlock, [3, 4, 6, 3], **kw)
def resnet101(**kw): return ResNet(BottleneckBlock, [3, 4, 23, 3], **kw)
def resnet152(**kw): return ResNet(BottleneckBlock, [3, 8, 36, 3], **kw)
