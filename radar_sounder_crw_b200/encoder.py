"""Encoders (reference: src/encoder.py).  OUT OF SCOPE as kernels: they stay plain PyTorch / cuDNN.

Only what the hot path's callers need is here: the two encoder families with *state-dict-compatible*
parameter names (``fc0``, ``bn0``, ``model.conv1`` ... as in the reference, encoder.py:62-89), so
checkpoints written by the reference's ``scripts/train.py:92`` load unchanged.  The ResNet trunk is
torchvision's own ``ResNet(BasicBlock, [1,1,1,1], num_classes=128)`` rather than a vendored copy.
"""
from __future__ import annotations

import torch.nn as nn
from torchvision.models.resnet import BasicBlock, ResNet


class Resnet(nn.Module):
    """1x1 conv (padding=1: 32x32 -> 34x34, as the reference) + BN + ReLU into a 1-block-per-stage ResNet."""

    def __init__(self, pos_embed=True, pretrained=None):
        super().__init__()
        self.fc0 = nn.Conv2d(2 if pos_embed else 1, 3, kernel_size=1, padding=1)
        self.bn0 = nn.BatchNorm2d(3)
        self.relu0 = nn.ReLU(inplace=True)
        self.model = ResNet(BasicBlock, [1, 1, 1, 1], num_classes=128)

    def forward(self, x):
        return self.model(self.relu0(self.bn0(self.fc0(x))))


class CNN(nn.Module):
    """Five-conv encoder with two stride-1 max-pools, global average pool and a linear head (128-d)."""

    def __init__(self, pos_embed):
        super().__init__()
        self.conv1 = nn.Conv2d(2 if pos_embed else 1, 8, kernel_size=5, padding=1)
        self.relu1 = nn.ReLU()
        self.pool1 = nn.MaxPool2d(kernel_size=2, stride=1)
        self.conv2 = nn.Conv2d(8, 32, kernel_size=5, padding=1)
        self.relu2 = nn.ReLU()
        self.pool2 = nn.MaxPool2d(kernel_size=2, stride=1)
        self.conv3 = nn.Conv2d(32, 64, kernel_size=3, padding=1)
        self.relu3 = nn.ReLU()
        self.conv4 = nn.Conv2d(64, 128, kernel_size=3, padding=1)
        self.relu4 = nn.ReLU()
        self.conv5 = nn.Conv2d(128, 128, kernel_size=3, padding=1)
        self.relu5 = nn.ReLU()
        self.global_avg_pool = nn.AdaptiveAvgPool2d(1)
        self.fc = nn.Linear(128, 128)

    def forward(self, x):
        x = self.pool1(self.relu1(self.conv1(x)))
        x = self.pool2(self.relu2(self.conv2(x)))
        x = self.relu5(self.conv5(self.relu4(self.conv4(self.relu3(self.conv3(x))))))
        return self.fc(self.global_avg_pool(x).flatten(1))
