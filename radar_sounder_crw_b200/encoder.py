"""Encoders (reference: src/encoder.py).  OUT OF SCOPE as kernels: they stay plain PyTorch / cuDNN.

Only what the hot path's callers need is here: the two encoder families with *state-dict-compatible*
parameter names (``fc0``, ``bn0``, ``model.conv1`` ... as in the reference, encoder.py:62-89), so
checkpoints written by the reference's ``scripts/train.py:92`` load unchanged.  The ResNet trunk is
torchvision's own ``ResNet(BasicBlock, [1,1,1,1], num_classes=128)`` rather than a vendored copy.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F
from torchvision.models.resnet import BasicBlock, ResNet


def batchnorm_few_channels(x: torch.Tensor, bn: nn.BatchNorm2d) -> torch.Tensor:
    """``bn(x)`` for a BatchNorm2d with a handful of channels, written with plain tensor reductions.

    Same parameters, buffers, statistics update and arithmetic as ``nn.BatchNorm2d`` (so state dicts and results are
    interchangeable); the only difference is the execution strategy.  cuDNN's spatial batch-norm kernels parallelise
    over CHANNELS, so the 3-channel ``bn0`` of the reference encoder (src/encoder.py:68-74) on a
    [B*T*N, 3, 34, 34] batch runs on three thread blocks: 42 of the 75 ms of the config-2 train step on a B200.
    """
    if not x.is_cuda or bn.num_features > 8:
        return bn(x)
    if bn.training or not bn.track_running_stats:
        var, mean = torch.var_mean(x, dim=(0, 2, 3), unbiased=False)
        if bn.training and bn.track_running_stats:
            with torch.no_grad():
                n = x.numel() // x.shape[1]
                bn.num_batches_tracked += 1
                m = bn.momentum if bn.momentum is not None else 1.0 / float(bn.num_batches_tracked)
                bn.running_mean.mul_(1 - m).add_(mean.detach(), alpha=m)
                bn.running_var.mul_(1 - m).add_(var.detach() * (n / max(n - 1, 1)), alpha=m)
    else:
        mean, var = bn.running_mean, bn.running_var
    scale = torch.rsqrt(var + bn.eps)
    shift = -mean * scale
    if bn.affine:
        scale = scale * bn.weight
        shift = shift * bn.weight + bn.bias
    return x * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)


class Resnet(nn.Module):
    """1x1 conv (padding=1: 32x32 -> 34x34, as the reference) + BN + ReLU into a 1-block-per-stage ResNet."""

    def __init__(self, pos_embed=True, pretrained=None):
        super().__init__()
        self.fc0 = nn.Conv2d(2 if pos_embed else 1, 3, kernel_size=1, padding=1)
        self.bn0 = nn.BatchNorm2d(3)
        self.relu0 = nn.ReLU(inplace=True)
        self.model = ResNet(BasicBlock, [1, 1, 1, 1], num_classes=128)

    def forward(self, x):
        return self.model(self.relu0(batchnorm_few_channels(self.fc0(x), self.bn0)))


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


# ------------------------------------------------------------------------------------------------
# UNet-as-encoder adapter (BASELINE config 4).  NOT reference behaviour: in the reference, ``UNet`` (src/unet.py) is only a
# supervised segmentation baseline (scripts/test/test_unet.py:27) and never feeds CRW.  BASELINE.json's config 4 names it as
# the encoder, so this adapter builds the same 3-down / 3-up bilinear U-Net with ``n_classes = 128`` output channels (module
# names as in src/unet.py:80-104, so a reference ``UNet(1,128)`` state dict loads) and reduces the per-pixel map to one
# 128-d vector per patch with a global average pool.
# ------------------------------------------------------------------------------------------------
def _double_conv(cin, cout, cmid=None):
    cmid = cmid or cout
    m = nn.Module()
    m.double_conv = nn.Sequential(nn.Conv2d(cin, cmid, 3, padding=1, bias=False), nn.BatchNorm2d(cmid), nn.ReLU(inplace=True),
                                  nn.Conv2d(cmid, cout, 3, padding=1, bias=False), nn.BatchNorm2d(cout), nn.ReLU(inplace=True))
    return m


class _Stage(nn.Module):
    """One U-Net stage: ``down`` = maxpool + double conv; ``up`` = bilinear x2 + skip concat + double conv."""

    def __init__(self, kind, cin, cout):
        super().__init__()
        self.kind = kind
        if kind == "in":
            self.double_conv = _double_conv(cin, cout).double_conv
        elif kind == "down":
            self.maxpool_conv = nn.Sequential(nn.MaxPool2d(2), _double_conv(cin, cout))
            self.maxpool_conv[1].forward = lambda x, m=self.maxpool_conv[1]: m.double_conv(x)
        else:
            self.up = nn.Upsample(scale_factor=2, mode="bilinear", align_corners=True)
            self.conv = _double_conv(cin, cout, cin // 2)

    def forward(self, x, skip=None):
        if self.kind == "in":
            return self.double_conv(x)
        if self.kind == "down":
            return self.maxpool_conv(x)
        x = self.up(x)
        dy, dx = skip.shape[2] - x.shape[2], skip.shape[3] - x.shape[3]
        if dy or dx:
            x = F.pad(x, [dx // 2, dx - dx // 2, dy // 2, dy - dy // 2])
        return self.conv.double_conv(torch.cat([skip, x], dim=1))


class UNetEncoder(nn.Module):
    """``UNet(1 or 2, 128)`` + global average pool -> [patches, 128].  ``chunk`` patches at a time go through the network under
    activation checkpointing (a B=32, T=20, N=47 step is 30 080 patches of 32x32: the full-resolution activations of the
    whole batch would not fit next to each other); BatchNorm statistics are therefore per chunk."""

    def __init__(self, pos_embed=False, out_dim=128, chunk=4096):
        super().__init__()
        self.chunk = chunk
        self.inc = _Stage("in", 2 if pos_embed else 1, 64)
        self.down1, self.down2, self.down3 = _Stage("down", 64, 128), _Stage("down", 128, 256), _Stage("down", 256, 256)
        self.up1, self.up2, self.up3 = _Stage("up", 512, 128), _Stage("up", 256, 64), _Stage("up", 128, 64)
        self.outc = nn.Module()
        self.outc.conv = nn.Conv2d(64, out_dim, kernel_size=1)

    def _net(self, x):
        x1 = self.inc(x)
        x2 = self.down1(x1)
        x3 = self.down2(x2)
        x4 = self.down3(x3)
        y = self.up3(self.up2(self.up1(x4, x3), x2), x1)
        return self.outc.conv(y).mean(dim=(2, 3))

    def forward(self, x):
        if x.shape[0] <= self.chunk:
            return self._net(x)
        from torch.utils.checkpoint import checkpoint
        outs = [checkpoint(self._net, xc, use_reentrant=False) if torch.is_grad_enabled() else self._net(xc)
                for xc in x.split(self.chunk)]
        return torch.cat(outs)
