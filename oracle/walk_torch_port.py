"""torch port of the reference CRW train step -- TEST INFRASTRUCTURE / CPU BASELINE ONLY.

``crw_loss_reference_order`` restates src/model.py:22-46 with the reference's own op sequence
(F.normalize, einsum / tau, row softmax, bmm in the (T-2)^2 loop order, cross_entropy with a
probability target) so that (a) autograd through it is what the reference's ``loss.backward()``
computes, used by the GPU parity test that checks gradients reaching encoder parameters, and
(b) timing it on the host cores is the reference-form CPU baseline of bench.py.  The palindrome is
addressed by index instead of being materialised with cat/flip (model.py:31,41): same values.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def crw_loss_reference_order(emb: torch.Tensor, tau: float):
    """emb [B,T,N,C] raw encoder output -> (loss/N, A [B,T-1,N,N])."""
    B, T, N, _ = emb.shape
    E = F.normalize(emb, dim=-1)                                        # model.py:22
    A = torch.einsum("btnc,btmc->btnm", E[:, :-1], E[:, 1:]) / tau      # model.py:26
    eye = torch.eye(N, dtype=emb.dtype, device=emb.device).expand(B, N, N)
    loss = emb.new_zeros(())
    for k in range(1, T - 1):                                           # model.py:35
        M = eye
        for t in range(1, 2 * k):                                       # model.py:42 (index 0 skipped)
            step = A[:, t] if t < k else A[:, 2 * k - 1 - t].transpose(1, 2)
            M = torch.bmm(F.softmax(step, dim=-1), M)                   # model.py:44
        loss = loss + F.cross_entropy(M.transpose(1, 2), eye)           # model.py:45
    return loss / N, A


def make_resnet_encoder(in_ch: int = 1) -> torch.nn.Module:
    """Same architecture as the reference's ``Resnet`` (src/encoder.py:62-89): 1x1 conv with padding=1,
    BN, ReLU, then torchvision's ResNet(BasicBlock, [1,1,1,1], num_classes=128).  CPU-baseline use only."""
    from torchvision.models.resnet import BasicBlock, ResNet
    return torch.nn.Sequential(torch.nn.Conv2d(in_ch, 3, kernel_size=1, padding=1), torch.nn.BatchNorm2d(3),
                               torch.nn.ReLU(inplace=True), ResNet(BasicBlock, [1, 1, 1, 1], num_classes=128))


def train_step(encoder, optimizer, seq: torch.Tensor, tau: float):
    """One reference-shaped optimisation step (scripts/train.py:66-72) on whatever device ``seq`` is on."""
    B, T, N, H, W = seq.shape
    emb = encoder(seq.reshape(-1, H, W).unsqueeze(1)).reshape(B, T, N, -1)
    loss, _ = crw_loss_reference_order(emb, tau)
    optimizer.zero_grad()
    loss.backward()
    optimizer.step()
    return loss.detach()
