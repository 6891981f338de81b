"""Drop-in ``CRW`` module (reference: src/model.py:6-46).

Same constructor and ``forward(seq) -> (loss, A)`` contract; everything after the encoder call
(normalise, affinities, palindrome walk, cycle cross-entropy and their backward) is one
``crw_b200::walk_loss`` custom op instead of ~240 ATen launches.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from .utils import pos_embed as _pos_embed


class CRW(nn.Module):
    """Contrastive random walk loss over a sequence of column-chunk frames.

    Args mirror the reference (src/model.py:7): ``encoder``, ``tau``, ``pos_embed``, ``only_a``.
    Extra keyword arguments (not in the reference, defaults keep its behaviour):
      need_A     return the affinities ``A [B,T-1,N,N]`` (model.py:46).  ``scripts/train.py:67``
                 discards them; pass False on the timed path to skip the N x N write.
      precision  ops.PREC_FP32 (default) | ops.PREC_BF16X3 (error-compensated bf16 pairs on the tensor cores)
    """

    def __init__(self, encoder, tau, pos_embed, only_a=False, need_A=True, precision=ops.PREC_FP32):
        super().__init__()
        self.encoder = encoder
        self.tau = tau
        self.pos_embed = pos_embed
        self.only_a = only_a
        self.need_A = need_A
        self.precision = precision

    def forward(self, seq):
        B, T, N, H, W = seq.shape
        x = seq.reshape(-1, H, W).unsqueeze(1)                 # model.py:17
        if self.pos_embed:
            x = _pos_embed(x)                                   # model.py:19
        emb = self.encoder(x).reshape(B, T, N, -1).float()      # model.py:20-21
        if self.only_a:                                         # model.py:27-28
            _, A, _ = ops.walk_loss(emb, float(self.tau), True, self.precision)
            return A
        loss, A, _ = ops.walk_loss(emb, float(self.tau), bool(self.need_A), self.precision)
        return loss, (A if self.need_A else None)
