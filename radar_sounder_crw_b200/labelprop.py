"""Drop-in ``LabelPropVOS_CRW`` (reference: src/imported/labelprop.py:42-116).

``predict`` keeps the reference's stepwise contract (lists of per-frame features / masks) for callers
that drive the loop themselves; ``propagate`` in utils.py uses the fused whole-sequence op instead.
"""
from __future__ import annotations

import torch

from . import ops


class LabelPropVOS_CRW(object):
    def __init__(self, cfg, precision=ops.PREC_AUTO, mode=ops.LP_REF_EXACT):
        """cfg: the reference's dict (CXT_SIZE, RADIUS, TEMP, KNN; labelprop.py:44-48).  ``precision`` (not in the reference):
        PREC_AUTO (default) = the exact tensor path where the shape allows, else the fp32 kernel -- identical results either
        way; PREC_BF16X3 = the approximate error-compensated bf16 kernel (>= 99.8 % of the labels on near-collinear features)."""
        self.cxt_size = cfg["CXT_SIZE"]
        self.radius = cfg["RADIUS"]
        self.temperature = cfg["TEMP"]
        self.topk = cfg["KNN"]
        self.precision = precision
        self.mode = mode
        self.mask = None
        self.mask_hw = None

    # reference labelprop.py:52-65 (never called by the reference either; kept for API parity)
    def context_long(self, t0, t):
        return [t0]

    def context_short(self, t0, t):
        return [max(tt, t0) for tt in range(t - self.cxt_size, t)]

    def context_index(self, t0, t):
        return self.context_long(t0, t) + self.context_short(t0, t)

    def predict(self, feats, masks, curr_feat, ref_index=None, t=None):
        """feats: list of n [1,C,h,1]; masks: list of n [1,M,h,1]; curr_feat [1,C,h,1] -> [1,M,h,1]."""
        h, w = curr_feat.shape[-2:]
        if w != 1:
            raise NotImplementedError("crw_b200 supports the reference's h=N, w=1 node grid only")
        n, ctx = len(feats), self.cxt_size
        # frames that survive the trim of maskedatt.py:166-167 ...
        kf = list(range(n)) if n <= ctx + 1 else [0] + list(range(n - ctx, n))
        # ... and the frames the ids are gathered from (labelprop.py:82,106: the untrimmed list -> SURVEY F5)
        lf = kf if (n <= ctx + 1 or self.mode == ops.LP_FIXED) else list(range(ctx + 1))
        keys = torch.stack([feats[f][0, :, :, 0].t() for f in kf]).contiguous().float()       # [F,N,C]
        lbl = torch.stack([masks[f][0, :, :, 0] for f in lf]).contiguous().float()            # [F,M,N]
        query = curr_feat[0, :, :, 0].t().contiguous().float()[None]                          # [1,N,C]
        F = len(kf)
        # the stepwise API always runs the fp32 kernel (the tensor-core kernel works on whole sequences: propagate())
        W, I = ops.affinity_topk(keys, query, F, max(F, 1), float(self.radius), float(self.temperature),
                                 int(self.topk), ops.PREC_FP32)
        pred = ops.label_gather_step(W[0], I[0], lbl)                                          # [M,N]
        return pred[None, :, :, None]
