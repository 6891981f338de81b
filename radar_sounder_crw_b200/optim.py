"""Train-loop glue (reference scripts/train.py:56,70-72): ``Adam(model.parameters(), lr)``, ``zero_grad()``, ``step()``.

``FlatAdam`` keeps every parameter and every gradient as a VIEW of one flat buffer each (same sizes and strides as before, so
channels_last weights stay channels_last) next to flat moment buffers; ``zero_grad()`` is one memset, ``step()`` is -- for N > 1
one SUM all-reduce of the gradient buffer and -- ONE launch of ``crw_b200::adam_step`` over all parameters (the 1 / world of the
average rides in the kernel).  Same update rule as torch.optim.Adam (amsgrad / maximize off), restated in oracle/adam_oracle.py.
There is no CPU path: the parameters must live on a CUDA device.

Measured at BASELINE config 2 (bench.py, CRW_BENCH_FLAT_ADAM): on 2 B200s 34.71 ms per step against 34.82 ms for
parallel.FlatGradients + torch.optim.Adam(fused=True) (weak-scaling efficiency 0.995 against 0.992) -- the default of bench.py for
N > 1; on ONE B200 34.66 against 34.53 ms for torch's optimizer with zero_grad(set_to_none=True): autograd accumulates into gradient
views that already exist (one add per parameter) where set_to_none lets it hand the fresh gradient over, which costs a little more
than the single launch saves -- bench.py keeps torch's optimizer at N = 1.  (Views that do not start on 256-byte boundaries cost
another 0.35 ms: every tensor is padded to 64 elements.)
"""
from __future__ import annotations

from typing import Iterable, Tuple

import torch
import torch.distributed as dist

from . import ops


def _flat_views(tensors, flat):
    """Re-seat each tensor's storage as a view of ``flat`` (dense tensors keep their strides); returns the views."""
    out, off = [], 0
    for t in tensors:
        dense = t.is_contiguous() or (t.dim() == 4 and t.is_contiguous(memory_format=torch.channels_last))
        v = torch.as_strided(flat, t.size(), t.stride(), off) if dense else flat[off:off + t.numel()].view(t.size())
        out.append(v)
        off += _padded(t.numel())
    return out


def _padded(n: int) -> int:
    """Every tensor starts on a 256-byte boundary of the flat buffer (vector loads / cuDNN alignment); the pad elements stay zero."""
    return (n + 63) // 64 * 64


class FlatAdam:
    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 1e-3, betas: Tuple[float, float] = (0.9, 0.999),
                 eps: float = 1e-8, weight_decay: float = 0.0, world: int = 1, group=None):
        self.ps = [p for p in params if p.requires_grad]
        if not self.ps:
            raise ValueError("FlatAdam: no parameters")
        ref = self.ps[0]
        if not ref.is_cuda or any(p.device != ref.device or p.dtype != torch.float32 for p in self.ps):
            raise RuntimeError("FlatAdam: all parameters must be f32 on one CUDA device (there is no CPU path)")
        self.lr, self.betas, self.eps, self.weight_decay = float(lr), (float(betas[0]), float(betas[1])), float(eps), float(weight_decay)
        self.world, self.group, self.steps = int(world), group, 0
        n = sum(_padded(p.numel()) for p in self.ps)
        self.flat_p = torch.zeros(n, device=ref.device, dtype=torch.float32)
        self.flat_g = torch.zeros(n, device=ref.device, dtype=torch.float32)
        self.exp_avg = torch.zeros(n, device=ref.device, dtype=torch.float32)
        self.exp_avg_sq = torch.zeros(n, device=ref.device, dtype=torch.float32)
        with torch.no_grad():
            for p, v in zip(self.ps, _flat_views([p.data for p in self.ps], self.flat_p)):
                v.copy_(p.data)
                p.data = v
        for p, v in zip(self.ps, _flat_views(self.ps, self.flat_g)):
            p.grad = v

    def zero_grad(self) -> None:
        """One memset (``set_to_none`` would drop the views the backward accumulates into)."""
        self.flat_g.zero_()

    @torch.no_grad()
    def step(self) -> None:
        if self.world > 1:
            dist.all_reduce(self.flat_g, op=dist.ReduceOp.SUM, group=self.group)
        self.steps += 1
        ops.adam_step(self.flat_p, self.flat_g, self.exp_avg, self.exp_avg_sq, self.lr, self.betas[0], self.betas[1], self.eps,
                      self.weight_decay, self.steps, 1.0 / self.world)

    def state_dict(self):
        return dict(step=self.steps, exp_avg=self.exp_avg.clone(), exp_avg_sq=self.exp_avg_sq.clone(), lr=self.lr, betas=self.betas,
                    eps=self.eps, weight_decay=self.weight_decay)

    def load_state_dict(self, sd) -> None:
        self.steps = int(sd["step"])
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
