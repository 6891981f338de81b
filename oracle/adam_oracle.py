"""TEST INFRASTRUCTURE ONLY (oracle): numpy restatement of the optimizer step of the reference's train loop,
scripts/train.py:56,72 -- ``torch.optim.Adam(model.parameters(), lr)`` with its defaults (betas (0.9, 0.999), eps 1e-8,
weight_decay 0, amsgrad False, maximize False).  The algorithm lives in PyTorch (torch/optim/adam.py, ``_single_tensor_adam``);
pinned in tests/test_oracle_adam.py against torch.optim.Adam itself run on the CPU in float64.
"""
import numpy as np


def adam_step(p, g, m, v, step, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0, grad_scale=1.0):
    """One update, ``step`` counting from 1; returns new (p, m, v) in the dtype of ``p``."""
    dt = p.dtype
    g = g.astype(dt) * dt.type(grad_scale)
    if weight_decay != 0.0:
        g = g + dt.type(weight_decay) * p
    m = m + (g - m) * dt.type(1.0 - beta1)
    v = v * dt.type(beta2) + dt.type(1.0 - beta2) * g * g
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    denom = np.sqrt(v) / dt.type(np.sqrt(bc2)) + dt.type(eps)
    p = p - dt.type(lr / bc1) * (m / denom)
    return p, m, v
