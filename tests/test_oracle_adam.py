"""(CPU) The optimizer-step oracle against torch.optim.Adam itself -- the reference's optimizer (scripts/train.py:56) --
run on the CPU in float64: the restatement must agree to rounding over several steps, with and without weight decay."""
import numpy as np
import pytest
import torch

from oracle import adam_oracle as ao


@pytest.mark.parametrize("wd", [0.0, 0.01])
def test_adam_oracle_matches_torch_adam_float64(wd):
    rs = np.random.RandomState(0)
    p0 = rs.randn(257)
    grads = [rs.randn(257) * (0.1 + i) for i in range(6)]
    tp = torch.nn.Parameter(torch.tensor(p0, dtype=torch.float64))
    opt = torch.optim.Adam([tp], lr=1e-3, weight_decay=wd)
    p, m, v = p0.copy(), np.zeros_like(p0), np.zeros_like(p0)
    for i, g in enumerate(grads):
        tp.grad = torch.tensor(g, dtype=torch.float64)
        opt.step()
        p, m, v = ao.adam_step(p, g, m, v, i + 1, lr=1e-3, weight_decay=wd)
        assert np.max(np.abs(p - tp.detach().numpy())) <= 1e-13 * max(1.0, np.max(np.abs(p)))
    st = opt.state[tp]
    assert np.allclose(m, st["exp_avg"].numpy(), rtol=1e-12, atol=1e-15)
    assert np.allclose(v, st["exp_avg_sq"].numpy(), rtol=1e-12, atol=1e-15)


def test_adam_oracle_grad_scale_is_a_gradient_factor():
    rs = np.random.RandomState(1)
    p0, g = rs.randn(64), rs.randn(64)
    a = ao.adam_step(p0, g * 0.25, np.zeros(64), np.zeros(64), 1)
    b = ao.adam_step(p0, g, np.zeros(64), np.zeros(64), 1, grad_scale=0.25)
    for x, y in zip(a, b):
        assert np.allclose(x, y, rtol=1e-15, atol=0)
