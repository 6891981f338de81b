"""(GPU) Train-loop glue: crw_b200::adam_step / optim.FlatAdam against the oracle and against torch.optim.Adam, the reference's
optimizer (scripts/train.py:56,70-72).  Floating point: 2e-6 relative on the parameters after several steps (the kernel uses
fused multiply-adds and a reciprocal of sqrt(1 - beta2^t); torch's own fused and unfused Adam differ by as much)."""
import copy

import numpy as np
import pytest
import torch

from oracle import adam_oracle as ao

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pkg():
    import radar_sounder_crw_b200 as crw
    crw._lib.lib()
    return crw


def _rel(a, b):
    return float(np.linalg.norm(a.astype(np.float64) - b.astype(np.float64)) / max(np.linalg.norm(b.astype(np.float64)), 1e-30))


@pytest.mark.parametrize("n", [1, 3, 4, 1027, 1 << 20])
@pytest.mark.parametrize("wd", [0.0, 0.01])
def test_adam_step_kernel_vs_oracle(pkg, n, wd):
    rs = np.random.RandomState(n % 97)
    p0 = rs.randn(n).astype(np.float32)
    p = torch.tensor(p0, device="cuda")
    m = torch.zeros(n, device="cuda")
    v = torch.zeros(n, device="cuda")
    po, mo, vo = p0.astype(np.float64), np.zeros(n), np.zeros(n)
    for step in range(1, 6):
        g = (rs.randn(n) * step).astype(np.float32)
        pkg.ops.adam_step(p, torch.tensor(g, device="cuda"), m, v, 1e-3, 0.9, 0.999, 1e-8, wd, step, 0.5)
        po, mo, vo = ao.adam_step(po, g.astype(np.float64), mo, vo, step, weight_decay=wd, grad_scale=0.5)
    torch.cuda.synchronize()
    assert _rel(p.cpu().numpy(), po) < 2e-6
    assert _rel(m.cpu().numpy(), mo) < 2e-6
    assert _rel(v.cpu().numpy(), vo) < 2e-6


def test_adam_step_rejects_bad_arguments(pkg):
    p = torch.zeros(8, device="cuda")
    with pytest.raises(RuntimeError):
        pkg.ops.adam_step(p, torch.zeros(7, device="cuda"), p.clone(), p.clone(), 1e-3, 0.9, 0.999, 1e-8, 0.0, 1, 1.0)
    with pytest.raises(RuntimeError):
        pkg.ops.adam_step(p, p.clone(), p.clone(), p.clone(), 1e-3, 0.9, 0.999, 1e-8, 0.0, 0, 1.0)      # step counts from 1
    with pytest.raises(RuntimeError):
        pkg.ops.adam_step(torch.zeros(8), torch.zeros(8), torch.zeros(8), torch.zeros(8), 1e-3, 0.9, 0.999, 1e-8, 0.0, 1, 1.0)   # no CPU path


def test_flat_adam_trains_like_torch_adam(pkg):
    """The reference's loop (train.py:70-72) on the Resnet encoder in channels_last: zero_grad / backward / step with FlatAdam;
    torch.optim.Adam steps a copy of the encoder with the SAME gradients (autograd accumulates them into the flat views) --
    same parameters after five steps, the loss goes through the fused walk."""
    torch.manual_seed(0)
    enc_a = pkg.Resnet(pos_embed=False).cuda().train().to(memory_format=torch.channels_last)
    enc_b = copy.deepcopy(enc_a)
    model_a = pkg.CRW(enc_a, 0.07, False, need_A=False)
    opt_a = pkg.optim.FlatAdam(model_a.parameters(), lr=1e-3)
    opt_b = torch.optim.Adam(enc_b.parameters(), lr=1e-3)
    for p in enc_a.parameters():
        assert p.data.untyped_storage().data_ptr() == opt_a.flat_p.untyped_storage().data_ptr()
        assert p.grad.untyped_storage().data_ptr() == opt_a.flat_g.untyped_storage().data_ptr()
        if p.dim() == 4 and min(p.shape[1:]) > 1:
            assert p.is_contiguous(memory_format=torch.channels_last)
    losses = []
    for it in range(5):
        seq = torch.randn(2, 4, 12, 32, 32, device="cuda")
        opt_a.zero_grad()
        la, _ = model_a(seq)
        la.backward()
        assert float(opt_a.flat_g.abs().max()) > 0.0
        for pa, pb in zip(enc_a.parameters(), enc_b.parameters()):
            pb.grad = pa.grad.detach().clone()
        opt_a.step()
        opt_b.step()
        losses.append(la.item())
        for (na, pa), (nb, pb) in zip(enc_a.named_parameters(), enc_b.named_parameters()):
            assert _rel(pa.detach().cpu().numpy(), pb.detach().cpu().numpy()) < 2e-6, (it, na)
    assert all(np.isfinite(losses))
    assert opt_a.steps == 5
    sd = opt_a.state_dict()
    opt_a.load_state_dict(sd)
    assert opt_a.steps == 5


def test_flat_adam_zero_grad_keeps_the_views(pkg):
    lin = torch.nn.Linear(16, 8).cuda()
    opt = pkg.optim.FlatAdam(lin.parameters(), lr=1e-2)
    w0 = lin.weight.detach().clone()
    for _ in range(3):
        opt.zero_grad()
        lin(torch.randn(4, 16, device="cuda")).square().mean().backward()
        assert lin.weight.grad.untyped_storage().data_ptr() == opt.flat_g.untyped_storage().data_ptr()
        opt.step()
    assert not torch.equal(w0, lin.weight.detach())
