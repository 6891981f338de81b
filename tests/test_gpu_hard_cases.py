"""Hard-case precision tests: ENCODER-DERIVED features.  A random-init Resnet maps any radargram to nearly
collinear embeddings (cos-sim 0.86 .. 0.97 between any two nodes), so top-k near-ties are frequent and operand precision
matters: plain bf16 operands reach only 92.6 .. 99.3 % label agreement there.  Checked here, through the whole drop-in path
(radargram -> RGDataset cut -> Resnet(eval) -> propagate kernels):

  * exact tensor path (PREC_TC_EXACT): ids, weights, masks, labels IDENTICAL to the fp32 path, no exemptions;
  * error-compensated bf16 path (PREC_BF16X3, the round-1 kernel, now opt-in): measured 99.86 .. 100 % of the pixels at (k=10,
    r=12) and (k=20, r=24) -- ONE of the four cases falls short of BASELINE's 99.9 % bar for "the bf16 path" (with near-uniform
    top-k weights the class sums are nearly tied and 1e-6 weight errors flip the argmax), which is why the default tensor path
    is the exact one; asserted here at >= 99.8 % and printed, together with the number of queries whose neighbouring top-k
    weights are within 1e-6 (the near ties that may legitimately flip);
  * BF16X3 walk (tcgen05 tiles at N > 64, warp-level MMAs at N = 47): gradient within 1e-3 relative of fp64 at tau = 0.01,
    the reference's training default (scripts/train.py:31), which amplifies operand error 7x more than tau = 0.07.
"""
import numpy as np
import pytest
import torch

from oracle import walk_oracle as wo

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pkg():
    import radar_sounder_crw_b200 as p
    return p


def _radargram(kind, cols, seed):
    g = torch.Generator().manual_seed(seed)
    if kind == "white_noise":
        return torch.randn(400, cols, generator=g)
    # layered: two Gaussian-profile horizons + sinusoidal in-ice texture + 0.3 randn (SURVEY 8d synthetic input ii)
    y = torch.arange(400.0)[:, None]
    x = torch.arange(float(cols))[None, :]
    surf = 90 + 12 * torch.sin(x / 230.0)
    bed = 290 + 25 * torch.sin(x / 410.0 + 1.0)
    rg = 2.5 * torch.exp(-0.5 * ((y - surf) / 3.0) ** 2) + 1.8 * torch.exp(-0.5 * ((y - bed) / 5.0) ** 2)
    rg = rg + 0.25 * torch.sin(y / 9.0 + x / 57.0) * ((y > surf) & (y < bed))
    return rg + 0.3 * torch.randn(400, cols, generator=g)


def _features(pkg, kind, T, patch, overlap, seed=11):
    """Encoder-derived, un-normalised features [1,T,N,128] of a synthetic radargram (cut exactly as src/dataset.py:34-39)."""
    (h, w), (oh, ow) = patch, overlap
    rg = _radargram(kind, T * (w - ow) + ow, seed).cuda()
    N = (400 - oh) // (h - oh)
    item = pkg.ops.patch_unfold(rg.contiguous(), 0, 0, 1, T, h, w, oh, ow, False)           # [1,T,N,h,w]
    torch.manual_seed(seed)
    enc = pkg.Resnet(pos_embed=False).cuda().eval()
    with torch.no_grad():
        emb = enc(item.reshape(-1, 1, h, w)).reshape(1, T, N, 128).float().contiguous()
    return emb, N


@pytest.mark.parametrize("kind", ["white_noise", "layered"])
@pytest.mark.parametrize("k,radius", [(10, 12.0), (20, 24.0)])
def test_lp_precision_on_encoder_features(pkg, kind, k, radius):
    T, M = 160, 4
    emb, N = _features(pkg, kind, T, (16, 16), (8, 0))
    en = torch.nn.functional.normalize(emb, dim=-1)
    cos = (en[0, 5] @ en[0, 40].T)
    assert float(cos.min()) > 0.5, "the fixture is meant to be near-collinear"
    g = torch.Generator(device="cuda").manual_seed(3)
    label0 = torch.randint(0, M, (1, N), device="cuda", generator=g)
    mask0 = torch.nn.functional.one_hot(label0, M).permute(0, 2, 1).float().contiguous()
    run = lambda prec: pkg.ops.labelprop(emb, mask0, 20, radius, 0.07, k, 0, prec, True, True)   # noqa: E731
    l32, m32, W32, I32 = run(pkg.ops.PREC_FP32)
    lx, mx, Wx, Ix = run(pkg.ops.PREC_TC_EXACT)
    assert torch.equal(Ix[:, 1:], I32[:, 1:]) and torch.equal(Wx[:, 1:], W32[:, 1:])
    assert torch.equal(mx, m32) and torch.equal(lx, l32)
    lb, _, Wb, Ib = run(pkg.ops.PREC_BF16X3)
    agree = float((lb == l32).float().mean())
    same_sets = float((Ib[:, 1:].sort(dim=2).values == I32[:, 1:].sort(dim=2).values).all(dim=2).float().mean())
    # near ties: queries whose smallest gap between neighbouring top-k weights is below 1e-6 relative (they may flip order)
    gaps = (W32[:, 1:, :-1] - W32[:, 1:, 1:]).abs() / W32[:, 1:, :-1].clamp_min(1e-30)
    near = int((gaps.min(dim=2).values < 1e-6).sum())
    print(f"[{kind} k={k} r={radius}] cos-sim {float(cos.min()):.3f}..{float(cos.max()):.3f}; bf16x3 label agreement {agree:.5f}, "
          f"identical top-k sets {same_sets:.5f}, queries with a near tie (< 1e-6) {near} of {W32[:, 1:].shape[1] * N}")
    assert agree >= 0.998


@pytest.mark.parametrize("N_geom", [((32, 32), (24, 0)), ((8, 8), (4, 0))])     # N = 47 (one-tile kernels), N = 99 (tcgen05 tiles)
def test_walk_bf16x3_gradient_at_tau_001_on_encoder_features(pkg, N_geom):
    patch, overlap = N_geom
    B, T = 2, 8
    embs = [_features(pkg, "layered", T, patch, overlap, seed=20 + b)[0] for b in range(B)]
    x = torch.cat(embs).detach().clone().requires_grad_(True)
    loss, _, _ = pkg.ops.walk_loss(x, 0.01, False, pkg.ops.PREC_BF16X3)
    loss.backward()
    l64, _, _, dx64 = wo.walk_backward_chain(x.detach().cpu().numpy().astype(np.float64), 0.01)
    rel = np.abs(x.grad.cpu().numpy() - dx64).max() / np.abs(dx64).max()
    print(f"[walk bf16x3 tau=0.01 N={x.shape[2]}] loss {loss.item():.6f} vs {l64:.6f}, gradient rel err {rel:.2e}")
    assert abs(loss.item() - l64) <= 1e-3 * abs(l64)
    assert rel <= 1e-3


def test_walk_loss_fake_kernel_matches_the_real_op(pkg):
    """torch.library.opcheck: output metadata of the fake (meta) implementation, incl. the size of the saved workspace."""
    x = torch.randn(2, 5, 12, 128, device="cuda", requires_grad=True)
    torch.library.opcheck(torch.ops.crw_b200.walk_loss.default, (x, 0.07, True, pkg.ops.PREC_FP32),
                          test_utils=("test_schema", "test_faketensor"))
