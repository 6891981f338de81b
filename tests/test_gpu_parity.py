"""GPU parity tests: the CUDA path (through the torch custom ops -> C ABI) against the oracle.

Run on the B200 box:  python -m pytest tests -m gpu -x -q

Bars (BASELINE.json north_star):
  * label propagation, fp32 path: top-k ids, weights, soft masks and labels BIT-EXACT against the
    pinned-order C oracle (oracle/crw_oracle.c), and labels identical to the live reference's
    outputs stored in tests/golden/lp_*.npz;
  * walk: loss and gradients within 1e-3 relative of the fp64 oracle / the reference's autograd
    (fp32 path measures ~1e-6).
"""
import numpy as np
import pytest
import torch

from helpers import load_golden, lp_case, order_mismatches_are_ties, rel_err, topk_sets_equal
from oracle import c_oracle, labelprop_oracle as lo, walk_oracle as wo

pytestmark = pytest.mark.gpu

WALK_TOL = 1e-3   # relative, stated by BASELINE.json


@pytest.fixture(scope="module")
def pkg():
    import radar_sounder_crw_b200 as p
    return p


def _dev(a, dtype=torch.float32):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype, device="cuda")


# ----------------------------------------------------------------------------------------------
# normalise
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("rows,C", [(1, 128), (49, 128), (1000, 128), (77, 16), (5, 100)])
def test_l2_normalize_bit_exact(pkg, rows, C):
    rs = np.random.RandomState(rows + C)
    x = rs.randn(rows, C).astype(np.float32)
    x[0] = 0.0  # eps branch of F.normalize
    got = pkg.ops.l2_normalize(_dev(x)).cpu().numpy()
    assert np.array_equal(got, c_oracle.l2_normalize(x))
    assert np.abs(got - wo.l2_normalize(x.astype(np.float64))).max() < 1e-6


# ----------------------------------------------------------------------------------------------
# label propagation
# ----------------------------------------------------------------------------------------------
def _run_lp(pkg, feats, label0, M, ctx, radius, temp, k, mode="ref_exact", normalize=True):
    R = feats.shape[0]
    mask0 = np.stack([lo.one_hot_mask(label0[r], M, np.float32) for r in range(R)])
    labels, masks, W, I = pkg.ops.labelprop(_dev(feats), _dev(mask0), ctx, float(radius), float(temp), k,
                                            1 if mode == "fixed" else 0, pkg.ops.PREC_FP32, normalize, True)
    torch.cuda.synchronize()
    return labels.cpu().numpy(), masks.cpu().numpy(), W.cpu().numpy(), I.cpu().numpy()


@pytest.mark.parametrize("name", ["lp_quirk.npz", "lp_cfg3_short.npz", "lp_masked_ties.npz", "lp_clustered.npz", "lp_cfg5_short.npz"])
def test_lp_golden_reference_outputs(pkg, name):
    """CUDA fp32 path vs the outputs of the LIVE reference (golden) and vs the C oracle (bit-exact)."""
    g = lp_case(name)
    labels, masks, W, I = _run_lp(pkg, g["feats"][None], g["label0"][None], g["M"], g["ctx"], g["radius"], g["temp"], g["k"])
    # reference: labels identical, top-k sets identical (tie-aware), weights/masks to fp32 noise
    assert np.array_equal(labels[0].T, g["labels"].astype(np.int32))
    frac, _ = topk_sets_equal(I[0, 1:], W[0, 1:], g["I"], g["W"])
    assert frac == 1.0
    _, ok = order_mismatches_are_ties(I[0, 1:], g["I"], g["W"])
    assert ok
    assert np.abs(W[0, 1:] - g["W"]).max() < 2e-6
    assert np.abs(masks[0, 1:] - g["masks"]).max() < 5e-6
    # C oracle: bit-exact everything
    o = c_oracle.labelprop(g["feats"][None], g["label0"][None], g["M"], g["ctx"], g["radius"], g["temp"], g["k"])
    assert np.array_equal(I[0, 1:], o["I"][0, 1:])
    assert np.array_equal(W[0, 1:], o["W"][0, 1:])
    assert np.array_equal(masks, o["masks"])
    assert np.array_equal(labels, o["labels"])


LP_CASES = [
    # R, T, N, C, M, ctx, k, radius, temp, mode
    (1, 30, 47, 128, 4, 20, 10, 12, 0.07, "ref_exact"),      # config-3 secondary geometry
    (2, 45, 49, 128, 4, 20, 20, 24, 0.07, "ref_exact"),      # config-5 parameters
    (1, 12, 113, 128, 5, 4, 10, 12, 0.07, "ref_exact"),      # SHARAD node count: two 64-node chunks
    (1, 9, 140, 64, 3, 3, 7, 70.5, 0.05, "fixed"),           # band wider than a chunk, fractional radius, 3 chunks
    (1, 40, 49, 128, 4, 6, 10, 12, 0.07, "fixed"),           # fixed mode, sequential gather
    (1, 2, 49, 128, 4, 20, 10, 12, 0.07, "ref_exact"),       # single step
    (1, 1, 49, 128, 4, 20, 10, 12, 0.07, "ref_exact"),       # no step at all
    (1, 6, 25, 16, 3, 20, 20, 10, 0.07, "ref_exact"),        # fewer than k in-band keys -> masked fill
    (1, 8, 33, 128, 2, 2, 32, 100, 0.1, "ref_exact"),        # k = 32 (lane limit), radius >= N
    (3, 7, 12, 32, 9, 1, 5, 3, 0.07, "ref_exact"),           # M > 8 (two class passes), ctx = 1
]


@pytest.mark.parametrize("case", LP_CASES)
def test_lp_bit_exact_vs_c_oracle(pkg, case):
    R, T, N, C, M, ctx, k, radius, temp, mode = case
    rs = np.random.RandomState(100 + LP_CASES.index(case))
    feats = (rs.randn(R, T, N, C) + 1.5 * rs.randn(R, 1, 1, C)).astype(np.float32)
    label0 = rs.randint(0, M, (R, N)).astype(np.int32)
    labels, masks, W, I = _run_lp(pkg, feats, label0, M, ctx, radius, temp, k, mode)
    o = c_oracle.labelprop(feats, label0, M, ctx, radius, temp, k, mode=mode)
    assert np.array_equal(I[:, 1:], o["I"][:, 1:])
    assert np.array_equal(W[:, 1:], o["W"][:, 1:])
    assert np.array_equal(masks, o["masks"])
    assert np.array_equal(labels, o["labels"])


def test_lp_config3_full_size(pkg):
    """BASELINE config 3 at full size: 400 x 20k columns -> T=1250, N=49, M=4, ctx=20, k=10, r=12."""
    T, N, C, M = 1250, 49, 128, 4
    rs = np.random.RandomState(11)
    feats = (rs.randn(1, T, N, C) + 2.0 * rs.randn(1, 1, 1, C)).astype(np.float32)
    label0 = rs.randint(0, M, (1, N)).astype(np.int32)
    labels, masks, W, I = _run_lp(pkg, feats, label0, M, 20, 12, 0.07, 10)
    o = c_oracle.labelprop(feats, label0, M, 20, 12, 0.07, 10)
    assert np.array_equal(labels, o["labels"])
    assert np.array_equal(I[:, 1:], o["I"][:, 1:])
    assert np.array_equal(W[:, 1:], o["W"][:, 1:])
    # size-independent properties: weights are a distribution, soft masks stay a distribution over classes
    assert np.abs(W[:, 1:].sum(2) - 1).max() < 1e-5
    assert np.abs(masks.sum(2) - 1).max() < 1e-4
    assert I[:, 1:].min() >= 0 and I[:, 1:].max() < 21 * N


def test_lp_radargrams_are_independent(pkg):
    """Sharding property: propagating R radargrams together == one at a time (no cross-talk)."""
    rs = np.random.RandomState(5)
    feats = rs.randn(3, 25, 49, 128).astype(np.float32)
    label0 = rs.randint(0, 4, (3, 49)).astype(np.int32)
    la, ma, _, _ = _run_lp(pkg, feats, label0, 4, 5, 12, 0.07, 10)
    for r in range(3):
        lb, mb, _, _ = _run_lp(pkg, feats[r:r + 1], label0[r:r + 1], 4, 5, 12, 0.07, 10)
        assert np.array_equal(la[r], lb[0]) and np.array_equal(ma[r], mb[0])


def test_stepwise_predict_matches_fused(pkg):
    """LabelPropVOS_CRW.predict driven frame by frame (the reference's loop, utils.py:152-160) == fused op."""
    g = lp_case("lp_quirk.npz")
    T, N, C = g["feats"].shape
    emb = pkg.ops.l2_normalize(_dev(g["feats"]))
    lp = pkg.LabelPropVOS_CRW({"CXT_SIZE": g["ctx"], "RADIUS": g["radius"], "TEMP": g["temp"], "KNN": g["k"]})
    mask0 = _dev(lo.one_hot_mask(g["label0"], g["M"], np.float32))
    feats, masks = [emb[0].t()[None, :, :, None]], [mask0[None, :, :, None]]
    pred = np.zeros((N, T), np.int64)
    pred[:, 0] = g["label0"]
    for n in range(1, T):
        cur = emb[n].t()[None, :, :, None]
        m = lp.predict(feats=feats, masks=masks, curr_feat=cur)
        assert m.shape == (1, g["M"], N, 1)
        feats.append(cur)
        masks.append(m)
        pred[:, n] = m.argmax(1).squeeze().cpu().numpy()
    assert np.array_equal(pred, g["labels"].astype(np.int64))


def test_batched_affinity_dropin(pkg):
    """batched_affinity with the reference's argument layout and a materialised 0/-1e10 bias tensor."""
    rs = np.random.RandomState(9)
    n, N, C, ctx, k, r, temp = 9, 49, 128, 4, 10, 12, 0.07
    emb = wo.l2_normalize(rs.randn(n + 1, N, C)).astype(np.float32)
    keys = _dev(emb[:n]).permute(2, 0, 1)[None, :, None]            # [1,C,1,n,hw]
    query = _dev(emb[n]).t()[None, :, None]                         # [1,C,1,hw]
    bias = _dev(lo.radius_bias(N, r, np.float32))[None, None]
    Ws, Is = pkg.batched_affinity(query, keys, bias, temp, k, [0], ctx, "cuda")
    assert Ws[0].shape == (k, N) and Is[0].dtype == torch.int64
    kf = lo.key_frames(n, ctx)
    W, I = lo.affinity_topk(emb[n].astype(np.float64), emb[kf].astype(np.float64), r, temp, k)
    frac, _ = topk_sets_equal(Is[0].cpu().numpy(), Ws[0].cpu().numpy(), I, W)
    assert frac == 1.0
    assert np.abs(Ws[0].cpu().numpy() - W).max() < 2e-6


def test_xent_matches_reference_golden(pkg):
    g = lp_case("lp_cfg3_short.npz")
    emb = pkg.ops.l2_normalize(_dev(g["feats"]))
    x = pkg.ops.horizontality_xent(emb).cpu().numpy()
    assert x.shape == g["xent"].shape
    assert np.abs(x - g["xent"]).max() < 1e-4


# ----------------------------------------------------------------------------------------------
# training walk
# ----------------------------------------------------------------------------------------------
def _walk_gpu(pkg, x, tau, need_A=True):
    xt = _dev(x).requires_grad_(True)
    loss, A, _ = pkg.ops.walk_loss(xt, float(tau), need_A, pkg.ops.PREC_FP32)
    if loss.requires_grad:
        loss.backward()
    g = xt.grad.cpu().numpy() if xt.grad is not None else np.zeros_like(x)
    return float(loss.item()), (A.detach().cpu().numpy() if need_A else None), g


@pytest.mark.parametrize("name", ["walk_small_f64.npz", "walk_t3_f64.npz", "walk_cfg1_f32.npz", "walk_tau001_f32.npz", "walk_cfg4_f32.npz"])
def test_walk_golden_reference_outputs(pkg, name):
    g = load_golden(name)
    x = g["x"].astype(np.float32)
    loss, A, dx = _walk_gpu(pkg, x, float(g["tau"]))
    assert abs(loss - float(g["loss"])) <= WALK_TOL * abs(float(g["loss"]))
    assert rel_err(A, g["A"]) < 1e-4
    # fp32 fixtures carry the reference's own fp32 rounding (tau=0.01 amplifies it); compare with the fp64 oracle too
    assert rel_err(dx, g["dx"]) < (WALK_TOL if str(g["dtype"]) == "float64" else 3e-3)
    l64, _, _, dx64 = wo.walk_backward_chain(g["x"].astype(np.float64), float(g["tau"]))
    assert abs(loss - l64) <= 1e-5 * abs(l64)
    assert rel_err(dx, dx64) < WALK_TOL


WALK_CASES = [
    # B, T, N, C, tau
    (4, 10, 47, 128, 0.07),    # config 1 / 2 geometry
    (2, 20, 47, 128, 0.07),    # config 4 geometry (T=20)
    (2, 6, 100, 128, 0.07),    # N > 64: several GEMM tiles
    (1, 5, 185, 128, 0.07),    # scaled geometry (32,30) -> N=185
    (3, 4, 12, 128, 0.07),     # K = 2: first chain step only
    (2, 3, 9, 32, 0.05),       # K = 1
    (2, 7, 24, 128, 0.01),     # the reference's train default tau
]


@pytest.mark.parametrize("case", WALK_CASES)
def test_walk_vs_f64_oracle(pkg, case):
    B, T, N, C, tau = case
    rs = np.random.RandomState(B * 1000 + T * 100 + N)
    x = (rs.randn(B, T, N, C) + 1.0 * rs.randn(B, 1, 1, C)).astype(np.float32)
    loss, A, dx = _walk_gpu(pkg, x, tau)
    l64, dA64, _, dx64 = wo.walk_backward_chain(x.astype(np.float64), tau)
    assert abs(loss - l64) <= 1e-5 * abs(l64)
    assert rel_err(A, wo.affinities(wo.l2_normalize(x.astype(np.float64)), tau)) < 1e-5
    assert rel_err(dx, dx64) < WALK_TOL, rel_err(dx, dx64)
    # reference-order loop gives the same loss (the chain form is only a re-association)
    if N <= 64:
        assert abs(loss - wo.walk_loss_reference_order(wo.affinities(wo.l2_normalize(x.astype(np.float64)), tau))) < 1e-5 * l64


def test_walk_t2_zero_loss_and_A(pkg):
    x = np.random.RandomState(2).randn(2, 2, 9, 16).astype(np.float32)
    xt = _dev(x)
    loss, A, _ = pkg.ops.walk_loss(xt, 0.07, True, pkg.ops.PREC_FP32)
    assert float(loss.item()) == 0.0
    assert rel_err(A.cpu().numpy(), wo.affinities(wo.l2_normalize(x.astype(np.float64)), 0.07)) < 1e-5


def test_walk_grad_through_returned_A(pkg):
    """The reference returns A as a differentiable tensor (model.py:46); dA flows back into x."""
    rs = np.random.RandomState(4)
    x = rs.randn(2, 5, 11, 32).astype(np.float32)
    Gext = rs.randn(2, 4, 11, 11).astype(np.float32)
    xt = _dev(x).requires_grad_(True)
    loss, A, _ = pkg.ops.walk_loss(xt, 0.07, True, pkg.ops.PREC_FP32)
    (loss + (A * _dev(Gext)).sum()).backward()
    xr = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    E = torch.nn.functional.normalize(xr, dim=-1)
    Ar = torch.einsum("btnc,btmc->btnm", E[:, :-1], E[:, 1:]) / 0.07
    (Ar * torch.tensor(Gext, dtype=torch.float64)).sum().backward()
    _, _, _, dx_loss = wo.walk_backward_chain(x.astype(np.float64), 0.07)
    assert rel_err(xt.grad.cpu().numpy(), xr.grad.numpy() + dx_loss) < WALK_TOL


def test_crw_module_dropin_with_encoder(pkg):
    """CRW(encoder, tau, pos_embed).forward(seq) -> (loss, A): gradients reach the encoder parameters and match a
    plain-torch restatement of model.py:22-46 (walk in fp64) driven by the same fp32 encoder and weights."""
    from oracle.walk_torch_port import crw_loss_reference_order
    torch.manual_seed(11)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.deterministic = True
    B, T, N, H, W = 2, 6, 12, 16, 16
    seq = torch.randn(B, T, N, H, W).cuda()
    enc_a = pkg.CNN(False).cuda()
    enc_b = pkg.CNN(False).cuda()
    enc_b.load_state_dict(enc_a.state_dict())
    loss, A = pkg.CRW(enc_a, 0.07, False)(seq)
    loss.backward()
    emb = enc_b(seq.reshape(-1, H, W).unsqueeze(1)).reshape(B, T, N, -1)
    ref_loss, ref_A = crw_loss_reference_order(emb.double(), 0.07)
    ref_loss.backward()
    assert abs(loss.item() - ref_loss.item()) < WALK_TOL * abs(ref_loss.item())
    assert rel_err(A.detach().cpu().numpy(), ref_A.detach().cpu().numpy()) < 1e-4
    for (n1, p1), (_, p2) in zip(enc_a.named_parameters(), enc_b.named_parameters()):
        assert rel_err(p1.grad.cpu().numpy(), p2.grad.cpu().numpy()) < WALK_TOL, n1


def test_propagate_dropin_end_to_end(pkg):
    """propagate(seq, seg_ref, encoder, lp, ...) with a real (random-init, eval) Resnet encoder."""
    torch.manual_seed(11)
    from radar_sounder_crw_b200.dataset import radargram_to_frames
    T, M = 24, 4
    rg = torch.randn(400, 16 * T)
    seq = radargram_to_frames(rg, 0, T, (16, 16), (8, 0))           # [T,49,16,16]
    N = seq.shape[1]
    seg_ref = torch.randint(0, M, (400, 16 * T))
    enc = pkg.Resnet(pos_embed=False).cuda().eval()
    lp = pkg.LabelPropVOS_CRW({"CXT_SIZE": 20, "RADIUS": 12, "TEMP": 0.07, "KNN": 10})
    pred, xent, _ = pkg.propagate(seq.cuda(), seg_ref, enc, lp, M, False, False)
    assert pred.shape == (N, T) and pred.is_cuda and xent.shape == (N, T - 1) and not xent.is_cuda
    with torch.no_grad():
        raw = enc(seq.cuda().reshape(-1, 16, 16).unsqueeze(1)).view(T, N, -1).float().cpu().numpy()
    label0 = lo.first_column_labels(seg_ref.numpy(), N).astype(np.int32)
    o = c_oracle.labelprop(raw[None], label0[None], M, 20, 12, 0.07, 10)
    assert np.array_equal(pred.cpu().numpy().astype(np.int32), o["labels"][0].T)
    assert np.abs(xent.numpy() - lo.horizontality_xent(c_oracle.l2_normalize(raw))).max() < 1e-3


def test_no_cpu_fallback(pkg):
    with pytest.raises(RuntimeError):
        pkg.ops.l2_normalize(torch.randn(4, 8))


def test_labels_upsample_matches_torchvision_nearest(pkg):
    """Post-processing row: Resize((seg_h, rg_len), NEAREST) of final_prediction[N,T] (test_all.py:79,96)."""
    from torchvision import transforms
    from torchvision.transforms import InterpolationMode
    for (T, N, H, W) in [(100, 49, 400, 1600), (37, 47, 410, 1184), (8, 113, 912, 128)]:
        labels = torch.randint(0, 5, (2, T, N), dtype=torch.int32)
        got = pkg.ops.labels_upsample(labels.cuda(), H, W).cpu()
        up = transforms.Resize((H, W), interpolation=InterpolationMode.NEAREST)
        for r in range(2):
            ref = up(labels[r].t().float()[None]).squeeze(0)       # final_prediction is [N,T]
            assert torch.equal(got[r], ref), (T, N, H, W)


def test_encoder_few_channel_batchnorm_is_equivalent(pkg):
    """The 3-channel bn0 of the encoder is evaluated with tensor reductions (cuDNN parallelises BN over channels);
    same parameters, running statistics and results as nn.BatchNorm2d."""
    import torch.nn as nn
    from radar_sounder_crw_b200.encoder import batchnorm_few_channels
    torch.manual_seed(3)
    x = torch.randn(200, 3, 18, 18, device="cuda", requires_grad=True)
    x2 = x.detach().clone().requires_grad_(True)
    bn_a, bn_b = nn.BatchNorm2d(3).cuda(), nn.BatchNorm2d(3).cuda()
    with torch.no_grad():
        bn_a.weight.uniform_(0.5, 1.5)
        bn_a.bias.uniform_(-1, 1)
    bn_b.load_state_dict(bn_a.state_dict())
    ya, yb = bn_a(x), batchnorm_few_channels(x2, bn_b)
    g = torch.randn_like(ya)
    ya.backward(g)
    yb.backward(g)
    assert torch.allclose(ya, yb, atol=1e-5) and torch.allclose(x.grad, x2.grad, atol=1e-5)
    assert torch.allclose(bn_a.weight.grad, bn_b.weight.grad, rtol=1e-4, atol=1e-4)
    assert torch.allclose(bn_a.running_var, bn_b.running_var, atol=1e-6) and bn_b.num_batches_tracked.item() == 1
    bn_a.eval(), bn_b.eval()
    assert torch.allclose(bn_a(x), batchnorm_few_channels(x2, bn_b), atol=1e-5)
