"""GPU tests of the tcgen05 / TMA / TMEM plumbing and the tensor-core kernels."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pkg():
    import radar_sounder_crw_b200 as p
    return p


@pytest.mark.parametrize("BN", [16, 64, 208, 256])
def test_umma_selftest_matches_torch(pkg, BN):
    """Pins the UMMA smem/instruction descriptor encodings + TMA swizzle + TMEM lane/column mapping."""
    torch.manual_seed(BN)
    A = torch.randn(128, 128, device="cuda").bfloat16()
    B = torch.randn(BN, 128, device="cuda").bfloat16()
    out = torch.full((128, BN), float("nan"), device="cuda")
    L = pkg._lib.lib()
    pkg._lib.check(L.crw_debug_umma_gemm(A.data_ptr(), B.data_ptr(), BN, out.data_ptr(),
                                         torch.cuda.current_stream().cuda_stream), "crw_debug_umma_gemm")
    torch.cuda.synchronize()
    ref = A.float() @ B.float().t()
    err = (out - ref).abs().max().item()
    assert err < 1e-3, err   # bf16 products are exact in fp32; only the accumulation order differs


@pytest.mark.parametrize("BN,a_mn,b_mn", [(64, 0, 0), (128, 0, 1), (64, 1, 0), (128, 1, 1)])
def test_umma_mn_major_selftest_matches_torch(pkg, BN, a_mn, b_mn):
    """MN-major (transposed) operands straight from a row-major [K][MN] matrix: pins LBO / SBO / the major bits."""
    torch.manual_seed(7 * BN + a_mn + 2 * b_mn)
    A = torch.randn(128, 64, device="cuda").bfloat16()          # logical A [M=128][K=64]
    B = torch.randn(64, BN, device="cuda").bfloat16()           # logical B [K=64][N=BN]
    A_in = A.t().contiguous() if a_mn else A                    # At [64][128]  vs  A [128][64]
    B_in = B if b_mn else B.t().contiguous()                    # Bkn [64][BN]  vs  Bt [BN][64]
    out = torch.full((128, BN), float("nan"), device="cuda")
    L = pkg._lib.lib()
    pkg._lib.check(L.crw_debug_umma_mn_gemm(A_in.data_ptr(), B_in.data_ptr(), BN, a_mn, b_mn, out.data_ptr(),
                                            torch.cuda.current_stream().cuda_stream), "crw_debug_umma_mn_gemm")
    torch.cuda.synchronize()
    err = (out - A.float() @ B.float()).abs().max().item()
    assert err < 1e-3, err


@pytest.mark.parametrize("BN", [16, 64, 256])
def test_umma_ts_selftest_matches_torch(pkg, BN):
    """A operand parked in TMEM with tcgen05.st and read by the TS form of tcgen05.mma: pins the A-in-TMEM layout."""
    torch.manual_seed(100 + BN)
    A = torch.randn(128, 128, device="cuda").bfloat16()
    B = torch.randn(BN, 128, device="cuda").bfloat16()
    out = torch.full((128, BN), float("nan"), device="cuda")
    L = pkg._lib.lib()
    pkg._lib.check(L.crw_debug_umma_ts_gemm(A.data_ptr(), B.data_ptr(), BN, out.data_ptr(),
                                            torch.cuda.current_stream().cuda_stream), "crw_debug_umma_ts_gemm")
    torch.cuda.synchronize()
    err = (out - A.float() @ B.float().t()).abs().max().item()
    assert err < 1e-3, err


@pytest.mark.parametrize("BN", [32, 64, 256])
def test_umma_tscp_selftest_matches_torch(pkg, BN):
    """A operand: TMA -> shared memory -> tcgen05.cp -> TMEM -> TS MMAs: pins the smem -> TMEM copy."""
    torch.manual_seed(300 + BN)
    A = torch.randn(128, 128, device="cuda").bfloat16()
    B = torch.randn(BN, 128, device="cuda").bfloat16()
    out = torch.full((128, BN), float("nan"), device="cuda")
    L = pkg._lib.lib()
    pkg._lib.check(L.crw_debug_umma_tscp_gemm(A.data_ptr(), B.data_ptr(), BN, out.data_ptr(),
                                              torch.cuda.current_stream().cuda_stream), "crw_debug_umma_tscp_gemm")
    torch.cuda.synchronize()
    err = (out - A.float() @ B.float().t()).abs().max().item()
    assert err < 1e-3, err


@pytest.mark.parametrize("BN", [32, 64, 128, 256])
def test_umma_pair_selftest_matches_torch(pkg, BN):
    """cta_group::2 on a 2-CTA cluster: M = 256 MMA, leader-credited TMA, multicast commit."""
    torch.manual_seed(200 + BN)
    A = torch.randn(256, 128, device="cuda").bfloat16()
    B = torch.randn(BN, 128, device="cuda").bfloat16()
    out = torch.full((256, BN), float("nan"), device="cuda")
    L = pkg._lib.lib()
    pkg._lib.check(L.crw_debug_umma_pair_gemm(A.data_ptr(), B.data_ptr(), BN, out.data_ptr(),
                                              torch.cuda.current_stream().cuda_stream), "crw_debug_umma_pair_gemm")
    torch.cuda.synchronize()
    err = (out - A.float() @ B.float().t()).abs().max().item()
    assert err < 1e-3, err


# ----------------------------------------------------------------------------------------------
# tensor-core label propagation (bf16 hi/lo x3, tcgen05): the "bf16 path" of BASELINE.json
# bar: >= 99.9 % pixel agreement with the fp32 path; top-k sets identical except across near-ties
# ----------------------------------------------------------------------------------------------
from helpers import lp_case, topk_sets_equal  # noqa: E402
from oracle import c_oracle, labelprop_oracle as lo  # noqa: E402

TC_CASES = [
    # R, T, N, M, ctx, k, radius, clustered
    (1, 3, 49, 4, 20, 10, 12, False),       # one key tile
    (1, 60, 49, 4, 20, 10, 12, False),      # config-3 parameters, frame-0 tile appears (n > ctx+1)
    (2, 40, 47, 4, 20, 20, 24, True),       # config-5 parameters, clustered (near-collinear) features, 2 radargrams
    (1, 30, 49, 4, 5, 10, 12, False),       # small ctx: trim active almost everywhere
    (1, 20, 113, 5, 4, 10, 12, True),       # SHARAD node count (N > 64)
    (1, 400, 49, 4, 20, 10, 12, True),      # > 148 query tiles: persistent loop, barrier phase wrap
    (1, 12, 25, 3, 20, 20, 10, False),      # fewer than k in-band keys -> masked fill
    (3, 9, 16, 9, 2, 16, 5, False),         # tiny N, several frames per tile, KT=16
    (1, 8, 64, 2, 3, 32, 100, False),       # k = 32, radius >= N
]


def _dev(a, dtype=torch.float32):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype, device="cuda")


@pytest.fixture(params=["ts", "ss", "pair"])
def lp_kernel(request, monkeypatch):
    """The tensor-path kernels: single-CTA with the query tile in tensor memory ("TS" MMAs, the default), single-CTA with
    the query tile in shared memory ("SS" MMAs, env CRW_LP_TS=0) and the CTA-pair one (cta_group::2, env CRW_LP_PAIR=1)."""
    monkeypatch.setenv("CRW_LP_PAIR", "1" if request.param == "pair" else "0")
    monkeypatch.setenv("CRW_LP_TS", "0" if request.param == "ss" else "1")
    return request.param


@pytest.mark.parametrize("case", TC_CASES)
def test_lp_tensorcore_vs_fp32_oracle(pkg, case, lp_kernel):
    R, T, N, M, ctx, k, radius, clustered = case
    rs = np.random.RandomState(200 + TC_CASES.index(case))
    feats = rs.randn(R, T, N, 128).astype(np.float32)
    if clustered:
        feats += 3.0 * rs.randn(R, 1, 1, 128).astype(np.float32)
    label0 = rs.randint(0, M, (R, N)).astype(np.int32)
    mask0 = np.stack([lo.one_hot_mask(label0[r], M, np.float32) for r in range(R)])
    labels, masks, W, I = pkg.ops.labelprop(_dev(feats), _dev(mask0), ctx, float(radius), 0.07, k, 0,
                                            pkg.ops.PREC_BF16X3, True, True)
    torch.cuda.synchronize()
    labels, masks, W, I = labels.cpu().numpy(), masks.cpu().numpy(), W.cpu().numpy(), I.cpu().numpy()
    o = c_oracle.labelprop(feats, label0, M, ctx, radius, 0.07, k)
    agree = (labels == o["labels"]).mean()
    frac, same = topk_sets_equal(I[:, 1:], W[:, 1:], o["I"][:, 1:], o["W"][:, 1:])
    assert agree >= 0.999, f"label agreement {agree:.5f}"
    assert frac >= 0.995, f"top-k set agreement {frac:.5f}"
    # where the id lists coincide exactly, the weights agree to bf16x3 accuracy
    eq = (I[:, 1:] == o["I"][:, 1:]).all(2, keepdims=True)
    assert np.abs(np.where(eq, W[:, 1:] - o["W"][:, 1:], 0)).max() < 2e-4
    assert np.abs(W[:, 1:].sum(2) - 1).max() < 1e-5


def test_lp_tensorcore_fork_join_and_graph_capture(pkg, monkeypatch):
    """The early query tiles + sequential gather run on a forked stream: same result as the single-stream order, also when
    the call is captured into a CUDA graph and replayed on new inputs."""
    rs = np.random.RandomState(77)
    R, T, N, M = 2, 300, 49, 4
    feats = _dev(rs.randn(R, T, N, 128).astype(np.float32))
    mask0 = _dev(np.stack([lo.one_hot_mask(rs.randint(0, M, N), M, np.float32) for _ in range(R)]))

    def run(f):
        return pkg.ops.labelprop(f, mask0, 20, 12.0, 0.07, 10, 0, pkg.ops.PREC_BF16X3, True, True)

    forked = [t.clone() for t in run(feats)]
    monkeypatch.setenv("CRW_LP_NO_FORK", "1")
    plain = run(feats)
    monkeypatch.delenv("CRW_LP_NO_FORK")
    for a_, b_ in zip(forked, plain):
        assert torch.equal(a_, b_)
    static_in = feats.clone()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        run(static_in)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out = run(static_in)
    feats2 = _dev(rs.randn(R, T, N, 128).astype(np.float32))
    static_in.copy_(feats2)
    graph.replay()
    torch.cuda.synchronize()
    ref2 = run(feats2)
    for a_, b_ in zip(out, ref2):
        assert torch.equal(a_, b_)


@pytest.mark.parametrize("shape", [(1, 1250, 49), (3, 200, 47), (2, 1, 49), (1, 30, 49)])
def test_lp_host_streamed_equals_device_path(pkg, shape, lp_kernel):
    """Features in pinned host memory, chunked H2D overlapped with the top-k: bit-identical to the device-resident call."""
    R, T, N = shape
    rs = np.random.RandomState(T + N)
    feats = torch.from_numpy(rs.randn(R, T, N, 128).astype(np.float32)).pin_memory()
    M = 4
    mask0 = _dev(np.stack([lo.one_hot_mask(rs.randint(0, M, N), M, np.float32) for _ in range(R)]))
    a_ = pkg.ops.labelprop_host(feats, mask0, 20, 12.0, 0.07, 10, 0, True, True)
    b_ = pkg.ops.labelprop(feats.cuda(), mask0, 20, 12.0, 0.07, 10, 0, pkg.ops.PREC_BF16X3, True, True)
    torch.cuda.synchronize()
    for x, y in zip(a_, b_):
        assert torch.equal(x, y)
    with pytest.raises(RuntimeError):
        pkg.ops.labelprop_host(feats.cuda(), mask0, 20, 12.0, 0.07, 10, 0, True, True)
    with pytest.raises(RuntimeError):
        pkg.ops.labelprop_host(torch.zeros(1, 4, N, 128), mask0[:1], 20, 12.0, 0.07, 10, 0, True, True)     # pageable


@pytest.mark.parametrize("host", [False, True])
def test_lp_tensorcore_fixed_mode(pkg, host):
    """mode = LP_FIXED (gather without the context-trim quirk: every frame is sequential, no fork) on the tensor path."""
    rs = np.random.RandomState(5)
    R, T, N, M, ctx, k = 2, 90, 49, 4, 6, 10
    feats = rs.randn(R, T, N, 128).astype(np.float32)
    label0 = rs.randint(0, M, (R, N)).astype(np.int32)
    mask0 = _dev(np.stack([lo.one_hot_mask(label0[r], M, np.float32) for r in range(R)]))
    if host:
        labels, _, W, I = pkg.ops.labelprop_host(torch.from_numpy(feats).pin_memory(), mask0, ctx, 12.0, 0.07, k,
                                                 pkg.ops.LP_FIXED, True, True)
    else:
        labels, _, W, I = pkg.ops.labelprop(_dev(feats), mask0, ctx, 12.0, 0.07, k, pkg.ops.LP_FIXED, pkg.ops.PREC_BF16X3, True, True)
    o = c_oracle.labelprop(feats, label0, M, ctx, 12, 0.07, k, mode="fixed")
    assert (labels.cpu().numpy() == o["labels"]).mean() >= 0.999
    frac, _ = topk_sets_equal(I.cpu().numpy()[:, 1:], W.cpu().numpy()[:, 1:], o["I"][:, 1:], o["W"][:, 1:])
    assert frac >= 0.995


def test_lp_config5_per_gpu_size_properties(pkg):
    """BASELINE config 5 as one GPU of eight sees it (8 radargrams x 400 x 50 000 columns -> R=8, T=3125, N=49, k=20, r=24):
    size-independent properties of the result, radargram independence, the host-streamed path, and a window against the oracle."""
    R, T, N, M, ctx, k, radius = 8, 3125, 49, 4, 20, 20, 24.0
    g = torch.Generator().manual_seed(5)
    feats_host = torch.randn(R, T, N, 128, generator=g).pin_memory()
    label0 = torch.randint(0, M, (R, N), generator=g)
    mask0 = torch.nn.functional.one_hot(label0, M).permute(0, 2, 1).float().contiguous().cuda()
    feats = feats_host.cuda()
    labels, masks, W, I = pkg.ops.labelprop(feats, mask0, ctx, radius, 0.07, k, 0, pkg.ops.PREC_BF16X3, True, True)
    torch.cuda.synchronize()
    W1, I1 = W[:, 1:], I[:, 1:]
    assert torch.all(W1 >= 0) and (W1.sum(2) - 1).abs().max().item() < 1e-5            # softmax over the k neighbours
    assert torch.all(W1[:, :, :-1] >= W1[:, :, 1:])                                     # sorted by descending weight
    n = torch.arange(1, T, device="cuda").view(1, -1, 1, 1)
    n_key = torch.clamp(n, max=ctx + 1)                                                 # frames in the (trimmed) key set
    assert torch.all(I1 >= 0) and torch.all(I1 < n_key * N)
    q = torch.arange(N, device="cuda").view(1, 1, 1, -1)
    assert torch.all(((I1 % N) - q).abs() <= 23)                                        # inside the radius band (ceil(r) - 1)
    srt = torch.sort(I1, dim=2).values
    assert torch.all(srt[:, :, 1:] != srt[:, :, :-1])                                   # k distinct candidates
    assert torch.all((labels >= 0) & (labels < M)) and torch.equal(labels[:, 0].cpu(), label0.int())
    assert (masks.sum(2) - 1).abs().max().item() < 1e-4                                 # soft masks stay distributions
    one = pkg.ops.labelprop(feats[3:4].contiguous(), mask0[3:4].contiguous(), ctx, radius, 0.07, k, 0, pkg.ops.PREC_BF16X3, True, True)
    for a_, b_ in zip(one, (labels[3:4], masks[3:4], W[3:4], I[3:4])):                  # radargrams are independent
        assert torch.equal(a_, b_)
    host = pkg.ops.labelprop_host(feats_host, mask0, ctx, radius, 0.07, k, 0, True, True)   # 32 chunks through the staging buffers
    torch.cuda.synchronize()
    for a_, b_ in zip(host, (labels, masks, W, I)):
        assert torch.equal(a_, b_)
    o = c_oracle.labelprop(feats_host[5:6, :64].numpy(), label0[5:6].numpy().astype(np.int32), M, ctx, radius, 0.07, k)
    assert (labels[5, :64].cpu().numpy() == o["labels"][0]).mean() >= 0.999


def test_lp_tensorcore_golden_reference_labels(pkg, lp_kernel):
    """bf16x3 path against the LIVE reference's outputs: >= 99.9 % of pixels."""
    for name in ["lp_quirk.npz", "lp_cfg3_short.npz", "lp_clustered.npz", "lp_cfg5_short.npz"]:
        g = lp_case(name)
        mask0 = lo.one_hot_mask(g["label0"], g["M"], np.float32)[None]
        labels, _, _, _ = pkg.ops.labelprop(_dev(g["feats"][None]), _dev(mask0), g["ctx"], g["radius"], g["temp"], g["k"], 0,
                                            pkg.ops.PREC_BF16X3, True, False)
        agree = (labels[0].t().cpu().numpy() == g["labels"]).mean()
        assert agree >= 0.999, (name, agree)


def test_lp_tensorcore_config3_full_size(pkg):
    T, N, M = 1250, 49, 4
    rs = np.random.RandomState(11)
    feats = (rs.randn(1, T, N, 128) + 2.0 * rs.randn(1, 1, 1, 128)).astype(np.float32)
    label0 = rs.randint(0, M, (1, N)).astype(np.int32)
    mask0 = lo.one_hot_mask(label0[0], M, np.float32)[None]
    labels, masks, W, I = pkg.ops.labelprop(_dev(feats), _dev(mask0), 20, 12.0, 0.07, 10, 0, pkg.ops.PREC_BF16X3, True, True)
    o = c_oracle.labelprop(feats, label0, M, 20, 12, 0.07, 10)
    assert (labels.cpu().numpy() == o["labels"]).mean() >= 0.999
    frac, _ = topk_sets_equal(I.cpu().numpy()[:, 1:], W.cpu().numpy()[:, 1:], o["I"][:, 1:], o["W"][:, 1:])
    assert frac >= 0.995


@pytest.mark.parametrize("shape", [(1, 400, 49, 10, 12.0), (1, 1250, 49, 10, 12.0), (2, 600, 47, 20, 24.0)])
def test_lp_tensorcore_variants_are_bit_identical(pkg, shape, monkeypatch):
    """Every form of the tensor-path top-k computes each dot product with the same MMAs and orders ties the same way, so
    the query tile in TMEM or shared memory, the tail split (partial lists merged through global memory) and the CTA-pair
    kernel must agree to the bit: weights, ids, soft masks, labels.  Shapes: 6 / 53 / 34 tiles in the partial round (the last one over two radargrams, k = 20)."""
    R, T, N, k, radius = shape
    rs = np.random.RandomState(T + k)
    feats = _dev((rs.randn(R, T, N, 128) + 1.5 * rs.randn(R, 1, 1, 128)).astype(np.float32))
    M = 4
    mask0 = _dev(np.stack([lo.one_hot_mask(rs.randint(0, M, N), M, np.float32) for _ in range(R)]))
    outs = {}
    for name, env in [("default", {}), ("no_split", {"CRW_LP_NO_SPLIT": "1"}), ("ss", {"CRW_LP_TS": "0"}),
                      ("ss_no_split", {"CRW_LP_TS": "0", "CRW_LP_NO_SPLIT": "1"}), ("pair", {"CRW_LP_PAIR": "1"}),
                      ("no_fork", {"CRW_LP_NO_FORK": "1"})]:
        for key in ("CRW_LP_NO_SPLIT", "CRW_LP_TS", "CRW_LP_PAIR", "CRW_LP_NO_FORK"):
            monkeypatch.delenv(key, raising=False)
        for key, val in env.items():
            monkeypatch.setenv(key, val)
        outs[name] = [x.clone() for x in pkg.ops.labelprop(feats, mask0, 20, radius, 0.07, k, 0, pkg.ops.PREC_BF16X3, True, True)]
        torch.cuda.synchronize()
    for name, o in outs.items():
        for a, b in zip(outs["default"], o):
            assert torch.equal(a[:, 1:], b[:, 1:]), name


# ----------------------------------------------------------------------------------------------
# tensor-core walk (tcgen05 bf16x3 GEMMs, fp32 accumulate in TMEM): loss / gradients within 1e-3 relative
# ----------------------------------------------------------------------------------------------
from helpers import load_golden, rel_err  # noqa: E402
from oracle import walk_oracle as wo  # noqa: E402

WALK_TC_CASES = [
    # B, T, N, C, tau
    (4, 10, 47, 128, 0.07),    # config 2 geometry on the tensor path
    (2, 20, 47, 128, 0.07),    # config 4 geometry
    (2, 6, 100, 128, 0.07),    # one 128-tile, K = 100 (two k-chunks)
    (1, 5, 185, 128, 0.07),    # scaled geometry: 2 x 2 output tiles, three k-chunks
    (2, 4, 12, 32, 0.07),      # K = 32 (half a chunk), K = 12
    (2, 3, 9, 16, 0.05),       # T = 3
    (2, 7, 24, 128, 0.01),     # the reference's train default tau
    (1, 4, 300, 128, 0.07),    # 3 x 3 output tiles, five k-chunks with a 44-wide tail, a third tile row of 44 rows
    (3, 4, 130, 96, 0.07),     # tiles that are almost empty (2 rows / 2 columns), C = 96: E boxes zero-filled past the channels
]


@pytest.fixture(params=["auto", "tiles"])
def walk_engine(request, monkeypatch):
    """BF16X3 engines: auto = warp-level MMA shared-memory kernels for one-tile sizes (N <= 64), tcgen05 tiles beyond;
    tiles = the tcgen05 tile path at every size."""
    if request.param == "tiles":
        monkeypatch.setenv("CRW_WALK_FORCE_TILES", "1")
    else:
        monkeypatch.delenv("CRW_WALK_FORCE_TILES", raising=False)
    return request.param


@pytest.mark.parametrize("case", WALK_TC_CASES)
def test_walk_tensorcore_vs_f64_oracle(pkg, case, walk_engine):
    B, T, N, C, tau = case
    rs = np.random.RandomState(B * 1000 + T * 100 + N)
    x = (rs.randn(B, T, N, C) + 1.0 * rs.randn(B, 1, 1, C)).astype(np.float32)
    xt = _dev(x).requires_grad_(True)
    loss, A, _ = pkg.ops.walk_loss(xt, float(tau), True, pkg.ops.PREC_BF16X3)
    loss.backward()
    torch.cuda.synchronize()
    l64, _, _, dx64 = wo.walk_backward_chain(x.astype(np.float64), tau)
    assert abs(loss.item() - l64) <= 1e-4 * abs(l64), (loss.item(), l64)
    assert rel_err(A.detach().cpu().numpy(), wo.affinities(wo.l2_normalize(x.astype(np.float64)), tau)) < 1e-4
    err = rel_err(xt.grad.cpu().numpy(), dx64)
    assert err < 1e-3, err


@pytest.mark.parametrize("N,T", [(21, 3), (100, 4), (150, 5)])
def test_walk_tile_engine_grad_through_returned_A(pkg, monkeypatch, N, T):
    """The reference returns A as a differentiable tensor (model.py:46): on the tile engine an incoming dA joins the softmax
    backward in t_dA_rows_kernel and reaches x through both dE products, also through the last affinity (which the loss never sees)."""
    monkeypatch.setenv("CRW_WALK_FORCE_TILES", "1")
    rs = np.random.RandomState(N + T)
    B = 2
    x = rs.randn(B, T, N, 128).astype(np.float32)
    Gext = rs.randn(B, T - 1, N, N).astype(np.float32)
    xt = _dev(x).requires_grad_(True)
    loss, A, _ = pkg.ops.walk_loss(xt, 0.07, True, pkg.ops.PREC_BF16X3)
    (loss + (A * _dev(Gext)).sum()).backward()
    xr = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    E = torch.nn.functional.normalize(xr, dim=-1)
    Ar = torch.einsum("btnc,btmc->btnm", E[:, :-1], E[:, 1:]) / 0.07
    (Ar * torch.tensor(Gext, dtype=torch.float64)).sum().backward()
    _, _, _, dx_loss = wo.walk_backward_chain(x.astype(np.float64), 0.07)
    assert rel_err(xt.grad.cpu().numpy(), xr.grad.numpy() + dx_loss) < 1e-3


@pytest.mark.parametrize("shape", [(8, 20, 185), (24, 6, 47), (2, 8, 369)])
def test_walk_tile_engine_many_tiles_per_cta_matches_fp32_engine(pkg, monkeypatch, shape):
    """More tiles than SMs: every persistent CTA walks several tiles, so the two TMEM accumulators, the stage ring and their
    barrier phases wrap around many times (the oracle cases above give each CTA at most one tile).  Checked against the fp32
    FMA engine (itself pinned to the fp64 oracle) at sizes the numpy oracle would take minutes for."""
    B, T, N = shape
    monkeypatch.setenv("CRW_WALK_FORCE_TILES", "1")
    torch.manual_seed(B + T + N)
    x = torch.randn(B, T, N, 128, device="cuda") + torch.randn(B, 1, 1, 128, device="cuda")
    Gext = torch.randn(B, T - 1, N, N, device="cuda") * 0.01
    out = {}
    for name, prec in [("tiles", pkg.ops.PREC_BF16X3), ("fp32", pkg.ops.PREC_FP32)]:
        xt = x.clone().requires_grad_(True)
        loss, A, _ = pkg.ops.walk_loss(xt, 0.07, True, prec)
        (loss + (A * Gext).sum()).backward()
        out[name] = (loss.item(), A.detach().cpu().numpy(), xt.grad.cpu().numpy())
    assert abs(out["tiles"][0] - out["fp32"][0]) <= 1e-4 * abs(out["fp32"][0])
    assert rel_err(out["tiles"][1], out["fp32"][1]) < 1e-4
    assert rel_err(out["tiles"][2], out["fp32"][2]) < 1e-3


def test_walk_tile_engine_is_deterministic(pkg, monkeypatch):
    """Every reduction of the tile engine has a fixed order (per-block loss partials, column statistics): two runs agree bit for bit."""
    monkeypatch.setenv("CRW_WALK_FORCE_TILES", "1")
    x = torch.randn(2, 5, 140, 128, device="cuda")
    outs = []
    for _ in range(2):
        xt = x.clone().requires_grad_(True)
        loss, A, _ = pkg.ops.walk_loss(xt, 0.07, True, pkg.ops.PREC_BF16X3)
        loss.backward()
        outs.append((loss.detach().clone(), A.detach().clone(), xt.grad.clone()))
    for a, b in zip(*outs):
        assert torch.equal(a, b)


def test_walk_tensorcore_golden_reference(pkg, walk_engine):
    for name in ["walk_cfg1_f32.npz", "walk_tau001_f32.npz"]:
        g = load_golden(name)
        xt = _dev(g["x"].astype(np.float32)).requires_grad_(True)
        loss, _, _ = pkg.ops.walk_loss(xt, float(g["tau"]), False, pkg.ops.PREC_BF16X3)
        loss.backward()
        assert abs(loss.item() - float(g["loss"])) <= 1e-3 * abs(float(g["loss"]))
        _, _, _, dx64 = wo.walk_backward_chain(g["x"].astype(np.float64), float(g["tau"]))
        assert rel_err(xt.grad.cpu().numpy(), dx64) < 1e-3
