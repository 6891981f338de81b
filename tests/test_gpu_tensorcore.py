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
