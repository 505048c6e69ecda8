"""GPU: the TMA -> fp32->16-bit split -> tcgen05.mma -> TMEM -> registers pipeline, as a plain GEMM."""
import pytest
import torch

import cases

pytestmark = pytest.mark.gpu

TOL = {"bf16x3": 2e-5, "fp16x3": 1e-5, "fp16": 3e-3, "bf16": 2e-2}


@pytest.fixture(scope="module")
def K():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import mhimk
    return mhimk.ops


@pytest.mark.parametrize("prec", ["fp16x3", "bf16x3", "fp16", "bf16"])
@pytest.mark.parametrize("M,N,Kd", [(128, 64, 32), (128, 128, 64), (128, 256, 128), (128, 512, 1024), (300, 512, 256), (1000, 128, 512), (20000, 512, 1024)])
def test_umma_gemm(K, prec, M, N, Kd):
    g = torch.Generator().manual_seed(M + N + Kd)
    A, B = torch.randn(M, Kd, generator=g), torch.randn(N, Kd, generator=g) * 0.05
    C = K.umma_selftest(A.cuda(), B.cuda(), prec)
    torch.cuda.synchronize()
    ref = A.double() @ B.double().t()
    e = cases.rel_err(C, ref)
    print(f"umma {prec} M={M} N={N} K={Kd}: {e:.2e}")
    assert e < TOL[prec]
