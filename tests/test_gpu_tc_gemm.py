"""tcgen05 / TMA GEMM (3xTF32) against a float64 matmul.  Stated tolerance: 1e-5 of max|C| -- two
orders of magnitude tighter than a plain tf32 GEMM would pass (5e-4), i.e. fp32-class accuracy."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M,N,K,lda,acc", [(15360, 384, 32, 32, 0), (15360, 192, 128, 128, 0), (64, 192, 128, 240 * 128, 0),
                                           (15360, 128, 192, 256, 0), (15360, 32, 384, 512, 0), (300, 48, 20, 20, 1),
                                           (129, 272, 36, 36, 0), (1, 16, 4, 4, 0)])
def test_tc_gemm_nt_3xtf32(M, N, K, lda, acc):
    from multimodalsignal_b200 import _ext
    lib = _ext.lib()
    torch.manual_seed(M + N + K)
    rows = M if lda >= K and lda < 4096 else M
    Abuf = torch.randn(M * lda + K, dtype=torch.float64)
    A = torch.as_strided(Abuf, (M, K), (lda, 1))
    W = torch.randn(N, K, dtype=torch.float64)
    bias = torch.randn(N, dtype=torch.float64)
    C0 = torch.randn(M, N, dtype=torch.float64)
    Ad, Wd, bd, Cd = Abuf.float().cuda(), W.float().cuda().contiguous(), bias.float().cuda(), C0.float().cuda().contiguous()
    A32 = torch.as_strided(Ad, (M, K), (lda, 1)).double().cpu()
    ref = A32 @ Wd.double().cpu().t() + bd.double().cpu() + (Cd.double().cpu() if acc else 0)
    _ext.check(lib.mms_tc_gemm_nt(Ad.data_ptr(), lda, Wd.data_ptr(), K, bd.data_ptr(), Cd.data_ptr(), N, M, N, K, acc,
                                  torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    err = (Cd.double().cpu() - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err <= 1e-5 * scale, (err, scale)
