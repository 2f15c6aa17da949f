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


@pytest.mark.parametrize("B,L,H,N2,lda_mul,ldb_mul,mode,shift", [
    (64, 240, 64, 128, 4, 1, "ih", 0),        # dW_ih of the top layer: D[:, r|z|n] x layer input
    (64, 240, 64, 64, 4, 1, "hh", -1),        # dW_hh forward direction: D[:, r|z|dq] x h_{t-1}
    (64, 240, 64, 64, 8, 2, "hh", 1),         # layer 0, reverse direction (operands are column slices of wider rows)
    (64, 240, 64, 32, 8, 1, "ih", 0),         # layer 0 input projection (I = cnn_out = 32)
    (7, 37, 64, 64, 4, 1, "hh", -1),          # ragged: M = 259 is not a multiple of the k-block
    (9, 50, 32, 32, 4, 1, "hh", 1),           # H = 32 (N1 = 96)
    (64, 240, 64, 0, 4, 1, "hh", 0),          # bias gradient only
])
def test_tc_gemm_tn_3xtf32(B, L, H, N2, lda_mul, ldb_mul, mode, shift):
    """Weight-gradient products on tcgen05 (MN-major operands, split-K, red.global.add) against float64:
    C[i, j] += sum_m A[m, acol(i)] * Bm[m + shift, j] with the sequence-boundary rows zeroed, and the bias gradient
    from the implicit ones column.  Tolerance 2e-5 of max|C| (fp32-class; plain tf32 would be ~5e-4)."""
    from multimodalsignal_b200 import _ext
    lib = _ext.lib()
    torch.manual_seed(B + L + N2 + shift)
    M = B * L
    lda, ldb = lda_mul * H, max(N2, 1) * ldb_mul
    D = torch.randn(M, lda, dtype=torch.float64)
    Bsrc = torch.randn(M, ldb, dtype=torch.float64)
    col_off = ldb - N2 if ldb_mul > 1 else 0         # read the right-hand half of wider rows
    Dd, Bd = D.float().cuda(), Bsrc.float().cuda()
    a_split, a_skip = (3 * H, 0) if mode == "ih" else (2 * H, H)
    A64 = Dd.double().cpu()
    Asel = A64[:, :3 * H] if mode == "ih" else torch.cat([A64[:, :2 * H], A64[:, 3 * H:4 * H]], dim=1)
    C0 = torch.randn(3 * H, max(N2, 1), dtype=torch.float64)
    b0 = torch.randn(3 * H, dtype=torch.float64)
    Cd, bd = C0.float().cuda(), b0.float().cuda()
    if N2:
        Bv = Bd.double().cpu()[:, col_off:col_off + N2].reshape(B, L, N2)
        Bsh = torch.zeros_like(Bv)
        if shift == -1:
            Bsh[:, 1:] = Bv[:, :-1]
        elif shift == 1:
            Bsh[:, :-1] = Bv[:, 1:]
        else:
            Bsh = Bv
        refC = Cd.double().cpu() + Asel.t() @ Bsh.reshape(M, N2)
    refb = bd.double().cpu() + Asel.sum(dim=0)
    _ext.check(lib.mms_tc_gemm_tn(Dd.data_ptr(), lda, a_split, a_skip, Bd.data_ptr() + 4 * col_off if N2 else None, ldb, shift, L,
                                  Cd.data_ptr() if N2 else None, N2, bd.data_ptr(), M, 3 * H, N2,
                                  torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    if N2:
        err = (Cd.double().cpu() - refC).abs().max().item()
        assert err <= 2e-5 * refC.abs().max().item(), (err, refC.abs().max().item())
    errb = (bd.double().cpu() - refb).abs().max().item()
    assert errb <= 2e-5 * refb.abs().max().item(), (errb, refb.abs().max().item())
    # the SIMT kernel it replaces agrees too
    C2, b2 = C0.float().cuda(), b0.float().cuda()
    _ext.check(lib.mms_gemm_tn_acc(Dd.data_ptr(), lda, a_split, a_skip, Bd.data_ptr() + 4 * col_off if N2 else None, ldb, shift, L,
                                   C2.data_ptr() if N2 else None, N2, b2.data_ptr(), M, 3 * H, N2, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    if N2:
        assert (C2 - Cd).abs().max().item() <= 4e-5 * refC.abs().max().item()
    assert (b2 - bd).abs().max().item() <= 4e-5 * refb.abs().max().item()


@pytest.mark.parametrize("stages", [4, 2, 1, 3])
def test_tc_gemm_tn_batch_equals_single_launches(stages):
    """mms_tc_gemm_tn_batch on the four weight-gradient products of a bidirectional GRU layer (both directions' dW_ih and
    dW_hh out of one D [M, 8H]) against float64 and against four single launches."""
    from multimodalsignal_b200 import _ext
    lib = _ext.lib()
    B, L, H, I = 64, 240, 64, 32
    M = B * L
    torch.manual_seed(7)
    D = torch.randn(M, 8 * H).cuda()
    X = torch.randn(M, I).cuda()
    Hs = torch.randn(M, 2 * H).cuda()
    st = torch.cuda.current_stream().cuda_stream
    prev = lib.mms_get_option(b"TN_STAGES", -1)
    _ext.check(lib.mms_set_option(b"TN_STAGES", stages))
    try:
        outs, refs, calls = [], [], (_ext.TnCall * 4)()
        for dd in range(2):
            Dd = D[:, dd * 4 * H:(dd + 1) * 4 * H].double().cpu()
            for which in range(2):
                N2 = I if which == 0 else H
                Cd = torch.zeros(3 * H, N2, device="cuda")
                bd = torch.zeros(3 * H, device="cuda")
                if which == 0:
                    Asel, Bm, ldb, shift, a_split, a_skip = Dd[:, :3 * H], X, I, 0, 3 * H, 0
                    Bsh = X.double().cpu()
                else:
                    Asel = torch.cat([Dd[:, :2 * H], Dd[:, 3 * H:]], dim=1)
                    Bm, ldb, shift, a_split, a_skip = Hs[:, dd * H:], 2 * H, (1 if dd else -1), 2 * H, H
                    Bv = Hs[:, dd * H:(dd + 1) * H].double().cpu().reshape(B, L, H)
                    Bsh = torch.zeros_like(Bv)
                    if shift == -1:
                        Bsh[:, 1:] = Bv[:, :-1]
                    else:
                        Bsh[:, :-1] = Bv[:, 1:]
                    Bsh = Bsh.reshape(M, H)
                refs.append((Asel.t() @ Bsh, Asel.sum(dim=0)))
                outs.append((Cd, bd))
                c = calls[2 * dd + which]
                c.A, c.lda, c.a_split, c.a_skip = D.data_ptr() + 4 * dd * 4 * H, 8 * H, a_split, a_skip
                c.Bm, c.ldb, c.shift, c.seq = Bm.data_ptr(), ldb, shift, L
                c.C, c.ldc, c.bias_grad = Cd.data_ptr(), N2, bd.data_ptr()
                c.M, c.N1, c.N2 = M, 3 * H, N2
        _ext.check(lib.mms_tc_gemm_tn_batch(calls, 4, st))
        torch.cuda.synchronize()
        for (Cd, bd), (refC, refb) in zip(outs, refs):
            assert (Cd.double().cpu() - refC).abs().max().item() <= 2e-5 * refC.abs().max().item()
            assert (bd.double().cpu() - refb).abs().max().item() <= 2e-5 * refb.abs().max().item()
        # and the same four products launched one by one
        for j in range(4):
            c = calls[j]
            C1 = torch.zeros_like(outs[j][0])
            b1 = torch.zeros_like(outs[j][1])
            _ext.check(lib.mms_tc_gemm_tn(c.A, c.lda, c.a_split, c.a_skip, c.Bm, c.ldb, c.shift, c.seq, C1.data_ptr(), c.ldc, b1.data_ptr(),
                                          c.M, c.N1, c.N2, st))
            torch.cuda.synchronize()
            assert (C1 - outs[j][0]).abs().max().item() <= 4e-5 * refs[j][0].abs().max().item()
            assert (b1 - outs[j][1]).abs().max().item() <= 4e-5 * refs[j][1].abs().max().item()
    finally:
        _ext.check(lib.mms_set_option(b"TN_STAGES", prev) if prev >= 0 else lib.mms_clear_option(b"TN_STAGES"))
