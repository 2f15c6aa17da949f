"""Operator-level parity: every CUDA kernel, called through the C ABI, against the CPU oracle
(oracle/model_oracle.py, pinned to the reference by tests/test_oracle_model.py).

Tolerances: fp32 kernels vs a float64 evaluation of the oracle -- |err| <= 1e-5 * scale for
single ops (scale = max|ref|), looser only where stated."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import model_oracle as mo

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    from multimodalsignal_b200 import _ext
    return _ext.lib()


def P(t):
    return None if t is None else t.data_ptr()


def ST():
    return torch.cuda.current_stream().cuda_stream


def ok(rc):
    from multimodalsignal_b200 import _ext
    _ext.check(rc)


def dev(t, dtype=torch.float32):
    return t.to(device="cuda", dtype=dtype).contiguous()


def close(got, ref, tol=1e-5, what=""):
    got = got.detach().double().cpu()
    ref = ref.detach().double().cpu()
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    if ref.numel() == 0:
        return
    scale = max(1e-6, ref.abs().max().item())
    err = (got - ref).abs().max().item()
    assert err <= tol * scale, f"{what}: max err {err:.3e} vs scale {scale:.3e} (tol {tol})"


# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("Cc,T", [(3, 640), (6, 3840), (8, 333), (14, 641)])
def test_chan_attn(lib, Cc, T):
    torch.manual_seed(Cc)
    B, A = 5, Cc // 4
    x = torch.randn(B, Cc, T, dtype=torch.float64) + 0.3
    w1 = torch.randn(A, Cc, dtype=torch.float64) * 0.5
    w2 = torch.randn(Cc, A, dtype=torch.float64) * 0.5
    dy = torch.randn(B, Cc, T, dtype=torch.float64)
    xr, w1r, w2r = x.clone().requires_grad_(), w1.clone().requires_grad_(A > 0), w2.clone().requires_grad_(A > 0)
    y, gate, mean = mo.channel_attention(xr, w1r, w2r)
    y.backward(dy)

    xd, w1d, w2d, dyd = dev(x), dev(w1), dev(w2), dev(dy)
    meand, gated, yd = torch.empty(B, Cc, device="cuda"), torch.empty(B, Cc, device="cuda"), torch.empty_like(xd)
    ok(lib.mms_chan_attn_fwd(P(xd), P(w1d) if A else None, P(w2d) if A else None, B, Cc, T, P(meand), P(gated), P(yd), ST()))
    close(meand, mean, 1e-5, "mean")
    close(gated, gate, 1e-5, "gate")
    close(yd, y, 1e-5, "y")
    dx, dw1, dw2 = torch.empty_like(xd), torch.zeros_like(w1d), torch.zeros_like(w2d)
    scratch = torch.empty(4 * B * Cc, device="cuda")
    ok(lib.mms_chan_attn_bwd(P(xd), P(dyd), P(w1d) if A else None, P(w2d) if A else None, P(meand), P(gated), B, Cc, T,
                             P(dx), P(dw1) if A else None, P(dw2) if A else None, P(scratch), ST()))
    close(dx, xr.grad, 2e-5, "dx")
    if A:
        close(dw1, w1r.grad, 5e-5, "dw1")
        close(dw2, w2r.grad, 5e-5, "dw2")


CONV = {1: (7, 2, 3), 2: (5, 2, 2)}


@pytest.mark.parametrize("which,ci,co,lin,gated", [(1, 6, 16, 3840, True), (1, 14, 16, 641, True), (1, 3, 16, 336, False),
                                                   (2, 16, 32, 960, False), (2, 16, 32, 85, False), (2, 16, 16, 160, False),
                                                   (2, 16, 64, 161, False), (1, 14, 16, 1000, True), (2, 16, 32, 100, False),
                                                   (1, 16, 16, 520, True)])
def test_conv1d_fwd_dgrad_wgrad(lib, which, ci, co, lin, gated):
    torch.manual_seed(which * 100 + ci)
    k, s, p = CONV[which]
    B = 3
    x = torch.randn(B, ci, lin, dtype=torch.float64)
    w = torch.randn(co, ci, k, dtype=torch.float64) * 0.2
    gate = torch.rand(B, ci, dtype=torch.float64) + 0.25 if gated else None
    xr, wr = x.clone().requires_grad_(), w.clone().requires_grad_()
    gr = gate.clone().requires_grad_() if gated else None
    xin = xr * gr[:, :, None] if gated else xr
    y = mo.conv1d_nobias(xin, wr, s, p)
    dy = torch.randn_like(y)
    y.backward(dy)
    lout = y.shape[2]

    xd, wd, dyd = dev(x), dev(w), dev(dy)
    gd = dev(gate) if gated else None
    yd = torch.empty(B, co, lout, device="cuda")
    stats = torch.zeros(2 * co, device="cuda", dtype=torch.float64)
    ok(lib.mms_conv1d_fwd(which, P(xd), P(wd), P(gd), B, ci, co, lin, P(yd), P(stats), ST()))
    close(yd, y, 1e-5, "conv fwd")
    close(stats[:co], y.detach().sum(dim=(0, 2)), 1e-5, "sum")
    close(stats[co:], (y.detach() ** 2).sum(dim=(0, 2)), 1e-5, "sumsq")

    if lin % 4 == 0 and lin >= 64 and co <= 32:      # the tcgen05 implicit-GEMM version of the same convolution
        yd2 = torch.empty(B, co, lout, device="cuda")
        stats2 = torch.zeros(2 * co, device="cuda", dtype=torch.float64)
        ok(lib.mms_conv1d_fwd_tc(which, P(xd), P(wd), P(gd), B, ci, co, lin, P(yd2), P(stats2), ST()))
        close(yd2, y, 1e-5, "conv fwd (tcgen05)")
        close(stats2[:co], y.detach().sum(dim=(0, 2)), 1e-5, "sum (tcgen05)")
        close(stats2[co:], (y.detach() ** 2).sum(dim=(0, 2)), 1e-5, "sumsq (tcgen05)")

    dxd = torch.empty_like(xd)
    dgate = torch.zeros(B, ci, device="cuda")
    ok(lib.mms_conv1d_dgrad(which, P(dyd), P(wd), B, ci, co, lin, P(dxd), P(xd), P(dgate), ST()))
    # dgrad is w.r.t. the (gated) conv input
    dxin = xr.grad / gr.detach()[:, :, None] if gated else xr.grad
    close(dxd, dxin, 1e-5, "conv dgrad")
    close(dgate, (dxin * x).sum(dim=2), 2e-5, "dgate dot")
    if gated:
        close(dgate, gr.grad, 2e-5, "dgate")
    dwd = torch.zeros_like(wd)
    ok(lib.mms_conv1d_wgrad(which, P(xd), P(dyd), P(gd), B, ci, co, lin, P(dwd), ST()))
    close(dwd, wr.grad, 2e-5, "conv wgrad")


@pytest.mark.parametrize("Cn,lin,tm", [(16, 1920, 0), (32, 480, 1), (16, 321, 0), (32, 43, 1), (64, 50, 1)])
@pytest.mark.parametrize("training", [1, 0])
def test_bn_relu_pool(lib, Cn, lin, tm, training):
    torch.manual_seed(Cn + lin)
    B = 4
    y = torch.randn(B, Cn, lin, dtype=torch.float64) * 1.5 + 0.2
    y[:, :, 10:20] = -3.0           # a run of ReLU zeros -> pooling ties (first index must win)
    gamma = torch.rand(Cn, dtype=torch.float64) + 0.5
    gamma[1] = -0.7                 # negative scale: max does not commute with the affine map
    beta = torch.randn(Cn, dtype=torch.float64) * 0.3
    rm, rv = torch.randn(Cn, dtype=torch.float64) * 0.1, torch.rand(Cn, dtype=torch.float64) + 0.5
    yr, gr, br = y.clone().requires_grad_(), gamma.clone().requires_grad_(), beta.clone().requires_grad_()
    if training:
        yb, mean, var = mo.batchnorm_train(yr, gr, br)
    else:
        yb = mo.batchnorm_eval(yr, gr, br, rm, rv)
    out = mo.maxpool3s2p1(torch.relu(yb))
    dout = torch.randn_like(out)
    out.backward(dout)
    lout = out.shape[2]

    yd, gd, bd = dev(y), dev(gamma), dev(beta)
    rmd, rvd = dev(rm), dev(rv)
    nbt = torch.zeros(1, dtype=torch.int64, device="cuda")
    stats = torch.stack([y.sum(dim=(0, 2)), (y ** 2).sum(dim=(0, 2))]).reshape(-1).to("cuda", torch.float64).contiguous()
    outd = torch.empty((B, lout, Cn) if tm else (B, Cn, lout), device="cuda")
    ok(lib.mms_bn_relu_pool_fwd(P(yd), P(stats), P(gd), P(bd), P(rmd), P(rvd), P(nbt), B, Cn, lin, training, tm, P(outd), ST()))
    ref_out = out.permute(0, 2, 1) if tm else out
    close(outd, ref_out, 2e-5, "bn_relu_pool fwd")
    if training:
        n = B * lin
        erm, erv = mo.running_stats_update(rm, rv, mean.detach(), var.detach(), n)
        close(rmd, erm, 1e-5, "running_mean")
        close(rvd, erv, 1e-5, "running_var")
        assert int(nbt.item()) == 1
    else:
        close(rmd, rm, 1e-7, "running_mean untouched")
        assert int(nbt.item()) == 0

    doutd = dev(dout.permute(0, 2, 1) if tm else dout)
    dyd = torch.empty_like(yd)
    dg, db = torch.zeros(Cn, device="cuda"), torch.zeros(Cn, device="cuda")
    red = torch.zeros(2 * Cn, dtype=torch.float64, device="cuda")
    ok(lib.mms_bn_relu_pool_bwd(P(yd), P(stats), P(gd), P(bd), P(rmd), P(rvd), P(doutd), B, Cn, lin, training, tm,
                                P(dyd), P(dg), P(db), P(red), ST()))
    close(dyd, yr.grad, 5e-5, "bn_relu_pool dy")
    close(dg, gr.grad, 5e-5, "dgamma")
    close(db, br.grad, 5e-5, "dbeta")


def test_gemms(lib):
    torch.manual_seed(0)
    for (M, N, K) in [(300, 384, 32), (130, 192, 128), (7, 50, 19)]:
        A = torch.randn(M, K, dtype=torch.float64)
        W = torch.randn(N, K, dtype=torch.float64)
        bias = torch.randn(N, dtype=torch.float64)
        Ad, Wd, bd = dev(A), dev(W), dev(bias)
        Cd = torch.empty(M, N, device="cuda")
        ok(lib.mms_gemm_nt_bias(P(Ad), K, P(Wd), K, P(bd), P(Cd), N, M, N, K, ST()))
        close(Cd, A @ W.t() + bias, 1e-5, f"nt {M}x{N}x{K}")
        Wn = torch.randn(K, N, dtype=torch.float64)
        Wnd = dev(Wn)
        C0 = torch.randn(M, N, dtype=torch.float64)
        Cd = dev(C0)
        ok(lib.mms_gemm_nn(P(Ad), K, P(Wnd), N, P(Cd), N, M, N, K, 1, ST()))
        close(Cd, C0 + A @ Wn, 1e-5, f"nn acc {M}x{N}x{K}")
        ok(lib.mms_gemm_nn(P(Ad), K, P(Wnd), N, P(Cd), N, M, N, K, 0, ST()))
        close(Cd, A @ Wn, 1e-5, f"nn {M}x{N}x{K}")
    # tn with the column remap (a_split / a_skip) and the h_{t-1} row shift
    B, L, H = 5, 37, 64
    M = B * L
    D = torch.randn(M, 4 * H, dtype=torch.float64)
    hs = torch.randn(B, L, H, dtype=torch.float64)
    Dd, hsd = dev(D), dev(hs.reshape(M, H))          # keep the device tensors alive across the launches
    for shift in (-1, 1, 0):
        hprev = torch.zeros_like(hs)
        if shift == -1:
            hprev[:, 1:] = hs[:, :-1]
        elif shift == 1:
            hprev[:, :-1] = hs[:, 1:]
        else:
            hprev = hs
        dgh = torch.cat([D[:, :2 * H], D[:, 3 * H:]], dim=1)
        ref = dgh.t() @ hprev.reshape(M, H)
        Cd = torch.zeros(3 * H, H, device="cuda")
        bgd = torch.zeros(3 * H, device="cuda")
        ok(lib.mms_gemm_tn_acc(P(Dd), 4 * H, 2 * H, H, P(hsd), H, shift, L, P(Cd), H, P(bgd), M, 3 * H, H, ST()))
        close(Cd, ref, 2e-5, f"tn shift {shift}")
        close(bgd, dgh.sum(dim=0), 2e-5, "tn bias")
    bgd = torch.zeros(3 * H, device="cuda")
    ok(lib.mms_gemm_tn_acc(P(Dd), 4 * H, 2 * H, H, None, 0, 0, 1, None, 0, P(bgd), M, 3 * H, 0, ST()))
    close(bgd, torch.cat([D[:, :2 * H], D[:, 3 * H:]], dim=1).sum(dim=0), 2e-5, "tn bias only")

@pytest.mark.parametrize("M,N,K", [(64, 192, 128), (64, 128, 192), (5, 50, 32), (70, 9, 256)])
def test_gemm_skinny(lib, M, N, K):
    """Few-row products of the single reverse step: both operand orders against float64, and the two dropout
    variants against the multipliers mms_dropout_apply draws for the same element ids."""
    torch.manual_seed(M + K)
    Lrows = 3                                              # the rows are a strided slice [m*Lrows + 2] of a taller operand
    Afull = torch.randn(M * Lrows, K, dtype=torch.float64)
    A = Afull.view(M, Lrows, K)[:, 2]
    Wnt = torch.randn(N, K, dtype=torch.float64)
    Wnn = torch.randn(K, (N + 3) // 4 * 4, dtype=torch.float64)        # row stride a multiple of 4 floats
    bias = torch.randn(N, dtype=torch.float64)
    Afd, Wntd, Wnnd, bd = dev(Afull), dev(Wnt), dev(Wnn), dev(bias)
    a_ptr = Afd.data_ptr() + 2 * K * 4
    Cd = torch.full((M, N), 7.0, device="cuda")
    ok(lib.mms_gemm_skinny(a_ptr, Lrows * K, P(Wntd), K, 1, P(bd), P(Cd), N, M, N, K, 0, 0, 0, 0.0, 0, 0, None, ST()))
    close(Cd, A @ Wnt.t() + bias, 1e-5, "skinny nt")
    ldw = Wnn.shape[1]
    ok(lib.mms_gemm_skinny(a_ptr, Lrows * K, P(Wnnd), ldw, 0, None, P(Cd), N, M, N, K, 0, 0, 0, 0.0, 0, 0, None, ST()))
    close(Cd, A @ Wnn[:, :N], 1e-5, "skinny nn")

    # dropout on the operand: ids base + m*row_stride + k, the multipliers of mms_dropout_apply over the whole tall operand
    p, seed, off, base = 0.5, 11, 5, 1 << 40
    mult = torch.ones(M * Lrows * K, device="cuda")
    ok(lib.mms_dropout_apply(P(mult), P(mult), M * Lrows * K, base, p, seed, off, None, ST()))
    mA = mult.view(M, Lrows, K)[:, 2].double().cpu()
    ok(lib.mms_gemm_skinny(a_ptr, Lrows * K, P(Wntd), K, 1, P(bd), P(Cd), N, M, N, K, 1, base + 2 * K, Lrows * K, p, seed, off, None, ST()))
    close(Cd, (A * mA) @ Wnt.t() + bias, 1e-5, "skinny nt, dropout on A")
    assert 0.3 < (mA == 0).double().mean().item() < 0.7
    # dropout on the result: ids base + m*row_stride + n
    multC = torch.ones(M * Lrows * N, device="cuda")
    ok(lib.mms_dropout_apply(P(multC), P(multC), M * Lrows * N, base, p, seed, off, None, ST()))
    mC = multC.view(M, Lrows, N)[:, 2].double().cpu()
    ok(lib.mms_gemm_skinny(a_ptr, Lrows * K, P(Wnnd), ldw, 0, None, P(Cd), N, M, N, K, 2, base + 2 * N, Lrows * N, p, seed, off, None, ST()))
    close(Cd, (A @ Wnn[:, :N]) * mC, 1e-5, "skinny nn, dropout on C")


def _gru_case(lib, B, L, H, I, reverse, steps=None, with_dout=True):
    from multimodalsignal_b200._ext import GruDirFwd, GruDirBwd
    torch.manual_seed(B * 7 + L + H + int(reverse))
    s = 1.0 / np.sqrt(H)
    x = torch.randn(B, L, I, dtype=torch.float64)
    w_ih = (torch.rand(3 * H, I, dtype=torch.float64) * 2 - 1) * s
    w_hh = (torch.rand(3 * H, H, dtype=torch.float64) * 2 - 1) * s
    b_ih = (torch.rand(3 * H, dtype=torch.float64) * 2 - 1) * s
    b_hh = (torch.rand(3 * H, dtype=torch.float64) * 2 - 1) * s
    leaves = [t.clone().requires_grad_() for t in (x, w_ih, w_hh, b_ih, b_hh)]
    out = mo.gru_direction(*leaves, reverse=reverse, steps=steps)
    dout = torch.randn_like(out) if with_dout else torch.zeros_like(out)
    nsteps = L if steps is None else steps
    t_last = (0 if reverse else L - 1) if steps is None else ((L - steps) if reverse else steps - 1)
    dlast = torch.randn(B, H, dtype=torch.float64)
    if steps is not None:      # only the visited time steps carry gradient
        mask = torch.zeros(L, dtype=torch.float64)
        idx = range(L - 1, L - 1 - steps, -1) if reverse else range(steps)
        mask[list(idx)] = 1
        dout = dout * mask[None, :, None]
    total = (out * dout).sum() + (out[:, t_last, :] * dlast).sum()
    total.backward()
    gi = (x @ w_ih.t() + b_ih)

    gid, whd, bhd = dev(gi), dev(w_hh), dev(b_hh)
    hs = torch.zeros(B, L, H, device="cuda")
    stash = torch.zeros(B, L, 4 * H, device="cuda")
    d = GruDirFwd()
    d.gi, d.gi_bs, d.gi_ts = P(gid), L * 3 * H, 3 * H
    d.w_hh, d.b_hh = P(whd), P(bhd)
    d.hs, d.hs_bs, d.hs_ts = P(hs), L * H, H
    d.hs_drop, d.drop_base = None, 0
    d.stash, d.st_bs, d.st_ts = P(stash), L * 4 * H, 4 * H
    d.t0, d.dt, d.nsteps = (L - 1 if reverse else 0), (-1 if reverse else 1), nsteps
    ok(lib.mms_gru_recur_fwd(C.byref(d), 1, B, H, 0.0, 0, 0, None, ST()))
    close(hs, out, 2e-5, f"gru fwd B{B} L{L} H{H} rev{reverse}")

    doutd, dlastd = dev(dout), dev(dlast)
    D = torch.zeros(B, L, 4 * H, device="cuda")
    bd = GruDirBwd()
    bd.w_hh = P(whd)
    bd.stash, bd.st_bs, bd.st_ts = P(stash), L * 4 * H, 4 * H
    bd.hs, bd.hs_bs, bd.hs_ts = P(hs), L * H, H
    bd.dout, bd.do_bs, bd.do_ts, bd.drop_base, bd.drop_mask = P(doutd), L * H, H, 0, 0
    bd.dout_last, bd.dl_ld = P(dlastd), H
    bd.dh_head, bd.w0, bd.w0_ld, bd.w0_col = None, None, 0, 0
    bd.D, bd.d_bs, bd.d_ts = P(D), L * 4 * H, 4 * H
    bd.t0, bd.dt, bd.nsteps = d.t0, d.dt, nsteps
    ok(lib.mms_gru_recur_bwd(C.byref(bd), 1, B, H, 0.0, 0, 0, None, ST()))
    Dc = D.double().cpu().reshape(B * L, 4 * H)
    dgi = Dc[:, :3 * H]
    dgh = torch.cat([Dc[:, :2 * H], Dc[:, 3 * H:]], dim=1)
    xf = x.reshape(B * L, I)
    close(dgi.t() @ xf, leaves[1].grad, 5e-5, "dW_ih")
    close(dgi.sum(0), leaves[3].grad, 5e-5, "db_ih")
    close(dgh.sum(0), leaves[4].grad, 5e-5, "db_hh")
    close((dgi @ w_ih).reshape(B, L, I), leaves[0].grad, 5e-5, "dx")
    hprev = torch.zeros(B, L, H, dtype=torch.float64)
    o = out.detach()
    if reverse:
        hprev[:, :-1] = o[:, 1:]
    else:
        hprev[:, 1:] = o[:, :-1]
    close(dgh.t() @ hprev.reshape(B * L, H), leaves[2].grad, 5e-5, "dW_hh")


@pytest.mark.parametrize("B,L,H,I,reverse", [(5, 40, 64, 32, False), (5, 40, 64, 32, True), (3, 33, 32, 32, False),
                                             (3, 33, 32, 16, True), (150, 12, 64, 32, False), (301, 9, 64, 32, True),
                                             (301, 9, 32, 32, False), (2, 240, 64, 128, False)])
def test_gru_recurrence(lib, B, L, H, I, reverse):
    _gru_case(lib, B, L, H, I, reverse)


@pytest.mark.parametrize("depth", [4, 8])
def test_gru_backward_shared_memory_ring(lib, depth):
    """The cp.async shared-memory-ring backward (GRU_BWD_RING = ring depth) against the same oracle, including
    sequences shorter than the ring, single steps and the fall-back to the register-ring kernel for R > 1 rows per CTA."""
    prev = lib.mms_get_option(b"GRU_BWD_RING", -1)
    ok(lib.mms_set_option(b"GRU_BWD_RING", depth))
    try:
        for case in [(5, 40, 64, 32, False), (5, 40, 64, 32, True), (3, 33, 32, 32, False), (3, 33, 32, 16, True),
                     (2, 240, 64, 128, False), (7, 2, 64, 32, True), (301, 9, 64, 32, True)]:
            _gru_case(lib, *case)
        _gru_case(lib, 4, 20, 64, 128, True, steps=1)
        _gru_case(lib, 4, 20, 64, 128, False, steps=3)
        _gru_case(lib, 4, 20, 64, 128, True, steps=depth + 1)
        _gru_case(lib, 4, 20, 32, 64, False, steps=depth, with_dout=False)
    finally:
        ok(lib.mms_set_option(b"GRU_BWD_RING", prev) if prev >= 0 else lib.mms_clear_option(b"GRU_BWD_RING"))


def test_gru_backward_register_ring(lib):
    """GRU_BWD_RING = 0 selects the register-ring kernel (also the path for R > 1 rows per CTA / unaligned operands)."""
    prev = lib.mms_get_option(b"GRU_BWD_RING", -1)
    ok(lib.mms_set_option(b"GRU_BWD_RING", 0))
    try:
        for case in [(5, 40, 64, 32, False), (5, 40, 64, 32, True), (3, 33, 32, 16, True), (2, 240, 64, 128, False)]:
            _gru_case(lib, *case)
        _gru_case(lib, 4, 20, 64, 128, True, steps=1)
        _gru_case(lib, 4, 20, 64, 128, False, steps=3)
    finally:
        ok(lib.mms_set_option(b"GRU_BWD_RING", prev) if prev >= 0 else lib.mms_clear_option(b"GRU_BWD_RING"))


def test_gru_single_reverse_step(lib):
    """The top layer's reverse direction is observed after ONE step only."""
    _gru_case(lib, 4, 20, 64, 128, True, steps=1)
    _gru_case(lib, 4, 20, 64, 128, False, steps=3)


@pytest.mark.parametrize("B,H2,nc", [(7, 128, 2), (64, 128, 3), (5, 64, 2), (300, 128, 2)])
def test_head_and_cross_entropy(lib, B, H2, nc):
    torch.manual_seed(B)
    last = torch.randn(B, H2, dtype=torch.float64)
    w0 = torch.randn(64, H2, dtype=torch.float64) * 0.1
    b0 = torch.randn(64, dtype=torch.float64) * 0.1
    w3 = torch.randn(nc, 64, dtype=torch.float64) * 0.2
    b3 = torch.randn(nc, dtype=torch.float64) * 0.1
    y = torch.randint(0, nc, (B,))
    leaves = [t.clone().requires_grad_() for t in (last, w0, b0, w3, b3)]
    hid = torch.relu(leaves[0] @ leaves[1].t() + leaves[2])
    logits = hid @ leaves[3].t() + leaves[4]
    loss = mo.cross_entropy_mean(logits, y)
    (dlog,) = torch.autograd.grad(loss, logits, retain_graph=True)
    loss.backward()

    lastd, w0d, b0d, w3d, b3d, yd = dev(last), dev(w0), dev(b0), dev(w3), dev(b3), y.cuda()
    hidd = torch.empty(B, 64, device="cuda")
    logd = torch.empty(B, nc, device="cuda")
    ok(lib.mms_head_fwd(P(lastd), P(w0d), P(b0d), P(w3d), P(b3d), B, H2, nc, 0.0, 0, 0, None, P(hidd), P(logd), ST()))
    close(logd, logits, 1e-5, "logits")
    lossd = torch.zeros(1, device="cuda")
    dlogd = torch.empty(B, nc, device="cuda")
    acc = torch.full((1,), 2.0, dtype=torch.float64, device="cuda")
    ok(lib.mms_cross_entropy(P(logd), P(yd), B, nc, P(lossd), P(dlogd), P(acc), ST()))
    close(lossd[0], loss, 1e-5, "loss")
    close(dlogd, dlog, 1e-5, "dlogits")
    assert abs(acc.item() - (2.0 + loss.item() * B)) < 1e-4 * B
    dhid = torch.empty(B, 64, device="cuda")
    g = [torch.zeros_like(t) for t in (w0d, b0d, w3d, b3d)]
    ok(lib.mms_head_bwd(P(lastd), P(hidd), P(dlogd), P(w3d), B, H2, nc, 0.0, 0, 0, None, P(dhid), P(g[0]), P(g[1]), P(g[2]), P(g[3]), ST()))
    for got, leaf, name in zip(g, leaves[1:], ("dw0", "db0", "dw3", "db3")):
        close(got, leaf.grad, 5e-5, name)
    close(dhid.double().cpu() @ w0, leaves[0].grad, 5e-5, "dlast")


def test_head_dropout_is_consistent_between_fwd_and_bwd(lib):
    """With p > 0 the mask is regenerated from (seed, offset, element id) in the backward."""
    torch.manual_seed(1)
    B, H2, nc, p = 50, 128, 2, 0.5
    lastd, w0d, b0d = dev(torch.randn(B, H2)), dev(torch.randn(64, H2) * 0.1), dev(torch.zeros(64) + 0.5)
    w3d, b3d = dev(torch.randn(nc, 64)), dev(torch.zeros(nc))
    hid, log = torch.empty(B, 64, device="cuda"), torch.empty(B, nc, device="cuda")
    ok(lib.mms_head_fwd(P(lastd), P(w0d), P(b0d), P(w3d), P(b3d), B, H2, nc, p, 11, 5, None, P(hid), P(log), ST()))
    # recover the multiplier from the logits: logits = (hid*m) @ w3^T  -> solve via probing dlogits
    dlog = torch.zeros(B, nc, device="cuda")
    dlog[:, 0] = 1.0
    dhid = torch.empty(B, 64, device="cuda")
    g = [torch.zeros(64, H2, device="cuda"), torch.zeros(64, device="cuda"), torch.zeros(nc, 64, device="cuda"), torch.zeros(nc, device="cuda")]
    ok(lib.mms_head_bwd(P(lastd), P(hid), P(dlog), P(w3d), B, H2, nc, p, 11, 5, None, P(dhid), P(g[0]), P(g[1]), P(g[2]), P(g[3]), ST()))
    m = torch.where(hid > 0, dhid / w3d[0][None, :], torch.zeros_like(hid))
    kept = m[hid > 0]
    assert set(torch.unique(torch.round(kept)).tolist()) <= {0.0, 2.0}
    frac = (kept > 1).float().mean().item()
    assert 0.4 < frac < 0.6
    relog = (hid * torch.where(hid > 0, m, torch.zeros_like(m))) @ w3d.t() + b3d
    close(log, relog, 1e-4, "dropout mask identical in fwd and bwd")


def test_adam_flat(lib):
    torch.manual_seed(3)
    n = 100003
    p = torch.randn(n, dtype=torch.float32)
    m, v = torch.zeros(n), torch.zeros(n)
    pd_, md, vd = p.cuda(), m.cuda(), v.cuda()
    lr = torch.tensor([1e-3], device="cuda")
    step = torch.zeros(1, dtype=torch.int64, device="cuda")
    scratch = torch.zeros(1, dtype=torch.int32, device="cuda")
    ref = torch.nn.Parameter(p.clone())
    opt = torch.optim.Adam([ref], lr=1e-3, weight_decay=1e-4)
    for it in range(4):
        g = torch.randn(n)
        ref.grad = g.clone()
        opt.step()
        gd = g.cuda()
        ok(lib.mms_adam_flat_step(P(pd_), P(gd), P(md), P(vd), n, P(lr), 0.9, 0.999, 1e-8, 1e-4, P(step), P(scratch), ST()))
        assert int(step.item()) == it + 1
        np.testing.assert_allclose(pd_.cpu().numpy(), ref.detach().numpy(), atol=2e-6)
        # oracle restatement agrees too
    po, mo_, vo = p.clone(), torch.zeros(n), torch.zeros(n)
    g = torch.randn(n)
    po2, _, _ = mo.adam_step(po, g, mo_, vo, 1)
    ref2 = torch.nn.Parameter(p.clone())
    opt2 = torch.optim.Adam([ref2], lr=1e-3, weight_decay=1e-4)
    ref2.grad = g.clone()
    opt2.step()
    np.testing.assert_allclose(po2.numpy(), ref2.detach().numpy(), atol=1e-6)


def test_window_gather_and_stats(lib):
    rng = np.random.default_rng(0)
    n_ch, n, win = 8, 50000, 3840
    streams = [torch.from_numpy(rng.standard_normal(n) + (3.0 if c == 4 else 0.0)).cuda() for c in range(n_ch)]
    starts_np = np.asarray([0, 640, 1280, 7777, n - win], dtype=np.int64)
    starts = torch.from_numpy(starts_np).cuda()
    arr = (C.c_void_p * n_ch)(*[s.data_ptr() for s in streams])
    out = torch.empty(len(starts_np), win, n_ch, dtype=torch.float64, device="cuda")
    ok(lib.mms_window_gather(arr, n_ch, n, P(starts), len(starts_np), win, 0, None, None, None, P(out), ST()))
    host = np.stack([s.cpu().numpy() for s in streams], axis=1)
    ref = np.stack([host[s:s + win] for s in starts_np])
    assert np.array_equal(out.cpu().numpy(), ref)                      # a pure copy: bit-exact
    flags_np = np.zeros(n_ch, dtype=np.int32)
    flags_np[4] = 1
    flags = torch.from_numpy(flags_np).cuda()
    sums = torch.zeros(n_ch, 2, dtype=torch.float64, device="cuda")
    ok(lib.mms_window_stats(arr, n_ch, n, P(starts), len(starts_np), win, P(flags), P(sums), ST()))
    refx = ref.copy()
    refx[:, :, 4] = np.log1p(refx[:, :, 4])
    np.testing.assert_allclose(sums.cpu().numpy()[:, 0], refx.sum(axis=(0, 1)), rtol=1e-12, atol=1e-9)
    np.testing.assert_allclose(sums.cpu().numpy()[:, 1], (refx ** 2).sum(axis=(0, 1)), rtol=1e-12)
    cnt = refx.shape[0] * refx.shape[1]
    mean = refx.mean(axis=(0, 1))
    std = refx.std(axis=(0, 1)) + 1e-8
    shift, scale = torch.from_numpy(mean).cuda(), torch.from_numpy(1.0 / std).cuda()
    outf = torch.empty(len(starts_np), n_ch, win, dtype=torch.float32, device="cuda")
    ok(lib.mms_window_gather(arr, n_ch, n, P(starts), len(starts_np), win, 1, P(shift), P(scale), P(flags), P(outf), ST()))
    want = ((refx - mean) / std).astype(np.float32).transpose(0, 2, 1)
    np.testing.assert_allclose(outf.cpu().numpy(), want, atol=2e-6)
    assert cnt == len(starts_np) * win


def test_dropout_apply_same_mask_forward_and_backward(lib):
    n, p = 100003, 0.5
    x = torch.randn(n, device="cuda")
    y = torch.empty_like(x)
    ok(lib.mms_dropout_apply(P(x), P(y), n, 12345, p, 7, 3, None, ST()))
    m = y / x
    kept = (m.abs() > 0).float().mean().item()
    assert 0.48 < kept < 0.52
    assert torch.allclose(m[m.abs() > 0], torch.full_like(m[m.abs() > 0], 2.0), atol=1e-5)
    g = torch.ones(n, device="cuda")
    ok(lib.mms_dropout_apply(P(g), P(g), n, 12345, p, 7, 3, None, ST()))          # in place, same ids
    assert torch.equal(g > 0, y != 0)
    g2 = torch.ones(n, device="cuda")
    ok(lib.mms_dropout_apply(P(g2), P(g2), n, 12345, p, 7, 4, None, ST()))        # next step: different mask
    assert not torch.equal(g2 > 0, g > 0)
