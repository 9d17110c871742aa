"""Kernel-level parity on the B200: every C-ABI entry point against the CPU oracle / plain fp32 torch
formulas on the same seeded inputs.  bf16 tensor-core kernels: <= 2e-2 relative (north-star);
fp32 kernels (LayerNorm, losses, reductions): <= 1e-4 relative; index work: bit-exact."""
import math

import numpy as np
import pytest
import torch

from oracle import clover_oracle as O

pytestmark = pytest.mark.gpu

BF16, F32 = torch.bfloat16, torch.float32


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from clover_b200 import ops as _ops
    return _ops


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def rnd(*shape, seed=0, scale=1.0, dtype=F32):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(dtype).cuda()


# ------------------------------------------------------------------------------------------------ GEMM
@pytest.mark.parametrize("a_t,b_t", [(False, False), (False, True), (True, False), (True, True)])
@pytest.mark.parametrize("M,N,K", [(256, 256, 128), (300, 136, 96), (64, 768, 1024), (1000, 384, 64), (128, 128, 1000)])
def test_gemm_layouts(ops, a_t, b_t, M, N, K):
    if (a_t and M % 8) or (b_t and N % 8) or (not a_t and K % 8) or (not b_t and K % 8):
        pytest.skip("TMA needs 16-byte row pitch")
    A = rnd(M, K, seed=1, dtype=BF16)
    B = rnd(N, K, seed=2, dtype=BF16)
    a = A.t().contiguous() if a_t else A
    b = B.t().contiguous() if b_t else B
    out = torch.empty(M, N, dtype=F32, device="cuda")
    ops.gemm(a, b, out, a_t=a_t, b_t=b_t)
    ref = A.float() @ B.float().t()
    assert rel(out, ref) < 2e-3


@pytest.fixture(params=[1, 0, 2], ids=["ts_default", "ts_off", "ts_forced"])
def tma_store_mode(request, ops):
    """The GEMM epilogue's output path: 1 = default (TMA stores for dual / fp32 outputs), 0 = per-lane stores only,
    2 = TMA stores wherever the layout allows -- every epilogue test runs under all three."""
    ops.set_tunable("gemm_tma_store", request.param)
    yield request.param
    ops.set_tunable("gemm_tma_store", -1)


def test_gemm_epilogues(ops, tma_store_mode):
    M, N, K = 392, 256, 192
    A, B = rnd(M, K, seed=3, dtype=BF16), rnd(N, K, seed=4, scale=0.1, dtype=BF16)
    bias = rnd(N, seed=5)
    base = A.float() @ B.float().t() + bias
    # bias + q-scale prefix, bf16 out
    out = torch.empty(M, N, dtype=BF16, device="cuda")
    ops.gemm(A, B, out, bias=bias, scale_cols=96, scale=0.25)
    ref = base.clone(); ref[:, :96] *= 0.25
    assert rel(out, ref) < 1e-2
    # GELU with pre-activation copy
    pre = torch.empty(M, N, dtype=BF16, device="cuda")
    ops.gemm(A, B, out, bias=bias, act="gelu", out_pre=pre)
    assert rel(pre, base) < 1e-2
    assert rel(out, O.gelu(base.cpu())) < 1e-2
    # multiply by GELU'(pre) (fc2 dgrad)
    out32 = torch.empty(M, N, dtype=F32, device="cuda")
    ops.gemm(A, B, out32, gelu_pre=pre)
    p = pre.float().cpu().requires_grad_(True)
    O.gelu(p).sum().backward()
    assert rel(out32, (A.float() @ B.float().t()).cpu() * p.grad) < 2e-3
    # residual fp32 -> fp32, residual bf16 -> bf16
    res = rnd(M, N, seed=6)
    ops.gemm(A, B, out32, bias=bias, residual=res)
    assert rel(out32, base + res) < 2e-3
    resb = res.to(BF16)
    ops.gemm(A, B, out, bias=bias, residual=resb)
    assert rel(out, base + resb.float()) < 1e-2


def test_gemm_specialised_epilogues_match_generic(ops):
    """The compile-time epilogue specialisations (fc1, fc2 dgrad, qkv, plain dgrad, fc2 forward, proj scatter) against the
    generic kernel on the same inputs: same arithmetic in the same order, so the outputs agree to the last bit."""
    M, N, K = 1000, 512, 256
    A, B = rnd(M, K, seed=3, dtype=BF16), rnd(N, K, seed=4, scale=0.1, dtype=BF16)
    Bt = B.t().contiguous()
    bias, res, rs = rnd(N, seed=5), rnd(M, N, seed=6), (0.5 + torch.rand(4, generator=torch.Generator().manual_seed(1))).cuda()
    gder = rnd(M, N, seed=7, dtype=BF16)
    B_, D, H, W = 2, 4, 7, 14
    wg = ops.Window(B_, D, H, W, (4, 7, 7), (2, 3, 3))
    Aw = rnd(wg.rows, K, seed=8, dtype=BF16)
    xw = rnd(wg.tokens, N, seed=9)

    def run_all():
        outs = []
        o1, p1 = torch.empty(M, N, dtype=BF16, device="cuda"), torch.empty(M, N, dtype=BF16, device="cuda")
        ops.gemm(A, B, o1, bias=bias, act="gelu", out_pre=p1); outs += [o1, p1]                  # SPEC 1
        o2 = torch.empty(M, N, dtype=BF16, device="cuda")
        ops.gemm(A, Bt, o2, b_t=True, gelu_pre=gder); outs.append(o2)                           # SPEC 2
        o3 = torch.empty(M, N, dtype=BF16, device="cuda")
        ops.gemm(A, B, o3, bias=bias, scale_cols=128, scale=0.125); outs.append(o3)                                # SPEC 3
        o4 = torch.empty(M, N, dtype=BF16, device="cuda")
        ops.gemm(A, Bt, o4, b_t=True); outs.append(o4)                                                            # SPEC 4
        o5 = torch.empty(M, N, dtype=F32, device="cuda")
        ops.gemm(A, B, o5, bias=bias, residual=res, row_scale=rs, row_scale_rows=250); outs.append(o5)             # SPEC 5
        o6 = torch.empty(wg.tokens, N, dtype=F32, device="cuda")
        ops.gemm(Aw, B, o6, bias=bias, residual=xw, window=wg, row_scale=rs[:2].contiguous(), row_scale_rows=wg.rows // 2)
        outs.append(o6)                                                                                            # SPEC 6
        return outs
    try:
        ops.set_tunable("gemm_spec", 0)
        ref = run_all()
        ops.set_tunable("gemm_spec", 1)
        got = run_all()
    finally:
        ops.set_tunable("gemm_spec", -1)
    for i, (g, r) in enumerate(zip(got, ref)):
        assert torch.equal(g, r), (i, rel(g, r))
    assert rel(got[5], (A.float() @ B.float().t() + bias) * rs.repeat_interleave(250)[:, None] + res) < 2e-3


def test_gemm_fc2_dgrad_tma_box_epilogue(ops):
    """fc2 dgrad on 128 x 256 tiles: the epilogue that moves the pre-activation rows in and the results out as swizzled TMA boxes
    against the per-lane epilogue (bit-identical), incl. a ragged last row tile, and against the fp32 reference."""
    for M in (40000, 40000 + 77):
        N, K = 512, 128
        A, Bt = rnd(M, K, seed=3, dtype=BF16), rnd(K, N, seed=4, scale=0.1, dtype=BF16)
        pre = rnd(M, N, seed=7, dtype=BF16)
        outs = []
        try:
            for mode in (0, 1):
                ops.set_tunable("gemm_box", mode)
                o = torch.full((M, N), float("nan"), dtype=BF16, device="cuda")
                ops.gemm(A, Bt, o, b_t=True, gelu_pre=pre)
                outs.append(o)
        finally:
            ops.set_tunable("gemm_box", -1)
        assert torch.equal(outs[0], outs[1]), rel(outs[1], outs[0])
        p = pre.float().cpu().requires_grad_(True)
        O.gelu(p).sum().backward()
        assert rel(outs[1], (A.float() @ Bt.float()).cpu() * p.grad) < 1e-2
        # bias + q-scale (qkv) and plain dgrad epilogues through the same boxes (gemm_box = 2) vs the per-lane stores
        Bk, bias = rnd(N, K, seed=5, scale=0.1, dtype=BF16), rnd(N, seed=6)
        res = []
        try:
            for mode in (0, 2):
                ops.set_tunable("gemm_box", mode)
                o1 = torch.full((M, N), float("nan"), dtype=BF16, device="cuda")
                ops.gemm(A, Bk, o1, bias=bias, scale_cols=128, scale=0.125)
                o2 = torch.full((M, N), float("nan"), dtype=BF16, device="cuda")
                ops.gemm(A, Bt, o2, b_t=True)
                res.append((o1, o2))
        finally:
            ops.set_tunable("gemm_box", -1)
        assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])


def test_gemm_window_scatter(ops):
    # proj GEMM + window_reverse + roll back + residual, with spatial padding
    B_, D, H, W, Cc = 2, 4, 10, 9, 64
    win, sh = O.get_window_size((D, H, W), (8, 7, 7), (4, 3, 3))
    wg = ops.Window(B_, D, H, W, win, sh)
    A = rnd(wg.rows, 64, seed=7, dtype=BF16)
    Wt = rnd(Cc, 64, seed=8, scale=0.1, dtype=BF16)
    x = rnd(wg.tokens, Cc, seed=9)
    out = torch.empty(wg.tokens, Cc, dtype=F32, device="cuda")
    ops.gemm(A, Wt, out, residual=x, window=wg)
    gi = torch.from_numpy(O.window_gather_index(B_, wg.Dp, wg.Hp, wg.Wp, win, sh))
    y = (A.float() @ Wt.float().t()).cpu()
    back = torch.zeros(B_ * wg.Dp * wg.Hp * wg.Wp, Cc).index_copy(0, gi.reshape(-1), y)
    back = back.view(B_, wg.Dp, wg.Hp, wg.Wp, Cc)[:, :D, :H, :W].reshape(-1, Cc)
    assert rel(out, back + x.cpu()) < 2e-3


@pytest.mark.parametrize("splits", [2, 7, 64])
def test_gemm_splitk_wgrad(ops, splits):
    T, Nout, Nin = 5000, 384, 128
    dY, X = rnd(T, Nout, seed=10, dtype=BF16), rnd(T, Nin, seed=11, dtype=BF16)
    dW = torch.empty(Nout, Nin, dtype=F32, device="cuda")
    ops.gemm(dY, X, dW, a_t=True, b_t=True, k_splits=splits)
    ref = dY.float().t() @ X.float()
    assert rel(dW, ref) < 2e-3
    ops.gemm(dY, X, dW, a_t=True, b_t=True, k_splits=splits, accumulate=True)
    assert rel(dW, 2 * ref) < 2e-3
    # bias gradient from the same GEMM: row sums of the A operand (= column sums of dY) by one extra MMA per K step
    db = torch.zeros(Nout, dtype=F32, device="cuda")
    dW2 = torch.empty_like(dW)
    ops.gemm(dY, X, dW2, a_t=True, b_t=True, k_splits=splits, rowsum=db)
    assert rel(dW2, ref) < 2e-3 and rel(db, dY.float().sum(0)) < 1e-5


@pytest.mark.parametrize("M,N,K,splits", [(2048, 512, 100352, 4), (512, 2048, 50176, 4), (200, 136, 333, 1), (1536, 512, 6000, 6),
                                          (96, 3 * 2 * 4 * 4, 25088, 8), (768, 768, 4096, 8), (336, 256, 9000, 2)])
def test_gemm_rowsum_shapes(ops, M, N, K, splits):
    """Row sums of the A operand (the bias gradient) on the production weight-gradient shapes and ragged M / K tails, with
    128x128 and 128x256 tiles; a K-major A is rejected loudly."""
    dY, X = rnd(K, M, seed=12, dtype=BF16), rnd(K, N, seed=13, dtype=BF16)
    dW = torch.empty(M, N, dtype=F32, device="cuda")
    db = torch.zeros(M, dtype=F32, device="cuda")
    ops.gemm(dY, X, dW, a_t=True, b_t=True, k_splits=splits, rowsum=db)
    assert rel(dW, dY.float().t() @ X.float()) < 3e-3
    assert rel(db, dY.float().sum(0)) < 1e-4
    if K % 8 == 0:
        with pytest.raises(RuntimeError):                            # K-major A: not a weight gradient
            ops.gemm(dY.t().contiguous(), X, torch.empty(M, N, dtype=F32, device="cuda"), b_t=True, rowsum=db)


# ------------------------------------------------------------------------------------------------ LayerNorm
def _ln_ref(x, g, b, eps):
    return torch.nn.functional.layer_norm(x, (x.shape[-1],), g, b, eps)


@pytest.mark.parametrize("C", [96, 128, 768, 1536, 2048])
def test_layernorm_plain(ops, C):
    rows = 333
    x, g, b = rnd(rows, C, seed=1, scale=3), 1 + 0.1 * rnd(C, seed=2), 0.1 * rnd(C, seed=3)
    y = torch.empty(rows, C, dtype=F32, device="cuda")
    mean, rstd = torch.empty(rows, device="cuda"), torch.empty(rows, device="cuda")
    ops.layernorm_fwd(x, g, b, 1e-5, y, mean=mean, rstd=rstd)
    xr = x.cpu().requires_grad_(True); gr = g.cpu().requires_grad_(True); br = b.cpu().requires_grad_(True)
    yr = _ln_ref(xr, gr, br, 1e-5)
    assert rel(y, yr.detach()) < 1e-5
    yb = torch.empty(rows, C, dtype=BF16, device="cuda")
    ops.layernorm_fwd(x, g, b, 1e-5, yb)
    assert rel(yb, yr.detach()) < 1e-2
    dy = rnd(rows, C, seed=4)
    dres = rnd(rows, C, seed=5)
    (yr * dy.cpu()).sum().backward()
    dx = torch.empty(rows, C, dtype=F32, device="cuda")
    dxc = torch.empty(rows, C, dtype=BF16, device="cuda")
    dg, db = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    ops.layernorm_bwd(x, g, b, 1e-5, mean, rstd, dy, rows=rows, dx=dx, dres=dres, dx_copy=dxc, dgamma=dg, dbeta=db)
    assert rel(dx, xr.grad + dres.cpu()) < 1e-4
    assert rel(dxc, xr.grad + dres.cpu()) < 1e-2
    assert rel(dg, gr.grad) < 1e-4 and rel(db, br.grad) < 1e-4


@pytest.mark.parametrize("C", [96, 128, 192, 256, 384, 512, 768, 1024])
@pytest.mark.parametrize("xdt", [F32, BF16])
def test_lnr_plain(ops, C, xdt):
    """Row-mapped LayerNorm without a map == nn.LayerNorm on dense rows (fp32 / bf16 activations)."""
    assert ops.lnr_supported(C)
    rows = 1003
    x32, g, b = rnd(rows, C, seed=1, scale=3), 1 + 0.1 * rnd(C, seed=2), 0.1 * rnd(C, seed=3)
    x = x32.to(xdt)
    y = torch.empty(rows, C, dtype=F32, device="cuda")
    mean, rstd = torch.empty(rows, device="cuda"), torch.empty(rows, device="cuda")
    ops.lnr_fwd(x, g, b, 1e-5, y, mean=mean, rstd=rstd)
    xr = x.float().cpu().requires_grad_(True); gr = g.cpu().requires_grad_(True); br = b.cpu().requires_grad_(True)
    yr = _ln_ref(xr, gr, br, 1e-5)
    assert rel(y, yr.detach()) < 1e-5
    yb = torch.empty(rows, C, dtype=BF16, device="cuda")
    ops.lnr_fwd(x, g, b, 1e-5, yb)
    assert rel(yb, yr.detach()) < 1e-2
    dy, dres = rnd(rows, C, seed=4), rnd(rows, C, seed=5)
    (yr * dy.cpu()).sum().backward()
    dx = torch.empty(rows, C, dtype=F32, device="cuda")
    dxc = torch.empty(rows, C, dtype=BF16, device="cuda")
    dg, db = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    ops.lnr_bwd(x, g, b, 1e-5, mean, rstd, dy, dx=dx, dres=dres, dx_bf16=dxc, dgamma=dg, dbeta=db)
    assert rel(dx, xr.grad + dres.cpu()) < 1e-4
    assert rel(dxc, xr.grad + dres.cpu()) < 1e-2
    assert rel(dg, gr.grad) < 1e-4 and rel(db, br.grad) < 1e-4
    # bf16 dy, in-place residual accumulation (dres aliases dx), no parameter gradients
    dyb = dy.to(BF16)
    acc = dres.clone()
    ops.lnr_bwd(x, g, b, 1e-5, mean, rstd, dyb, dx=acc, dres=acc)
    xr.grad = None
    (_ln_ref(xr, g.cpu(), b.cpu(), 1e-5) * dyb.float().cpu()).sum().backward()
    assert rel(acc, xr.grad + dres.cpu()) < 1e-4


@pytest.mark.parametrize("dims,Bc,C,shifted", [((4, 14, 14), 3, 128, True), ((8, 14, 7), 2, 256, True),
                                               ((4, 7, 7), 5, 512, False), ((2, 14, 14), 2, 96, True)])
def test_lnr_window_map(ops, dims, Bc, C, shifted):
    """norm1 / norm2 of a Swin block through the int32 row map: forward writes window order, backward reads dy in
    window order (norm1) or emits the window-ordered bf16 copy of dx (norm2); bit-exact placement."""
    D, H, W = dims
    win, sh = O.get_window_size(dims, (8, 7, 7), (4, 3, 3) if shifted else (0, 0, 0))
    wg = ops.Window(Bc, D, H, W, win, sh)
    assert not wg.padded
    rmap = wg.row_map("cuda")
    gi = torch.from_numpy(O.window_gather_index(Bc, D, H, W, win, sh)).reshape(-1)      # window row -> source row
    T = wg.tokens
    x, g, b = rnd(T, C, seed=1, scale=2), 1 + 0.1 * rnd(C, seed=2), 0.1 * rnd(C, seed=3)
    y = torch.empty(T, C, dtype=F32, device="cuda")
    mean, rstd = torch.empty(T, device="cuda"), torch.empty(T, device="cuda")
    ops.lnr_fwd(x, g, b, 1e-5, y, mean=mean, rstd=rstd, row_map=rmap, y_mapped=True)
    xr = x.cpu().requires_grad_(True)
    h = _ln_ref(xr, g.cpu(), b.cpu(), 1e-5)
    yr = h[gi]
    assert rel(y, yr.detach()) < 1e-5
    # the permutation itself is index work: bit-exact against an unmapped run
    y0 = torch.empty(T, C, dtype=F32, device="cuda")
    ops.lnr_fwd(x, g, b, 1e-5, y0)
    assert torch.equal(y.cpu(), y0.cpu()[gi])
    # norm1 backward: dy in window order, accumulated onto the residual gradient in place
    dy, dres = rnd(T, C, seed=4), rnd(T, C, seed=5)
    acc = dres.clone()
    c16 = torch.empty(T, C, dtype=BF16, device="cuda")
    dg, db = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    dyb = dy.to(BF16)
    xs = torch.zeros(C, device="cuda")
    ops.lnr_bwd(x, g, b, 1e-5, mean, rstd, dyb, dx=acc, dres=acc, dx_bf16=c16, row_map=rmap, dy_mapped=True, dgamma=dg, dbeta=db,
                dxsum=xs)
    xr.grad = None
    (yr * dyb.float().cpu()).sum().backward()
    want = xr.grad + dres.cpu()
    assert rel(acc, want) < 1e-4
    assert torch.equal(c16.cpu(), acc.cpu().to(BF16))
    assert rel(xs, want.sum(0)) < 1e-4                      # fused column sums (bias gradient of the producing nn.Linear)
    # norm2 backward: dense dy, bf16 copy of dx emitted in window order
    dx = torch.empty(T, C, dtype=F32, device="cuda")
    cw = torch.empty(T, C, dtype=BF16, device="cuda")
    ops.lnr_bwd(x, g, b, 1e-5, mean, rstd, dy, dx=dx, dres=dres, dx_bf16=cw, row_map=rmap, dx_bf16_mapped=True,
                dgamma=torch.zeros(C, device="cuda"), dbeta=torch.zeros(C, device="cuda"))
    assert torch.equal(cw.cpu(), dx.cpu().to(BF16)[gi])


@pytest.mark.parametrize("dims,Bc", [((4, 14, 14), 2), ((16, 14, 7), 1), ((3, 10, 9), 2)])
def test_layernorm_window_gather(ops, dims, Bc):
    D, H, W = dims
    C = 64
    win, sh = O.get_window_size(dims, (8, 7, 7), (4, 3, 3))
    wg = ops.Window(Bc, D, H, W, win, sh)
    x, g, b = rnd(wg.tokens, C, seed=1, scale=2), 1 + 0.1 * rnd(C, seed=2), 0.1 * rnd(C, seed=3)
    y = torch.empty(wg.rows, C, dtype=F32, device="cuda")
    mean, rstd = torch.empty(wg.rows, device="cuda"), torch.empty(wg.rows, device="cuda")
    ops.layernorm_fwd(x, g, b, 1e-5, y, mean=mean, rstd=rstd, window=wg)
    xr = x.cpu().requires_grad_(True)
    h = _ln_ref(xr, g.cpu(), b.cpu(), 1e-5).view(Bc, D, H, W, C)
    h = torch.nn.functional.pad(h, (0, 0, 0, wg.Wp - W, 0, wg.Hp - H, 0, wg.Dp - D))
    gi = torch.from_numpy(O.window_gather_index(Bc, wg.Dp, wg.Hp, wg.Wp, win, sh)).reshape(-1)
    yr = h.reshape(-1, C)[gi]
    assert rel(y, yr.detach()) < 1e-5
    # backward: scatter + residual gradient + window-ordered bf16 copy of the result
    dy, dres = rnd(wg.rows, C, seed=4), rnd(wg.tokens, C, seed=5)
    (yr * dy.cpu()).sum().backward()
    dx = torch.empty(wg.tokens, C, dtype=F32, device="cuda")
    dg, db = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    ops.layernorm_bwd(x, g, b, 1e-5, mean, rstd, dy, rows=wg.rows, dx=dx, dres=dres, dgamma=dg, dbeta=db, window=wg)
    assert rel(dx, xr.grad + dres.cpu()) < 1e-4
    # plain LN backward whose dx copy is emitted in window order (LN2 -> proj operand)
    y2 = torch.empty(wg.tokens, C, dtype=F32, device="cuda")
    m2, r2 = torch.empty(wg.tokens, device="cuda"), torch.empty(wg.tokens, device="cuda")
    ops.layernorm_fwd(x, g, b, 1e-5, y2, mean=m2, rstd=r2)
    dy2 = rnd(wg.tokens, C, seed=6)
    copy = torch.zeros(wg.rows, C, dtype=F32, device="cuda")
    dx2 = torch.empty(wg.tokens, C, dtype=F32, device="cuda")
    ops.layernorm_bwd(x, g, b, 1e-5, m2, r2, dy2, rows=wg.tokens, dx=dx2, dx_copy=copy, copy_window=wg)
    padded = torch.nn.functional.pad(dx2.cpu().view(Bc, D, H, W, C), (0, 0, 0, wg.Wp - W, 0, wg.Hp - H, 0, wg.Dp - D))
    assert torch.equal(copy.cpu(), padded.reshape(-1, C)[gi])


@pytest.mark.parametrize("H,W", [(8, 6), (7, 5)])
def test_layernorm_merge(ops, H, W):
    Bc, D, C = 2, 3, 32
    x = rnd(Bc * D * H * W, C, seed=1, scale=2)
    g, b = 1 + 0.1 * rnd(4 * C, seed=2), 0.1 * rnd(4 * C, seed=3)
    H2, W2 = (H + 1) // 2, (W + 1) // 2
    rows = Bc * D * H2 * W2
    y = torch.empty(rows, 4 * C, dtype=F32, device="cuda")
    mean, rstd = torch.empty(rows, device="cuda"), torch.empty(rows, device="cuda")
    ops.layernorm_fwd(x, g, b, 1e-5, y, mean=mean, rstd=rstd, merge=(Bc, D, H, W, C))
    xr = x.cpu().requires_grad_(True)
    xx = torch.nn.functional.pad(xr.view(Bc, D, H, W, C), (0, 0, 0, W % 2, 0, H % 2))
    cat = torch.cat([xx[:, :, i::2, j::2] for (i, j) in ((0, 0), (1, 0), (0, 1), (1, 1))], -1).reshape(rows, 4 * C)
    yr = _ln_ref(cat, g.cpu(), b.cpu(), 1e-5)
    assert rel(y, yr.detach()) < 1e-5
    dy = rnd(rows, 4 * C, seed=4)
    (yr * dy.cpu()).sum().backward()
    dx = torch.zeros(Bc * D * H * W, C, dtype=F32, device="cuda")
    dg, db = torch.zeros(4 * C, device="cuda"), torch.zeros(4 * C, device="cuda")
    ops.layernorm_bwd(x, g, b, 1e-5, mean, rstd, dy, rows=rows, dx=dx, dgamma=dg, dbeta=db, merge=(Bc, D, H, W, C))
    assert rel(dx, xr.grad) < 1e-4


@pytest.mark.parametrize("C,H,W", [(128, 8, 6), (128, 7, 5), (256, 6, 6), (256, 5, 7), (512, 4, 6), (512, 3, 5)])
def test_layernorm_merge_swin_widths(ops, C, H, W):
    """PatchMerging LN at the Swin widths (4C = 512 / 1024 / 2048, fp32 tokens in, bf16 out): the dedicated kernels
    (T = C/4 threads per merged row) against the fp32 reference incl. odd H / W padding, dgamma / dbeta and several
    grid-stride iterations (rows > rows per wave is forced by the small clip count only at C = 512; larger below)."""
    Bc, D = 3, 5
    x = rnd(Bc * D * H * W, C, seed=1, scale=2)
    g, b = 1 + 0.1 * rnd(4 * C, seed=2), 0.1 * rnd(4 * C, seed=3)
    H2, W2 = (H + 1) // 2, (W + 1) // 2
    rows = Bc * D * H2 * W2
    y = torch.empty(rows, 4 * C, dtype=BF16, device="cuda")
    mean, rstd = torch.empty(rows, device="cuda"), torch.empty(rows, device="cuda")
    ops.layernorm_fwd(x, g, b, 1e-5, y, mean=mean, rstd=rstd, merge=(Bc, D, H, W, C))
    xr = x.cpu().requires_grad_(True)
    gr, br = g.cpu().requires_grad_(True), b.cpu().requires_grad_(True)
    xx = torch.nn.functional.pad(xr.view(Bc, D, H, W, C), (0, 0, 0, W % 2, 0, H % 2))
    cat = torch.cat([xx[:, :, i::2, j::2] for (i, j) in ((0, 0), (1, 0), (0, 1), (1, 1))], -1).reshape(rows, 4 * C)
    yr = _ln_ref(cat, gr, br, 1e-5)
    assert rel(y, yr.detach()) < 4e-3                                           # bf16 rounding of the output only
    assert rel(mean, cat.detach().mean(-1)) < 1e-5
    dy = rnd(rows, 4 * C, seed=4, dtype=BF16)
    (yr * dy.float().cpu()).sum().backward()
    dx = torch.full((Bc * D * H * W, C), float("nan"), dtype=F32, device="cuda")  # every source row must be written
    dg, db = torch.zeros(4 * C, device="cuda"), torch.zeros(4 * C, device="cuda")
    ops.layernorm_bwd(x, g, b, 1e-5, mean, rstd, dy, rows=rows, dx=dx, dgamma=dg, dbeta=db, merge=(Bc, D, H, W, C))
    assert rel(dx, xr.grad) < 1e-4
    assert rel(dg, gr.grad) < 1e-4 and rel(db, br.grad) < 1e-4
    # the same call also emitting the bf16 copy of dx times a per-clip factor and the column sums of that copy
    scale = (0.5 + torch.arange(Bc, dtype=F32)).cuda()
    dx2 = torch.empty_like(dx)
    dx16 = torch.full((Bc * D * H * W, C), float("nan"), dtype=BF16, device="cuda")
    dsum = torch.zeros(C, device="cuda")
    dg2, db2 = torch.zeros(4 * C, device="cuda"), torch.zeros(4 * C, device="cuda")
    ops.layernorm_bwd(x, g, b, 1e-5, mean, rstd, dy, rows=rows, dx=dx2, dgamma=dg2, dbeta=db2, merge=(Bc, D, H, W, C),
                      dx_copy=dx16, dxsum=dsum, copy_scale=scale, copy_scale_rows=D * H * W)
    want = xr.grad.view(Bc, -1, C) * scale.cpu()[:, None, None]
    assert torch.equal(dx2, dx) and rel(dx16, want.reshape(-1, C)) < 4e-3
    assert rel(dsum, want.reshape(-1, C).sum(0)) < 1e-4


def test_layernorm_merge_many_rows(ops):
    """More merged rows than one wave of CTAs covers (grid-stride loop, shared-memory parity buffers)."""
    Bc, D, H, W, C = 8, 4, 28, 28, 256
    x = rnd(Bc * D * H * W, C, seed=1, scale=2)
    g, b = 1 + 0.1 * rnd(4 * C, seed=2), 0.1 * rnd(4 * C, seed=3)
    rows = Bc * D * (H // 2) * (W // 2)
    y = torch.empty(rows, 4 * C, dtype=BF16, device="cuda")
    mean, rstd = torch.empty(rows, device="cuda"), torch.empty(rows, device="cuda")
    ops.layernorm_fwd(x, g, b, 1e-5, y, mean=mean, rstd=rstd, merge=(Bc, D, H, W, C))
    xr = x.requires_grad_(True)
    cat = torch.cat([xr.view(Bc, D, H, W, C)[:, :, i::2, j::2] for (i, j) in ((0, 0), (1, 0), (0, 1), (1, 1))], -1).reshape(rows, 4 * C)
    yr = torch.nn.functional.layer_norm(cat, (4 * C,), g, b, 1e-5)
    assert rel(y, yr.detach()) < 4e-3
    dy = rnd(rows, 4 * C, seed=4, dtype=BF16)
    (yr * dy.float()).sum().backward()
    dx = torch.full((Bc * D * H * W, C), float("nan"), dtype=F32, device="cuda")
    dg, db = torch.zeros(4 * C, device="cuda"), torch.zeros(4 * C, device="cuda")
    ops.layernorm_bwd(x.detach(), g, b, 1e-5, mean, rstd, dy, rows=rows, dx=dx, dgamma=dg, dbeta=db, merge=(Bc, D, H, W, C))
    assert rel(dx, xr.grad) < 1e-4
    gg, gb = torch.autograd.grad((torch.nn.functional.layer_norm(cat.detach(), (4 * C,), g.requires_grad_(True), b.requires_grad_(True), 1e-5)
                                  * dy.float()).sum(), (g, b))
    assert rel(dg, gg) < 2e-4 and rel(db, gb) < 2e-4


def test_lnr_mask_token_blend(ops):
    """Patch-embed LayerNorm + SimMIM mask-token blend on the lean row kernels (C = 128, fp32 in / out, per-row weights):
    y = LN(x) (1 - w) + token w; backward returns d x (+ bf16 copy), dgamma / dbeta and d token."""
    rows, C = 5000, 128
    x = rnd(rows, C, seed=1, scale=2)
    g, b, tok = 1 + 0.1 * rnd(C, seed=2), 0.1 * rnd(C, seed=3), rnd(C, seed=4)
    w = (torch.rand(rows, generator=torch.Generator().manual_seed(5)) < 0.4).float().cuda()
    y = torch.empty_like(x)
    mean, rstd = torch.empty(rows, device="cuda"), torch.empty(rows, device="cuda")
    ops.lnr_fwd(x, g, b, 1e-5, y, mean=mean, rstd=rstd, blend=(w, tok))
    xr, gr, br, tr = (t.detach().clone().requires_grad_(True) for t in (x, g, b, tok))
    yr = torch.nn.functional.layer_norm(xr, (C,), gr, br, 1e-5) * (1 - w[:, None]) + tr[None] * w[:, None]
    assert rel(y, yr.detach()) < 1e-5
    dy = rnd(rows, C, seed=6)
    (yr * dy).sum().backward()
    dx = torch.empty_like(x)
    dx16 = torch.empty(rows, C, dtype=BF16, device="cuda")
    small = torch.zeros(3 * C, device="cuda")
    ops.lnr_bwd(x, g, b, 1e-5, mean, rstd, dy, dx=dx, dx_bf16=dx16, dgamma=small[:C], dbeta=small[C:2 * C], blend=(w, tok),
                dtoken=small[2 * C:])
    assert rel(dx, xr.grad) < 1e-4 and rel(dx16, xr.grad) < 4e-3
    assert rel(small[:C], gr.grad) < 1e-4 and rel(small[C:2 * C], br.grad) < 1e-4 and rel(small[2 * C:], tr.grad) < 1e-4


def test_layernorm_fusion_adds_and_blend_and_lookup(ops):
    # fusion: LN(v + space[s] + tempor[t] + type0) written into rows of the concat buffer
    Bc, T, S, C, L = 3, 2, 49, 128, 16
    v = rnd(Bc * T * S, C, seed=1, dtype=BF16)
    space, tempor, type0 = rnd(S, C, seed=2), rnd(T, C, seed=3), rnd(C, seed=4)
    g, b = 1 + 0.1 * rnd(C, seed=5), 0.1 * rnd(C, seed=6)
    tot = T * S + L
    z = torch.zeros(Bc * tot, C, dtype=F32, device="cuda")
    ops.layernorm_fwd(v, g, b, 1e-5, z, rows=Bc * T * S, add0=type0, add1=(space, 1, S), add2=(tempor, S, T),
                      group=(T * S, tot, 0))
    ref = v.float().cpu().view(Bc, T, S, C) + space.cpu()[None, None] + tempor.cpu()[None, :, None] + type0.cpu()
    ref = _ln_ref(ref.view(Bc, T * S, C), g.cpu(), b.cpu(), 1e-5)
    assert rel(z.view(Bc, tot, C)[:, :T * S], ref) < 1e-5
    assert float(z.view(Bc, tot, C)[:, T * S:].abs().max()) == 0.0
    # mask-token blend (patch-embed LN epilogue)
    Bc, D, H, W, C = 2, 2, 14, 14, 32
    x = rnd(Bc * D * H * W, C, seed=7)
    g, b, tok = 1 + 0.1 * rnd(C, seed=8), 0.1 * rnd(C, seed=9), rnd(C, seed=10)
    mask = (torch.rand(Bc, 7, 7, generator=torch.Generator().manual_seed(3)) < 0.3).long().cuda()
    y = torch.empty_like(x)
    mean, rstd = torch.empty(x.shape[0], device="cuda"), torch.empty(x.shape[0], device="cuda")
    ops.layernorm_fwd(x, g, b, 1e-5, y, mean=mean, rstd=rstd, blend=(mask, tok, (D, H, W)))
    st = {"mask_token": tok.cpu().view(1, C, 1, 1, 1)}
    yr, _ = O.mask_token_blend(st, _ln_ref(x.cpu(), g.cpu(), b.cpu(), 1e-5).view(Bc, D, H, W, C), mask.cpu()[:, None])
    assert rel(y, yr.reshape(-1, C)) < 1e-5
    # embedding lookup + position add (HF BertEmbeddings)
    V, H_, Lq, Bq = 50, 64, 8, 4
    table, pos, typ = rnd(V, H_, seed=11), rnd(Lq, H_, seed=12), rnd(H_, seed=13)
    ids = torch.randint(0, V, (Bq * Lq,), generator=torch.Generator().manual_seed(5)).cuda()
    g, b = 1 + 0.1 * rnd(H_, seed=14), 0.1 * rnd(H_, seed=15)
    y = torch.empty(Bq * Lq, H_, dtype=F32, device="cuda")
    ops.layernorm_fwd(table, g, b, 1e-12, y, rows=Bq * Lq, row_index=ids, add0=typ, add1=(pos, 1, Lq))
    ref = table.cpu()[ids.cpu()].view(Bq, Lq, H_) + pos.cpu()[None] + typ.cpu()
    assert rel(y, _ln_ref(ref, g.cpu(), b.cpu(), 1e-12).view(-1, H_)) < 1e-5


# ------------------------------------------------------------------------------------------------ attention
def _attn_inputs(batch, seq, heads, hd, seed):
    qkv = rnd(batch * seq, 3 * heads * hd, seed=seed, scale=0.7, dtype=BF16)
    dout = rnd(batch * seq, heads * hd, seed=seed + 1, dtype=BF16)
    return qkv, dout


def _attn_ref(qkv, dout, batch, seq, heads, hd, bias=None):
    """fp32 reference on CPU with autograd.  bias: (batch|1, heads|1, seq, seq) additive or None."""
    x = qkv.float().cpu().requires_grad_(True)
    q, k, v = x.view(batch, seq, 3, heads, hd).permute(2, 0, 3, 1, 4)
    s = q @ k.transpose(-1, -2)
    if bias is not None:
        s = s + bias
    p = torch.softmax(s, -1)
    o = (p @ v).transpose(1, 2).reshape(batch * seq, heads * hd)
    lse = torch.logsumexp(s, -1)
    return x, o, lse


@pytest.mark.parametrize("dims,shifted,heads", [((4, 14, 14), False, 3), ((4, 14, 14), True, 3), ((8, 14, 7), True, 3),
                                               ((16, 7, 7), True, 3), ((4, 14, 14), True, 4), ((4, 14, 7), True, 8),
                                               ((4, 7, 7), False, 16), ((4, 7, 7), False, 32)])
def test_window_attention_core(ops, dims, shifted, heads):
    hd, Bc = 32, 2
    win, sh = O.get_window_size(dims, (8, 7, 7), (4, 3, 3) if shifted else (0, 0, 0))
    N = win[0] * win[1] * win[2]
    nwin = (dims[0] // win[0]) * (dims[1] // win[1]) * (dims[2] // win[2])
    batch = Bc * nwin
    qkv, dout = _attn_inputs(batch, N, heads, hd, 20)
    table = rnd(2535, heads, seed=22, scale=0.5)
    from clover_b200.tables import rel_code, region_ids
    code, off = rel_code(N, (8, 7, 7))
    code = torch.from_numpy(code).cuda()
    region = torch.from_numpy(region_ids(*dims, win, sh)).cuda() if shifted else None
    out = torch.empty(batch * N, heads * hd, dtype=BF16, device="cuda")
    lse = torch.empty(batch, heads, N, dtype=F32, device="cuda")
    kw = dict(bias_table=table, rel_code=code, code_off=off, region=region)
    ops.attention_fwd(qkv, batch, N, heads, hd, out, lse, **kw)
    idx = torch.from_numpy(O.relative_position_index((8, 7, 7))[:N, :N].reshape(-1))
    tb = table.cpu().requires_grad_(True)
    bias = tb[idx].view(N, N, heads).permute(2, 0, 1)[None]
    if shifted:
        m = torch.from_numpy(O.compute_mask(*dims, win, sh))            # (nwin, N, N)
        bias = bias + m.repeat(Bc, 1, 1)[:, None]
    x, o_ref, lse_ref = _attn_ref(qkv, dout, batch, N, heads, hd, bias)
    assert rel(out, o_ref.detach()) < 1e-2
    assert rel(lse, lse_ref.detach()) < 1e-4
    (o_ref * dout.float().cpu()).sum().backward()
    dqkv = torch.empty_like(qkv)
    dtab = torch.zeros(2535, heads, dtype=F32, device="cuda")
    ops.attention_bwd(qkv, out, dout, lse, batch, N, heads, hd, dqkv, 1.0, dbias_table=dtab, **kw)
    assert rel(dqkv, x.grad) < 2e-2
    assert rel(dtab, tb.grad) < 2e-2


@pytest.mark.parametrize("dims,shifted,Bc", [((4, 14, 14), False, 2), ((4, 14, 14), True, 2), ((4, 14, 14), True, 40),
                                            ((2, 14, 14), True, 3), ((8, 14, 7), True, 2), ((16, 7, 7), True, 2),
                                            ((6, 7, 7), False, 5), ((8, 14, 14), False, 20), ((8, 14, 14), True, 9)])
@pytest.mark.parametrize("heads", [3, 4, 8, 16, 32])
def test_window_attention_w7(ops, dims, shifted, Bc, heads):
    """Specialised (wd, 7, 7) window attention (attention_w7.cu): static bias gather, shift mask / lse / D folded
    into the MMA K-extension.  Same oracle and tolerances as the generic kernels.  heads 4 / 8 / 16 / 32 are the four
    Swin-B stages (production), 3 the Swin-T stage-1 width."""
    hd = 32
    if heads > 4 and (Bc > 9 or dims[0] > 8):          # keep the CPU reference of the wide cases to a few seconds
        pytest.skip("wide-head case covered at the smaller batch")
    win, sh = O.get_window_size(dims, (8, 7, 7), (4, 3, 3) if shifted else (0, 0, 0))
    N = win[0] * win[1] * win[2]
    nwin = (dims[0] // win[0]) * (dims[1] // win[1]) * (dims[2] // win[2])
    batch = Bc * nwin
    qkv, dout = _attn_inputs(batch, N, heads, hd, 30)
    table = rnd(2535, heads, seed=32, scale=0.5)
    from clover_b200 import swin
    from clover_b200.tables import rel_code, region_ids, w7_ext_tables
    code, off = rel_code(N, (8, 7, 7))
    code = torch.from_numpy(code).cuda()
    masked = any(s > 0 for s in sh)
    rid = region_ids(*dims, win, sh) if masked else None
    region = torch.from_numpy(rid).cuda() if masked else None
    spec = swin._w7_spec(dims, win, sh, (8, 7, 7), "cuda")
    assert spec is not None and spec.fwd_ok() and (spec.q_ext is not None) == masked
    if masked:
        # the K-extension reproduces compute_mask: 0 where the reference has 0, <= -100 where it has -100 (exact integers)
        q, k = w7_ext_tables(rid)
        ext = np.einsum("wik,wjk->wij", q, k)
        ref_mask = O.compute_mask(*dims, win, sh)
        assert np.array_equal(ext == 0, ref_mask == 0) and np.all(ext[ref_mask != 0] <= -100.0)
        assert np.array_equal(spec.q_ext.float().cpu().numpy(), q) and np.array_equal(spec.k_ext.float().cpu().numpy(), k)
    out = torch.empty(batch * N, heads * hd, dtype=BF16, device="cuda")
    lse = torch.empty(batch, heads, N, dtype=F32, device="cuda")
    kw = dict(bias_table=table, rel_code=code, code_off=off, region=region)
    ops.attention_fwd(qkv, batch, N, heads, hd, out, lse, w7=spec, **kw)
    idx = torch.from_numpy(O.relative_position_index((8, 7, 7))[:N, :N].reshape(-1))
    tb = table.cpu().requires_grad_(True)
    bias = tb[idx].view(N, N, heads).permute(2, 0, 1)[None]
    if masked:
        m = torch.from_numpy(O.compute_mask(*dims, win, sh))            # (nwin, N, N)
        bias = bias + m.repeat(Bc, 1, 1)[:, None]
    x, o_ref, lse_ref = _attn_ref(qkv, dout, batch, N, heads, hd, bias)
    assert rel(out, o_ref.detach()) < 1e-2
    assert rel(lse, lse_ref.detach()) < 1e-4
    # the generic tcgen05 kernel and the specialised one agree closely (same bf16 operands, same fp32 math)
    out2 = torch.empty_like(out); lse2 = torch.empty_like(lse)
    ops.attention_fwd(qkv, batch, N, heads, hd, out2, lse2, **kw)
    assert rel(out, out2) < 4e-3 and rel(lse, lse2) < 1e-5
    # ... and so do the two generations of the specialised forward (gen 1: 4 softmax warps, 98-row tiles; gen 2: 8 softmax
    # warps, tiles cut at multiples of 32 rows -- the default from 294 tokens up)
    outs = []
    for gen in (0, 1):
        ops.set_tunable("w7_fwd2", gen)
        try:
            o_, l_ = torch.empty_like(out), torch.empty_like(lse)
            ops.attention_fwd(qkv, batch, N, heads, hd, o_, l_, w7=spec, **kw)
            outs.append((o_, l_))
        finally:
            ops.set_tunable("w7_fwd2", -1)
    assert rel(outs[0][0], outs[1][0]) < 2e-3 and rel(outs[0][1], outs[1][1]) < 1e-6
    assert rel(out, outs[1][0]) < 2e-3 and rel(lse, outs[1][1]) < 1e-6
    # the two issue orders of the forward (w7_fwd_early: score MMAs of the next tile before the O read-out of the current one is
    # acknowledged) run the same arithmetic: bit-identical results, repeated to shake out ordering hazards
    # w7_fwd_pvsplit: P V products of the first key bodies issued early; w7_fwd_qtile: 98- or 128-row query tiles
    for early, split, qtile in ((0, 0, 0), (1, 0, 0), (0, 1, 1), (1, 1, 0), (1, 1, 1), (0, 0, 1)):
        ops.set_tunable("w7_fwd_early", early)
        ops.set_tunable("w7_fwd_pvsplit", split)
        ops.set_tunable("w7_fwd_qtile", qtile)
        try:
            for _ in range(3):
                o_, l_ = torch.full_like(out, float("nan")), torch.empty_like(lse)
                ops.attention_fwd(qkv, batch, N, heads, hd, o_, l_, w7=spec, **kw)
                assert torch.equal(o_, outs[0][0]) and torch.equal(l_, outs[0][1])
        finally:
            ops.set_tunable("w7_fwd_early", -1)
            ops.set_tunable("w7_fwd_pvsplit", -1)
            ops.set_tunable("w7_fwd_qtile", -1)
    (o_ref * dout.float().cpu()).sum().backward()
    dqkv = torch.empty_like(qkv)
    dtab = torch.zeros(2535, heads, dtype=F32, device="cuda")
    ops.attention_bwd(qkv, out, dout, lse, batch, N, heads, hd, dqkv, 1.0, dbias_table=dtab, w7=spec, **kw)
    assert rel(dqkv, x.grad) < 2e-2
    assert rel(dtab, tb.grad) < 2e-2
    dq2 = torch.empty_like(qkv)                                          # without the bias-table gradient
    ops.attention_bwd(qkv, out, dout, lse, batch, N, heads, hd, dq2, 1.0, w7=spec, **kw)
    assert torch.equal(dq2, dqkv)


@pytest.mark.parametrize("seq", [32, 228, 432])
def test_bert_attention_core(ops, seq):
    heads, hd, batch = 2, 64, 3
    qkv, dout = _attn_inputs(batch, seq, heads, hd, 30)
    lens = [seq, seq - 5, max(3, seq // 2)]
    keep = torch.zeros(batch, seq)
    for i, n in enumerate(lens):
        keep[i, :n] = 1
    km = ((1 - keep) * -10000.0).cuda()
    out = torch.empty(batch * seq, heads * hd, dtype=BF16, device="cuda")
    lse = torch.empty(batch, heads, seq, dtype=F32, device="cuda")
    ops.attention_fwd(qkv, batch, seq, heads, hd, out, lse, key_mask=km)
    x, o_ref, lse_ref = _attn_ref(qkv, dout, batch, seq, heads, hd, km.cpu()[:, None, None, :])
    assert rel(out, o_ref.detach()) < 1e-2
    (o_ref * dout.float().cpu()).sum().backward()
    dqkv = torch.empty_like(qkv)
    ops.attention_bwd(qkv, out, dout, lse, batch, seq, heads, hd, dqkv, 0.125, key_mask=km)
    g = x.grad.clone().view(batch * seq, 3, heads * hd)
    g[:, 0] *= 0.125
    assert rel(dqkv, g.view(batch * seq, -1)) < 2e-2


@pytest.mark.parametrize("seq,batch,heads,p", [(228, 3, 2, 0.1), (40, 5, 12, 0.1), (432, 2, 3, 0.25), (1, 2, 2, 0.0),
                                               (33, 2, 1, 0.0), (448, 1, 2, 0.0), (228, 30, 12, 0.0)])
def test_bert_attention_tc64_mask_dropout_and_legacy_agreement(ops, seq, batch, heads, p):
    """attention_h64.cu (tcgen05, head_dim 64): key mask + attention-probability dropout against the fp32 reference that
    applies the SAME keep mask (clv_keep_mask stream, index (b, h, i, j)), and against the mma.sync kernels it replaces;
    30 x 12 x two query tiles = 720 units exercises the persistent loop (more units than resident CTAs)."""
    hd = 64
    qkv, dout = _attn_inputs(batch, seq, heads, hd, 77)
    keep = torch.ones(batch, seq)
    for b in range(batch):
        keep[b, max(1, seq - 1 - (b * 7) % max(1, seq // 2)):] = 0 if seq > 2 else 1
    km = ((1 - keep) * -10000.0).cuda()
    drop = (p, 1234567, 96) if p > 0 else None
    out = torch.empty(batch * seq, heads * hd, dtype=BF16, device="cuda")
    lse = torch.empty(batch, heads, seq, dtype=F32, device="cuda")
    assert ops.USE_TC64_ATTENTION
    ops.attention_fwd(qkv, batch, seq, heads, hd, out, lse, key_mask=km, drop=drop)
    # fp32 reference with the identical keep mask
    x = qkv.float().cpu().requires_grad_(True)
    q, k, v = x.view(batch, seq, 3, heads, hd).permute(2, 0, 3, 1, 4)
    s_ = q @ k.transpose(-1, -2) + km.cpu()[:, None, None, :]
    pr = torch.softmax(s_, -1)
    if p > 0:
        kmask = ops.keep_mask(batch * heads * seq * seq, p, drop[1], drop[2], "cuda").cpu().view(batch, heads, seq, seq).float()
        pr = pr * kmask / (1.0 - p)
    o_ref = (pr @ v).transpose(1, 2).reshape(batch * seq, heads * hd)
    assert rel(out, o_ref.detach()) < 1e-2
    assert rel(lse, torch.logsumexp(s_, -1).detach()) < 1e-4
    (o_ref * dout.float().cpu()).sum().backward()
    dqkv = torch.empty_like(qkv)
    ops.attention_bwd(qkv, out, dout, lse, batch, seq, heads, hd, dqkv, 0.125, key_mask=km, drop=drop)
    g = x.grad.clone().view(batch * seq, 3, heads * hd)
    g[:, 0] *= 0.125
    assert rel(dqkv, g.view(batch * seq, -1)) < 2e-2
    # the legacy mma.sync kernels on the same inputs (same lse convention, same dropout stream)
    ops.USE_TC64_ATTENTION = False
    try:
        out2 = torch.empty_like(out); lse2 = torch.empty_like(lse); dq2 = torch.empty_like(dqkv)
        ops.attention_fwd(qkv, batch, seq, heads, hd, out2, lse2, key_mask=km, drop=drop)
        ops.attention_bwd(qkv, out2, dout, lse2, batch, seq, heads, hd, dq2, 0.125, key_mask=km, drop=drop)
    finally:
        ops.USE_TC64_ATTENTION = True
    assert rel(out, out2.float()) < 6e-3 and rel(lse, lse2) < 1e-5 and rel(dqkv, dq2.float()) < 1.5e-2


# ------------------------------------------------------------------------------------------------ elementwise
def test_cast_patchify_colsum_affine_scatter(ops):
    x = rnd(4096, seed=1)
    assert torch.equal(ops.to_bf16(x), x.to(BF16))
    imgs = rnd(2, 3, 6, 30, 28, seed=2)
    cols, (D, Hp, Wp) = ops.patchify(imgs, (2, 4, 4))
    xp = torch.nn.functional.pad(imgs.cpu(), (0, 0, 0, 2))
    ref = xp.view(2, 3, D, 2, Hp, 4, Wp, 4).permute(0, 2, 4, 6, 1, 3, 5, 7).reshape(-1, 96)
    assert torch.equal(cols.cpu(), ref.to(BF16))
    y = rnd(1000, 256, seed=3)
    out = torch.empty(1, 256, device="cuda")
    ops.grouped_colsum(y, out)
    assert rel(out[0], y.sum(0)) < 1e-5
    out = torch.empty(7, 256, device="cuda")
    ops.grouped_colsum(y, out, div=3, mod=7, scale=0.5)
    grp = (torch.arange(1000) // 3) % 7
    ref = torch.stack([y.cpu()[grp == g].sum(0) * 0.5 for g in range(7)])
    assert rel(out, ref) < 1e-5
    yb = y.to(BF16)
    ops.grouped_colsum(yb, out, div=1, mod=7)
    ref = torch.stack([yb.float().cpu()[torch.arange(1000) % 7 == g].sum(0) for g in range(7)])
    assert rel(out, ref) < 1e-5
    # text half of the fusion concat: z[b, off + l] = t[b, l] + type1
    Bc, L, tot, off, C = 3, 5, 12, 7, 64
    t, ty = rnd(Bc * L, C, seed=4, dtype=BF16), rnd(C, seed=5)
    z = torch.zeros(Bc * tot, C, dtype=BF16, device="cuda")
    ops.rows_affine(z, Bc * L, C, x=t, out_group=(L, tot, off), add0=ty)
    ref = (t.float() + ty).to(BF16).view(Bc, L, C)
    assert torch.equal(z.view(Bc, tot, C)[:, off:], ref)
    # broadcast add of a pooled gradient
    a, bv = rnd(Bc * L, C, seed=6), rnd(Bc, C, seed=7)
    o = torch.empty(Bc * L, C, device="cuda")
    ops.rows_affine(o, Bc * L, C, x=a, bvec=bv, bdiv=L, bscale=0.2)
    assert rel(o, a + 0.2 * bv.repeat_interleave(L, 0)) < 1e-6
    src = rnd(40, 64, seed=8)
    idx = torch.randint(0, 9, (40,), generator=torch.Generator().manual_seed(1)).cuda()
    dst = torch.zeros(9, 64, device="cuda")
    ops.scatter_add_rows(src, idx, dst)
    assert rel(dst, torch.zeros(9, 64).index_add_(0, idx.cpu(), src.cpu())) < 1e-5


# ------------------------------------------------------------------------------------------------ losses
@pytest.mark.parametrize("Bg,D", [(6, 24), (130, 768)])
def test_exclusive_nce_ranking(ops, Bg, D):
    embs = [rnd(Bg, D, seed=40 + i) for i in range(4)]
    embs[1] = embs[0] * 0.5 + embs[1] * 0.5           # make positives non-trivial (hinge partly active)
    losses, ws = ops.nce_rank_fwd(embs, 0.05, 5.0, True)
    cpu = [e.cpu().requires_grad_(True) for e in embs]
    d = O.exclusive_nce_ranking(*cpu, t=0.05, margin=5.0)
    assert abs(float(losses[0]) - float(d["nce_loss"])) < 1e-4 * abs(float(d["nce_loss"]))
    assert abs(float(losses[1]) - float(d["rank_t_tm_loss"])) < 1e-4 * max(1.0, abs(float(d["rank_t_tm_loss"])))
    (0.7 * d["nce_loss"] + 1.3 * d["rank_t_tm_loss"]).backward()
    g = ops.nce_rank_bwd(ws, 3, Bg, D, 0.05, True, torch.tensor([0.7], device="cuda"), torch.tensor([1.3], device="cuda"))
    for i in range(4):
        assert rel(g[i], cpu[i].grad) < 1e-3, i


def test_norm_softmax_loss(ops):
    Bg, D = 65, 768
    embs = [rnd(Bg, D, seed=50 + i) for i in range(2)]
    losses, ws = ops.nce_rank_fwd(embs, 0.05, 0.0, False)
    cpu = [e.cpu().requires_grad_(True) for e in embs]
    v = O.norm_softmax_loss(*cpu, 0.05, True)
    assert abs(float(losses[0]) - float(v)) < 1e-4 * abs(float(v))
    v.backward()
    g = ops.nce_rank_bwd(ws, 1, Bg, D, 0.05, False, torch.ones(1, device="cuda"), None)
    for i in range(2):
        assert rel(g[i], cpu[i].grad) < 1e-3


@pytest.mark.parametrize("gamma,V", [(2.0, 30522), (0.0, 1500)])
def test_softmax_focal(ops, gamma, V):
    rows, Vpad = 37, (V + 7) // 8 * 8
    logits = rnd(rows, Vpad, seed=60, scale=3)
    logits[:, V:] = -1e30
    tgt = torch.randint(0, V, (rows,), generator=torch.Generator().manual_seed(2))
    tgt[::5] = -100
    tgt = tgt.cuda()
    loss, stats, sums = ops.softmax_focal_fwd(logits, tgt, V, gamma)
    lr = logits[:, :V].cpu().requires_grad_(True)
    keep = tgt.cpu() != -100
    ref = O.softmax_focal_multiclass(lr[keep], tgt.cpu()[keep], gamma) if gamma else O.cross_entropy(lr[keep], tgt.cpu()[keep])
    assert abs(float(loss) - float(ref)) < 1e-5 * abs(float(ref))
    (ref * 1.7).backward()
    dl = torch.empty(rows, Vpad, dtype=F32, device="cuda")
    ops.softmax_focal_bwd(logits, tgt, V, gamma, stats, sums, torch.tensor([1.7], device="cuda"), dl)
    assert rel(dl[:, :V], lr.grad) < 1e-4
    assert float(dl[:, V:].abs().max()) == 0.0 if Vpad > V else True


# ------------------------------------------------------------------------------------------------ stochastic regularisers
def _rng_ref():
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import rng_ref
    return rng_ref


def test_keep_mask_bit_exact_and_dropout(ops):
    """The device stream equals its numpy restatement bit for bit (integer work); dropout = x * keep / (1 - p) + residual
    for every dtype combination, and the backward call regenerates the same mask."""
    R = _rng_ref()
    for p, seed, off, n in ((0.1, 7, 0, 4096), (0.5, 2 ** 63 + 11, 2 ** 33 + 8, 10000), (0.3, 99, 123456, 260)):
        got = ops.keep_mask(n, p, seed, off, "cuda").cpu().numpy().astype(bool)
        assert np.array_equal(got, R.keep_mask(n, p, seed, off)), (p, seed, off)
    p, seed, off = 0.1, 1234, 40
    keep = torch.from_numpy(R.keep_mask(6 * 768, p, seed, off)).view(6, 768)
    for xdt, rdt, ydt in ((BF16, BF16, F32), (BF16, None, BF16), (F32, F32, F32), (F32, None, BF16)):
        x = rnd(6, 768, seed=3, dtype=xdt)
        r = rnd(6, 768, seed=4, dtype=rdt) if rdt is not None else None
        y = ops.dropout(x, torch.empty(6, 768, dtype=ydt, device="cuda"), p, seed, off, residual=r)
        want = x.float().cpu() * keep / (1.0 - float(np.float32(p))) + (r.float().cpu() if r is not None else 0.0)
        assert rel(y.float(), want) < (1e-6 if ydt == F32 else 4e-3), (xdt, rdt, ydt)
    assert abs(float(keep.float().mean()) - 0.9) < 0.02


def test_rows_scale(ops):
    scale = torch.tensor([0.0, 1.0 / 0.7, 1.0 / 0.7], device="cuda")
    for xdt, ydt in ((F32, BF16), (BF16, BF16), (F32, F32)):
        x = rnd(3 * 50, 96, seed=8, dtype=xdt)
        y = ops.rows_scale(x, torch.empty(150, 96, dtype=ydt, device="cuda"), scale, 50)
        want = x.float().cpu().view(3, 50, 96) * scale.cpu().view(3, 1, 1)
        assert rel(y.float(), want.view(150, 96)) < (1e-6 if ydt == F32 else 4e-3)
    x = rnd(150, 96, seed=9, dtype=BF16)
    want = x.float().cpu().view(3, 50, 96) * scale.cpu().view(3, 1, 1)
    ops.rows_scale(x, x, scale, 50)                                  # in place
    assert rel(x.float(), want.view(150, 96)) < 4e-3


def test_gemm_row_scale_drop_path(ops):
    """DropPath factor in the GEMM epilogue: out = residual + factor[row / rows_per_sample] * (A W^T + b)."""
    M, N, K, per = 4 * 196, 128, 256, 196
    a, w = rnd(M, K, seed=1, dtype=BF16), rnd(N, K, seed=2, scale=0.1, dtype=BF16)
    bias, res = rnd(N, seed=3), rnd(M, N, seed=4)
    fac = torch.tensor([1 / 0.8, 0.0, 1 / 0.8, 0.0], device="cuda")
    out = torch.empty(M, N, dtype=F32, device="cuda")
    ops.gemm(a, w, out, bias=bias, residual=res, row_scale=fac, row_scale_rows=per)
    want = res.cpu() + fac.cpu().repeat_interleave(per)[:, None] * (a.float().cpu() @ w.float().cpu().T + bias.cpu())
    assert rel(out, want) < 2e-3
    assert torch.equal(out[per:2 * per], res[per:2 * per])          # a dropped sample passes the residual through exactly


@pytest.mark.parametrize("seq,p", [(32, 0.1), (228, 0.1), (432, 0.3)])
def test_bert_attention_dropout(ops, seq, p):
    """Attention-probability dropout inside the BERT attention kernels against the fp32 formula with the SAME mask
    (stream element ((b*heads+h)*seq+i)*seq+j), forward and backward."""
    R = _rng_ref()
    heads, hd, batch = 2, 64, 3
    seed, off = 4242, 1000
    qkv, dout = _attn_inputs(batch, seq, heads, hd, 31)
    keep = torch.zeros(batch, seq)
    for i, n in enumerate([seq, seq - 5, max(3, seq // 2)]):
        keep[i, :n] = 1
    km = ((1 - keep) * -10000.0).cuda()
    out = torch.empty(batch * seq, heads * hd, dtype=BF16, device="cuda")
    lse = torch.empty(batch, heads, seq, dtype=F32, device="cuda")
    ops.attention_fwd(qkv, batch, seq, heads, hd, out, lse, key_mask=km, drop=(p, seed, off))
    mask = torch.from_numpy(R.keep_mask(batch * heads * seq * seq, p, seed, off)).view(batch, heads, seq, seq).float()
    x = qkv.float().cpu().requires_grad_(True)
    q, k, v = x.view(batch, seq, 3, heads, hd).permute(2, 0, 3, 1, 4)
    s = q @ k.transpose(-1, -2) + km.cpu()[:, None, None, :]
    pr = torch.softmax(s, -1) * mask / (1.0 - float(np.float32(p)))
    o_ref = (pr @ v).transpose(1, 2).reshape(batch * seq, heads * hd)
    assert rel(out, o_ref.detach()) < 1e-2
    assert rel(lse, torch.logsumexp(s, -1).detach()) < 1e-4        # the normaliser ignores the dropout
    (o_ref * dout.float().cpu()).sum().backward()
    dqkv = torch.empty_like(qkv)
    ops.attention_bwd(qkv, out, dout, lse, batch, seq, heads, hd, dqkv, 0.125, key_mask=km, drop=(p, seed, off))
    g = x.grad.clone().view(batch * seq, 3, heads * hd)
    g[:, 0] *= 0.125
    assert rel(dqkv, g.view(batch * seq, -1)) < 2e-2


@pytest.mark.parametrize("seq,heads,hd", [(40, 2, 64), (412, 2, 64), (196, 3, 32)])
def test_attention_probs_mean(ops, seq, heads, hd):
    batch = 2
    qkv, _ = _attn_inputs(batch, seq, heads, hd, 33)
    keep = torch.ones(batch, seq)
    keep[1, seq - 7:] = 0
    km = ((1 - keep) * -10000.0).cuda()
    got = ops.attention_probs_mean(qkv, batch, seq, heads, hd, key_mask=km)
    q, k, v = qkv.float().cpu().view(batch, seq, 3, heads, hd).permute(2, 0, 3, 1, 4)
    want = torch.softmax(q @ k.transpose(-1, -2) + km.cpu()[:, None, None, :], -1).mean(1)
    assert rel(got, want) < 1e-4
    assert float((got.sum(-1) - 1).abs().max()) < 1e-4


# ------------------------------------------------------------------------------------------------ optimizer
def test_fused_adamw_vs_torch(ops):
    """clv_adamw_step against torch.optim.AdamW + clip_grad_norm_ (the reference's optimizer + grad_clip) over several
    steps: odd sizes (scalar tail path, unaligned views), per-group lr / weight decay, clipping active, bf16 refresh,
    and the device-side skip of a step with a non-finite gradient."""
    from clover_b200 import functional as Fn
    from clover_b200.optim import FusedAdamW
    torch.manual_seed(0)
    shapes = [(512, 96), (30522,), (2535, 4), (7,), (3 * 16384 + 5,)]
    base = torch.randn(sum(int(np.prod(s)) for s in shapes) + 1, device="cuda")
    ours, ref, off = [], [], 1
    for k, s in enumerate(shapes):
        n = int(np.prod(s))
        src = base[off:off + n].view(s)
        # odd k: the parameter is a view at a 4-byte-aligned (not 16-byte-aligned) address -> the kernel's scalar path
        ours.append(torch.nn.Parameter(src.detach().clone() if k % 2 == 0 else src.detach()))
        ref.append(torch.nn.Parameter(src.detach().clone()))
        off += n
    groups = lambda ps: [dict(params=[ps[0], ps[2]], lr=1e-2, weight_decay=0.05), dict(params=ps[1:2] + ps[3:], lr=3e-3, weight_decay=0.0)]
    o1 = FusedAdamW(groups(ours), betas=(0.9, 0.98), eps=1e-8, max_grad_norm=2.0)
    o2 = torch.optim.AdamW(groups(ref), betas=(0.9, 0.98), eps=1e-8)
    w16 = Fn.w16(ours[0])                                     # a cached bf16 operand copy that the optimizer must refresh
    for it in range(4):
        for a, b in zip(ours, ref):
            g = torch.randn(a.shape, device="cuda", generator=torch.Generator("cuda").manual_seed(100 + it)) * (3.0 if it % 2 else 0.2)
            a.grad, b.grad = g.clone(), g.clone()
        o1.step()
        total = torch.nn.utils.clip_grad_norm_(ref, 2.0)
        o2.step()
        nrm, skipped = o1.grad_norm()
        assert not skipped and abs(nrm - float(total)) < 1e-3 * float(total)
        for a, b in zip(ours, ref):
            assert rel(a.detach(), b.detach()) < 2e-6, (it, tuple(a.shape))
        assert Fn.w16(ours[0]) is w16 and torch.equal(w16, ours[0].detach().to(BF16))
    before = [p.detach().clone() for p in ours]
    for a in ours:
        a.grad = torch.randn_like(a)
    ours[1].grad[17] = float("inf")
    o1.step()
    nrm, skipped = o1.grad_norm()
    assert skipped and all(torch.equal(a.detach(), b) for a, b in zip(ours, before))
    assert o1.applied_steps == 4                              # the skipped step did not age the bias correction
    # checkpoint / resume: state_dict() carries the step in torch.optim.AdamW's layout, a fresh optimizer
    # resumed from it continues exactly like the torch reference resumed from ITS state -- and the two are interchangeable
    sd = o1.state_dict()
    assert all(set(st) == {"step", "exp_avg", "exp_avg_sq"} and float(st["step"]) == 4.0 for st in sd["state"].values())
    ours2 = [torch.nn.Parameter(a.detach().clone()) for a in ours]
    o3 = FusedAdamW(groups(ours2), betas=(0.9, 0.98), eps=1e-8, max_grad_norm=2.0)
    import copy
    o3.load_state_dict(copy.deepcopy(sd))                     # (load_state_dict keeps same-device tensors by reference)
    ref2 = [torch.nn.Parameter(b.detach().clone()) for b in ref]
    o4 = torch.optim.AdamW(groups(ref2), betas=(0.9, 0.98), eps=1e-8)
    o4.load_state_dict(copy.deepcopy(sd))                     # our state loads into the reference optimizer as is
    for it in range(2):
        for a, b in zip(ours2, ref2):
            g = torch.randn(a.shape, device="cuda", generator=torch.Generator("cuda").manual_seed(200 + it))
            a.grad, b.grad = g.clone(), g.clone()
        o3.step()
        torch.nn.utils.clip_grad_norm_(ref2, 2.0)
        o4.step()
        for a, b in zip(ours2, ref2):
            assert rel(a.detach(), b.detach()) < 2e-6, (it, tuple(a.shape))
    assert o3.applied_steps == 6


# ------------------------------------------------------------------------------------------------ input staging
def test_patchify_uint8_fused_normalise(ops):
    """Raw uint8 frames normalised inside the patch gather == GPUNormalize (x.float().sub_(mean).div_(std),
    utils/module_hooks.py:80-83) followed by the fp32 patch gather; odd sizes exercise the zero padding."""
    mean = torch.tensor([123.675, 116.28, 103.53], device="cuda")
    std = torch.tensor([58.395, 57.12, 57.375], device="cuda")
    for shape in ((2, 3, 8, 56, 56), (1, 3, 5, 30, 26)):
        x8 = torch.randint(0, 256, shape, dtype=torch.uint8, generator=torch.Generator().manual_seed(7)).cuda()
        xf = (x8.float() - mean.view(1, 3, 1, 1, 1)) / std.view(1, 3, 1, 1, 1)
        want, dims = ops.patchify(xf.contiguous(), (2, 4, 4))
        got, dims2 = ops.patchify(x8, (2, 4, 4), norm=(mean, 1.0 / std))
        assert dims == dims2 and got.shape == want.shape
        assert rel(got.float(), want.float()) < 4e-3              # (v - m) * (1/s) vs (v - m) / s before the bf16 rounding
    with pytest.raises(ValueError):
        ops.patchify(x8, (2, 4, 4))


# ------------------------------------------------------------------------------------------------ full-size properties (c2)
@pytest.mark.parametrize("shifted", [False, True])
def test_window_attention_c2_full_size_properties(ops, shifted):
    """BASELINE config c2 at its full size (4096 windows of N = 392 tokens, 4 heads x 32): too big for the CPU oracle, so
    the check uses size-independent properties of softmax attention.  (1) softmax rows sum to 1: with V == 1 the output
    is 1 everywhere, whatever Q, K, bias and shift mask are; (2) dV = P^T dO, so sum_j dV_j == sum_i dO_i per (window,
    head); (3) d(bias table) sums to 0 over each head (each softmax row's dS sums to 0); (4) the windows are
    independent: permuting the windows of one mask class permutes outputs and gradients bit for bit."""
    from clover_b200 import swin
    from clover_b200.tables import rel_code, region_ids
    dims, heads, hd, clips = (8, 56, 56), 4, 32, 64
    win, sh = O.get_window_size(dims, (8, 7, 7), (4, 3, 3) if shifted else (0, 0, 0))
    N, nwin = 392, 64
    batch = clips * nwin
    g = torch.Generator(device="cuda").manual_seed(5)
    qkv = (torch.randn(batch * N, 3 * heads * hd, generator=g, device="cuda") * 0.7).to(BF16)
    dout = torch.randn(batch * N, heads * hd, generator=g, device="cuda").to(BF16)
    table = torch.randn(2535, heads, generator=g, device="cuda") * 0.5
    code, off = rel_code(N, (8, 7, 7))
    code = torch.from_numpy(code).cuda()
    masked = any(s > 0 for s in sh)
    region = torch.from_numpy(region_ids(*dims, win, sh)).cuda() if masked else None
    kw = dict(bias_table=table, rel_code=code, code_off=off, region=region, w7=swin._w7_spec(dims, win, sh, (8, 7, 7), "cuda"))
    out = torch.empty(batch * N, heads * hd, dtype=BF16, device="cuda")
    lse = torch.empty(batch, heads, N, dtype=F32, device="cuda")
    qkv1 = qkv.clone()
    qkv1.view(batch * N, 3, heads * hd)[:, 2] = 1.0                     # V == 1
    ops.attention_fwd(qkv1, batch, N, heads, hd, out, lse, **kw)
    assert float((out.float() - 1).abs().max()) < 1e-2                                  # (1)
    del qkv1
    ops.attention_fwd(qkv, batch, N, heads, hd, out, lse, **kw)         # random V for the gradient properties
    dqkv = torch.empty_like(qkv)
    dtab = torch.zeros(2535, heads, dtype=F32, device="cuda")
    ops.attention_bwd(qkv, out, dout, lse, batch, N, heads, hd, dqkv, hd ** -0.5, dbias_table=dtab, **kw)
    dv = dqkv.view(batch, N, 3, heads, hd)[:, :, 2].float().sum(1)
    want = dout.view(batch, N, heads, hd).float().sum(1)
    assert rel(dv, want) < 1e-2                                                         # (2)
    assert float(dtab.sum(0).abs().max()) < 2e-2 * float(dtab.abs().sum(0).max())       # (3)
    # (5) windows are independent, so a sample of them against the fp32 oracle is a full-size numerical check: 16 windows
    # spread over the clips and over every shift-mask class (corner / edge / interior windows of the 8x8 grid)
    sample = [0, 7, 56, 63, 9, 27, 36, 62] + [nwin * c + w for c, w in ((1, 0), (5, 63), (17, 8), (31, 15), (40, 33), (50, 55), (63, 62), (63, 63))]
    idx = torch.from_numpy(O.relative_position_index((8, 7, 7))[:N, :N].reshape(-1))
    bias_ref = table.cpu()[idx].view(N, N, heads).permute(2, 0, 1)
    mask_ref = torch.from_numpy(O.compute_mask(*dims, win, sh)) if masked else None
    worst = [0.0, 0.0, 0.0]
    for wdx in sample:
        rows_w = slice(wdx * N, (wdx + 1) * N)
        b = bias_ref[None] + (mask_ref[wdx % nwin][None, None] if masked else 0.0)
        xw, o_ref, lse_ref = _attn_ref(qkv[rows_w], dout[rows_w], 1, N, heads, hd, b)
        (o_ref * dout[rows_w].float().cpu()).sum().backward()
        gq = xw.grad.clone().view(N, 3, heads * hd)
        gq[:, 0] *= hd ** -0.5                                          # the kernel returns d(unscaled q) = scale * d(q)
        worst[0] = max(worst[0], rel(out[rows_w], o_ref.detach()))
        worst[1] = max(worst[1], rel(lse[wdx], lse_ref.detach()[0]))
        worst[2] = max(worst[2], rel(dqkv[rows_w], gq.view(N, -1)))
    from conftest import record_parity
    record_parity(f"window_attention_c2_full_size_sampled_windows_{'shifted' if shifted else 'unshifted'}",
                  {"windows": len(sample), "out_rel": worst[0], "lse_rel": worst[1], "dqkv_rel": worst[2]})
    assert worst[0] < 1e-2 and worst[1] < 1e-4 and worst[2] < 2e-2, worst
    # (4) swap clip 3 and clip 17 (same window classes, row-block permutation)
    rows = nwin * N
    perm = torch.arange(batch * N, device="cuda").view(clips, rows)
    perm[[3, 17]] = perm[[17, 3]]
    perm = perm.reshape(-1)
    out2 = torch.empty_like(out); lse2 = torch.empty_like(lse)
    ops.attention_fwd(qkv[perm].contiguous(), batch, N, heads, hd, out2, lse2, **kw)
    assert torch.equal(out2, out[perm])
    dq2 = torch.empty_like(qkv)
    ops.attention_bwd(qkv[perm].contiguous(), out2, dout[perm].contiguous(), lse2, batch, N, heads, hd, dq2, hd ** -0.5, **kw)
    assert torch.equal(dq2, dqkv[perm])
