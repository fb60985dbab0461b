"""Op-level parity of every uwr CUDA kernel family against fp64 PyTorch math on the same GPU
(the oracle's functions are device-agnostic torch code, so the attention check calls
oracle.ast_oracle.window_attention directly).  Tolerance: TF32 operands, fp32 accumulate ->
relative L2 <= 1e-3 (north star); pure fp32 kernels <= 2e-5."""
import math
import os

import pytest
import torch
import torch.nn.functional as F

from conftest import ROOT, rel_l2

pytestmark = pytest.mark.gpu
TOL_TF32 = 1e-3
TOL_FP32 = 2e-5


@pytest.fixture(scope="module")
def ops():
    from uwr import ops as o
    return o


@pytest.fixture
def exact(ops):
    """tf32x3 mode: producers keep full fp32 outputs (no TF32 rounding at the stores)."""
    ops.set_gemm_precision("tf32x3")
    yield ops
    ops.set_gemm_precision("tf32")


def _r(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).cuda()


# ------------------------------------------------------------------------------------------ GEMM
@pytest.mark.parametrize("M,N,K", [(256, 128, 64), (200, 96, 72), (1024, 32, 128), (128, 2048, 512), (64, 64, 32),
                                   (4096, 256, 64)])
def test_gemm_nt_bias(ops, M, N, K):
    x, w, b = _r(M, K, seed=1), _r(N, K, seed=2, scale=0.1), _r(N, seed=3)
    y = ops.linear(x, w, b)
    ref = x.double() @ w.double().t() + b.double()
    assert rel_l2(y, ref) < TOL_TF32


def test_gemm_segmented_and_strided(ops):
    M, C = 512, 64
    x = _r(M, 2 * C, seed=4)[:, :C]  # strided view, lda = 2C
    wq, wkv = _r(C, C, seed=5, scale=0.1), _r(2 * C, C, seed=6, scale=0.1)
    bq, bkv = _r(C, seed=7), _r(2 * C, seed=8)
    y = ops.linear(x, wq, bq, weight2=wkv, bias2=bkv)
    ref = torch.cat([x.double() @ wq.double().t() + bq.double(), x.double() @ wkv.double().t() + bkv.double()], 1)
    assert rel_l2(y, ref) < TOL_TF32
    dy = _r(M, 3 * C, seed=9)
    dx = ops.linear_dgrad(dy, wq, weight2=wkv)
    refdx = dy.double() @ torch.cat([wq, wkv], 0).double()
    assert rel_l2(dx, refdx) < TOL_TF32


def test_gemm_residual_rowscale(ops):
    B, L, K, N = 4, 256, 128, 64
    x, w, b, r = _r(B * L, K, seed=1), _r(N, K, seed=2, scale=0.1), _r(N, seed=3), _r(B * L, N, seed=4)
    s = torch.tensor([0.0, 1.25, 1.25, 0.0]).cuda()
    y = ops.linear(x, w, b, residual=r, rowscale=s, rows_per_group=L)
    ref = r.double() + s.double().repeat_interleave(L)[:, None] * (x.double() @ w.double().t() + b.double())
    assert rel_l2(y, ref) < TOL_TF32
    dx = ops.linear_dgrad(r, w, rowscale=s, rows_per_group=L)
    refdx = (s.double().repeat_interleave(L)[:, None] * r.double()) @ w.double()
    assert rel_l2(dx, refdx) < TOL_TF32


@pytest.mark.parametrize("M,N,K", [(8192, 128, 32), (65536, 256, 64), (1000, 96, 72), (512, 2048, 512)])
def test_gemm_wgrad_splitk_colsum(ops, M, N, K):
    dy, x = _r(M, N, seed=1), _r(M, K, seed=2)
    dW, db = ops.linear_wgrad(dy, x)
    assert rel_l2(dW, dy.double().t() @ x.double()) < TOL_TF32
    assert rel_l2(db, dy.double().sum(0)) < TOL_FP32 * 10


def test_gemm_wgrad_kscale(ops):
    B, L, N, K = 4, 1024, 64, 32
    dy, x = _r(B * L, N, seed=1), _r(B * L, K, seed=2)
    s = torch.tensor([1.1, 0.0, 1.1, 1.1]).cuda()
    dW, db = ops.linear_wgrad(dy, x, rowscale=s, rows_per_group=L)
    sd = s.double().repeat_interleave(L)[:, None] * dy.double()
    assert rel_l2(dW, sd.t() @ x.double()) < TOL_TF32
    assert rel_l2(db, sd.sum(0)) < TOL_FP32 * 10


def test_gemm_dgelu_epilogue(ops):
    M, N, K = 512, 128, 64
    x, w, v = _r(M, K, seed=1), _r(N, K, seed=2, scale=0.1), _r(M, N, seed=3)
    out = torch.empty(M, N, device="cuda")
    ops.gemm(x, w, out, M, N, K, lda=K, ldb=K, ldc=N, b_nk=True, epilogue=ops.EPI_MUL_DGELU, R=v, ldr=N)
    vd = v.double().requires_grad_()
    F.gelu(vd).sum().backward()
    assert rel_l2(out, (x.double() @ w.double().t()) * vd.grad) < TOL_TF32


# ------------------------------------------------------------------------------------- LayerNorm
@pytest.mark.parametrize("rows,C", [(4096, 32), (1000, 64), (777, 128), (512, 256), (256, 512), (64, 48), (333, 16),
                                     (65, 1024), (5, 32)])
def test_layernorm(exact, rows, C):
    ops = exact
    x = _r(rows, C, seed=1) * 2 + 0.3
    g, b = _r(C, seed=2) * 0.1 + 1, _r(C, seed=3) * 0.1
    y, mean, rstd = ops.layernorm_fwd(x, g, b)
    xd, gd, bd = (t.double().requires_grad_() for t in (x, g, b))
    ref = F.layer_norm(xd, (C,), gd, bd, 1e-5)
    assert rel_l2(y, ref) < TOL_FP32
    dy, dres = _r(rows, C, seed=4), _r(rows, C, seed=5)
    ref.backward(dy.double())
    dx, dg, db = ops.layernorm_bwd(dy, x, g, mean, rstd, dres=dres)
    assert rel_l2(dx, xd.grad + dres.double()) < TOL_FP32
    assert rel_l2(dg, gd.grad) < TOL_FP32 * 5
    assert rel_l2(db, bd.grad) < TOL_FP32 * 5


def test_torch_psnr_metric():
    """uwr.metrics.torchPSNR (native clamped-MSE pass) vs the reference formula (ModelTrainer.py:17-21)."""
    from uwr.metrics import torchPSNR
    tar, prd = _r(3, 3, 64, 64, seed=7) * 0.5 + 0.5, _r(3, 3, 64, 64, seed=8) * 0.5 + 0.5   # values outside [0, 1] too
    got = torchPSNR(tar, prd).item()
    t, p = tar.double().clamp(0, 1), prd.double().clamp(0, 1)
    ref = (20 * torch.log10(1.0 / torch.sqrt(((p - t) ** 2).mean()))).item()
    assert abs(got - ref) < 1e-4


# ------------------------------------------------------------------------------ window attention
def _attn_ref(qkv, table, w, B, H, W, heads, shift, sparse=True):
    """fp64 reference from the oracle's window_attention with identity projections."""
    from oracle import ast_oracle as ao
    C = qkv.shape[1] // 3
    hd = C // heads
    q, k, v = (qkv[:, i * C:(i + 1) * C].double().view(B, H, W, C) for i in range(3))

    def win(t):
        if shift:
            t = torch.roll(t, (-shift, -shift), (1, 2))
        return ao.to_windows(t, B, H, W, C)
    qw, kw, vw = win(q), win(k), win(v)
    N = 64
    qh = qw.view(-1, N, heads, hd).transpose(1, 2) * hd ** -0.5
    kh = kw.view(-1, N, heads, hd).transpose(1, 2)
    vh = vw.view(-1, N, heads, hd).transpose(1, 2)
    s = qh @ kh.transpose(-2, -1)
    from uwr.ast import _relative_position_index
    idx = _relative_position_index(8).view(-1).cuda()
    s = s + table.double()[idx].view(N, N, heads).permute(2, 0, 1).unsqueeze(0)
    if shift:
        m = ao.shift_mask(H, W, shift, torch.float64, 'cuda')
        nW = m.shape[0]
        s = (s.view(B, nW, heads, N, N) + m[None, :, None]).view(-1, heads, N, N)
    p = torch.softmax(s, -1)
    if sparse:
        e = torch.exp(w.double())
        p = p * (e[0] / e.sum()) + torch.relu(s) ** 2 * (e[1] / e.sum())
    o = (p @ vh).transpose(1, 2).reshape(-1, N, C)
    o = ao.from_windows(o, B, H, W, C)
    if shift:
        o = torch.roll(o, (shift, shift), (1, 2))
    return o.reshape(B * H * W, C)


@pytest.mark.parametrize("B,H,heads,hd,shift", [(2, 16, 2, 32, 0), (2, 16, 2, 32, 4), (1, 32, 4, 32, 4),
                                                 (1, 16, 4, 8, 4), (1, 16, 2, 16, 0), (1, 16, 2, 64, 4),
                                                 (1, 16, 1, 128, 4), (3, 8, 1, 32, 4)])
@pytest.mark.parametrize("sparse", [True, False])
def test_window_attention_fwd_bwd(ops, B, H, heads, hd, shift, sparse):
    W = H
    C = heads * hd
    qkv = _r(B * H * W, 3 * C, seed=1)
    table = _r(225, heads, seed=2, scale=0.5)
    w = torch.tensor([0.3, -0.2]).cuda()
    scale = hd ** -0.5
    o = ops.window_attn_fwd(qkv, 0, qkv, C, 2 * C, table, w if sparse else None, B, H, W, heads, hd, shift, scale)
    qd, td, wd = qkv.double().requires_grad_(), table.double().requires_grad_(), w.double().requires_grad_()
    ref = _attn_ref(qd, td, wd, B, H, W, heads, shift, sparse)
    assert rel_l2(o, ref) < TOL_TF32
    do = _r(B * H * W, C, seed=3)
    ref.backward(do.double())
    cs = torch.full((3 * C,), float("nan"), device="cuda")   # projection bias gradient, accumulated inside the kernel
    dqkv, _, dtable, dw = ops.window_attn_bwd(do, qkv, 0, qkv, C, 2 * C, table, w if sparse else None, B, H, W,
                                              heads, hd, shift, scale, colsum_q=cs, colsum_kv=cs)
    assert rel_l2(dqkv, qd.grad) < 2 * TOL_TF32
    assert rel_l2(dtable, td.grad) < 2 * TOL_TF32
    if sparse:
        assert rel_l2(dw, wd.grad) < 5 * TOL_TF32
    assert rel_l2(cs, dqkv.double().sum(0)) < 5e-4            # sums of the unrounded values vs the stored (TF32-rounded) ones
    assert rel_l2(cs, qd.grad.sum(0)) < 5 * TOL_TF32
    # not requested: same gradients, nothing else written
    dqkv2, _, _, _ = ops.window_attn_bwd(do, qkv, 0, qkv, C, 2 * C, table, w if sparse else None, B, H, W, heads, hd,
                                         shift, scale)
    assert torch.equal(dqkv2, dqkv)


@pytest.mark.parametrize("B,H,heads,hd,shift,t5", [(2, 16, 2, 32, 0, False), (1, 32, 2, 32, 4, False), (2, 16, 4, 16, 4, False),
                                                   (1, 16, 1, 64, 0, False), (2, 16, 2, 32, 4, True), (1, 32, 4, 32, 0, True)])
def test_window_attention_exact_tf32_operands(ops, B, H, heads, hd, shift, t5):
    """operands_rounded = 1: q, k, v, dout arrive as exact TF32 values (rounded by the producing GEMM epilogues), the
    kernels then run ONE tensor-core pass per product — exact products, so the bounds of the 3xTF32 path still hold
    against fp64 math on the same (rounded) operands.  t5: forward through the tcgen05 / TMA kernel."""
    W = H
    C = heads * hd
    qkv = ops.scale_round(_r(B * H * W, 3 * C, seed=1), 3 * C)
    table = _r(225, heads, seed=2, scale=0.5)
    w = torch.tensor([0.3, -0.2]).cuda()
    scale = hd ** -0.5
    ops.set_attn_tcgen05(t5)
    try:
        o = ops.window_attn_fwd(qkv, 0, qkv, C, 2 * C, table, w, B, H, W, heads, hd, shift, scale, rounded=True)
    finally:
        ops.set_attn_tcgen05("auto")
    qd, td, wd = qkv.double().requires_grad_(), table.double().requires_grad_(), w.double().requires_grad_()
    ref = _attn_ref(qd, td, wd, B, H, W, heads, shift, True)
    assert rel_l2(o, ref) < TOL_TF32
    do = ops.scale_round(_r(B * H * W, C, seed=3), C)
    ref.backward(do.double())
    dqkv, _, dtable, dw = ops.window_attn_bwd(do, qkv, 0, qkv, C, 2 * C, table, w, B, H, W, heads, hd, shift, scale,
                                              rounded=True)
    assert rel_l2(dqkv, qd.grad) < 2 * TOL_TF32
    assert rel_l2(dtable, td.grad) < 2 * TOL_TF32
    # dw: two scalars, w0 * (g1 - (w0 g1 + w1 g2)) over every score of the tensor — the most cancelling sum there is
    assert rel_l2(dw, wd.grad) < 1e-2


# ------------------------------------------------------------------------------------ dwconv+GELU
@pytest.mark.parametrize("B,H,Ch,mode", [(2, 16, 64, 0), (1, 32, 128, 0), (2, 24, 48, 0), (1, 16, 64, 1),
                                         (1, 40, 32, 1)])
def test_dwconv_gelu(exact, B, H, Ch, mode):
    ops = exact
    W = H
    width = Ch * (2 if mode else 1)
    u = _r(B * H * W, width, seed=1)
    wt, bs = _r(Ch, 1, 3, 3, seed=2, scale=0.3), _r(Ch, seed=3, scale=0.1)
    v, h2 = ops.dwconv_gelu_fwd(u, wt, bs, B, H, W, Ch, mode=mode)
    ud, wd, bd = u.double().requires_grad_(), wt.double().requires_grad_(), bs.double().requires_grad_()
    h1 = F.gelu(ud[:, :Ch]).view(B, H, W, Ch).permute(0, 3, 1, 2)
    vr = F.conv2d(h1, wd, bd, padding=1, groups=Ch).permute(0, 2, 3, 1).reshape(-1, Ch)
    ref = F.gelu(vr)
    if mode:
        ref = ref * F.gelu(ud[:, Ch:])
    assert rel_l2(v, vr) < TOL_FP32
    assert rel_l2(h2, ref) < TOL_FP32
    dh2 = _r(B * H * W, Ch, seed=4)
    ref.backward(dh2.double())
    du = torch.empty_like(u)
    dv = ops.gelu_gate_bwd(dh2, u, v, Ch, mode, du=du)
    du, dwt, dbs = ops.dwconv_gelu_bwd(dv, u, wt, B, H, W, Ch, du=du)
    assert rel_l2(du, ud.grad) < TOL_FP32 * 5
    assert rel_l2(dwt, wd.grad) < TOL_FP32 * 10
    assert rel_l2(dbs, bd.grad) < TOL_FP32 * 10


# ------------------------------------------------------------------------------- boundary convs
@pytest.mark.parametrize("precision", ["tf32", "tf32x3"])
@pytest.mark.parametrize("Cout,H", [(32, 48), (64, 40), (32, 24)])
def test_input_proj(ops, Cout, H, precision):
    """Default mode: tensor-core kernels (3xTF32: fp32-level, forward and weight gradient); tf32x3: the scalar fp32
    kernels.  H = 40 and 24 leave ragged 16-pixel tiles."""
    from uwr.blocks import InputProjFn
    ops.set_gemm_precision(precision)
    try:
        B = 2
        img = _r(B, 3, H, H, seed=1)
        w, b = _r(Cout, 3, 3, 3, seed=2, scale=0.2).requires_grad_(), _r(Cout, seed=3, scale=0.1).requires_grad_()
        tok = InputProjFn.apply(img, w, b, 0.01)
        wd, bd = w.detach().double().requires_grad_(), b.detach().double().requires_grad_()
        ref = F.leaky_relu(F.conv2d(img.double(), wd, bd, padding=1), 0.01).flatten(2).transpose(1, 2)
        assert rel_l2(tok, ref) < TOL_FP32
        g = _r(B, H * H, Cout, seed=4)
        tok.backward(g)
        ref.backward(g.double())
        assert rel_l2(w.grad, wd.grad) < TOL_FP32 * 10
        assert rel_l2(b.grad, bd.grad) < TOL_FP32 * 10
    finally:
        ops.set_gemm_precision("tf32")


@pytest.mark.parametrize("precision", ["tf32", "tf32x3"])
@pytest.mark.parametrize("Cin,H", [(64, 40), (32, 40), (64, 16), (32, 24)])
def test_output_proj(ops, Cin, H, precision):
    """Forward: direct fp32 kernel.  Backward: tensor-core kernel (both gradients from the im2col of the 3-channel
    cotangent, 3xTF32) in the default mode, the scalar fp32 kernel in tf32x3; ragged
    tiles (40 = 2.5 x 16) included."""
    from uwr.blocks import OutputProjFn
    ops.set_gemm_precision(precision)
    try:
        B = 2
        tok = _r(B, H * H, Cin, seed=1).requires_grad_()
        img = _r(B, 3, H, H, seed=5)
        w, b = _r(3, Cin, 3, 3, seed=2, scale=0.1).requires_grad_(), _r(3, seed=3, scale=0.1).requires_grad_()
        out = OutputProjFn.apply(tok, w, b, img, H, H)
        td, wd, bd = (t.detach().double().requires_grad_() for t in (tok, w, b))
        ref = img.double() + F.conv2d(td.transpose(1, 2).reshape(B, Cin, H, H), wd, bd, padding=1)
        assert rel_l2(out, ref) < TOL_FP32
        g = _r(B, 3, H, H, seed=4)
        out.backward(g)
        ref.backward(g.double())
        # both gradients are 3xTF32 on the tensor cores: fp32-level in either mode (every gradient of the network flows
        # through this data gradient, and the boundary weight gradients carry a large share of the gradient norm)
        assert rel_l2(tok.grad, td.grad) < TOL_FP32 * 10
        assert rel_l2(w.grad, wd.grad) < TOL_FP32 * 10
        assert rel_l2(b.grad, bd.grad) < TOL_FP32 * 10
    finally:
        ops.set_gemm_precision("tf32")


def test_downsample():
    from uwr.blocks import DownsampleFn
    B, H, C = 2, 16, 32
    x = _r(B, H * H, C, seed=1).requires_grad_()
    w, b = _r(2 * C, C, 4, 4, seed=2, scale=0.1).requires_grad_(), _r(2 * C, seed=3, scale=0.1).requires_grad_()
    y = DownsampleFn.apply(x, w, b, H, H)
    xd, wd, bd = (t.detach().double().requires_grad_() for t in (x, w, b))
    ref = F.conv2d(xd.transpose(1, 2).reshape(B, C, H, H), wd, bd, stride=2, padding=1).flatten(2).transpose(1, 2)
    assert rel_l2(y, ref) < TOL_TF32
    g = _r(B, H * H // 4, 2 * C, seed=4)
    y.backward(g)
    ref.backward(g.double())
    assert rel_l2(x.grad, xd.grad) < TOL_TF32
    assert rel_l2(w.grad, wd.grad) < TOL_TF32
    assert rel_l2(b.grad, bd.grad) < TOL_FP32 * 10


def test_upsample_cat():
    from uwr.blocks import UpsampleCatFn
    B, H, Cin, Cout = 2, 8, 64, 32
    x = _r(B, H * H, Cin, seed=1).requires_grad_()
    skip = _r(B, 4 * H * H, Cout, seed=5).requires_grad_()
    w, b = _r(Cin, Cout, 2, 2, seed=2, scale=0.1).requires_grad_(), _r(Cout, seed=3, scale=0.1).requires_grad_()
    y = UpsampleCatFn.apply(x, w, b, skip, H, H)
    xd, sd_, wd, bd = (t.detach().double().requires_grad_() for t in (x, skip, w, b))
    up = F.conv_transpose2d(xd.transpose(1, 2).reshape(B, Cin, H, H), wd, bd, stride=2).flatten(2).transpose(1, 2)
    ref = torch.cat([up, sd_], -1)
    assert rel_l2(y, ref) < TOL_TF32
    g = _r(B, 4 * H * H, 2 * Cout, seed=4)
    y.backward(g)
    ref.backward(g.double())
    assert rel_l2(x.grad, xd.grad) < TOL_TF32
    assert rel_l2(skip.grad, sd_.grad) < 1e-7
    assert rel_l2(w.grad, wd.grad) < TOL_TF32
    assert rel_l2(b.grad, bd.grad) < TOL_FP32 * 10


# ------------------------------------------------------------------------------------- losses
@pytest.mark.parametrize("kind", ["L1", "L2", "L1withColor", "charbonnier"])
def test_pixel_losses(kind):
    from oracle import losses_oracle as lo
    from uwr.losses import LossFunction
    torch.manual_seed(0)
    p = torch.rand(2, 3, 256, 256).cuda().requires_grad_()
    t = torch.rand(2, 3, 256, 256).cuda()
    loss = LossFunction(kind, "cuda").getloss(p, t)
    loss.backward()
    pd = p.detach().double().cpu().requires_grad_()
    fn = {"L1": lo.l1, "L2": lo.l2, "L1withColor": lo.l1_with_color, "charbonnier": lo.charbonnier}[kind]
    ref = fn(pd, t.double().cpu())
    ref.backward()
    assert abs(loss.item() - ref.item()) <= 1e-6 * abs(ref.item()) + 1e-9
    assert rel_l2(p.grad, pd.grad) < 1e-5


def test_loss_known_answers():
    """src/Loss.ipynb:45 known answer: Charbonnier(x, x) = eps = 0.0010; L1(x,x) = 0."""
    from uwr.losses import LossFunction
    x = torch.rand(1, 3, 256, 256).cuda()
    assert abs(LossFunction("charbonnier", "cuda").getloss(x, x).item() - 1e-3) < 1e-7
    assert LossFunction("L1", "cuda").getloss(x, x).item() == 0.0
    with pytest.raises(ValueError):
        LossFunction("nope", "cuda").getloss(x, x)


# ----------------------------------------------------------------------------------- optimizer
@pytest.mark.parametrize("decoupled,wd", [(False, 0.0), (True, 0.01)])
def test_fused_clip_adam(decoupled, wd):
    from uwr.optim import FusedClipAdam
    torch.manual_seed(0)
    shapes = [(64, 32), (7,), (3, 5, 3, 3), (2048, 512), (1,)]
    ps = [torch.nn.Parameter(torch.randn(s).cuda()) for s in shapes]
    qs = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    opt = FusedClipAdam(ps, lr=1e-3, weight_decay=wd, decoupled=decoupled, max_norm=1.0)
    ref = (torch.optim.AdamW if decoupled else torch.optim.Adam)(qs, lr=1e-3, weight_decay=wd)
    for it in range(3):
        for p, q in zip(ps, qs):
            g = torch.randn_like(p) * (0.001 if it == 1 else 1.0)
            p.grad = g.clone()
            q.grad = g.clone()
        norm = opt.step()
        rn = torch.nn.utils.clip_grad_norm_(qs, 1.0)
        ref.step()
        assert abs(norm[0].item() - rn.item()) < 1e-4 * rn.item()
        for p, q in zip(ps, qs):
            assert rel_l2(p, q) < 1e-6


# ------------------------------------------------------------------- whole transformer blocks
@pytest.mark.parametrize("att,shift,scales", [(True, 0, (1.25, 0.0)), (True, 4, (0.0, 1.25)), (False, 0, (1.25, 1.25)),
                                              (True, 4, None)])
def test_transformer_block_vs_oracle(att, shift, scales):
    """uwr TransformerBlock (AttnBlockFn + LeFFBlockFn incl. DropPath scaling) vs the oracle's
    transformer_block in fp64, tf32x3 GEMMs so that only summation order differs."""
    from oracle import ast_oracle as ao
    from uwr import ops
    from uwr.ast import TransformerBlock, DropPath
    torch.manual_seed(3)
    B, H, C, heads = 2, 16, 64, 2
    blk = TransformerBlock(C, (H, H), heads, win_size=8, shift_size=shift, drop_path=0.1, att=att, sparseAtt=att)
    for p in blk.parameters():
        torch.nn.init.normal_(p, std=0.08) if p.ndim > 1 else torch.nn.init.normal_(p, mean=0.5, std=0.2)
    blk = blk.cuda().train()
    sd = {k: v.detach().double().clone().requires_grad_() if v.is_floating_point() else v.detach().clone()
          for k, v in blk.state_dict().items()}
    x = _r(B, H * H, C, seed=5).requires_grad_()
    sa = sm = None
    if scales is not None:
        sa = torch.tensor([scales[0], scales[1]]).cuda()
        sm = torch.tensor([scales[1], scales[0]]).cuda()
    seq = ([sa] if att else []) + [sm]
    blk.drop_path.scale = lambda batch, device: seq.pop(0)
    ops.set_gemm_precision("tf32x3")
    try:
        y = blk(x)
        g = _r(B, H * H, C, seed=6)
        y.backward(g)
    finally:
        ops.set_gemm_precision("tf32")
    xd = x.detach().double().requires_grad_()
    yo = ao.transformer_block(sd, "", xd, heads, shift, att, "leff",
                              sa.double() if (att and sa is not None) else None, sm.double() if sm is not None else None)
    yo.backward(g.double())
    errs = {"out": rel_l2(y, yo), "dx": rel_l2(x.grad, xd.grad)}
    for n, p in blk.named_parameters():
        errs[n] = rel_l2(p.grad, sd[n].grad)
    print({k: f"{v:.1e}" for k, v in errs.items()})
    # P V and dV stay single-pass TF32 inside the attention kernel (P >= 0: no cancellation)
    assert errs.pop("out") < (3e-4 if att else 2e-5)
    assert max(errs.values()) < (2e-3 if att else 5e-4), errs


def test_producers_round_to_tf32_in_fast_mode(ops):
    """tf32 mode: LayerNorm output is exactly the TF32 rounding (cvt.rna) of the fp32 result."""
    x = _r(512, 64, seed=1)
    g, b = _r(64, seed=2) * 0.1 + 1, _r(64, seed=3) * 0.1
    y_fast, _, _ = ops.layernorm_fwd(x, g, b)
    ops.set_gemm_precision("tf32x3")
    try:
        y_full, _, _ = ops.layernorm_fwd(x, g, b)
    finally:
        ops.set_gemm_precision("tf32")
    i = y_full.view(torch.int32)
    expect = ((i + 0x1000) & ~0x1FFF).view(torch.float32)
    assert torch.equal(y_fast, expect)
    assert (y_fast.view(torch.int32) & 0x1FFF).abs().max().item() == 0


@pytest.mark.parametrize("B,S", [(2, 256), (1, 128), (3, 64)])
def test_focal_frequency_loss(B, S):
    """uwr_ffl_loss (smem FFT) vs the oracle restatement of focal_frequency_loss; FFL(x, x) = 0
    (src/Loss.ipynb:49)."""
    from oracle import losses_oracle as lo
    from uwr.losses import LossFunction
    torch.manual_seed(0)
    p = torch.rand(B, 3, S, S).cuda().requires_grad_()
    t = torch.rand(B, 3, S, S).cuda()
    loss = LossFunction("ffl", "cuda").getloss(p, t)
    loss.backward()
    pd = p.detach().double().cpu().requires_grad_()
    ref = lo.focal_frequency(pd, t.double().cpu())
    ref.backward()
    assert abs(loss.item() - ref.item()) <= 2e-5 * abs(ref.item())
    assert rel_l2(p.grad, pd.grad) < 2e-5
    assert LossFunction("ffl", "cuda").getloss(t, t).item() == 0.0
    both = LossFunction("fflCharbonnier", "cuda").getloss(p.detach(), t).item()
    assert abs(both - (ref.item() + lo.charbonnier(pd.detach(), t.double().cpu()).item())) < 1e-5


@pytest.mark.parametrize("C,scales", [(64, (1.25, 0.0)), (32, None), (128, (1.0, 1.0))])
def test_frfn_block_vs_oracle(C, scales):
    """FRFN token MLP (AST.py:329-372): partial conv + gated dwconv block vs the oracle, tf32x3."""
    from oracle import ast_oracle as ao
    from uwr import ops
    from uwr.ast import TransformerBlock
    torch.manual_seed(4)
    B, H = 2, 16
    blk = TransformerBlock(C, (H, H), 2, win_size=8, shift_size=0, drop_path=0.1, att=False, token_mlp="frfn")
    for p in blk.parameters():
        torch.nn.init.normal_(p, std=0.08) if p.ndim > 1 else torch.nn.init.normal_(p, mean=0.5, std=0.2)
    blk = blk.cuda().train()
    sd = {k: v.detach().double().clone().requires_grad_() for k, v in blk.state_dict().items()}
    x = _r(B, H * H, C, seed=5).requires_grad_()
    sm = torch.tensor(scales).cuda() if scales is not None else None
    blk.drop_path.scale = lambda batch, device: sm
    ops.set_gemm_precision("tf32x3")
    try:
        y = blk(x)
        g = _r(B, H * H, C, seed=6)
        y.backward(g)
    finally:
        ops.set_gemm_precision("tf32")
    xd = x.detach().double().requires_grad_()
    yo = ao.transformer_block(sd, "", xd, 2, 0, False, "frfn", None, sm.double() if sm is not None else None)
    yo.backward(g.double())
    errs = {"out": rel_l2(y, yo), "dx": rel_l2(x.grad, xd.grad)}
    for n, p in blk.named_parameters():
        errs[n] = rel_l2(p.grad, sd[n].grad)
    assert max(errs.values()) < 2e-4, errs
    # default (tf32, tcgen05) mode within the north-star tolerance
    x2 = x.detach().clone().requires_grad_()
    blk.zero_grad()
    y2 = blk(x2)
    y2.backward(g)
    assert rel_l2(y2, yo) < 1e-3 and rel_l2(x2.grad, xd.grad) < 2e-3


def test_fflmix_loss_tuple_matches_reference_golden():
    """LossFunction("fflMix").getloss returns the reference's 6-tuple (ModelTrainer.py:82-85); values vs
    the golden produced by the reference (patch P3), gradient flows through every term."""
    import os
    from conftest import ROOT
    from uwr.losses import LossFunction
    g = torch.load(os.path.join(ROOT, "tests", "golden", "losses_metrics.pt"), weights_only=False)["fflMix"]
    torch.manual_seed(0)
    p = torch.rand(2, 3, 256, 256)
    t = torch.rand(2, 3, 256, 256)
    pc = p.cuda().requires_grad_()
    out = LossFunction("fflMix", "cuda", vgg_weights="random").getloss(pc, t.cuda())   # patch P3, explicit opt-in
    assert isinstance(out, tuple) and len(out) == 6
    for got, want in zip(out, g):
        assert abs(got.item() - want) <= 2e-4 * abs(want), (got.item(), want)
    out[0].backward()
    assert torch.isfinite(pc.grad).all() and pc.grad.abs().sum().item() > 0


# ------------------------------------------------------------------ SpectralTransformer pieces
def test_gdfn_gate():
    """GeluMulFn (uwr_gelu_mul_fwd/bwd) vs gelu(t[:, :h]) * t[:, h:] (SpectralTransformer.py:126-129)."""
    from uwr import fn, ops
    ops.set_gemm_precision("tf32x3")   # no TF32 rounding at the store
    try:
        rows, h = 777, 44
        t = _r(rows, 2 * h, seed=1).requires_grad_()
        g = _r(rows, h, seed=2)
        y = fn.GeluMulFn.apply(t, h)
        y.backward(g)
        td = t.detach().double().requires_grad_()
        ref = F.gelu(td[:, :h]) * td[:, h:]
        ref.backward(g.double())
        assert rel_l2(y, ref) < TOL_FP32
        assert rel_l2(t.grad, td.grad) < TOL_FP32
    finally:
        ops.set_gemm_precision("tf32")


@pytest.mark.parametrize("B,L,C,heads", [(2, 256, 32, 2), (3, 1024, 16, 1)])
def test_mdta_channel_attention(B, L, C, heads):
    """MDTAAttnFn / ChannelApplyFn (batched Gram GEMMs + tiny softmax algebra) vs the reference formulation
    (SpectralTransformer.py:92-109) in fp64, values and gradients, incl. the second use of the attention."""
    from uwr import fn
    qkv = _r(B * L, 3 * C, seed=1).requires_grad_()
    kvf = _r(B * L, 2 * C, seed=2).requires_grad_()
    temp = (torch.ones(1, heads, 1, 1, device="cuda") * 1.3).requires_grad_()
    g1, g2 = _r(B * L, C, seed=3), _r(B * L, C, seed=4)
    out, attn = fn.MDTAAttnFn.apply(qkv, temp, B, L, C, heads)
    outf = fn.ChannelApplyFn.apply(kvf, attn, C, B, L)
    (out * g1).sum().add((outf * g2).sum()).backward()

    qd, kd, td = qkv.detach().double().requires_grad_(), kvf.detach().double().requires_grad_(), temp.detach().double().requires_grad_()
    c = C // heads
    x = qd.view(B, L, 3 * C).transpose(1, 2)                       # (B, 3C, L) as the reference's NCHW flatten
    q, k, v = (t.reshape(B, heads, c, L) for t in x.chunk(3, dim=1))
    q, k = F.normalize(q, dim=-1), F.normalize(k, dim=-1)
    a = torch.softmax(q @ k.transpose(-2, -1) * td, dim=-1)        # (B, h, c, c)
    ro = (a @ v).reshape(B, C, L).transpose(1, 2).reshape(B * L, C)
    vf = kd.view(B, L, 2 * C)[:, :, C:].transpose(1, 2).reshape(B, heads, c, L)
    rof = (a @ vf).reshape(B, C, L).transpose(1, 2).reshape(B * L, C)
    (ro * g1.double()).sum().add((rof * g2.double()).sum()).backward()
    assert rel_l2(out, ro) < TOL_TF32 and rel_l2(outf, rof) < TOL_TF32
    assert rel_l2(qkv.grad, qd.grad) < 2 * TOL_TF32
    assert rel_l2(kvf.grad, kd.grad) < 2 * TOL_TF32
    assert rel_l2(temp.grad, td.grad) < 2 * TOL_TF32


@pytest.mark.parametrize("B,H,heads,shift", [(2, 16, 2, 0), (1, 32, 1, 4), (2, 48, 4, 4), (1, 64, 2, 4)])
@pytest.mark.parametrize("sparse", [True, False])
def test_window_attention_fwd_tcgen05_vs_mma_sync(ops, B, H, heads, shift, sparse):
    """head_dim 32: the tcgen05/TMA forward (two windows per 128-row MMA tile) against the mma.sync kernel
    and the fp64 oracle on the same inputs (both use 3xTF32 scores, so they agree far below TF32 level)."""
    hd, W = 32, H
    C = heads * hd
    qkv = _r(B * H * W, 3 * C, seed=11)
    table = _r(225, heads, seed=12, scale=0.5)
    w = torch.tensor([0.3, -0.2]).cuda() if sparse else None
    args = (qkv, 0, qkv, C, 2 * C, table, w, B, H, W, heads, hd, shift, hd ** -0.5)
    try:
        ops.set_attn_tcgen05(False)
        o_ref = ops.window_attn_fwd(*args)
        ops.set_attn_tcgen05(True)
        o_t5 = ops.window_attn_fwd(*args)
    finally:
        ops.set_attn_tcgen05("auto")   # library default
    torch.cuda.synchronize()
    ref = _attn_ref(qkv.double(), table.double(), (w if sparse else torch.zeros(2).cuda()).double(), B, H, W, heads, shift, sparse)
    assert rel_l2(o_t5, ref) < TOL_TF32
    assert rel_l2(o_t5, o_ref) < 3e-4
    assert not torch.equal(o_t5, o_ref)   # two different kernels did run


# ------------------------------------------------------------------------------------------ MDTA
@pytest.mark.parametrize("B,L,heads,c", [(2, 1024, 1, 16), (3, 4096, 2, 16), (2, 1024, 8, 16), (2, 4096, 1, 32),
                                         (1, 256, 4, 8), (1, 128, 1, 64), (2, 96, 2, 16), (1, 40, 1, 32)])
def test_mdta_gram_apply(ops, B, L, heads, c):
    """uwr_mdta_gram / uwr_mdta_apply (SpectralTransformer.py:99-101,109,113) vs fp64 einsum, on column slices of a
    wider token matrix (the q | k | v layout of the qkv projection)."""
    C = heads * c
    qkv = _r(B * L, 3 * C, seed=31)
    G, sq, sk = ops.mdta_gram(qkv, 0, qkv, C, B, L, heads, c, want_sq=True)
    q = qkv[:, :C].double().view(B, L, heads, c)
    k = qkv[:, C:2 * C].double().view(B, L, heads, c)
    v = qkv[:, 2 * C:].double().view(B, L, heads, c)
    refG = torch.einsum("blhi,blhj->bhij", q, k)
    # Gram entries of uncorrelated data cancel to ~sqrt(L): judge against the un-cancelled scale
    scale = torch.einsum("blhi,blhj->bhij", q.abs(), k.abs()).norm()
    assert ((G.double() - refG).norm() / scale).item() < TOL_TF32
    assert rel_l2(sq, (q * q).sum(1).reshape(B, C)) < TOL_FP32
    assert rel_l2(sk, (k * k).sum(1).reshape(B, C)) < TOL_FP32
    A = torch.softmax(_r(B, heads, c, c, seed=32), dim=-1)
    out = ops.mdta_apply(qkv, 2 * C, A, B, L, heads, c)
    ref = torch.einsum("bhij,blhj->blhi", A.double(), v).reshape(B * L, C)
    assert rel_l2(out, ref) < TOL_TF32
    # transposed apply with the diagonal term, written into a column slice of a wider output
    d = _r(B, C, seed=33)
    dst = torch.zeros(B * L, 3 * C, device="cuda")
    ops.mdta_apply(qkv, C, A, B, L, heads, c, transpose=True, yd=qkv, ycol=0, diag=d, out=dst, ocol=C)
    ref2 = torch.einsum("bhji,blhj->blhi", A.double(), k).reshape(B, L, C) + d.double()[:, None, :] * q.reshape(B, L, C)
    assert rel_l2(dst[:, C:2 * C], ref2.reshape(B * L, C)) < TOL_TF32
    assert dst[:, :C].abs().max().item() == 0 and dst[:, 2 * C:].abs().max().item() == 0


def test_mdta_attention_fn_vs_autograd(ops):
    """MDTAAttnFn + ChannelApplyFn (forward and hand-written backward) vs plain autograd on the reference formula."""
    from uwr import fn
    B, L, heads, c = 2, 1024, 2, 16
    C = heads * c
    qkv = _r(B * L, 3 * C, seed=41).requires_grad_()
    vf = _r(B * L, 2 * C, seed=42).requires_grad_()
    temp = (torch.ones(1, heads, 1, 1, device="cuda") * 1.7).requires_grad_()
    out, A = fn.MDTAAttnFn.apply(qkv, temp, B, L, C, heads)
    outf = fn.ChannelApplyFn.apply(vf, A, C, B, L)
    g1, g2 = _r(B * L, C, seed=43), _r(B * L, C, seed=44)
    (out * g1).sum().add((outf * g2).sum()).backward()

    q64 = qkv.detach().double().requires_grad_()
    vf64 = vf.detach().double().requires_grad_()
    t64 = temp.detach().double().requires_grad_()
    x = q64.view(B, L, 3, heads, c).permute(2, 0, 3, 4, 1)          # (3, B, h, c, L)
    qn, kn = F.normalize(x[0], dim=-1), F.normalize(x[1], dim=-1)
    A64 = torch.softmax(qn @ kn.transpose(-2, -1) * t64, dim=-1)
    o64 = (A64 @ x[2]).permute(0, 3, 1, 2).reshape(B * L, C)
    vfh = vf64[:, C:].view(B, L, heads, c).permute(0, 2, 3, 1)
    of64 = (A64 @ vfh).permute(0, 3, 1, 2).reshape(B * L, C)
    (o64 * g1.double()).sum().add((of64 * g2.double()).sum()).backward()
    errs = dict(out=rel_l2(out, o64), outf=rel_l2(outf, of64), dq=rel_l2(qkv.grad[:, :C], q64.grad[:, :C]),
                dk=rel_l2(qkv.grad[:, C:2 * C], q64.grad[:, C:2 * C]), dv=rel_l2(qkv.grad[:, 2 * C:], q64.grad[:, 2 * C:]),
                dvf=rel_l2(vf.grad, vf64.grad), dtemp=rel_l2(temp.grad, t64.grad))
    print("mdta fn errors", {k: f"{v:.2e}" for k, v in errs.items()})
    assert errs["out"] < TOL_TF32 and errs["outf"] < TOL_TF32 and errs["dv"] < TOL_TF32 and errs["dvf"] < TOL_TF32, errs
    # dq, dk, dtemp pass through the softmax / normalisation Jacobians of a (c x c) matrix whose entries come from
    # TF32 Gram products that cancel to O(sqrt(L)): a few 1e-3 on random data
    assert errs["dq"] < 5 * TOL_TF32 and errs["dk"] < 5 * TOL_TF32 and errs["dtemp"] < 5 * TOL_TF32, errs


# ------------------------------------------------------------------- up-sampler elementwise chain, thin convs
def test_polar_split_join_cabs_vs_autograd():
    """csrc/spectral_ew.cu vs torch.abs / angle / cos / sin on complex tensors (SpectralTransformer.py:176-186)."""
    from uwr import fn
    f = _r(2, 16, 16, 8, 2, seed=51).requires_grad_()
    mag, pha = fn.PolarSplitFn.apply(f)
    z = fn.PolarJoinFn.apply(mag * 1.3, pha + 0.2)
    a = fn.CAbsFn.apply(z + f)
    g = _r(*a.shape, seed=52)
    (a * g).sum().backward()
    f64 = f.detach().double().requires_grad_()
    fc = torch.view_as_complex(f64)
    m64, p64 = torch.abs(fc), torch.angle(fc)
    z64 = torch.complex(m64 * 1.3 * torch.cos(p64 + 0.2), m64 * 1.3 * torch.sin(p64 + 0.2))
    a64 = torch.abs(z64 + fc)
    (a64 * g.double()).sum().backward()
    assert rel_l2(mag, m64) < TOL_FP32 and rel_l2(pha, p64) < TOL_FP32
    assert rel_l2(a, a64) < TOL_FP32
    assert rel_l2(f.grad, f64.grad) < 5 * TOL_FP32
    # abs'(0) = angle'(0) = 0 (torch's convention)
    zero = torch.zeros(4, 2, device="cuda", requires_grad=True)
    m0, p0 = fn.PolarSplitFn.apply(zero)
    (m0.sum() + p0.sum() + fn.CAbsFn.apply(zero).sum()).backward()
    assert zero.grad.abs().max().item() == 0


def test_leaky_gelu_even_scatter_shuffle(exact):
    from uwr import fn
    x = _r(512, 64, seed=53).requires_grad_()
    y = fn.LeakyReluFn.apply(x, 0.1) + fn.GeluFn.apply(x)
    g = _r(512, 64, seed=54)
    (y * g).sum().backward()
    x64 = x.detach().double().requires_grad_()
    y64 = F.leaky_relu(x64, 0.1) + F.gelu(x64)
    (y64 * g.double()).sum().backward()
    assert rel_l2(y, y64) < TOL_FP32 and rel_l2(x.grad, x64.grad) < TOL_FP32
    B, H, W, Cc = 2, 8, 12, 16
    t = _r(B * H * W, Cc, seed=55).requires_grad_()
    bias = _r(Cc, seed=56).requires_grad_()
    out = fn.EvenScatterFn.apply(t, bias, B, H, W)
    go = _r(B * 4 * H * W, Cc, seed=57)
    (out * go).sum().backward()
    ref = bias.detach().view(1, 1, 1, Cc).expand(B, 2 * H, 2 * W, Cc).clone()
    ref[:, ::2, ::2] = t.detach().view(B, H, W, Cc)
    assert torch.equal(out.view(B, 2 * H, 2 * W, Cc), ref)
    g4 = go.view(B, 2 * H, 2 * W, Cc)
    assert torch.equal(t.grad.view(B, H, W, Cc), g4[:, ::2, ::2])
    assert rel_l2(bias.grad, g4.double().sum((0, 1, 2)) - g4[:, ::2, ::2].double().sum((0, 1, 2))) < 1e-5
    # PixelShuffle / PixelUnshuffle on tokens vs torch on NCHW
    tok = _r(B * H * W, 4 * Cc, seed=58).requires_grad_()
    up = fn.PixelShuffleFn.apply(tok, B, H, W)
    ref_up = F.pixel_shuffle(tok.detach().view(B, H, W, 4 * Cc).permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1)
    assert torch.equal(up.view(B, 2 * H, 2 * W, Cc), ref_up)
    back = fn.PixelUnshuffleFn.apply(up, B, H, W)
    assert torch.equal(back, tok.detach())
    (back * tok.detach()).sum().backward()
    assert torch.equal(tok.grad, tok.detach())


@pytest.mark.parametrize("Cout", [8, 16])
def test_conv_small_img2tok(exact, Cout):
    from uwr import fn
    B, H, W = 2, 40, 24
    img = _r(B, 3, H, W, seed=61)
    w = _r(Cout, 3, 3, 3, seed=62, scale=0.3).requires_grad_()
    b = _r(Cout, seed=63).requires_grad_() if Cout == 8 else None
    tok = fn.ConvImg2TokFn.apply(img, w, b)
    g = _r(B * H * W, Cout, seed=64)
    (tok * g).sum().backward()
    w64 = w.detach().double().requires_grad_()
    b64 = b.detach().double().requires_grad_() if b is not None else None
    ref = F.conv2d(img.double(), w64, b64, padding=1).permute(0, 2, 3, 1).reshape(B * H * W, Cout)
    (ref * g.double()).sum().backward()
    assert rel_l2(tok, ref) < TOL_FP32
    assert rel_l2(w.grad, w64.grad) < TOL_FP32
    if b is not None:
        assert rel_l2(b.grad, b64.grad) < TOL_FP32


def test_conv_small_tok2img(exact):
    from uwr import fn
    B, H, W = 2, 24, 40
    tok = _r(B * H * W, 8, seed=65).requires_grad_()
    w = _r(3, 8, 3, 3, seed=66, scale=0.3).requires_grad_()
    b = _r(3, seed=67).requires_grad_()
    res = _r(B, 3, H, W, seed=68)
    out = fn.ConvTok2ImgFn.apply(tok, w, b, res, B, H, W)
    g = _r(B, 3, H, W, seed=69)
    (out * g).sum().backward()
    t64 = tok.detach().double().requires_grad_()
    w64, b64 = w.detach().double().requires_grad_(), b.detach().double().requires_grad_()
    ref = F.conv2d(t64.view(B, H, W, 8).permute(0, 3, 1, 2), w64, b64, padding=1) + res.double()
    (ref * g.double()).sum().backward()
    assert rel_l2(out, ref) < TOL_FP32
    assert rel_l2(tok.grad, t64.grad) < TOL_FP32
    assert rel_l2(w.grad, w64.grad) < TOL_FP32 and rel_l2(b.grad, b64.grad) < TOL_FP32


def test_fused_clip_adam_is_a_torch_optimizer_with_checkpoint_and_scheduler():
    """lr lives on the device (graph replays follow a scheduler), state_dict round-trips in
    torch.optim.Adam's layout (the reference checkpoint's `optimizer_state_dict`, ModelTrainer.py:172-190)."""
    from uwr.optim import FusedClipAdam
    torch.manual_seed(0)
    ps = [torch.nn.Parameter(torch.randn(64, 32).cuda()), torch.nn.Parameter(torch.randn(7).cuda())]
    qs = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    opt = FusedClipAdam(ps, lr=1e-3, max_norm=1.0)
    ref = torch.optim.Adam(qs, lr=1e-3)
    sched = torch.optim.lr_scheduler.MultiStepLR(opt, milestones=[1, 3], gamma=0.25)   # ModelTrainer.py:55
    rsched = torch.optim.lr_scheduler.MultiStepLR(ref, milestones=[1, 3], gamma=0.25)
    grads = [[torch.randn_like(p) for p in ps] for _ in range(5)]

    def both(it):
        for p, q, g in zip(ps, qs, grads[it]):
            p.grad, q.grad = g.clone(), g.clone()
        opt.step()
        torch.nn.utils.clip_grad_norm_(qs, 1.0)
        ref.step()
        sched.step()
        rsched.step()
    for it in range(3):
        both(it)
    assert opt.param_groups[0]["lr"] == ref.param_groups[0]["lr"] and abs(opt.lr - 1e-3 * 0.25 ** 2) < 1e-12
    for p, q in zip(ps, qs):
        assert rel_l2(p, q) < 1e-6
    # checkpoint in torch layout: loads into a fresh torch.optim.Adam and back into a fresh FusedClipAdam
    sd = opt.state_dict()
    assert set(sd) == {"state", "param_groups"} and set(sd["state"][0]) == {"step", "exp_avg", "exp_avg_sq"}
    assert float(sd["state"][0]["step"]) == 3.0 and opt.step_count == 3
    ps2 = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    qs2 = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    opt2, ref2 = FusedClipAdam(ps2, lr=123.0, max_norm=1.0), torch.optim.Adam(qs2, lr=123.0)
    import copy
    opt2.load_state_dict(copy.deepcopy(sd))      # torch's loader aliases the tensors of the dict it is given
    ref2.load_state_dict(copy.deepcopy(sd))
    assert opt2.lr == ref2.param_groups[0]["lr"] == opt.lr and opt2.step_count == 3
    for p, q, g in zip(ps2, qs2, grads[3]):
        p.grad, q.grad = g.clone(), g.clone()
    opt2.step()
    torch.nn.utils.clip_grad_norm_(qs2, 1.0)
    ref2.step()
    for p, q in zip(ps2, qs2):
        assert rel_l2(p, q) < 1e-6
    # excluded (never-touched) parameters keep their value under AdamW, as torch's grad-None parameters do
    a, b = torch.nn.Parameter(torch.ones(8).cuda()), torch.nn.Parameter(torch.ones(8).cuda())
    o3 = FusedClipAdam([a, b], lr=1e-2, weight_decay=0.1, decoupled=True)
    a.grad, b.grad = torch.ones_like(a), torch.zeros_like(b)
    o3.exclude([b])
    o3.step()
    assert torch.equal(b.detach(), torch.ones_like(b)) and not torch.equal(a.detach(), torch.ones_like(a))
    assert b not in o3.state or "exp_avg" not in o3.state[b]


def test_graphed_step_follows_lr_changes():
    """A captured training step replays with the CURRENT learning rate (device scalar) — ADVICE r1."""
    import uwr
    from uwr.graph import GraphedTrainStep
    from uwr.train import TrainStep
    torch.manual_seed(3)
    model = uwr.AST(img_size=128).cuda().train()
    step = TrainStep(model, "L1", lr=1e-3)
    g = torch.Generator().manual_seed(1)
    raw = (torch.rand(1, 3, 128, 128, generator=g) * 2 - 1).cuda()
    ref = (torch.rand(1, 3, 128, 128, generator=g) * 2 - 1).cuda()
    graphed = GraphedTrainStep(step, raw, ref, warmup=3)
    p = model.output_proj.proj[0].weight
    before = p.detach().clone()
    graphed.replay()
    d1 = (p.detach() - before).abs().max().item()
    step.opt.param_groups[0]["lr"] = 0.0          # what a scheduler does
    before = p.detach().clone()
    graphed.replay()
    assert d1 > 0 and torch.equal(p.detach(), before)


# ------------------------------------------------------------------- fp16 storage of the LeFF hidden tensors
@pytest.mark.parametrize("B,H,C", [(2, 16, 64), (1, 32, 32), (2, 16, 256)])
def test_leff_block_half_storage_vs_oracle(ops, B, H, C):
    """LeFF block (LeFFBlockFn) with u / gelu'(v) stored as float16 (the default in single-pass mode) vs the fp64 oracle,
    next to the same block with fp32 storage: both must sit at TF32 level, and the half path must really be taken."""
    from oracle import ast_oracle as ao
    from uwr.ast import TransformerBlock
    torch.manual_seed(11)
    blk = TransformerBlock(C, (H, H), max(1, C // 32), win_size=8, shift_size=0, drop_path=0.0, att=False, sparseAtt=False)
    for p in blk.parameters():
        torch.nn.init.normal_(p, std=0.08) if p.ndim > 1 else torch.nn.init.normal_(p, mean=0.3, std=0.2)
    blk = blk.cuda().train()
    sd = {k: v.detach().double().clone().requires_grad_() if v.is_floating_point() else v.detach().clone()
          for k, v in blk.state_dict().items()}
    x = _r(B, H * H, C, seed=12)
    g = _r(B, H * H, C, seed=13)
    xd = x.double().requires_grad_()
    yo = ao.transformer_block(sd, "", xd, max(1, C // 32), 0, False, "leff", None, None)
    yo.backward(g.double())
    res = {}
    for half in (True, False):
        ops.set_half_storage(half)
        try:
            blk.zero_grad(set_to_none=True)
            xx = x.clone().requires_grad_()
            n0 = {}
            with ops.KernelProfile() as prof:
                y = blk(xx)
                y.backward(g)
            names = {r["kernel"] for r in prof.table()}
            assert ("uwr_dwconv_gelu_fwd_half" in names) == half and ("uwr_dwconv_gelu_bwd_half" in names) == half
            errs = {"out": rel_l2(y, yo), "dx": rel_l2(xx.grad, xd.grad)}
            for n, p in blk.named_parameters():
                errs[n] = rel_l2(p.grad, sd[n].grad)
            res[half] = errs
        finally:
            ops.set_half_storage(True)
    print("half", {k: f"{v:.1e}" for k, v in res[True].items()})
    print("fp32", {k: f"{v:.1e}" for k, v in res[False].items()})
    for half in (True, False):
        assert res[half]["out"] < 3e-4 and max(res[half].values()) < 1e-3, (half, res[half])
    # half storage costs at most a TF32-sized extra rounding of u
    assert max(res[True].values()) < 2 * max(res[False].values()) + 2e-4


def test_gemm_half_output_and_multiplier(ops):
    """tcgen05 GEMM with c_half (float16 C) and r_half (float16 UWR_EPI_MUL operand)."""
    M, N, K = 4096, 256, 64
    x = ops.scale_round(_r(M, K, seed=71), K)
    w = ops.scale_round(_r(N, K, seed=72, scale=0.1), K)
    b = _r(N, seed=73)
    y = ops.linear(x, w, b, t5=True, out_half=True)
    assert y.dtype == torch.float16
    ref = x.double() @ w.double().t() + b.double()
    assert rel_l2(y, ref) < 6e-4          # 2^-11 output rounding on top of exact TF32 products
    big = ops.linear(x * 1e4, w * 1e3, None, t5=True, out_half=True)     # |y| > 65504 saturates instead of inf
    assert torch.isfinite(big.float()).all() and big.float().abs().max().item() == 65504.0
    d = ops.scale_round(_r(M, K, seed=74), K)
    w2 = ops.scale_round(_r(K, N, seed=75, scale=0.1), N)                # stored [N_in=K][K_out=N]: dx = d w2
    mul = torch.rand(M, N, device="cuda").half()
    out = ops.linear_dgrad(d, w2, mul_by=mul, t5=True)
    ref2 = (d.double() @ w2.double()) * mul.double()
    assert out.dtype == torch.float32 and rel_l2(out, ref2) < 1e-5


# ------------------------------------------------------------------- fflMix terms and SSIM on the device (csrc/ssim.cu)
def _shim(name):
    import importlib
    import sys
    sys.path.insert(0, os.path.join(ROOT, "oracle", "shims"))
    try:
        return importlib.import_module(name)
    finally:
        sys.path.pop(0)


@pytest.mark.parametrize("B,S", [(2, 256), (1, 192)])
def test_laplacian_and_ms_ssim_vs_reference_formulas(B, S):
    """uwr.ssim (device kernels) vs Gradient_Loss restated from losses.py:162-181 and the pytorch_msssim stand-in of
    oracle/shims (the published algorithm) evaluated in fp64 on the same GPU."""
    from uwr import ssim as dev
    msssim = _shim("pytorch_msssim")
    g = torch.Generator().manual_seed(5)
    t = torch.rand(B, 3, S, S, generator=g)
    p = (t * 0.8 + 0.1 + 0.05 * torch.randn(B, 3, S, S, generator=g)).clamp(0, 1)
    pc, tc = p.cuda().requires_grad_(), t.cuda()
    lap = dev.gradient_loss(pc, tc)
    ms = dev.ms_ssim(pc, tc)
    (0.3 * lap + 0.7 * (1 - ms)).backward()
    p64, t64 = p.cuda().double().requires_grad_(), t.cuda().double()
    k = torch.tensor([[0.0, 1.0, 0.0], [1.0, -4.0, 1.0], [0.0, 1.0, 0.0]], dtype=torch.float64, device="cuda")
    k = k.view(1, 1, 3, 3).repeat(3, 1, 1, 1)
    lap64 = F.l1_loss(F.conv2d(p64, k, groups=3), F.conv2d(t64, k, groups=3))
    ms64 = msssim.MS_SSIM(data_range=1.0, size_average=True, channel=3)(p64, t64)
    (0.3 * lap64 + 0.7 * (1 - ms64)).backward()
    print(f"lap {lap.item():.6f}/{lap64.item():.6f} ms_ssim {ms.item():.6f}/{ms64.item():.6f} grad {rel_l2(pc.grad, p64.grad):.2e}")
    assert abs(lap.item() - lap64.item()) < 1e-5 * abs(lap64.item())
    assert abs(ms.item() - ms64.item()) < 1e-5
    assert rel_l2(pc.grad, p64.grad) < 1e-3      # fp32 variances e - a^2 cancel, as in the reference's own fp32 path
    s1 = dev.ssim(tc, pc.detach())
    s64 = msssim.ssim(t64, p64.detach(), data_range=1.0, size_average=True)
    assert abs(s1.item() - s64.item()) < 1e-5
    import uwr
    assert abs(uwr.torchSSIM(tc, pc.detach()).item() - s64.item()) < 1e-5
