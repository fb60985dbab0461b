"""tcgen05/TMA GEMM (uwr_gemm_tcgen05) vs fp64 on TF32-pre-rounded operands: all three layouts,
tile edges, epilogues, split contraction."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

from conftest import rel_l2

pytestmark = pytest.mark.gpu


def tf32_round(x):
    """round-to-nearest (ties away) to 10 mantissa bits, like cvt.rna.tf32.f32"""
    i = x.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


def _r(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return tf32_round((torch.randn(*shape, generator=g) * scale).cuda())


def _t5(A, B, Cout, M, N, K, lda, ldb, ldc, a_km=False, b_nk=True, bias=None, epilogue=0, R=None, ldr=0,
        rowscale=None, rpg=0):
    from uwr._lib import GemmDesc, check, fn
    d = GemmDesc()
    d.A, d.lda, d.a_km = A.data_ptr(), lda, int(a_km)
    d.B, d.ldb, d.b_nk = B.data_ptr(), ldb, int(b_nk)
    d.C, d.ldc, d.M, d.N, d.K = Cout.data_ptr(), ldc, M, N, K
    d.bias = bias.data_ptr() if bias is not None else None
    d.epilogue = epilogue
    d.R, d.ldr = (R.data_ptr() if R is not None else None), ldr
    d.rowscale, d.rows_per_group = (rowscale.data_ptr() if rowscale is not None else None), rpg
    ws = None
    nbytes = fn["uwr_gemm_tcgen05_workspace_bytes"](M, N, K, int(a_km))
    if nbytes:
        ws = torch.empty(nbytes // 4, device="cuda")
        d.workspace, d.workspace_bytes = ws.data_ptr(), nbytes
    assert fn["uwr_gemm_tcgen05_supported"](C.byref(d)) == 1
    check(fn["uwr_gemm_tcgen05"](C.byref(d), torch.cuda.current_stream().cuda_stream), "uwr_gemm_tcgen05")
    torch.cuda.synchronize()
    return Cout


@pytest.mark.parametrize("M,N,K", [(128, 32, 32), (256, 64, 64), (1000, 256, 64), (4096, 192, 64), (512, 512, 128),
                                   (300, 2048, 512), (65536, 256, 64), (2048, 64, 256), (128, 128, 72)])
def test_t5_nt(M, N, K):
    x, w, b = _r(M, K, seed=1), _r(N, K, seed=2, scale=0.1), _r(N, seed=3)
    out = torch.full((M, N), float("nan"), device="cuda")
    _t5(x, w, out, M, N, K, K, K, N, bias=b)
    assert rel_l2(out, x.double() @ w.double().t() + b.double()) < 1e-5


def test_t5_nt_epilogues():
    B, L, K, N = 4, 256, 128, 64
    x, w, b, r = _r(B * L, K, seed=1), _r(N, K, seed=2, scale=0.1), _r(N, seed=3), _r(B * L, N, seed=4)
    s = torch.tensor([0.0, 1.25, 1.25, 0.0]).cuda()
    out = torch.empty(B * L, N, device="cuda")
    _t5(x, w, out, B * L, N, K, K, K, N, bias=b, epilogue=1, R=r, ldr=N, rowscale=s, rpg=L)
    ref = r.double() + s.double().repeat_interleave(L)[:, None] * (x.double() @ w.double().t() + b.double())
    assert rel_l2(out, ref) < 2e-6
    _t5(x, w, out, B * L, N, K, K, K, N, epilogue=2, R=r, ldr=N)
    vd = r.double().requires_grad_()
    F.gelu(vd).sum().backward()
    assert rel_l2(out, (x.double() @ w.double().t()) * vd.grad) < 2e-6
    _t5(x, w, out, B * L, N, K, K, K, N, epilogue=3, R=r, ldr=N, rowscale=s, rpg=L)
    assert rel_l2(out, s.double().repeat_interleave(L)[:, None] * (x.double() @ w.double().t()) * r.double()) < 2e-6


@pytest.mark.parametrize("M,N,K", [(256, 64, 256), (1000, 32, 128), (4096, 64, 256), (512, 128, 512),
                                   (300, 512, 2048), (65536, 64, 256),
                                   # widths that are not multiples of 32 (GDFN hidden 2*88 / 2*172, 3-channel ends):
                                   # the ragged last 32-float group of the MN-major operand is zero-filled by TMA
                                   (4096, 88, 32), (1000, 176, 32), (2048, 344, 64), (512, 36, 176), (256, 300, 88)])
def test_t5_nn(M, N, K):
    """dx[M,N] = dy[M,K] W[K,N]  (W stored [K][N]: MN-major B operand)"""
    dy, w = _r(M, K, seed=1), _r(K, N, seed=2, scale=0.1)
    out = torch.full((M, N), float("nan"), device="cuda")
    _t5(dy, w, out, M, N, K, K, N, N, b_nk=False)
    assert rel_l2(out, dy.double() @ w.double()) < 1e-5


@pytest.mark.parametrize("M,N,K", [(128, 32, 512), (256, 64, 8192), (64, 256, 65536), (2048, 512, 4096),
                                   (192, 64, 1000), (32, 128, 100000),
                                   (176, 32, 8192), (32, 88, 8192), (88, 36, 4100), (344, 64, 2048), (300, 172, 1024)])
def test_t5_tn(M, N, K):
    """dW[M,N] = dy[K,M]^T x[K,N]  (both MN-major, contraction split across CTAs)"""
    dy, x = _r(K, M, seed=1), _r(K, N, seed=2)
    out = torch.full((M, N), float("nan"), device="cuda")
    _t5(dy, x, out, M, N, K, M, N, N, a_km=True, b_nk=False)
    assert rel_l2(out, dy.double().t() @ x.double()) < 5e-6


@pytest.mark.parametrize("lay,M,N,K", [("nt", 384, 512, 1024), ("nt", 4096, 2048, 512), ("nt", 1152, 256, 2048),
                                       ("nn", 640, 1024, 2048), ("nn", 2048, 512, 1024), ("tn", 512, 2048, 16384),
                                       ("tn", 384, 256, 65536), ("tn", 2048, 512, 4096)])
def test_t5_cluster_multicast_matches_single_cta(lay, M, N, K):
    """Tensor-bound shapes run as cluster pairs with TMA multicast of the B tile (gemm_tcgen05_kernel<..., CL = 2>):
    bit-identical to the single-CTA kernel (same products, same accumulation order per tile) and within TF32-exact
    distance of fp64, including an odd number of row tiles (the unpaired CTA computes on zero-filled rows)."""
    from uwr import ops
    bias = _r(N, seed=5) if lay == "nt" else None
    if lay == "nt":
        A, Bm = _r(M, K, seed=1), _r(N, K, seed=2, scale=0.1)
        ref = A.double() @ Bm.double().t() + bias.double()
        kw = dict(lda=K, ldb=K, ldc=N, bias=bias)
    elif lay == "nn":
        A, Bm = _r(M, K, seed=1), _r(K, N, seed=2, scale=0.1)
        ref = A.double() @ Bm.double()
        kw = dict(lda=K, ldb=N, ldc=N, b_nk=False)
    else:
        A, Bm = _r(K, M, seed=1), _r(K, N, seed=2, scale=0.1)
        ref = A.double().t() @ Bm.double()
        kw = dict(lda=M, ldb=N, ldc=N, a_km=True, b_nk=False)
    outs = {}
    for on in (True, False):
        ops.set_gemm_cluster(bool(on))
        try:
            outs[on] = _t5(A, Bm, torch.full((M, N), float("nan"), device="cuda"), M, N, K, **kw).clone()
        finally:
            ops.set_gemm_cluster(True)
    assert torch.equal(outs[True], outs[False])
    assert rel_l2(outs[True], ref) < 2e-5


@pytest.mark.parametrize("B,H,W,Cin,Cout,k,stride,pad", [
    (2, 16, 16, 32, 64, 3, 1, 1),       # 2 rows of 16 x 8 = one 128-pixel box
    (1, 32, 32, 64, 32, 3, 1, 1),
    (2, 64, 64, 32, 16, 3, 1, 1),       # thin output (PixelUnshuffle downsample of SpectralTransformer)
    (1, 128, 128, 32, 32, 3, 1, 1),     # one box = one full row
    (1, 256, 256, 32, 8, 3, 1, 1),      # half a row per box
    (2, 32, 32, 64, 128, 4, 2, 1),      # AST Downsample: Conv4x4 stride 2 (element strides in the tensor map)
    (1, 64, 64, 32, 64, 4, 2, 1),
    (1, 256, 256, 32, 64, 4, 2, 1),
    (2, 16, 48, 32, 32, 3, 1, 1),       # width neither a power of two nor a multiple of the box: must be refused
])
def test_conv_implicit_gemm(B, H, W, Cin, Cout, k, stride, pad):
    """uwr_convgemm_tcgen05 (TMA-materialised im2col tiles, no im2col buffer) vs fp64 F.conv2d on the same TF32-rounded
    operands: forward (+ bias) and weight gradient."""
    from uwr import ops
    x = _r(B * H * W, Cin, seed=1)                                         # tokens, NHWC
    w = _r(Cout, Cin, k, k, seed=2, scale=0.1)
    bias = _r(Cout, seed=3)
    wmat = w.permute(0, 2, 3, 1).reshape(Cout, k * k * Cin).contiguous()   # K index = (ky, kx, ci)
    y = ops.conv_gemm_fwd(x, wmat, bias, B, H, W, k, k, stride, pad)
    if W == 48:
        assert y is None
        return
    assert y is not None
    xd = x.double().view(B, H, W, Cin).permute(0, 3, 1, 2)
    ref = F.conv2d(xd, w.double(), bias.double(), stride=stride, padding=pad)
    OH, OW = ref.shape[2], ref.shape[3]
    torch.cuda.synchronize()
    assert rel_l2(y.view(B, OH, OW, Cout), ref.permute(0, 2, 3, 1)) < 1e-5
    dy = _r(B * OH * OW, Cout, seed=4)
    dw = ops.conv_gemm_wgrad(dy, x, B, H, W, k, k, stride, pad)
    assert dw is not None
    wd = w.double().requires_grad_()
    F.conv2d(xd, wd, None, stride=stride, padding=pad).backward(dy.double().view(B, OH, OW, Cout).permute(0, 3, 1, 2))
    torch.cuda.synchronize()
    assert rel_l2(dw.reshape(Cout, k, k, Cin), wd.grad.permute(0, 2, 3, 1)) < 1e-5
