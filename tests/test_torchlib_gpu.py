"""torch.library surface (uwr/torchlib.py): every `torch.ops.uwr.*` op passes torch.library.opcheck (schema, fake
tensor, autograd registration) and its autograd matches fp64 PyTorch math / the oracle."""
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_l2

pytestmark = pytest.mark.gpu
CHECKS = ("test_schema", "test_autograd_registration", "test_faketensor")


def _r(*shape, seed=0, scale=1.0, grad=False):
    g = torch.Generator(device="cpu").manual_seed(seed)
    t = (torch.randn(*shape, generator=g) * scale).cuda()
    return t.requires_grad_() if grad else t


def _cases():
    import uwr  # noqa: F401  (registers the ops)
    B, H, W, heads, Cc = 2, 16, 16, 2, 64
    M = B * H * W
    qkv = _r(M, 3 * Cc, seed=1, grad=True)
    table = _r(225, heads, seed=2, scale=0.02, grad=True)
    w = torch.ones(2, device="cuda", requires_grad=True)
    img = lambda s: _r(2, 3, 64, 64, seed=s)
    blk = [_r(B, H * W, Cc, seed=20, grad=True), torch.ones(Cc, device="cuda", requires_grad=True),
           torch.zeros(Cc, device="cuda", requires_grad=True), _r(Cc, Cc, seed=21, scale=0.1, grad=True),
           _r(Cc, seed=22, scale=0.1, grad=True), _r(2 * Cc, Cc, seed=23, scale=0.1, grad=True),
           _r(2 * Cc, seed=24, scale=0.1, grad=True), table, w, _r(Cc, Cc, seed=25, scale=0.1, grad=True),
           _r(Cc, seed=26, scale=0.1, grad=True), H, W, heads, 4]
    return {
        "linear": (_r(256, 64, seed=3, grad=True), _r(128, 64, seed=4, scale=0.1, grad=True), _r(128, seed=5, grad=True)),
        "layernorm": (_r(512, 64, seed=6, grad=True), _r(64, seed=7, grad=True), _r(64, seed=8, grad=True), 1e-5),
        "window_attn_sparse": (qkv, table, w, B, H, W, heads, 4, 32 ** -0.5),
        "dwconv3x3_gelu": (_r(M, 128, seed=9, grad=True), _r(128, 1, 3, 3, seed=10, scale=0.3, grad=True),
                           _r(128, seed=11, grad=True), B, H, W),
        "l1_family_loss": (img(12).requires_grad_(), img(13), "L1withColor", 2),
        "charbonnier_loss": (img(12).requires_grad_(), img(13)),
        "ffl_loss": (img(12).requires_grad_(), img(13)),
        "fft2_real": (_r(2, 16, 16, 32, seed=14, grad=True), 1.0),
        "fft_lc_real": (_r(2, 16, 16, 32, seed=14, grad=True), 0.5),
        "mdta_gram": (_r(2 * 256, 32, seed=15, grad=True), _r(2 * 256, 32, seed=16, grad=True), 2, 256, 2),
        "mdta_apply": (_r(2 * 256, 32, seed=15, grad=True), _r(2, 2, 16, 16, seed=17, grad=True), False, 2, 256),
        "polar_split": (_r(128, 8, 2, seed=18, grad=True),),
        "polar_join": (_r(128, 8, seed=18, grad=True), _r(128, 8, seed=19, grad=True)),
        "fused_window_block": tuple(blk),
    }


def test_every_registered_op_has_a_case():
    from uwr import torchlib
    assert set(_cases()) == set(torchlib.OPS)


@pytest.mark.parametrize("name", ["linear", "layernorm", "window_attn_sparse", "dwconv3x3_gelu", "l1_family_loss",
                                  "charbonnier_loss", "ffl_loss", "fft2_real", "fft_lc_real", "mdta_gram", "mdta_apply",
                                  "polar_split", "polar_join", "fused_window_block"])
def test_opcheck(name):
    args = _cases()[name]
    torch.library.opcheck(getattr(torch.ops.uwr, name).default, args, test_utils=CHECKS)


def test_cpu_tensor_is_refused():
    import uwr  # noqa: F401
    with pytest.raises(NotImplementedError):
        torch.ops.uwr.linear(torch.zeros(4, 4), torch.zeros(4, 4), None)


def test_linear_layernorm_autograd_vs_fp64():
    x, w, b = _cases()["linear"]
    y = torch.ops.uwr.linear(x, w, b)
    g = _r(*y.shape, seed=30)
    (y * g).sum().backward()
    x64, w64, b64 = (t.detach().double().requires_grad_() for t in (x, w, b))
    (F.linear(x64, w64, b64) * g.double()).sum().backward()
    assert rel_l2(y, F.linear(x64, w64, b64)) < 1e-3
    assert rel_l2(x.grad, x64.grad) < 1e-3 and rel_l2(w.grad, w64.grad) < 1e-3 and rel_l2(b.grad, b64.grad) < 1e-3
    from uwr import ops
    ops.set_gemm_precision("tf32x3")     # LayerNorm keeps full fp32 outputs (no TF32 rounding at the store)
    try:
        x, gw, gb, eps = _cases()["layernorm"]
        y, _, _ = torch.ops.uwr.layernorm(x, gw, gb, eps)
        g = _r(*y.shape, seed=31)
        (y * g).sum().backward()
    finally:
        ops.set_gemm_precision("tf32")
    x64, w64, b64 = (t.detach().double().requires_grad_() for t in (x, gw, gb))
    y64 = F.layer_norm(x64, (64,), w64, b64, eps)
    (y64 * g.double()).sum().backward()
    assert rel_l2(y, y64) < 2e-5 and rel_l2(x.grad, x64.grad) < 2e-5
    assert rel_l2(gw.grad, w64.grad) < 2e-5 and rel_l2(gb.grad, b64.grad) < 2e-5


def test_fused_window_block_vs_oracle():
    """torch.ops.uwr.fused_window_block == the attention half of TransformerBlock.forward (AST.py:590-619)."""
    from oracle import ast_oracle
    from uwr import ops
    ops.set_gemm_precision("tf32x3")
    try:
        args = list(_cases()["fused_window_block"])
        x, n1w, n1b, wq, bq, wkv, bkv, table, w, wp, bp, H, W, heads, shift = args
        out = torch.ops.uwr.fused_window_block(*args)[0]
        g = _r(*out.shape, seed=40)
        (out * g).sum().backward()
        names = ["x", "norm1.weight", "norm1.bias", "attn.qkv.to_q.weight", "attn.qkv.to_q.bias", "attn.qkv.to_kv.weight",
                 "attn.qkv.to_kv.bias", "attn.relative_position_bias_table", "attn.w", "attn.proj.weight", "attn.proj.bias"]
        sd = {n: t.detach().double().requires_grad_() for n, t in zip(names, args[:11])}
        from uwr.ast import _relative_position_index
        sd["attn.relative_position_index"] = _relative_position_index(8).cuda()
        B, L, C = x.shape
        x64 = sd["x"]
        y = F.layer_norm(x64, (C,), sd["norm1.weight"], sd["norm1.bias"], 1e-5).view(B, H, W, C)
        y = torch.roll(y, shifts=(-shift, -shift), dims=(1, 2))
        mask = ast_oracle.shift_mask(H, W, shift, torch.float64, "cuda")
        yw = ast_oracle.window_attention(sd, "attn.", ast_oracle.to_windows(y, B, H, W, C), heads, mask, sparse=True)
        y = torch.roll(ast_oracle.from_windows(yw, B, H, W, C), shifts=(shift, shift), dims=(1, 2)).reshape(B, L, C)
        ref = x64 + y
        (ref * g.double()).sum().backward()
        # P V and dV = P^T dO stay single-pass TF32 inside the attention kernel: TF32-level bounds
        assert rel_l2(out, ref) < 1e-3
        for n, t in zip(names, args[:11]):
            assert rel_l2(t.grad, sd[n].grad) < 2e-3, n
    finally:
        ops.set_gemm_precision("tf32")
