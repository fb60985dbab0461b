"""`torch.library` registration of the uwr kernels: `torch.ops.uwr.<name>` (SURVEY.md §8b "New C-ABI": each op
"registered as torch.library op uwr::<name> with a fake/meta impl and register_autograd").

Every op here is a thin shim over the same C ABI the modules use (uwr.ops -> ctypes -> libuwr_b200.so):
  * CUDA implementation only — a CPU tensor raises NotImplementedError from the dispatcher (no fallback);
  * a fake (meta) implementation, so the ops trace under FakeTensorMode / torch.export / AOT autograd;
  * autograd through a second registered op (`*_bwd`) that launches the hand-written backward kernels.
Forward ops that need activations in backward return them as extra outputs (the standard custom-op pattern).
The nn.Modules keep their fused autograd.Functions (uwr/blocks.py: in-place gradient-bucket writes, LayerNorm
links); these ops are the public operator surface on top of the same kernels.

Reference lines each op replaces are cited in include/uwr_b200.h next to the C entry it calls.
"""
from typing import Optional, Tuple

import torch
from torch import Tensor
from torch.library import custom_op, register_autograd

from . import ops

_CUDA = "cuda"


def _c(t):
    return t if t.is_contiguous() else t.contiguous()


# ------------------------------------------------------------------------------------------------ linear
@custom_op("uwr::linear", mutates_args=(), device_types=_CUDA)
def linear(x: Tensor, weight: Tensor, bias: Optional[Tensor]) -> Tensor:
    """y = x W^T + b on a token matrix x (M, K); nn.Linear (AST.py:47-48,104,297,302)."""
    return ops.linear(_c(x), _c(weight), bias)


@linear.register_fake
def _(x, weight, bias):
    return x.new_empty(x.shape[0], weight.shape[0])


@custom_op("uwr::linear_bwd", mutates_args=(), device_types=_CUDA)
def linear_bwd(dy: Tensor, x: Tensor, weight: Tensor, has_bias: bool) -> Tuple[Tensor, Tensor, Tensor]:
    dy = _c(dy)
    dx = ops.linear_dgrad(dy, _c(weight))
    dw, db = ops.linear_wgrad(dy, _c(x), want_bias=True)
    return dx, dw, db


@linear_bwd.register_fake
def _(dy, x, weight, has_bias):
    return torch.empty_like(x), torch.empty_like(weight), dy.new_empty(weight.shape[0])


def _linear_setup(ctx, inputs, output):
    x, weight, bias = inputs
    ctx.save_for_backward(x, weight)
    ctx.has_bias = bias is not None


def _linear_backward(ctx, dy):
    x, weight = ctx.saved_tensors
    dx, dw, db = torch.ops.uwr.linear_bwd(dy, x, weight, ctx.has_bias)
    return dx, dw, (db if ctx.has_bias else None)


register_autograd("uwr::linear", _linear_backward, setup_context=_linear_setup)


# --------------------------------------------------------------------------------------------- layernorm
@custom_op("uwr::layernorm", mutates_args=(), device_types=_CUDA)
def layernorm(x: Tensor, weight: Tensor, bias: Tensor, eps: float) -> Tuple[Tensor, Tensor, Tensor]:
    """nn.LayerNorm over the last axis of a token matrix (AST.py:521,534) -> (y, mean, rstd)."""
    return ops.layernorm_fwd(_c(x), weight, bias, eps)


@layernorm.register_fake
def _(x, weight, bias, eps):
    return torch.empty_like(x), x.new_empty(x.shape[0]), x.new_empty(x.shape[0])


@custom_op("uwr::layernorm_bwd", mutates_args=(), device_types=_CUDA)
def layernorm_bwd(dy: Tensor, x: Tensor, weight: Tensor, mean: Tensor, rstd: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    return ops.layernorm_bwd(_c(dy), _c(x), weight, mean, rstd)


@layernorm_bwd.register_fake
def _(dy, x, weight, mean, rstd):
    return torch.empty_like(x), torch.empty_like(weight), torch.empty_like(weight)


def _ln_setup(ctx, inputs, output):
    x, weight, bias, eps = inputs
    _, mean, rstd = output
    ctx.save_for_backward(x, weight, mean, rstd)


def _ln_backward(ctx, dy, dmean, drstd):
    x, weight, mean, rstd = ctx.saved_tensors
    dx, dw, db = torch.ops.uwr.layernorm_bwd(dy, x, weight, mean, rstd)
    return dx, dw, db, None


register_autograd("uwr::layernorm", _ln_backward, setup_context=_ln_setup)


# ------------------------------------------------------------------------------------- window attention
@custom_op("uwr::window_attn_sparse", mutates_args=(), device_types=_CUDA)
def window_attn_sparse(qkv: Tensor, bias_table: Tensor, w: Optional[Tensor], B: int, H: int, W: int, heads: int,
                       shift: int, scale: float) -> Tensor:
    """Adaptive sparse 8x8 window attention on a packed (B*H*W, 3C) [q | k | v] token matrix, cyclic shift and mask as
    address arithmetic (WindowAttention_sparse.forward, AST.py:187-219; w = None: plain softmax, AST.py:109-137)."""
    qkv = _c(qkv)
    Cc = qkv.shape[1] // 3
    return ops.window_attn_fwd(qkv, 0, qkv, Cc, 2 * Cc, bias_table, w, B, H, W, heads, Cc // heads, shift, scale)


@window_attn_sparse.register_fake
def _(qkv, bias_table, w, B, H, W, heads, shift, scale):
    return qkv.new_empty(qkv.shape[0], qkv.shape[1] // 3)


@custom_op("uwr::window_attn_sparse_bwd", mutates_args=(), device_types=_CUDA)
def window_attn_sparse_bwd(dout: Tensor, qkv: Tensor, bias_table: Tensor, w: Optional[Tensor], B: int, H: int, W: int,
                           heads: int, shift: int, scale: float) -> Tuple[Tensor, Tensor, Tensor]:
    qkv = _c(qkv)
    Cc = qkv.shape[1] // 3
    dqkv, _, dtable, dw = ops.window_attn_bwd(_c(dout), qkv, 0, qkv, Cc, 2 * Cc, bias_table, w, B, H, W, heads,
                                              Cc // heads, shift, scale)
    return dqkv, dtable, dw


@window_attn_sparse_bwd.register_fake
def _(dout, qkv, bias_table, w, B, H, W, heads, shift, scale):
    return torch.empty_like(qkv), torch.empty_like(bias_table), qkv.new_empty(2)


def _attn_setup(ctx, inputs, output):
    qkv, table, w, B, H, W, heads, shift, scale = inputs
    ctx.save_for_backward(qkv, table, w)
    ctx.meta = (B, H, W, heads, shift, scale)


def _attn_backward(ctx, dout):
    qkv, table, w = ctx.saved_tensors
    dqkv, dtable, dw = torch.ops.uwr.window_attn_sparse_bwd(dout, qkv, table, w, *ctx.meta)
    return dqkv, dtable, (dw if w is not None else None), None, None, None, None, None, None


register_autograd("uwr::window_attn_sparse", _attn_backward, setup_context=_attn_setup)


# ----------------------------------------------------------------------------------- dwconv3x3 + GELU
@custom_op("uwr::dwconv3x3_gelu", mutates_args=(), device_types=_CUDA)
def dwconv3x3_gelu(u: Tensor, weight: Tensor, bias: Tensor, B: int, H: int, W: int) -> Tuple[Tensor, Tensor]:
    """h2 = GELU(dwconv3x3(GELU(u)) + b) on tokens (LeFF, AST.py:312-321) -> (h2, gelu'(conv pre-activation))."""
    u = _c(u)
    dg, h2 = ops.dwconv_gelu_fwd(u, weight, bias, B, H, W, u.shape[1], mode=0, save_v=True, v_is_dgelu=True)
    return h2, dg


@dwconv3x3_gelu.register_fake
def _(u, weight, bias, B, H, W):
    return torch.empty_like(u), torch.empty_like(u)


@custom_op("uwr::dwconv3x3_gelu_bwd", mutates_args=(), device_types=_CUDA)
def dwconv3x3_gelu_bwd(dh2: Tensor, dgelu: Tensor, u: Tensor, weight: Tensor, B: int, H: int, W: int) -> Tuple[Tensor, Tensor, Tensor]:
    u = _c(u)
    dv = _c(dh2) * dgelu        # tiny glue: the fused path folds this product into the GEMM epilogue (UWR_EPI_MUL)
    du, dw, db = ops.dwconv_gelu_bwd(dv, u, weight, B, H, W, u.shape[1])
    return du, dw, db


@dwconv3x3_gelu_bwd.register_fake
def _(dh2, dgelu, u, weight, B, H, W):
    return torch.empty_like(u), torch.empty_like(weight), u.new_empty(u.shape[1])


def _dw_setup(ctx, inputs, output):
    u, weight, bias, B, H, W = inputs
    ctx.save_for_backward(u, weight, output[1])
    ctx.dims = (B, H, W)


def _dw_backward(ctx, dh2, ddg):
    u, weight, dg = ctx.saved_tensors
    du, dw, db = torch.ops.uwr.dwconv3x3_gelu_bwd(dh2, dg, u, weight, *ctx.dims)
    return du, dw, db, None, None, None


register_autograd("uwr::dwconv3x3_gelu", _dw_backward, setup_context=_dw_setup)


# ------------------------------------------------------------------------------------------------ losses
@custom_op("uwr::l1_family_loss", mutates_args=(), device_types=_CUDA)
def l1_family_loss(pred: Tensor, truth: Tensor, kind: str, batch_divisor: int) -> Tuple[Tensor, Tensor]:
    """(loss, dloss/dpred) in one pass; kind in {"L1", "L2", "L1withColor", "charbonnier"} (losses.py:54-81,182-213)."""
    loss, grad = ops.pixel_loss(_c(pred), _c(truth), kind, batch_divisor)
    return loss.view(()), grad


@l1_family_loss.register_fake
def _(pred, truth, kind, batch_divisor):
    return pred.new_empty(()), torch.empty_like(pred)


def _loss_setup(ctx, inputs, output):
    ctx.save_for_backward(output[1])


def _loss_backward(ctx, dloss, dgrad):
    (grad,) = ctx.saved_tensors
    return grad * dloss, None, None, None


register_autograd("uwr::l1_family_loss", _loss_backward, setup_context=_loss_setup)


@custom_op("uwr::charbonnier_loss", mutates_args=(), device_types=_CUDA)
def charbonnier_loss(pred: Tensor, truth: Tensor) -> Tuple[Tensor, Tensor]:
    """CharbonnierLoss (losses.py:182-193): mean(sqrt(d^2 + 1e-6)) and its gradient."""
    loss, grad = ops.pixel_loss(_c(pred), _c(truth), "charbonnier", pred.shape[0])
    return loss.view(()), grad


@charbonnier_loss.register_fake
def _(pred, truth):
    return pred.new_empty(()), torch.empty_like(pred)


register_autograd("uwr::charbonnier_loss", lambda ctx, dl, dg: (ctx.saved_tensors[0] * dl, None),
                  setup_context=_loss_setup)


@custom_op("uwr::ffl_loss", mutates_args=(), device_types=_CUDA)
def ffl_loss(pred: Tensor, truth: Tensor) -> Tuple[Tensor, Tensor]:
    """focal frequency loss (focal_frequency_loss 0.3.0, ctor losses.py:48) and its gradient (shared-memory FFTs)."""
    from .ffl import ffl_value_and_grad
    loss, grad = ffl_value_and_grad(_c(pred), _c(truth))
    return loss.view(()), grad


@ffl_loss.register_fake
def _(pred, truth):
    return pred.new_empty(()), torch.empty_like(pred)


register_autograd("uwr::ffl_loss", lambda ctx, dl, dg: (ctx.saved_tensors[0] * dl, None), setup_context=_loss_setup)


# ------------------------------------------------------------------------------------- frequency mixing
@custom_op("uwr::fft2_real", mutates_args=(), device_types=_CUDA)
def fft2_real(x: Tensor, scale: float) -> Tensor:
    """scale * Re(FFT2 over (H, W)) of real (B, H, W, C) tokens (FDFP, block.py:532-556); self-adjoint."""
    B, H, W, Cc = x.shape
    return ops.dft_real(_c(x), B, H, W, Cc, scale, "hw").view(x.shape)


@fft2_real.register_fake
def _(x, scale):
    return torch.empty_like(x)


def _fft_setup(ctx, inputs, output):
    ctx.scale = inputs[1]


register_autograd("uwr::fft2_real", lambda ctx, dy: (torch.ops.uwr.fft2_real(dy, ctx.scale), None),
                  setup_context=_fft_setup)


@custom_op("uwr::fft_lc_real", mutates_args=(), device_types=_CUDA)
def fft_lc_real(x: Tensor, scale: float) -> Tensor:
    """scale * Re(FFT2 over (L = H*W tokens, C channels)) (EncoderBlock, model.py:72-88); self-adjoint."""
    B, H, W, Cc = x.shape
    return ops.dft_real(_c(x), B, H, W, Cc, scale, "lc").view(x.shape)


@fft_lc_real.register_fake
def _(x, scale):
    return torch.empty_like(x)


register_autograd("uwr::fft_lc_real", lambda ctx, dy: (torch.ops.uwr.fft_lc_real(dy, ctx.scale), None),
                  setup_context=_fft_setup)


# --------------------------------------------------------------------------------------------------- MDTA
@custom_op("uwr::mdta_gram", mutates_args=(), device_types=_CUDA)
def mdta_gram(x: Tensor, y: Tensor, B: int, L: int, heads: int) -> Tuple[Tensor, Tensor, Tensor]:
    """per-head X^T Y over the L tokens of each image + squared column norms (SpectralTransformer.py:99-100)."""
    Cc = x.shape[1]
    return ops.mdta_gram(_c(x), 0, _c(y), 0, B, L, heads, Cc // heads, want_sq=True)


@mdta_gram.register_fake
def _(x, y, B, L, heads):
    Cc = x.shape[1]
    c = Cc // heads
    return x.new_empty(B, heads, c, c), x.new_empty(B, Cc), x.new_empty(B, Cc)


@custom_op("uwr::mdta_apply", mutates_args=(), device_types=_CUDA)
def mdta_apply(x: Tensor, m: Tensor, transpose: bool, B: int, L: int) -> Tensor:
    """out[b,l,h*c+i] = sum_j M[b,h,i,j] x[b,l,h*c+j] (attn @ v, SpectralTransformer.py:101,109,113)."""
    heads, c = m.shape[1], m.shape[2]
    return ops.mdta_apply(_c(x), 0, _c(m), B, L, heads, c, transpose=transpose)


@mdta_apply.register_fake
def _(x, m, transpose, B, L):
    return torch.empty_like(x)


def _gram_setup(ctx, inputs, output):
    x, y, B, L, heads = inputs
    ctx.save_for_backward(x, y)
    ctx.meta = (B, L)


def _gram_backward(ctx, dG, dsqx, dsqy):
    x, y = ctx.saved_tensors
    B, L = ctx.meta
    dx = torch.ops.uwr.mdta_apply(y, dG, False, B, L) + 2 * dsqx.repeat_interleave(L, 0) * x
    dy = torch.ops.uwr.mdta_apply(x, dG, True, B, L) + 2 * dsqy.repeat_interleave(L, 0) * y
    return dx, dy, None, None, None


register_autograd("uwr::mdta_gram", _gram_backward, setup_context=_gram_setup)


def _apply_setup(ctx, inputs, output):
    x, m, transpose, B, L = inputs
    ctx.save_for_backward(x, m)
    ctx.meta = (transpose, B, L)


def _apply_backward(ctx, dout):
    x, m = ctx.saved_tensors
    transpose, B, L = ctx.meta
    heads = m.shape[1]
    dx = torch.ops.uwr.mdta_apply(dout, m, not transpose, B, L)
    G = torch.ops.uwr.mdta_gram(dout, x, B, L, heads)[0]
    return dx, (G.transpose(-2, -1) if transpose else G), None, None, None


register_autograd("uwr::mdta_apply", _apply_backward, setup_context=_apply_setup)


# -------------------------------------------------------------------------- spectral up-sampler pieces
@custom_op("uwr::polar_split", mutates_args=(), device_types=_CUDA)
def polar_split(f: Tensor) -> Tuple[Tensor, Tensor]:
    """(abs, angle) of interleaved complex (..., 2) (SpectralTransformer.py:176-177)."""
    return ops.polar_split_fwd(_c(f))


@polar_split.register_fake
def _(f):
    return f.new_empty(f.shape[:-1]), f.new_empty(f.shape[:-1])


@custom_op("uwr::polar_split_bwd", mutates_args=(), device_types=_CUDA)
def polar_split_bwd(f: Tensor, dmag: Tensor, dpha: Tensor) -> Tensor:
    return ops.polar_split_bwd(_c(f), _c(dmag), _c(dpha))


@polar_split_bwd.register_fake
def _(f, dmag, dpha):
    return torch.empty_like(f)


register_autograd("uwr::polar_split",
                  lambda ctx, dm, dp: torch.ops.uwr.polar_split_bwd(ctx.saved_tensors[0], dm, dp),
                  setup_context=lambda ctx, inputs, output: ctx.save_for_backward(inputs[0]))


@custom_op("uwr::polar_join", mutates_args=(), device_types=_CUDA)
def polar_join(mag: Tensor, pha: Tensor) -> Tensor:
    """mag * exp(i pha) as interleaved complex (SpectralTransformer.py:181-183)."""
    return ops.polar_join_fwd(_c(mag), _c(pha))


@polar_join.register_fake
def _(mag, pha):
    return mag.new_empty(tuple(mag.shape) + (2,))


@custom_op("uwr::polar_join_bwd", mutates_args=(), device_types=_CUDA)
def polar_join_bwd(mag: Tensor, pha: Tensor, dz: Tensor) -> Tuple[Tensor, Tensor]:
    return ops.polar_join_bwd(_c(mag), _c(pha), _c(dz))


@polar_join_bwd.register_fake
def _(mag, pha, dz):
    return torch.empty_like(mag), torch.empty_like(mag)


register_autograd("uwr::polar_join",
                  lambda ctx, dz: torch.ops.uwr.polar_join_bwd(ctx.saved_tensors[0], ctx.saved_tensors[1], dz),
                  setup_context=lambda ctx, inputs, output: ctx.save_for_backward(inputs[0], inputs[1]))


def spectral_upsample(t, B, H, W, amp_fuse, pha_fuse, post_weight, post_bias):
    """SpectralTransformer.UpSample.forward (SpectralTransformer.py:174-188) composed from torch.ops.uwr pieces and the
    FFT autograd function: tokens (B*H*W, C) -> tokens (B*2H*2W, C_out).  amp_fuse / pha_fuse: callables on tokens."""
    from . import fn
    Cc = t.shape[1]
    f = fn.Fft2Fn.apply(t.view(B, H, W, Cc), B, H, W, Cc, False, False, 1.0)
    mag0, pha0 = torch.ops.uwr.polar_split(f)
    mag = amp_fuse(mag0.view(B * H * W, Cc))
    pha = pha_fuse(pha0.view(B * H * W, Cc))
    z = torch.ops.uwr.polar_join(mag, pha).view(B, H, W, Cc, 2)
    small = fn.CAbsFn.apply(fn.Fft2Fn.apply(z, B, H, W, Cc, True, True, 1.0 / (H * W)))
    y = torch.ops.uwr.linear(small.view(B * H * W, Cc), post_weight, post_bias)
    return fn.EvenScatterFn.apply(y, post_bias, B, H, W)


# ---------------------------------------------------------------------------------- fused window block
@custom_op("uwr::fused_window_block", mutates_args=(), device_types=_CUDA)
def fused_window_block(x: Tensor, n1w: Tensor, n1b: Tensor, wq: Tensor, bq: Tensor, wkv: Tensor, bkv: Tensor,
                       table: Tensor, w: Optional[Tensor], wp: Tensor, bp: Tensor, H: int, W: int, heads: int,
                       shift: int) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor, Tensor]:
    """x + proj(W-MSA(LN1(x))) on tokens x (B, L, C): LayerNorm, packed q|kv projection, shifted-window adaptive sparse
    attention, output projection with the residual add in the GEMM epilogue (TransformerBlock.forward, AST.py:590-619).
    Returns (out, y1, mean, rstd, qkv, o); the last five are the activations the backward op consumes."""
    x = _c(x)
    B, L, Cc = x.shape
    x2 = x.view(B * L, Cc)
    y1, mean, rstd = ops.layernorm_fwd(x2, n1w, n1b)
    qkv = ops.linear(y1, _c(wq), bq, weight2=_c(wkv), bias2=bkv)
    hd = Cc // heads
    o = ops.window_attn_fwd(qkv, 0, qkv, Cc, 2 * Cc, table, w, B, H, W, heads, hd, shift, hd ** -0.5)
    out = ops.linear(o, _c(wp), bp, residual=x2)
    return out.view(B, L, Cc), y1, mean, rstd, qkv, o


@fused_window_block.register_fake
def _(x, n1w, n1b, wq, bq, wkv, bkv, table, w, wp, bp, H, W, heads, shift):
    B, L, Cc = x.shape
    M = B * L
    return (torch.empty_like(x), x.new_empty(M, Cc), x.new_empty(M), x.new_empty(M), x.new_empty(M, 3 * Cc),
            x.new_empty(M, Cc))


@custom_op("uwr::fused_window_block_bwd", mutates_args=(), device_types=_CUDA)
def fused_window_block_bwd(dout: Tensor, x: Tensor, n1w: Tensor, wq: Tensor, wkv: Tensor, table: Tensor,
                           w: Optional[Tensor], wp: Tensor, y1: Tensor, mean: Tensor, rstd: Tensor, qkv: Tensor,
                           o: Tensor, H: int, W: int, heads: int,
                           shift: int) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor,
                                                Tensor, Tensor]:
    B, L, Cc = x.shape
    d = _c(dout).view(B * L, Cc)
    x2 = _c(x).view(B * L, Cc)
    hd = Cc // heads
    d_o = ops.linear_dgrad(d, _c(wp))
    dwp, dbp = ops.linear_wgrad(d, o)
    dqkv, _, dtable, dw = ops.window_attn_bwd(d_o, qkv, 0, qkv, Cc, 2 * Cc, table, w, B, H, W, heads, hd, shift,
                                              hd ** -0.5)
    dy1 = ops.linear_dgrad(dqkv, _c(wq), weight2=_c(wkv))
    dwqkv, dbqkv = ops.linear_wgrad(dqkv, y1)
    dx, dg, db = ops.layernorm_bwd(dy1, x2, n1w, mean, rstd, dres=d)
    return (dx.view(B, L, Cc), dg, db, dwqkv[:Cc].clone(), dbqkv[:Cc].clone(), dwqkv[Cc:].clone(), dbqkv[Cc:].clone(),
            dtable, dw, dwp, dbp)


@fused_window_block_bwd.register_fake
def _(dout, x, n1w, wq, wkv, table, w, wp, y1, mean, rstd, qkv, o, H, W, heads, shift):
    Cc = x.shape[2]
    e = torch.empty_like
    return (e(x), e(n1w), e(n1w), e(wq), x.new_empty(Cc), e(wkv), x.new_empty(2 * Cc), e(table), x.new_empty(2), e(wp),
            x.new_empty(Cc))


def _fwb_setup(ctx, inputs, output):
    x, n1w, n1b, wq, bq, wkv, bkv, table, w, wp, bp, H, W, heads, shift = inputs
    _, y1, mean, rstd, qkv, o = output
    ctx.save_for_backward(x, n1w, wq, wkv, table, w, wp, y1, mean, rstd, qkv, o)
    ctx.meta = (H, W, heads, shift)


def _fwb_backward(ctx, dout, *unused):
    x, n1w, wq, wkv, table, w, wp, y1, mean, rstd, qkv, o = ctx.saved_tensors
    g = torch.ops.uwr.fused_window_block_bwd(dout, x, n1w, wq, wkv, table, w, wp, y1, mean, rstd, qkv, o, *ctx.meta)
    dx, dg, db, dwq, dbq, dwkv, dbkv, dtable, dw, dwp, dbp = g
    return dx, dg, db, dwq, dbq, dwkv, dbkv, dtable, (dw if w is not None else None), dwp, dbp, None, None, None, None


register_autograd("uwr::fused_window_block", _fwb_backward, setup_context=_fwb_setup)

OPS = ("linear", "layernorm", "window_attn_sparse", "dwconv3x3_gelu", "l1_family_loss", "charbonnier_loss", "ffl_loss",
       "fft2_real", "fft_lc_real", "mdta_gram", "mdta_apply", "polar_split", "polar_join", "fused_window_block")
