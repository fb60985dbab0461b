"""Generic autograd functions over the uwr kernels (used by the NewBig/FRFN family, where blocks are
composed with a few ATen pieces — FFTs, pixel shuffles — that are not hand-written yet):

  LinearFn      y = x W^T + b on token matrices (tcgen05 path when `rounded` promises TF32 operands)
  LayerNormFn   nn.LayerNorm over the channel axis
  Conv3x3Fn     dense 3x3 s1 p1 convolution on tokens = im2col + GEMM
  AttnFn        sparse window attention incl. q / kv / output projections (self or cross)
"""
import torch
from torch.autograd.function import once_differentiable

from . import ops


def _c(t):
    return t if t.is_contiguous() else t.contiguous()


def _rows(x):
    """2-D row-major view (row stride may exceed the width: column slices of token matrices are fine)."""
    if x.dim() == 2 and x.stride(1) == 1 and x.stride(0) % 4 == 0 and x.data_ptr() % 16 == 0:
        return x
    x = _c(x)
    return x.view(-1, x.shape[-1])


class LinearFn(torch.autograd.Function):
    """In tf32 mode every product runs on the tcgen05 kernel: operands that are not known to be
    TF32-rounded (`rounded=False`) get a rounded copy first (one extra pass, still far cheaper than
    the legacy mma.sync kernel at the tall-skinny shapes of these models)."""

    @staticmethod
    def forward(ctx, x, w, b, rounded):
        x2 = _rows(x)
        fast = ops.fast_path() and x2.shape[1] % 4 == 0 and w.shape[0] % 4 == 0
        if fast and not rounded:
            # a rounded copy does not pay off on tiny token matrices (a launch for a few KB); those take the legacy
            # kernel, which rounds its fragments in flight.  From 8 192 rows on the tcgen05 kernel wins even with the
            # copy (NT M65536 N512 K256: 0.18 ms legacy vs ~0.05 ms)
            fast = x2.shape[0] >= 8192
            if fast:
                x2 = ops.scale_round(x2, x2.shape[1])
        y = ops.linear(x2, ops.rounded_weight(w) if fast else w, b, t5=fast)
        ctx.save_for_backward(x2, w)
        ctx.meta = (x.shape, b is not None, fast)
        return y.view(*x.shape[:-1], w.shape[0])

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        x2, w = ctx.saved_tensors
        shape, has_b, fast = ctx.meta
        d = _c(dy).view(-1, w.shape[0])
        db = None
        if fast:
            d, db = ops.scale_round_colsum(d, d.shape[1]) if has_b else (ops.scale_round(d, d.shape[1]), None)
            dx = ops.linear_dgrad(d, ops.rounded_weight(w), t5=True) if ctx.needs_input_grad[0] else None
            dw, _ = ops.linear_wgrad(d, x2, want_bias=False, t5=True)
        else:
            dx = ops.linear_dgrad(d, w) if ctx.needs_input_grad[0] else None
            dw, db = ops.linear_wgrad(d, x2, want_bias=has_b)
        return (dx.view(shape) if dx is not None else None), dw, db, None


def linear(x, w, b=None, rounded=False):
    return LinearFn.apply(x, w, b, rounded)


class LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, g, b):
        x = _c(x)
        x2 = x.view(-1, x.shape[-1])
        y, mean, rstd = ops.layernorm_fwd(x2, g, b)
        ctx.save_for_backward(x2, g, mean, rstd)
        ctx.shape = x.shape
        return y.view(x.shape)

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        x2, g, mean, rstd = ctx.saved_tensors
        dx, dg, db = ops.layernorm_bwd(_c(dy).view(x2.shape), x2, g, mean, rstd)
        return dx.view(ctx.shape), dg, db


def layernorm(x, mod):
    return LayerNormFn.apply(x, mod.weight, mod.bias)


class Conv3x3Fn(torch.autograd.Function):
    """tokens (B, H*W, Cin) -> (B, H*W, Cout); weight (Cout, Cin, 3, 3), Cin % 4 == 0."""

    @staticmethod
    def forward(ctx, x, weight, bias, H, W):
        x = _c(x)
        B, L, Cin = x.shape
        Cout = weight.shape[0]
        wmat = ops.scale_round(weight.permute(0, 2, 3, 1).reshape(Cout, 9 * Cin), 9 * Cin)
        ctx.meta = (B, H, W, Cin, Cout, bias is not None)
        ctx.implicit = False
        if ops.fast_path() and Cin % 32 == 0:
            # implicit GEMM: TMA builds the im2col tiles inside the tcgen05 kernel from a TF32-rounded copy of x
            xr = ops.scale_round(x.view(B * L, Cin), Cin)
            y = ops.conv_gemm_fwd(xr, wmat, bias, B, H, W, 3, 3, 1, 1)
            if y is not None:
                ctx.save_for_backward(xr, wmat, weight)
                ctx.implicit = True
                return y.view(B, L, Cout)
        col = ops.im2col_3x3(x.view(B * L, Cin), B, H, W, Cin)
        y = ops.linear(col, wmat, bias, t5=True)
        ctx.save_for_backward(x, wmat, weight)
        return y.view(B, L, Cout)

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        x, wmat, weight = ctx.saved_tensors
        B, H, W, Cin, Cout, has_b = ctx.meta
        d = _c(dy).view(-1, Cout)
        fast = ops.fast_path() and Cout % 4 == 0 and d.shape[0] >= 4096
        db = dwm = dx = None
        if fast:  # tcgen05 path: TF32-rounded copy of the cotangent (x / col and wmat already are rounded)
            d, db = ops.scale_round_colsum(d, Cout) if has_b else (ops.scale_round(d, Cout), None)
            if ctx.implicit:
                dwm = ops.conv_gemm_wgrad(d, x.view(-1, Cin), B, H, W, 3, 3, 1, 1)
                if ctx.needs_input_grad[0] and Cout % 32 == 0:
                    # data gradient = the same convolution of dy with the flipped, transposed weights
                    wflip = ops.scale_round(weight.detach().flip(2, 3).permute(1, 2, 3, 0).reshape(Cin, 9 * Cout), 9 * Cout)
                    dx = ops.conv_gemm_fwd(d, wflip, None, B, H, W, 3, 3, 1, 1)
        col = None
        if dwm is None:
            col = ops.im2col_3x3(x.view(-1, Cin), B, H, W, Cin)  # recomputed (9x the input)
            if fast:
                dwm, _ = ops.linear_wgrad(d, col, want_bias=False, t5=True)
            else:
                dwm, db = ops.linear_wgrad(d, col, want_bias=has_b)
            del col
        if dx is None and ctx.needs_input_grad[0]:
            dcol = ops.linear_dgrad(d, wmat, t5=fast)
            dx = ops._empty((B * H * W, Cin), x)
            ops.col2im_3x3(dcol, dx, B, H, W, Cin)
        dw = dwm.reshape(Cout, 3, 3, Cin).permute(0, 3, 1, 2).contiguous()
        return (dx.view(B, H * W, Cin) if dx is not None else None), dw, db, None, None


class AttnFn(torch.autograd.Function):
    """WindowAttention_Sparse (block.py:325-370) on un-partitioned token matrices: q projection of xq,
    kv projection of xkv (cross, block.py:185-188) or of xq (self), 8x8 window attention (windows and
    shift are address arithmetic), output projection."""

    @staticmethod
    def forward(ctx, xq, xkv, wq, bq, wkv, bkv, table, wparam, wp, bp, B, H, W, heads, shift):
        xq = _c(xq)
        M, Cc = xq.shape
        hd = Cc // heads
        scale = hd ** -0.5
        # tf32 mode: every projection on the tcgen05 kernel; inputs of unknown provenance get a TF32-rounded copy
        # (one extra pass over a C-wide matrix, far cheaper than the legacy kernel at these tall-skinny shapes)
        fast = ops.fast_path() and Cc % 32 == 0 and M >= 4096
        if fast:
            xq = ops.scale_round(xq, Cc)
        if xkv is None:
            if fast:
                w, bias = ops.packed_qkv(wq, bq, wkv, bkv)
                qkv = ops.linear(xq, w, bias, t5=True, round_out=True)
            else:
                qkv = ops.linear(xq, wq, bq, weight2=wkv, bias2=bkv)
            q_buf, q_off, kv_buf, k_off, v_off = qkv, 0, qkv, Cc, 2 * Cc
        else:
            xkv = _c(xkv)
            if fast:
                xkv = ops.scale_round(xkv, xkv.shape[1])
                q_buf = ops.linear(xq, ops.rounded_weight(wq), bq, t5=True, round_out=True)
                kv_buf = ops.linear(xkv, ops.rounded_weight(wkv), bkv, t5=True, round_out=True)
            else:
                q_buf = ops.linear(xq, wq, bq)
                kv_buf = ops.linear(xkv, wkv, bkv)
            q_off, k_off, v_off = 0, 0, Cc
        o = ops.window_attn_fwd(q_buf, q_off, kv_buf, k_off, v_off, table, wparam, B, H, W, heads, hd, shift, scale,
                                rounded=fast)
        y = ops.linear(o, ops.rounded_weight(wp), bp, t5=True)
        ctx.save_for_backward(xq, xkv, wq, wkv, table, wparam, wp, q_buf, kv_buf, o, bq, bkv)
        ctx.meta = (B, H, W, heads, hd, shift, scale, Cc, bq is not None, fast)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        xq, xkv, wq, wkv, table, wparam, wp, q_buf, kv_buf, o, bq, bkv = ctx.saved_tensors
        B, H, W, heads, hd, shift, scale, Cc, has_b, fast = ctx.meta
        d = _c(dy)
        if fast:   # rounded cotangent + proj bias gradient in one pass; o was rounded by the attention kernel
            d, dbp = ops.scale_round_colsum(d, Cc)
            d_o = ops.linear_dgrad(d, ops.rounded_weight(wp), t5=True, round_out=True)
            dwp, _ = ops.linear_wgrad(d, o, want_bias=False, t5=True)
        else:
            d_o = ops.linear_dgrad(d, wp)
            dwp, dbp = ops.linear_wgrad(d, o)
        cross = xkv is not None
        q_off, k_off, v_off = (0, 0, Cc) if cross else (0, Cc, 2 * Cc)
        # bias gradients of the projections = column sums of dq | dk | dv, accumulated inside the attention backward
        csq = cskv = None
        if fast and has_b:
            csq = ops._empty((Cc,) if cross else (3 * Cc,), q_buf)
            cskv = ops._empty((2 * Cc,), q_buf) if cross else csq
        dq_buf, dkv_buf, dtable, dw = ops.window_attn_bwd(d_o, q_buf, q_off, kv_buf, k_off, v_off, table, wparam,
                                                          B, H, W, heads, hd, shift, scale, rounded=fast,
                                                          colsum_q=csq, colsum_kv=cskv)
        if fast:   # dq / dk / dv leave the attention kernel TF32-rounded; xq / xkv were saved rounded
            if cross:
                dxq = ops.linear_dgrad(dq_buf, ops.rounded_weight(wq), t5=True)
                dwq, _ = ops.linear_wgrad(dq_buf, xq, want_bias=False, t5=True)
                dxkv = ops.linear_dgrad(dkv_buf, ops.rounded_weight(wkv), t5=True)
                dwkv, _ = ops.linear_wgrad(dkv_buf, xkv, want_bias=False, t5=True)
                dbq, dbkv = csq, cskv
            else:
                w, _ = ops.packed_qkv(wq, bq, wkv, bkv)
                dxq = ops.linear_dgrad(dq_buf, w, t5=True)
                dwqkv, _ = ops.linear_wgrad(dq_buf, xq, want_bias=False, t5=True)
                dbqkv = csq
                dwq, dwkv = dwqkv[:Cc], dwqkv[Cc:]
                dbq, dbkv = (dbqkv[:Cc], dbqkv[Cc:]) if has_b else (None, None)
                dxkv = None
        elif cross:
            dxq = ops.linear_dgrad(dq_buf, wq)
            dwq, dbq = ops.linear_wgrad(dq_buf, xq, want_bias=has_b)
            dxkv = ops.linear_dgrad(dkv_buf, wkv)
            dwkv, dbkv = ops.linear_wgrad(dkv_buf, xkv, want_bias=has_b)
        else:
            dxq = ops.linear_dgrad(dq_buf, wq, weight2=wkv)
            dwqkv, dbqkv = ops.linear_wgrad(dq_buf, xq, want_bias=has_b)
            dwq, dwkv = dwqkv[:Cc], dwqkv[Cc:]
            dbq, dbkv = (dbqkv[:Cc], dbqkv[Cc:]) if has_b else (None, None)
            dxkv = None
        return (dxq, dxkv, dwq, dbq, dwkv, dbkv, dtable, dw if wparam is not None else None, dwp, dbp,
                None, None, None, None, None)


class PlainDWConvFn(torch.autograd.Function):
    """depthwise 3x3 conv (pad 1, no bias, no activation) on tokens (B*H*W, C): MDTA qkv_conv / kv_conv and
    GDFN conv of the SpectralTransformer (SpectralTransformer.py:82,89,123)."""

    @staticmethod
    def forward(ctx, u, weight, B, H, W):
        u = _c(u)
        Ch = u.shape[1]
        _, y = ops.dwconv_gelu_fwd(u, weight, None, B, H, W, Ch, mode=2, save_v=False)
        ctx.save_for_backward(u, weight)
        ctx.dims = (B, H, W, Ch)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        u, weight = ctx.saved_tensors
        B, H, W, Ch = ctx.dims
        du, dw, _ = ops.dwconv_gelu_bwd(_c(dy), u, weight, B, H, W, Ch, plain=True)
        return du, dw, None, None, None


class LinearDWConvFn(torch.autograd.Function):
    """1x1 convolution followed by a depthwise 3x3 (MDTA qkv -> qkv_conv, kv -> kv_conv and GDFN project_in -> conv,
    SpectralTransformer.py:81-89,121-123) as ONE autograd node: the depthwise backward already emits its du
    TF32-rounded in tf32 mode, so the 1x1's data / weight gradient GEMMs read it as it is -- as two nodes the Linear
    backward cannot know that and spends a pass on a rounded copy of the widest gradient of the block."""

    @staticmethod
    def forward(ctx, x, w, wdw, B, H, W, rounded):
        x2 = _rows(x)
        Ch = w.shape[0]
        fast = ops.fast_path() and x2.shape[1] % 4 == 0 and Ch % 4 == 0
        if fast and not rounded:
            fast = x2.shape[0] >= 8192        # as LinearFn: a rounded copy does not pay off on tiny token matrices
            if fast:
                x2 = ops.scale_round(x2, x2.shape[1])
        u = ops.linear(x2, ops.rounded_weight(w) if fast else w, None, t5=fast)
        _, y = ops.dwconv_gelu_fwd(u, wdw, None, B, H, W, Ch, mode=2, save_v=False)
        ctx.save_for_backward(x2, w, u, wdw)
        ctx.meta = (B, H, W, Ch, fast, x.shape)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        x2, w, u, wdw = ctx.saved_tensors
        B, H, W, Ch, fast, xshape = ctx.meta
        du, dwdw, _ = ops.dwconv_gelu_bwd(_c(dy), u, wdw, B, H, W, Ch, plain=True)   # rounded in tf32 mode
        del u
        if fast:
            dx = ops.linear_dgrad(du, ops.rounded_weight(w), t5=True) if ctx.needs_input_grad[0] else None
            dw, _ = ops.linear_wgrad(du, x2, want_bias=False, t5=True)
        else:
            dx = ops.linear_dgrad(du, w) if ctx.needs_input_grad[0] else None
            dw, _ = ops.linear_wgrad(du, x2, want_bias=False)
        return (dx.view(xshape) if dx is not None else None), dw, dwdw, None, None, None, None


class DftRealFn(torch.autograd.Function):
    """y = scale * Re(FFT2(x)) over the spatial axes ('hw', FDFP block.py:532-556) or over
    (tokens, channels) ('lc', EncoderBlock model.py:72-88).  x -> Re(F x) is symmetric, so the
    backward is the same kernel on the cotangent."""

    @staticmethod
    def forward(ctx, x, B, H, W, C, scale, axes):
        ctx.args = (B, H, W, C, scale, axes)
        return ops.dft_real(x, B, H, W, C, scale, axes).view(x.shape)

    @staticmethod
    def backward(ctx, dy):
        B, H, W, C, scale, axes = ctx.args
        return ops.dft_real(_c(dy), B, H, W, C, scale, axes).view(dy.shape), None, None, None, None, None, None


def dft_real(x, B, H, W, C, scale, axes):
    return DftRealFn.apply(x, B, H, W, C, scale, axes)


class Fft2Fn(torch.autograd.Function):
    """Complex FFT2 over the spatial axes of a (B, H, W, C) real or (B, H, W, C, 2) interleaved-complex
    tensor on the shared-memory FFT passes (csrc/fft.cu); returns (B, H, W, C, 2).
    y = scale * F x (forward) or scale * conj(F) x (inverse); the adjoint of one is the other, and for
    a real input the cotangent is the real part."""

    @staticmethod
    def forward(ctx, x, B, H, W, C, in_complex, inverse, scale):
        ctx.args = (B, H, W, C, in_complex, inverse, scale)
        return ops.fft2_hw(x, B, H, W, C, in_complex, inverse, scale)

    @staticmethod
    def backward(ctx, dy):
        B, H, W, C, in_complex, inverse, scale = ctx.args
        dx = ops.fft2_hw(_c(dy), B, H, W, C, True, not inverse, scale)
        if not in_complex:
            dx = dx[..., 0].contiguous()
        return dx, None, None, None, None, None, None, None


def _mdta_matrices(G, sq_q, sq_k, temperature, heads):
    """per-head Gram blocks G (B, h, c, c) of q^T k and squared norms (B, C) of q, k -> channel-attention
    matrices softmax(normalize(q) normalize(k)^T * temperature) (SpectralTransformer.py:97-101); tiny batched
    ATen ops, differentiated by autograd inside MDTAAttnFn."""
    B, h, c, _ = G.shape
    nq = sq_q.clamp_min(0).sqrt().clamp_min(1e-12).view(B, h, c, 1)      # F.normalize eps (line 99)
    nk = sq_k.clamp_min(0).sqrt().clamp_min(1e-12).view(B, h, 1, c)
    return torch.softmax(G / (nq * nk) * temperature.view(1, heads, 1, 1), dim=-1)


class MDTAAttnFn(torch.autograd.Function):
    """MDTA channel attention (SpectralTransformer.py:92-109) for a whole batch: qkv (B*L, 3C) tokens ->
    out = attn @ v (B*L, C) and the attention matrices A (B, heads, c, c) (re-used for the frequency
    branch's `attn @ vf`).  Two batched per-head kernels (csrc/mdta.cu): uwr_mdta_gram (q^T k and the L2 norms
    in one pass over the tokens) and uwr_mdta_apply; the (heads, c, c)-sized normalise / softmax algebra runs
    batched and is differentiated by autograd on those tiny tensors only.  The backward is the same two
    kernels: dA = gram(dout, v), dv = apply(dout, A^T), dq = apply(k, dG) + 2 dsq_q q, dk = apply(q, dG^T) + 2 dsq_k k."""

    @staticmethod
    def forward(ctx, qkv, temperature, B, L, C, heads):
        qkv = _c(qkv)
        c = C // heads
        G, sq_q, sq_k = ops.mdta_gram(qkv, 0, qkv, C, B, L, heads, c, want_sq=True)
        with torch.enable_grad():
            leaves = tuple(t.detach().requires_grad_() for t in (G, sq_q, sq_k, temperature))
            A = _mdta_matrices(*leaves, heads)
        Ad = A.detach().contiguous()
        out = ops.mdta_apply(qkv, 2 * C, Ad, B, L, heads, c)
        ctx.save_for_backward(qkv, Ad)
        ctx.graph = (leaves, A)
        ctx.meta = (B, L, C, heads, c)
        return out, Ad

    @staticmethod
    @once_differentiable
    def backward(ctx, dout, dA_ext):
        qkv, Ad = ctx.saved_tensors
        leaves, A = ctx.graph
        B, L, C, heads, c = ctx.meta
        dout = _c(dout)
        dqkv = torch.empty_like(qkv)
        dA, _, _ = ops.mdta_gram(dout, 0, qkv, 2 * C, B, L, heads, c)                        # dout^T v
        ops.mdta_apply(dout, 0, Ad, B, L, heads, c, transpose=True, out=dqkv, ocol=2 * C)    # dv = A^T dout
        if dA_ext is not None:
            dA = dA + dA_ext
        with torch.enable_grad():
            dG, dsq_q, dsq_k, dt = torch.autograd.grad(A, leaves, dA)
        dG = _c(dG)
        ops.mdta_apply(qkv, C, dG, B, L, heads, c, yd=qkv, ycol=0, diag=_c(2 * dsq_q), out=dqkv, ocol=0)
        ops.mdta_apply(qkv, 0, dG, B, L, heads, c, transpose=True, yd=qkv, ycol=C, diag=_c(2 * dsq_k), out=dqkv,
                       ocol=C)
        ctx.graph = None
        return dqkv, dt.view_as(leaves[3]), None, None, None, None


class ChannelApplyFn(torch.autograd.Function):
    """out[b] = A[b] applied per head to the channels x[b][:, col:col+C] of token slabs x (B*L, ld); A (B, heads, c, c)."""

    @staticmethod
    def forward(ctx, x, A, col, B, L):
        x = _c(x)
        A = _c(A)
        heads, c = A.shape[1], A.shape[2]
        out = ops.mdta_apply(x, col, A, B, L, heads, c)
        ctx.save_for_backward(x, A)
        ctx.meta = (col, B, L, heads, c)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        x, A = ctx.saved_tensors
        col, B, L, heads, c = ctx.meta
        dout = _c(dout)
        dx = torch.zeros_like(x)
        dA, _, _ = ops.mdta_gram(dout, 0, x, col, B, L, heads, c)
        ops.mdta_apply(dout, 0, A, B, L, heads, c, transpose=True, out=dx, ocol=col)
        return dx, dA, None, None, None


class GeluMulFn(torch.autograd.Function):
    """GDFN gate (SpectralTransformer.py:126-129): gelu(t[:, :h]) * t[:, h:2h] in one pass (no autograd slicing)."""

    @staticmethod
    def forward(ctx, t, h):
        t = _c(t)
        out = ops._empty((t.shape[0], h), t)
        ops._run("uwr_gelu_mul_fwd", f"rows{t.shape[0]} h{h}", 12 * t.shape[0] * h, 0.0, ops._ptr(t), t.stride(0),
                 ops._ptr(out), t.shape[0], h)
        ctx.save_for_backward(t)
        ctx.h = h
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        (t,) = ctx.saved_tensors
        h = ctx.h
        dt = torch.empty_like(t) if t.shape[1] == 2 * h else torch.zeros_like(t)
        ops._run("uwr_gelu_mul_bwd", f"rows{t.shape[0]} h{h}", 20 * t.shape[0] * h, 0.0, ops._ptr(_c(dout)), ops._ptr(t),
                 t.stride(0), ops._ptr(dt), t.shape[0], h)
        return dt, None


# ---- SpectralTransformer.UpSample elementwise chain (SpectralTransformer.py:174-188) on csrc/spectral_ew.cu -----------
class PolarSplitFn(torch.autograd.Function):
    """(abs(f), angle(f)) of an interleaved-complex tensor f (..., 2) in one pass."""

    @staticmethod
    def forward(ctx, f):
        f = _c(f)
        ctx.save_for_backward(f)
        return ops.polar_split_fwd(f)

    @staticmethod
    @once_differentiable
    def backward(ctx, dmag, dpha):
        (f,) = ctx.saved_tensors
        return ops.polar_split_bwd(f, _c(dmag), _c(dpha))


class PolarJoinFn(torch.autograd.Function):
    """z = mag * exp(i pha) as interleaved complex (..., 2)."""

    @staticmethod
    def forward(ctx, mag, pha):
        mag, pha = _c(mag), _c(pha)
        ctx.save_for_backward(mag, pha)
        return ops.polar_join_fwd(mag, pha)

    @staticmethod
    @once_differentiable
    def backward(ctx, dz):
        mag, pha = ctx.saved_tensors
        return ops.polar_join_bwd(mag, pha, _c(dz))


class CAbsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z):
        z = _c(z)
        ctx.save_for_backward(z)
        return ops.cabs_fwd(z)

    @staticmethod
    @once_differentiable
    def backward(ctx, da):
        (z,) = ctx.saved_tensors
        return ops.cabs_bwd(z, _c(da))


class GeluFn(torch.autograd.Function):
    """exact (erf) GELU; `rounded`: the result feeds a GEMM (TF32-rounded at the store in single-pass mode)."""

    @staticmethod
    def forward(ctx, x, rounded=False):
        x = _c(x)
        ctx.save_for_backward(x)
        return ops.gelu_fwd(x, rounded)

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        return ops.gelu_bwd(x, _c(dy)), None


class LeakyReluFn(torch.autograd.Function):
    """LeakyReLU(slope > 0); the output is what is saved.  `rounded`: the result feeds a GEMM."""

    @staticmethod
    def forward(ctx, x, slope, rounded=False):
        y = ops.leaky_relu_fwd(_c(x), slope, rounded)
        ctx.save_for_backward(y)
        ctx.slope = slope
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        return ops.leaky_relu_bwd(y, _c(dy), ctx.slope), None, None


class EvenScatterFn(torch.autograd.Function):
    """tokens y (B*H*W, C) -> tokens (B*2H*2W, C): y on the even pixels, `bias` on all the others."""

    @staticmethod
    def forward(ctx, y, bias, B, H, W):
        y = _c(y)
        ctx.dims = (B, H, W, y.shape[1])
        return ops.even_scatter(y, bias, B, H, W, y.shape[1])

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        B, H, W, Cc = ctx.dims
        dout = _c(dout)
        dy = ops.even_gather(dout, B, H, W, Cc)
        dbias = ops.colsum(dout, Cc) - ops.colsum(dy, Cc)       # the gradient of the non-even pixels
        return dy, dbias, None, None, None


class PixelShuffleFn(torch.autograd.Function):
    """nn.PixelShuffle(2) on tokens (B*H*W, 4C) -> (B*2H*2W, C); the backward is PixelUnshuffle."""

    @staticmethod
    def forward(ctx, t, B, H, W):
        Cc = t.shape[1] // 4
        ctx.dims = (B, H, W, Cc)
        return ops.pixel_shuffle2(_c(t), B, H, W, Cc)

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        B, H, W, Cc = ctx.dims
        return ops.pixel_unshuffle2(_c(dout), B, H, W, Cc), None, None, None


class PixelUnshuffleFn(torch.autograd.Function):
    """nn.PixelUnshuffle(2) on tokens (B*2H*2W, C) -> (B*H*W, 4C); H, W = the coarse grid."""

    @staticmethod
    def forward(ctx, t, B, H, W):
        Cc = t.shape[1]
        ctx.dims = (B, H, W, Cc)
        return ops.pixel_unshuffle2(_c(t), B, H, W, Cc)

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        B, H, W, Cc = ctx.dims
        return ops.pixel_shuffle2(_c(dout), B, H, W, Cc), None, None, None


class ConvImg2TokFn(torch.autograd.Function):
    """3x3 s1 p1 conv of an NCHW image (3 ch) to tokens (B*H*W, Cout), Cout in {8, 16}; the image gets no gradient."""

    @staticmethod
    def forward(ctx, img, weight, bias):
        img = _c(img)
        B, _, H, W = img.shape
        ctx.save_for_backward(img, weight)
        ctx.has_bias = bias is not None
        return ops.conv3x3_small_fwd(img, False, weight, bias, B, H, W)

    @staticmethod
    @once_differentiable
    def backward(ctx, dtok):
        img, weight = ctx.saved_tensors
        if ctx.needs_input_grad[0]:
            raise NotImplementedError("ConvImg2TokFn: gradient w.r.t. the input image is not implemented")
        B, _, H, W = img.shape
        dw, db = ops.conv3x3_small_wgrad(img, False, weight, _c(dtok), B, H, W, want_bias=ctx.has_bias)
        return None, dw, db


class ConvTok2ImgFn(torch.autograd.Function):
    """3x3 s1 p1 conv of tokens (B*H*W, 8) to an NCHW image (B, 3, H, W) [+ residual image]."""

    @staticmethod
    def forward(ctx, tok, weight, bias, residual, B, H, W):
        tok = _c(tok)
        ctx.save_for_backward(tok, weight)
        ctx.meta = (B, H, W, bias is not None, residual is not None)
        return ops.conv3x3_small_fwd(tok, True, weight, bias, B, H, W, residual=residual)

    @staticmethod
    @once_differentiable
    def backward(ctx, dimg):
        tok, weight = ctx.saved_tensors
        B, H, W, has_bias, has_res = ctx.meta
        dimg = _c(dimg)
        dw, db = ops.conv3x3_small_wgrad(tok, True, weight, dimg, B, H, W, want_bias=has_bias)
        # data gradient = the image -> tokens kernel with flipped, transposed weights
        wt = weight.detach().flip(2, 3).transpose(0, 1).contiguous()
        dtok = ops.conv3x3_small_fwd(dimg, False, wt, None, B, H, W)
        return dtok, dw, db, (dimg if has_res and ctx.needs_input_grad[3] else None), None, None, None


# ---- Haar "wavelet" mode (src/model/wave_modules.py) on csrc/wavelet.cu ----------------------------------------------
class HaarDWTFn(torch.autograd.Function):
    """DWT_2D on tokens (B, 2h*2w, C) -> (B, h*w, C); the backward is the REFERENCE's formula (DWT_function.backward)."""

    @staticmethod
    def forward(ctx, x, B, h, w):
        x = _c(x)
        C = x.shape[-1]
        ctx.dims = (B, h, w, C)
        out = ops._empty((B, h * w, C), x)
        ops._run("uwr_haar_dwt_fwd", f"B{B} h{h} C{C}", 5 * out.numel() * 4, 0.0, ops._ptr(x), ops._ptr(out), B, h, w, C)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        B, h, w, C = ctx.dims
        dout = _c(dout)
        din = ops._empty((B, 4 * h * w, C), dout)
        ops._run("uwr_haar_dwt_bwd", f"B{B} h{h} C{C}", 5 * dout.numel() * 4, 0.0, ops._ptr(dout), ops._ptr(din), B, h, w, C)
        return din, None, None, None


class HaarIDWTFn(torch.autograd.Function):
    """IDWT_2D on tokens (B, h*w, C) -> (B, 2h*2w, C); the backward is the REFERENCE's formula (IDWT_function.backward,
    raw NCHW-memory reshapes included)."""

    @staticmethod
    def forward(ctx, x, B, h, w):
        x = _c(x)
        C = x.shape[-1]
        ctx.dims = (B, h, w, C)
        out = ops._empty((B, 4 * h * w, C), x)
        ops._run("uwr_haar_idwt_fwd", f"B{B} h{h} C{C}", 5 * x.numel() * 4, 0.0, ops._ptr(x), ops._ptr(out), B, h, w, C)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        B, h, w, C = ctx.dims
        dout = _c(dout)
        din = ops._empty((B, h * w, C), dout)
        ws = ops._ws(B * 4 * (h * w // 16) * 4, dout)
        ops._run("uwr_haar_idwt_bwd", f"B{B} h{h} C{C}", 5 * din.numel() * 4, 0.0, ops._ptr(dout), ops._ptr(din), ops._ptr(ws),
                 B, h, w, C)
        return din, None, None, None
