"""Hand-written forward/backward of the building blocks of the restoration transformers.

Each autograd.Function below is one fused "block" of the reference with an explicit backward made
of uwr kernels (no ATen autograd graph inside):

  InputProjFn     AST.InputProj            (AST.py:447-466)
  OutputProjFn    AST.OutputProj + x + y   (AST.py:470-493, 921)
  DownsampleFn    AST.Downsample           (AST.py:408-424)
  UpsampleCatFn   AST.Upsample + torch.cat (AST.py:428-443, 903-916)
  AttnBlockFn     x + DropPath(W-MSA(LN1(x)))   (TransformerBlock.forward, AST.py:590-619)
  LeFFBlockFn     x + DropPath(LeFF(LN2(x)))    (AST.py:307-326, 622)

Token tensors are (B, L, C) fp32 contiguous; DropPath arrives as a per-sample scale vector
(mask / keep_prob) or None.
"""
import torch
from torch.autograd.function import once_differentiable

from . import ops


def _c(t):
    return t if t.is_contiguous() else t.contiguous()


class InputProjFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img, weight, bias, slope):
        img = _c(img)
        tokens = ops.input_proj_fwd(img, weight, bias, slope)
        ctx.save_for_backward(img, weight, tokens)
        ctx.slope = slope
        return tokens

    @staticmethod
    @once_differentiable
    def backward(ctx, dtokens):
        img, weight, tokens = ctx.saved_tensors
        if ctx.needs_input_grad[0]:
            raise NotImplementedError("InputProjFn: the gradient w.r.t. the input image is not implemented "
                                      "(the training path never needs it, SURVEY.md §8a row 2)")
        dweight, dbias = ops.input_proj_bwd(_c(dtokens), tokens, img, weight, ctx.slope)
        # the image itself never needs a gradient on the training path
        return None, dweight, dbias, None


class OutputProjFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, tokens, weight, bias, residual_img, H, W):
        tokens = _c(tokens)
        B = tokens.shape[0]
        out = ops.output_proj_fwd(tokens, weight, bias, residual_img, B, H, W)
        ctx.save_for_backward(tokens, weight)
        ctx.hw = (H, W)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        tokens, weight = ctx.saved_tensors
        H, W = ctx.hw
        dtokens, dweight, dbias = ops.output_proj_bwd(_c(dout), tokens, weight, tokens.shape[0], H, W)
        # out = conv(tokens) + residual_img: the residual's gradient is the cotangent itself
        return dtokens, dweight, dbias, (dout if ctx.needs_input_grad[3] else None), None, None


class DownsampleFn(torch.autograd.Function):
    """Conv4x4 stride 2 pad 1 on tokens (AST.py:408-424).  tf32 mode: implicit GEMM -- the im2col tiles are materialised
    by TMA inside the tcgen05 kernel (uwr_convgemm_tcgen05), forward and weight gradient read a TF32-rounded copy of x;
    otherwise (tf32x3, geometry not served) im2col (tap-major K) + GEMM."""

    @staticmethod
    def forward(ctx, x, weight, bias, H, W):
        x = _c(x)
        B, L, Cc = x.shape
        Cout = weight.shape[0]
        # (Cout, Cin, 4, 4) -> (Cout, 4, 4, Cin): K index = (ky, kx, ci) matches the im2col rows
        wmat = weight.permute(0, 2, 3, 1).reshape(Cout, 16 * Cc)
        ctx.dims = (B, H, W, Cc, Cout)
        if ops.fast_path():
            xr = ops.scale_round(x.view(B * L, Cc), Cc)
            wmat_r = ops.scale_round(wmat, 16 * Cc)
            y = ops.conv_gemm_fwd(xr, wmat_r, bias, B, H, W, 4, 4, 2, 1)
            if y is not None:
                ctx.save_for_backward(xr, wmat_r)   # x itself (1/4 of the im2col matrix) is all the weight gradient needs
                ctx.implicit = True
                return y.view(B, L // 4, Cout)
        ctx.implicit = False
        col = ops.im2col_4x4s2(x.view(B * L, Cc), B, H, W, Cc)  # TF32-rounded at the store in tf32 mode
        wmat_r = ops.scale_round(wmat, 16 * Cc)
        y = ops.linear(col, wmat_r, bias, t5=True)
        # the im2col matrix (4x the input) is kept for the weight gradient instead of being rebuilt in backward
        ctx.save_for_backward(col, wmat_r)
        return y.view(B, L // 4, Cout)

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        col, wmat = ctx.saved_tensors
        B, H, W, Cc, Cout = ctx.dims
        dy2 = _c(dy).view(-1, Cout)
        fast = ops.fast_path()
        if fast:  # tcgen05 path: both operands rounded to TF32 (col / x and wmat already are); bias gradient =
            dy2, dbias = ops.scale_round_colsum(dy2, Cout)  # column sums taken in the same pass
            dwmat = ops.conv_gemm_wgrad(dy2, col, B, H, W, 4, 4, 2, 1) if ctx.implicit else None
            if dwmat is None:
                if ctx.implicit:
                    col = ops.im2col_4x4s2(col, B, H, W, Cc)
                dwmat, _ = ops.linear_wgrad(dy2, col, want_bias=False, t5=True)
        else:
            dwmat, dbias = ops.linear_wgrad(dy2, col)
        del col
        dcol = ops.linear_dgrad(dy2, wmat, t5=fast)
        dx = ops.col2im_4x4s2(dcol, B, H, W, Cc)
        dweight = dwmat.reshape(Cout, 4, 4, Cc).permute(0, 3, 1, 2).contiguous()
        return dx.view(B, H * W, Cc), dweight, dbias, None, None


class UpsampleCatFn(torch.autograd.Function):
    """ConvTranspose2x2 stride 2 (GEMM + 2x2 pixel scatter) written straight into the left half of
    the concatenated decoder input; the encoder skip is copied into the right half."""

    @staticmethod
    def forward(ctx, x, weight, bias, skip, H, W):
        x = _c(x)
        skip = _c(skip)
        B, L, Cin = x.shape
        Cout = weight.shape[1]
        Cs = skip.shape[2]
        fast = ops.fast_path()
        wmat = (ops.rounded_weight(weight) if fast else weight).view(Cin, Cout * 4)  # stored [K=Cin][N=(co,dy,dx)]
        if fast:  # tcgen05 path: TF32-rounded copy of the input, also what the weight gradient reads
            x = ops.scale_round(x.view(B * L, Cin), Cin).view(B, L, Cin)
        g = ops._empty((B * L, Cout * 4), x)
        ops.gemm(x.view(B * L, Cin), wmat, g, B * L, Cout * 4, Cin, lda=Cin, ldb=Cout * 4, ldc=Cout * 4,
                 b_nk=False, t5=fast)
        out = ops._empty((B, 4 * L, Cout + Cs), x)
        out2 = out.view(B * 4 * L, Cout + Cs)
        ops.pixel_scatter_2x2(g, bias, out2, B, H, W, Cout)
        ops.copy2d(skip.view(B * 4 * L, Cs), out2[:, Cout:], Cs)
        ctx.save_for_backward(x, weight)
        ctx.dims = (B, H, W, Cin, Cout, Cs)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        x, weight = ctx.saved_tensors
        B, H, W, Cin, Cout, Cs = ctx.dims
        L = H * W
        d2 = _c(dout).view(B * 4 * L, Cout + Cs)
        dskip = ops._empty((B, 4 * L, Cs), x)
        ops.copy2d(d2[:, Cout:], dskip.view(B * 4 * L, Cs), Cs)
        dbias = ops.colsum(d2, Cout)
        dg = ops.pixel_gather_2x2(d2, B, H, W, Cout)
        fast = ops.fast_path()  # x (saved) and dg are TF32-rounded in fast mode
        wmat = (ops.rounded_weight(weight) if fast else weight).view(Cin, Cout * 4)
        # dx[M,Cin] = dg[M,4Cout] wmat^T : B stored [N=Cin][K=4Cout]
        dx = ops._empty((B * L, Cin), x)
        ops.gemm(dg, wmat, dx, B * L, Cin, Cout * 4, lda=Cout * 4, ldb=Cout * 4, ldc=Cin, b_nk=True, t5=fast)
        # dW[Cin,4Cout] = x^T dg
        dwmat = ops._empty((Cin, Cout * 4), x)
        ops.gemm(x.view(B * L, Cin), dg, dwmat, Cin, Cout * 4, B * L, lda=Cin, ldb=Cout * 4, ldc=Cout * 4,
                 a_km=True, b_nk=False, t5=fast)
        return dx.view(B, L, Cin), dwmat.view_as(weight), dbias, dskip, None, None


def _qkv_forward(y1, wq, bq, wkv, bkv):
    if ops.fast_path():
        w, b = ops.packed_qkv(wq, bq, wkv, bkv)
        # q | k | v leave the epilogue TF32-rounded: the attention kernels' products are then exact in one pass
        return ops.linear(y1, w, b, t5=True, round_out=ops.attn_rounded())
    return ops.linear(y1, wq, bq, weight2=wkv, bias2=bkv)


def _qkv_dgrad(dqkv, wq, bq, wkv, bkv):
    if ops.fast_path():
        w, _ = ops.packed_qkv(wq, bq, wkv, bkv)
        return ops.linear_dgrad(dqkv, w, t5=True)
    return ops.linear_dgrad(dqkv, wq, weight2=wkv)


# ---- links between consecutive block functions -----------------------------------------------------------------
# The backward of a block function starts by turning the incoming residual-stream gradient d into the GEMM operand
# d_s = tf32(droppath_scale * d) (+ its column sums = a bias gradient).  d is produced by the LayerNorm backward of the
# function that ran AFTER it in the forward pass, so that kernel can emit d_s in the same pass (uwr_layernorm_bwd_ds).
# A link is a plain dict shared by the two apply() calls: the consumer registers what it will need at forward time,
# the producer leaves the result there at backward time; anything missing simply falls back to the separate pass.
def _link_register(link, dp_scale, L, bias):
    if link is not None:
        link.clear()
        link.update(dp=dp_scale, L=L, bias=bias)


def _link_take(link, d):
    """(d_s, colsum, colsum_is_grad_slot) left by the producer for exactly this gradient tensor, else None."""
    if link is None:
        return None
    pre = link.pop("ds", None)
    if pre is None or pre[3] != d.data_ptr():
        return None
    return pre[0], pre[1], pre[2]


def _ln_bwd_linked(link, dy, x2, nw, mean, rstd, d, g_w, g_b):
    """LayerNorm backward; when a consumer is registered on `link` the kernel also emits its d_s / column sums."""
    if link is not None and "dp" in link:
        bias = link["bias"]
        slot = ops.grad_slot(bias) if bias is not None else None
        res = ops.layernorm_bwd_ds(dy, x2, nw, mean, rstd, d, link["dp"], link["L"], dgamma_out=g_w, dbeta_out=g_b,
                                   cs_out=slot)
        if res is not None:
            dx, dg, db, d_s, cs = res
            link["ds"] = (d_s, cs, slot is not None, dx.data_ptr())
            return dx, dg, db
    return ops.layernorm_bwd(dy, x2, nw, mean, rstd, dres=d, dgamma_out=g_w, dbeta_out=g_b)


class AttnBlockFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, n1w, n1b, wq, bq, wkv, bkv, table, wparam, wp, bp, dp_scale, H, W, heads, shift, scale,
                link_in=None, link_out=None):
        x = _c(x)
        B, L, Cc = x.shape
        M = B * L
        x2 = x.view(M, Cc)
        hd = Cc // heads
        ctx.link_in, ctx.link_out = link_in, link_out   # producer side (our LN1 backward) / consumer side (our d)
        _link_register(link_out, dp_scale, L, bp)
        y1, mean, rstd = ops.layernorm_fwd(x2, n1w, n1b)
        qkv = _qkv_forward(y1, wq, bq, wkv, bkv)
        o = ops.window_attn_fwd(qkv, 0, qkv, Cc, 2 * Cc, table, wparam, B, H, W, heads, hd, shift, scale,
                                rounded=ops.attn_rounded())
        x1 = ops.linear(o, ops.rounded_weight(wp), bp, residual=x2, rowscale=dp_scale, rows_per_group=L, t5=True)
        ctx.save_for_backward(x2, n1w, mean, rstd, y1, qkv, o, wq, bq, wkv, bkv, table, wparam, wp, dp_scale)
        ctx.n1b, ctx.bp = n1b, bp   # only their gradient slots are needed in backward
        ctx.meta = (B, L, Cc, H, W, heads, hd, shift, scale)
        return x1.view(B, L, Cc)

    @staticmethod
    @once_differentiable
    def backward(ctx, dx1):
        x2, n1w, mean, rstd, y1, qkv, o, wq, bq, wkv, bkv, table, wparam, wp, dp = ctx.saved_tensors
        B, L, Cc, H, W, heads, hd, shift, scale = ctx.meta
        d = _c(dx1).view(B * L, Cc)
        # parameter gradients are written straight into their bucket slots when uwr.train owns them (ops.grad_slot);
        # the Function then returns None for them and autograd's per-parameter `.grad += g` kernel disappears
        bp = ctx.bp
        g_bp, g_wp, g_tab, g_w = (ops.grad_slot(t) if t is not None else None for t in (bp, wp, table, wparam))
        g_n1w, g_n1b = ops.grad_slot(n1w), ops.grad_slot(ctx.n1b)
        # DropPath-scaled (and, in tf32 mode, TF32-rounded) branch gradient: operand of three GEMMs; the column sums
        # are the proj bias gradient.  Usually already emitted by the LayerNorm backward that produced d (link).
        pre = _link_take(ctx.link_out, d)
        if pre is not None:
            d_s, dbp, direct = pre
            if direct:
                g_bp = dbp
        else:
            d_s, dbp = ops.scale_round_colsum(d, Cc, dp, L, cs_out=g_bp)
        fast = ops.fast_path()
        d_o = ops.linear_dgrad(d_s, ops.rounded_weight(wp), t5=True, round_out=fast and ops.attn_rounded())
        dwp, _ = ops.linear_wgrad(d_s, o, want_bias=False, t5=True, out=g_wp)
        del d_s
        # [to_q; to_kv] weight / bias gradients come out of one GEMM / one vector of column sums: written in place when
        # the two bucket slots are adjacent (GradBuckets(adjacent=...)).  The column sums of dq | dk | dv (bias gradient)
        # are accumulated inside the attention backward kernel.
        g_wqkv = ops.fused_grad_slot(wq, wkv)
        g_bqkv = ops.fused_grad_slot(bq, bkv) if bq is not None else None
        dbqkv = None
        if bq is not None:
            dbqkv = g_bqkv if g_bqkv is not None else ops._empty((3 * Cc,), qkv)
        dqkv, _, dtable, dw = ops.window_attn_bwd(d_o, qkv, 0, qkv, Cc, 2 * Cc, table, wparam, B, H, W, heads, hd,
                                                  shift, scale, dtable_out=g_tab, dw_out=g_w, rounded=fast and ops.attn_rounded(),
                                                  colsum_q=dbqkv, colsum_kv=dbqkv)
        del d_o
        dy1 = _qkv_dgrad(dqkv, wq, bq, wkv, bkv)
        dwqkv, _ = ops.linear_wgrad(dqkv, y1, want_bias=False, t5=True, out=g_wqkv)
        del dqkv
        dx, dg, db = _ln_bwd_linked(ctx.link_in, dy1, x2, n1w, mean, rstd, d, g_n1w, g_n1b)
        dwq, dwkv = (None, None) if g_wqkv is not None else (dwqkv[:Cc], dwqkv[Cc:])
        dbq, dbkv = (None, None) if (bq is None or g_bqkv is not None) else (dbqkv[:Cc], dbqkv[Cc:])
        nz = lambda g, slot: None if slot is not None else g
        return (dx.view(B, L, Cc), nz(dg, g_n1w), nz(db, g_n1b), dwq, dbq, dwkv, dbkv, nz(dtable, g_tab),
                nz(dw, g_w) if wparam is not None else None, nz(dwp, g_wp), nz(dbp, g_bp), None, None, None, None, None,
                None, None, None)


class LeFFBlockFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, n2w, n2b, w1, b1, dww, dwb, w2, b2, dp_scale, H, W, link_in=None, link_out=None):
        x = _c(x)
        B, L, Cc = x.shape
        M = B * L
        Ch = w1.shape[0]
        x2 = x.view(M, Cc)
        ctx.link_in, ctx.link_out = link_in, link_out
        _link_register(link_out, dp_scale, L, b2)
        y2, mean, rstd = ops.layernorm_fwd(x2, n2w, n2b)
        need_bwd = any(ctx.needs_input_grad)
        if need_bwd and ops.half_storage_ok(H, W, Ch) and dwb is not None:
            # fp16 storage of u and gelu'(v) (neither is a tensor-core operand): half the bytes of 5 of the 13 passes
            # this block makes over 4C-wide tensors, at the accuracy a TF32 operand has anyway
            u = ops.linear(y2, ops.rounded_weight(w1), b1, t5=True, out_half=True)
            v, h2 = ops.dwconv_gelu_fwd_half(u, dww, dwb, B, H, W, Ch)
        else:
            u = ops.linear(y2, ops.rounded_weight(w1), b1, t5=True)
            # the conv pre-activation v is only ever needed as gelu'(v): store that instead
            v, h2 = ops.dwconv_gelu_fwd(u, dww, dwb, B, H, W, Ch, mode=0, save_v=need_bwd, v_is_dgelu=True)
        out = ops.linear(h2, ops.rounded_weight(w2), b2, residual=x2, rowscale=dp_scale, rows_per_group=L, t5=True)
        if need_bwd:
            ctx.save_for_backward(x2, n2w, mean, rstd, y2, u, v, h2, w1, dww, w2, dp_scale)
            ctx.biases = (n2b, b1, dwb, b2)   # only their gradient slots are needed in backward
        ctx.meta = (B, L, Cc, Ch, H, W)
        return out.view(B, L, Cc)

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        x2, n2w, mean, rstd, y2, u, v, h2, w1, dww, w2, dp = ctx.saved_tensors
        B, L, Cc, Ch, H, W = ctx.meta
        d = _c(dout).view(B * L, Cc)
        n2b, b1, dwb, b2 = ctx.biases
        g = {k: (ops.grad_slot(t) if t is not None else None)
             for k, t in dict(n2w=n2w, n2b=n2b, w1=w1, b1=b1, dww=dww, dwb=dwb, w2=w2, b2=b2).items()}
        pre = _link_take(ctx.link_out, d)   # d_s / column sums (= linear2 bias gradient) from the producer of d
        if pre is not None:
            d_s, db2, direct = pre
            if direct:
                g["b2"] = db2
        else:
            d_s, db2 = ops.scale_round_colsum(d, Cc, dp, L, cs_out=g["b2"])
        # dv = (d_s W2) * gelu'(v): the second GELU's derivative (saved by the forward) rides in the
        # GEMM epilogue
        dv = ops.linear_dgrad(d_s, ops.rounded_weight(w2), mul_by=v, t5=True)
        dw2, _ = ops.linear_wgrad(d_s, h2, want_bias=False, t5=True, out=g["w2"])
        del d_s
        du, ddww, ddwb, db1 = ops.dwconv_gelu_bwd(dv, u, dww, B, H, W, Ch, want_du_colsum=True, dweight_out=g["dww"],
                                                  dbias_out=g["dwb"], dusum_out=g["b1"])
        del dv
        dy2 = ops.linear_dgrad(du, ops.rounded_weight(w1), t5=True)
        dw1, _ = ops.linear_wgrad(du, y2, want_bias=False, t5=True, out=g["w1"])
        del du
        dx, dg, db = _ln_bwd_linked(ctx.link_in, dy2, x2, n2w, mean, rstd, d, g["n2w"], g["n2b"])
        nz = lambda grad, k: None if g[k] is not None else grad
        dg, db, dw1, db1, ddww, ddwb, dw2, db2 = (nz(dg, "n2w"), nz(db, "n2b"), nz(dw1, "w1"), nz(db1, "b1"),
                                                  nz(ddww, "dww"), nz(ddwb, "dwb"), nz(dw2, "w2"), nz(db2, "b2"))
        return dx.view(B, L, Cc), dg, db, dw1, db1, ddww, ddwb, dw2, db2, None, None, None, None, None
