"""FRFN feed-forward (src/Models/AST.py:329-372 == src/model/block.py:249-282) on the uwr kernels:

    x' = [conv3x3(x[:, :C/4]) | x[:, C/4:]]           partial conv = im2col + TF32 GEMM
    u  = x' W1^T + b1  (C -> 8C)                      tcgen05 GEMM
    h  = gelu(dwconv3x3(gelu(u[:, :4C])) + bd) * gelu(u[:, 4C:])      gated depthwise kernel (mode 1)
    y  = h W2^T + b2   (4C -> C), out = x_res + DropPath(y)           tcgen05 GEMM, residual epilogue
"""
import torch
import torch.nn as nn
from torch.autograd.function import once_differentiable

from . import ops


class FRFNBlockFn(torch.autograd.Function):
    """[x +] DropPath(FRFN([LN](x))) with a hand-written backward.  `nw is None` skips the LayerNorm,
    `residual=False` returns the bare FRFN output (EncoderBlock of the NewBig family, model.py:57-72)."""

    @staticmethod
    def forward(ctx, x, nw, nb, wpc, w1, b1, dww, dwb, w2, b2, dp_scale, H, W, residual):
        x = x if x.is_contiguous() else x.contiguous()
        B, L, Cc = x.shape
        M, Cq, Ch = B * L, Cc // 4, w2.shape[1]
        x2 = x.view(M, Cc)
        if nw is not None:
            y, mean, rstd = ops.layernorm_fwd(x2, nw, nb)
        else:
            y, mean, rstd = ops.scale_round(x2, Cc), None, None   # private (TF32-rounded) copy
        # partial conv on the first C/4 channels, written in place of them (y is a private buffer)
        wpm = ops.scale_round(wpc.permute(0, 2, 3, 1).reshape(Cq, 9 * Cq), 9 * Cq)
        col = ops.im2col_3x3(y, B, H, W, Cq)
        xp = y  # the conv result overwrites y[:, :Cq] after col was gathered
        ops.linear(col, wpm, None, out=xp[:, :Cq], t5=True, round_out=True)
        u = ops.linear(xp, ops.rounded_weight(w1), b1, t5=True)
        need_bwd = any(ctx.needs_input_grad)
        v, h = ops.dwconv_gelu_fwd(u, dww, dwb, B, H, W, Ch, mode=1, save_v=need_bwd)
        out = ops.linear(h, ops.rounded_weight(w2), b2, residual=x2 if residual else None, rowscale=dp_scale,
                         rows_per_group=L, t5=True)
        if need_bwd:
            ctx.save_for_backward(x2, nw, mean, rstd, xp, col, wpm, u, v, h, w1, dww, w2, dp_scale)
        ctx.meta = (B, L, Cc, Cq, Ch, H, W, residual)
        return out.view(B, L, Cc)

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        x2, nw, mean, rstd, xp, col, wpm, u, v, h, w1, dww, w2, dp = ctx.saved_tensors
        B, L, Cc, Cq, Ch, H, W, residual = ctx.meta
        M = B * L
        d = (dout if dout.is_contiguous() else dout.contiguous()).view(M, Cc)
        d_s, db2 = ops.scale_round_colsum(d, Cc, dp, L)          # rounded cotangent + linear2 bias gradient, one pass
        dh = ops.linear_dgrad(d_s, ops.rounded_weight(w2), t5=True)
        dw2, _ = ops.linear_wgrad(d_s, h, want_bias=False, t5=True)
        del d_s
        du = torch.empty_like(u)
        dv = ops.gelu_gate_bwd(dh, u, v, Ch, 1, du=du)          # also fills du[:, Ch:]
        del dh
        # linear1 bias gradient = column sums of du: the conv half comes out of the depthwise backward itself
        db1 = ops._empty((2 * Ch,), du)
        _, ddww, ddwb, _ = ops.dwconv_gelu_bwd(dv, u, dww, B, H, W, Ch, du=du, want_du_colsum=True, dusum_out=db1[:Ch])
        ops.colsum(du[:, Ch:], Ch, out=db1[Ch:])
        del dv
        dxp = ops.linear_dgrad(du, ops.rounded_weight(w1), t5=True)
        dw1, _ = ops.linear_wgrad(du, xp, want_bias=False, t5=True)
        del du
        # partial conv backward: a dense TF32-rounded copy of the C/4-wide slice feeds the tcgen05 kernel (the strided,
        # un-rounded view used to take the legacy mma.sync GEMM: 4 ms per NewBigFRFN step)
        fast = ops.fast_path()
        g1 = ops.scale_round(dxp[:, :Cq], Cq) if fast else dxp[:, :Cq]
        dwpm, _ = ops.linear_wgrad(g1, col, want_bias=False, t5=fast)
        dcol = ops.linear_dgrad(g1, wpm, t5=fast)
        ops.col2im_3x3(dcol, dxp, B, H, W, Cq)                   # overwrites dxp[:, :Cq] with d/d(FRFN input)
        dg = db = None
        if nw is not None:
            dx, dg, db = ops.layernorm_bwd(dxp, x2, nw, mean, rstd, dres=d if residual else None)
        else:
            dx = dxp
            if residual:
                ops.copy2d(d, dx, Cc, accumulate=True)
        dwpc = dwpm.view(Cq, 3, 3, Cq).permute(0, 3, 1, 2).contiguous()
        return dx.view(B, L, Cc), dg, db, dwpc, dw1, db1, ddww, ddwb, dw2, db2, None, None, None, None


class FRFN(nn.Module):
    def __init__(self, dim=32, hidden_dim=128, act_layer=nn.GELU, drop=0.0, use_eca=False):
        super().__init__()
        self.linear1 = nn.Sequential(nn.Linear(dim, hidden_dim * 2), act_layer())
        self.dwconv = nn.Sequential(
            nn.Conv2d(hidden_dim, hidden_dim, groups=hidden_dim, kernel_size=3, stride=1, padding=1), act_layer())
        self.linear2 = nn.Sequential(nn.Linear(hidden_dim, dim))
        self.dim = dim
        self.hidden_dim = hidden_dim
        self.dim_conv = self.dim // 4
        self.dim_untouched = self.dim - self.dim_conv
        self.partial_conv3 = nn.Conv2d(self.dim_conv, self.dim_conv, 3, 1, 1, bias=False)

    def block_forward(self, x, norm, dp_scale, H, W, residual=True):
        """[x +] DropPath(FRFN([norm](x))) — LayerNorm and residual are fused into the block function."""
        return FRFNBlockFn.apply(x, norm.weight if norm is not None else None, norm.bias if norm is not None else None,
                                 self.partial_conv3.weight, self.linear1[0].weight, self.linear1[0].bias,
                                 self.dwconv[0].weight, self.dwconv[0].bias, self.linear2[0].weight,
                                 self.linear2[0].bias, dp_scale, H, W, residual)
