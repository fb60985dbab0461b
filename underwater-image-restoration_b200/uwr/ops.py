"""Thin tensor-level wrappers over the C ABI (one Python function per exported kernel family).

Every function takes CUDA fp32 tensors, extracts raw pointers / leading dimensions and launches
on torch's current stream.  Nothing here computes on the CPU or through ATen: a non-CUDA tensor
is an error (the product path has no fallback).
"""
import ctypes as C

import torch

from . import _lib
from ._lib import AttnDesc, GemmDesc, check, fn

EPI_NONE, EPI_RESID, EPI_MUL_DGELU = 0, 1, 2


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    if t is None:
        return None
    if not t.is_cuda or t.dtype != torch.float32:
        raise TypeError(f"uwr ops need CUDA float32 tensors, got {t.device} {t.dtype}")
    return t.data_ptr()


def _empty(shape, like):
    return torch.empty(shape, device=like.device, dtype=torch.float32)


def _ws(nbytes, like):
    n = max(1, (int(nbytes) + 3) // 4)
    return torch.empty(n, device=like.device, dtype=torch.float32)


# --------------------------------------------------------------------------------------------
def gemm(A, B, C_out, M, N, K, *, lda, ldb, ldc, a_km=False, b_nk=True, B2=None, n_split=0,
         bias=None, bias2=None, epilogue=EPI_NONE, R=None, ldr=0, rowscale=None, rows_per_group=0,
         colsum=None):
    """C[M,N] = epi(opA(A) opB(B)); see uwr_gemm_tf32 in include/uwr_b200.h."""
    d = GemmDesc()
    d.A, d.lda, d.a_km = _ptr(A), lda, int(a_km)
    d.B, d.ldb, d.b_nk = _ptr(B), ldb, int(b_nk)
    d.B2, d.n_split = _ptr(B2), n_split
    d.C, d.ldc = _ptr(C_out), ldc
    d.M, d.N, d.K = M, N, K
    d.bias, d.bias2 = _ptr(bias), _ptr(bias2)
    d.epilogue = epilogue
    d.R, d.ldr = _ptr(R), ldr
    d.rowscale, d.rows_per_group = _ptr(rowscale), rows_per_group
    d.colsum = _ptr(colsum)
    ws = None
    if a_km:
        nbytes = fn["uwr_gemm_workspace_bytes"](M, N, K, 1)
        if nbytes:
            ws = _ws(nbytes, A)
            d.workspace, d.workspace_bytes = ws.data_ptr(), nbytes
    check(fn["uwr_gemm_tf32"](C.byref(d), _stream()), "uwr_gemm_tf32")
    return C_out


def linear(x2d, weight, bias=None, *, weight2=None, bias2=None, residual=None, rowscale=None,
           rows_per_group=0, out=None):
    """y = x W^T + b  (optionally [W;W2], optionally residual + s*(.)). x2d: (M,K) view, row stride lda."""
    M, K = x2d.shape
    lda = x2d.stride(0)
    N = weight.shape[0] + (weight2.shape[0] if weight2 is not None else 0)
    if out is None:
        out = _empty((M, N), x2d)
    epi = EPI_RESID if residual is not None else EPI_NONE
    gemm(x2d, weight, out, M, N, K, lda=lda, ldb=weight.stride(0), ldc=out.stride(0), b_nk=True,
         B2=weight2, n_split=weight.shape[0] if weight2 is not None else 0, bias=bias, bias2=bias2,
         epilogue=epi, R=residual, ldr=residual.stride(0) if residual is not None else 0,
         rowscale=rowscale, rows_per_group=rows_per_group)
    return out


def linear_dgrad(dy2d, weight, *, weight2=None, rowscale=None, rows_per_group=0, out=None):
    """dx = (s*dy) [W;W2]   (dy: (M,N), W: (N1,K), W2: (N2,K))."""
    M, N = dy2d.shape
    K = weight.shape[1]
    if out is None:
        out = _empty((M, K), dy2d)
    gemm(dy2d, weight, out, M, K, N, lda=dy2d.stride(0), ldb=weight.stride(0), ldc=out.stride(0),
         b_nk=False, B2=weight2, n_split=weight.shape[0] if weight2 is not None else 0,
         rowscale=rowscale, rows_per_group=rows_per_group)
    return out


def linear_wgrad(dy2d, x2d, *, want_bias=True, rowscale=None, rows_per_group=0):
    """dW[N,K] = (s*dy)^T x ; db[N] = colsum(s*dy).  dy: (M,N) view, x: (M,K) view."""
    M, N = dy2d.shape
    K = x2d.shape[1]
    dW = _empty((N, K), dy2d)
    db = _empty((N,), dy2d) if want_bias else None
    gemm(dy2d, x2d, dW, N, K, M, lda=dy2d.stride(0), ldb=x2d.stride(0), ldc=K, a_km=True, b_nk=False,
         rowscale=rowscale, rows_per_group=rows_per_group, colsum=db)
    return dW, db


# --------------------------------------------------------------------------------------------
def layernorm_fwd(x2d, gamma, beta, eps=1e-5, save_stats=True):
    rows, Cc = x2d.shape
    y = torch.empty_like(x2d)
    mean = _empty((rows,), x2d) if save_stats else None
    rstd = _empty((rows,), x2d) if save_stats else None
    check(fn["uwr_layernorm_fwd"](_ptr(x2d), _ptr(gamma), _ptr(beta), _ptr(y), _ptr(mean), _ptr(rstd),
                                  rows, Cc, eps, _stream()), "uwr_layernorm_fwd")
    return y, mean, rstd


def layernorm_bwd(dy2d, x2d, gamma, mean, rstd, dres=None):
    rows, Cc = x2d.shape
    dx = torch.empty_like(x2d)
    dgamma = torch.empty_like(gamma)
    dbeta = torch.empty_like(gamma)
    ws = _ws(fn["uwr_layernorm_bwd_workspace_bytes"](rows, Cc), x2d)
    check(fn["uwr_layernorm_bwd"](_ptr(dy2d), _ptr(x2d), _ptr(gamma), _ptr(mean), _ptr(rstd), _ptr(dres),
                                  _ptr(dx), _ptr(dgamma), _ptr(dbeta), _ptr(ws), rows, Cc, _stream()),
          "uwr_layernorm_bwd")
    return dx, dgamma, dbeta


# --------------------------------------------------------------------------------------------
def _attn_desc(q_buf, q_off, kv_buf, k_off, v_off, table, w_param, B, H, W, heads, head_dim, shift, scale):
    d = AttnDesc()
    d.q, d.ld_q, d.q_off = _ptr(q_buf), q_buf.stride(0), q_off
    d.kv, d.ld_kv, d.k_off, d.v_off = _ptr(kv_buf), kv_buf.stride(0), k_off, v_off
    d.bias_table, d.w_param = _ptr(table), _ptr(w_param)
    d.B, d.H, d.W, d.heads, d.head_dim, d.shift, d.scale = B, H, W, heads, head_dim, shift, scale
    return d


def window_attn_fwd(q_buf, q_off, kv_buf, k_off, v_off, table, w_param, B, H, W, heads, head_dim, shift, scale):
    """q_buf/kv_buf: 2-D token matrices (B*H*W, ld). Returns O (B*H*W, heads*head_dim)."""
    d = _attn_desc(q_buf, q_off, kv_buf, k_off, v_off, table, w_param, B, H, W, heads, head_dim, shift, scale)
    out = _empty((B * H * W, heads * head_dim), q_buf)
    check(fn["uwr_window_attn_fwd"](C.byref(d), _ptr(out), out.stride(0), _stream()), "uwr_window_attn_fwd")
    return out


def window_attn_bwd(dout, q_buf, q_off, kv_buf, k_off, v_off, table, w_param, B, H, W, heads, head_dim, shift,
                    scale, dq_buf=None, dkv_buf=None):
    d = _attn_desc(q_buf, q_off, kv_buf, k_off, v_off, table, w_param, B, H, W, heads, head_dim, shift, scale)
    if dq_buf is None:
        dq_buf = torch.empty_like(q_buf)
    if dkv_buf is None:
        dkv_buf = dq_buf if kv_buf.data_ptr() == q_buf.data_ptr() else torch.empty_like(kv_buf)
    dtable = torch.empty_like(table)
    dw = _empty((2,), q_buf)
    ws = _ws(fn["uwr_window_attn_bwd_workspace_bytes"](C.byref(d)), q_buf)
    check(fn["uwr_window_attn_bwd"](C.byref(d), _ptr(dout), dout.stride(0), _ptr(dq_buf), _ptr(dkv_buf),
                                    _ptr(dtable), _ptr(dw), _ptr(ws), _stream()), "uwr_window_attn_bwd")
    return dq_buf, dkv_buf, dtable, dw


# --------------------------------------------------------------------------------------------
def dwconv_gelu_fwd(u2d, weight, bias, B, H, W, Ch, mode=0, save_v=True):
    v = _empty((B * H * W, Ch), u2d) if save_v else None
    h2 = _empty((B * H * W, Ch), u2d)
    check(fn["uwr_dwconv_gelu_fwd"](_ptr(u2d), u2d.stride(0), _ptr(weight), _ptr(bias), _ptr(v), _ptr(h2),
                                    B, H, W, Ch, mode, _stream()), "uwr_dwconv_gelu_fwd")
    return v, h2


def dwconv_gelu_bwd(dh2, u2d, v, weight, B, H, W, Ch, mode=0):
    du = torch.empty_like(u2d)
    dweight = torch.empty_like(weight)
    dbias = _empty((Ch,), u2d)
    ws = _ws(fn["uwr_dwconv_gelu_bwd_workspace_bytes"](B, H, W, Ch), u2d)
    check(fn["uwr_dwconv_gelu_bwd"](_ptr(dh2), _ptr(u2d), u2d.stride(0), _ptr(v), _ptr(weight), _ptr(du),
                                    _ptr(dweight), _ptr(dbias), _ptr(ws), B, H, W, Ch, mode, _stream()),
          "uwr_dwconv_gelu_bwd")
    return du, dweight, dbias


# --------------------------------------------------------------------------------------------
def input_proj_fwd(img, weight, bias, slope=0.01):
    B, Cin, H, W = img.shape
    Cout = weight.shape[0]
    tokens = _empty((B, H * W, Cout), img)
    check(fn["uwr_input_proj_fwd"](_ptr(img), _ptr(weight), _ptr(bias), _ptr(tokens), B, H, W, Cin, Cout,
                                   slope, _stream()), "uwr_input_proj_fwd")
    return tokens


def input_proj_bwd(dtokens, tokens, img, weight, slope=0.01):
    B, Cin, H, W = img.shape
    Cout = weight.shape[0]
    dweight = torch.empty_like(weight)
    dbias = _empty((Cout,), img)
    ws = _ws(fn["uwr_input_proj_bwd_workspace_bytes"](B, H, W, Cin, Cout), img)
    check(fn["uwr_input_proj_bwd"](_ptr(dtokens), _ptr(tokens), _ptr(img), _ptr(dweight), _ptr(dbias), _ptr(ws),
                                   B, H, W, Cin, Cout, slope, _stream()), "uwr_input_proj_bwd")
    return dweight, dbias


def output_proj_fwd(tokens, weight, bias, residual_img, B, H, W):
    Cin = weight.shape[1]
    out = _empty((B, 3, H, W), tokens)
    check(fn["uwr_output_proj_fwd"](_ptr(tokens), tokens.stride(-2), _ptr(weight), _ptr(bias),
                                    _ptr(residual_img), _ptr(out), B, H, W, Cin, _stream()),
          "uwr_output_proj_fwd")
    return out


def output_proj_bwd(dout_img, tokens, weight, B, H, W):
    Cin = weight.shape[1]
    dtokens = _empty((B, H * W, Cin), tokens)
    dweight = torch.empty_like(weight)
    dbias = _empty((3,), tokens)
    ws = _ws(fn["uwr_output_proj_bwd_workspace_bytes"](B, H, W, Cin), tokens)
    check(fn["uwr_output_proj_bwd"](_ptr(dout_img), _ptr(tokens), tokens.stride(-2), _ptr(weight), _ptr(dtokens),
                                    _ptr(dweight), _ptr(dbias), _ptr(ws), B, H, W, Cin, _stream()),
          "uwr_output_proj_bwd")
    return dtokens, dweight, dbias


def im2col_4x4s2(tokens2d, B, H, W, Cc):
    col = _empty((B * (H // 2) * (W // 2), 16 * Cc), tokens2d)
    check(fn["uwr_im2col_4x4s2"](_ptr(tokens2d), tokens2d.stride(0), _ptr(col), B, H, W, Cc, _stream()),
          "uwr_im2col_4x4s2")
    return col


def col2im_4x4s2(dcol, B, H, W, Cc):
    dx = _empty((B * H * W, Cc), dcol)
    check(fn["uwr_col2im_4x4s2"](_ptr(dcol), _ptr(dx), B, H, W, Cc, _stream()), "uwr_col2im_4x4s2")
    return dx


def pixel_scatter_2x2(g, bias, out2d, B, H, W, Cout):
    check(fn["uwr_pixel_scatter_2x2"](_ptr(g), _ptr(bias), _ptr(out2d), out2d.stride(0), B, H, W, Cout,
                                      _stream()), "uwr_pixel_scatter_2x2")


def pixel_gather_2x2(dout2d, B, H, W, Cout):
    dg = _empty((B * H * W, 4 * Cout), dout2d)
    check(fn["uwr_pixel_gather_2x2"](_ptr(dout2d), dout2d.stride(0), _ptr(dg), B, H, W, Cout, _stream()),
          "uwr_pixel_gather_2x2")
    return dg


def copy2d(src2d, dst2d, cols, accumulate=False):
    rows = src2d.shape[0]
    check(fn["uwr_copy2d"](_ptr(src2d), src2d.stride(0), _ptr(dst2d), dst2d.stride(0), rows, cols,
                           int(accumulate), _stream()), "uwr_copy2d")


def colsum(x2d, cols):
    rows = x2d.shape[0]
    out = _empty((cols,), x2d)
    ws = _ws(1024 * cols * 4, x2d)
    check(fn["uwr_colsum"](_ptr(x2d), x2d.stride(0), _ptr(out), _ptr(ws), rows, cols, _stream()), "uwr_colsum")
    return out


# --------------------------------------------------------------------------------------------
LOSS_KINDS = {"L1": 0, "L1withColor": 1, "charbonnier": 2, "L2": 3}


def pixel_loss(pred, truth, kind, batch_divisor=None, want_grad=True):
    B, Cc, H, W = pred.shape
    out = _empty((1,), pred)
    grad = torch.empty_like(pred) if want_grad else None
    ws = _ws(4 * 1024 * 4, pred)
    check(fn["uwr_pixel_loss"](_ptr(pred), _ptr(truth), _ptr(out), _ptr(grad), _ptr(ws), LOSS_KINDS[kind],
                               B, Cc, H, W, batch_divisor or B, _stream()), "uwr_pixel_loss")
    return out, grad
