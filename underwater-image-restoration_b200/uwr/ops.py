"""Thin tensor-level wrappers over the C ABI (one Python function per exported kernel family).

Every function takes CUDA fp32 tensors, extracts raw pointers / leading dimensions and launches
on torch's current stream.  Nothing here computes on the CPU or through ATen: a non-CUDA tensor
is an error (the product path has no fallback).
"""
import ctypes as C
import math
import os

import torch

from . import _lib
from ._lib import AttnDesc, ConvGemmDesc, ConvSmallDesc, GemmDesc, check, fn

EPI_NONE, EPI_RESID, EPI_MUL_DGELU, EPI_MUL = 0, 1, 2, 3


_PASSES = 1
WEIGHT_EPOCH = 0  # bumped by optimizers that update parameters behind torch's version counter


def set_gemm_precision(mode):
    """"tf32" (default: one TF32 product, fp32 accumulate; GEMM operands are rounded where they are
    produced and the Linear layers run on the tcgen05 path) or "tf32x3" (error-compensated 3xTF32 on
    full fp32 operands, fp32-level accuracy, legacy mma.sync kernel)."""
    global _PASSES
    _PASSES = {"tf32": 1, "tf32x3": 3}[mode]
    check(fn["uwr_set_gemm_precision"](_PASSES), "uwr_set_gemm_precision")


def set_attn_tcgen05(mode):
    """window attention forward (head_dim 32) on the tcgen05/TMA kernel: False / 0 = never, True / 1 = whenever the shape
    is eligible, "auto" / 2 (library default) = when the operands are exact TF32 values (the models' path)."""
    m = 2 if mode in ("auto", 2) else int(bool(mode))
    check(fn["uwr_set_attn_tcgen05"](m), "uwr_set_attn_tcgen05")


def set_gemm_cluster(mode):
    """tcgen05 GEMM: pair CTAs into clusters of 2 with TMA multicast of the B tile.  False = never, True / "auto" / "all"
    = every tensor-bound shape (default), "tn" = the weight-gradient layout only."""
    m = {False: 0, True: 1, "auto": 1, "all": 1, "tn": 2}[mode]
    check(fn["uwr_set_gemm_cluster"](m), "uwr_set_gemm_cluster")


def set_pdl(on):
    """Programmatic dependent launch of the library's kernels (csrc/uwr_common.cuh): every kernel is launched as soon as
    all CTAs of its predecessor have started and waits on the device for the predecessor to complete, so launch latency
    and per-CTA set-up overlap the predecessor's tail.  Same results as plain stream order.  Off by default (AST: no gain);
    +2 % for SpectralTransformer, whose 2 256 launches per step are ~15 us each.  UWR_PDL=1 enables it from the environment.
    Set it before a step is captured into a CUDA graph: the setting is baked into the graph's edges."""
    check(fn["uwr_set_pdl"](1 if on else 0), "uwr_set_pdl")


if os.environ.get("UWR_GEMM_CLUSTER"):      # A/B knob for benchmarks: off | auto | tn
    set_gemm_cluster({"off": False, "0": False, "auto": "auto", "1": "auto", "all": "all", "tn": "tn"}[os.environ["UWR_GEMM_CLUSTER"]])


def fast_path():
    return _PASSES == 1


# Precision knob (debugging / accuracy studies): UWR_ATTN_ROUNDED=0 keeps q | k | v and d_o in full fp32 and runs the attention
# products with 3xTF32 compensation instead (AST gradient error 2e-4 instead of 3e-4, attention ~0.5 ms per step slower).
_ATTN_ROUNDED = os.environ.get("UWR_ATTN_ROUNDED", "1") != "0"


def attn_rounded():
    """True when the attention operands are rounded to TF32 where they are produced (default in tf32 mode)."""
    return _PASSES == 1 and _ATTN_ROUNDED


_HALF_STORAGE = True


def set_half_storage(on):
    """fp16 storage of the two 4C-wide LeFF tensors that are not tensor-core operands: u (linear1 output) and gelu'(v)
    (DESIGN.md §3).  A 10-bit mantissa is what a TF32 operand keeps anyway, so this halves their HBM bytes at TF32-level
    accuracy; only in single-pass mode, on the tcgen05 path, for full 16x16 tiles (ops.half_storage_ok)."""
    global _HALF_STORAGE
    _HALF_STORAGE = bool(on)


def half_storage_ok(H, W, Ch):
    return _HALF_STORAGE and _PASSES == 1 and bool(fn["uwr_dwconv_half_supported"](H, W, Ch))


def bump_weight_epoch():
    global WEIGHT_EPOCH
    WEIGHT_EPOCH += 1


def launch_count():
    return int(fn["uwr_launch_count"]())


def _stream():
    return torch.cuda.current_stream().cuda_stream


class KernelProfile:
    """Per-call CUDA-event timing of the C-ABI launches (used by bench.py for the roofline table;
    never active inside a timed region).  Records (family, shape key, algorithmic bytes, flops)."""

    def __init__(self):
        self.records = []

    def __enter__(self):
        global _PROF
        _PROF = self
        return self

    def __exit__(self, *exc):
        global _PROF
        _PROF = None
        torch.cuda.synchronize()

    def table(self):
        agg = {}
        for name, key, nbytes, flops, e0, e1 in self.records:
            ms = e0.elapsed_time(e1)
            a = agg.setdefault((name, key), [0, 0.0, 0.0, 0.0])
            a[0] += 1
            a[1] += ms
            a[2] += nbytes
            a[3] += flops
        rows = [dict(kernel=k[0], shape=k[1], launches=v[0], ms_total=v[1], ms_avg=v[1] / v[0],
                     bytes_per_launch=v[2] / v[0], flops_per_launch=v[3] / v[0],
                     gbs=(v[2] / v[1] / 1e6) if v[1] > 0 else 0.0, tflops=(v[3] / v[1] / 1e9) if v[1] > 0 else 0.0)
                for k, v in agg.items()]
        rows.sort(key=lambda r: -r["ms_total"])
        return rows


_PROF = None


def _run(name, key, nbytes, flops, *args):
    """Launch one C-ABI entry on the current stream (optionally bracketed by CUDA events)."""
    if _PROF is None:
        check(fn[name](*args, _stream()), name)
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    check(fn[name](*args, _stream()), name)
    e1.record()
    _PROF.records.append((name, key, nbytes, flops, e0, e1))


def _ptr(t):
    if t is None:
        return None
    if not t.is_cuda or t.dtype != torch.float32:
        raise TypeError(f"uwr ops need CUDA float32 tensors, got {t.device} {t.dtype}")
    return t.data_ptr()


def _hptr(t):
    """device pointer of a CUDA float16 tensor (fp16-storage paths)"""
    if not t.is_cuda or t.dtype != torch.float16:
        raise TypeError(f"expected a CUDA float16 tensor, got {t.device} {t.dtype}")
    return t.data_ptr()


def _fptr(t):
    """float32 or float16 CUDA tensor -> (pointer, is_half)"""
    if t is None:
        return None, 0
    if t.dtype == torch.float16:
        return _hptr(t), 1
    return _ptr(t), 0


def _empty(shape, like):
    return torch.empty(shape, device=like.device, dtype=torch.float32)


def _ws(nbytes, like):
    n = max(1, (int(nbytes) + 3) // 4)
    return torch.empty(n, device=like.device, dtype=torch.float32)


# --------------------------------------------------------------------------------------------
def gemm(A, B, C_out, M, N, K, *, lda, ldb, ldc, a_km=False, b_nk=True, B2=None, n_split=0,
         bias=None, bias2=None, epilogue=EPI_NONE, R=None, ldr=0, rowscale=None, rows_per_group=0,
         colsum=None, t5=False, round_out=False):
    """C[M,N] = epi(opA(A) opB(B)); see uwr_gemm_tf32 in include/uwr_b200.h.  t5=True: the operands are
    TF32-rounded already, use the tcgen05/TMA kernel when the descriptor is supported."""
    d = GemmDesc()
    d.A, d.lda, d.a_km = _ptr(A), lda, int(a_km)
    d.B, d.ldb, d.b_nk = _ptr(B), ldb, int(b_nk)
    d.B2, d.n_split = _ptr(B2), n_split
    d.C, d.c_half = _fptr(C_out)
    d.ldc = ldc
    d.M, d.N, d.K = M, N, K
    d.bias, d.bias2 = _ptr(bias), _ptr(bias2)
    d.epilogue = epilogue
    d.R, d.r_half = _fptr(R)
    d.ldr = ldr
    d.rowscale, d.rows_per_group = _ptr(rowscale), rows_per_group
    d.colsum = _ptr(colsum)
    d.round_out = int(round_out and _PASSES == 1)
    nbytes = 4 * (M * K + K * N) + (2 if d.c_half else 4) * M * N + ((2 if d.r_half else 4) * M * N if R is not None else 0)
    lay = ("TN" if a_km else ("NT" if b_nk else "NN"))
    ws = None
    if t5 and _PASSES == 1 and fn["uwr_gemm_tcgen05_supported"](C.byref(d)):
        wbytes = fn["uwr_gemm_tcgen05_workspace_bytes"](M, N, K, int(a_km))
        if wbytes:
            ws = _ws(wbytes, A)
            d.workspace, d.workspace_bytes = ws.data_ptr(), wbytes
        _run("uwr_gemm_tcgen05", f"{lay} M{M} N{N} K{K} epi{epilogue}" + ("h" if (d.c_half or d.r_half) else ""), nbytes,
             2.0 * M * N * K, C.byref(d))
        return C_out
    if d.c_half or d.r_half:
        raise ValueError("fp16 C / R storage needs the tcgen05 GEMM path (t5=True, single-pass mode, supported shape)")
    if a_km:
        wbytes = fn["uwr_gemm_workspace_bytes"](M, N, K, 1)
        if wbytes:
            ws = _ws(wbytes, A)
            d.workspace, d.workspace_bytes = ws.data_ptr(), wbytes
    _run("uwr_gemm_tf32", f"{lay} M{M} N{N} K{K} epi{epilogue}", nbytes, 2.0 * M * N * K, C.byref(d))
    return C_out


def linear(x2d, weight, bias=None, *, weight2=None, bias2=None, residual=None, rowscale=None,
           rows_per_group=0, out=None, t5=False, round_out=False, out_half=False):
    """y = x W^T + b  (optionally [W;W2], optionally residual + s*(.)). x2d: (M,K) view, row stride lda.
    out_half: y is stored as float16 (tcgen05 path only)."""
    M, K = x2d.shape
    lda = x2d.stride(0)
    N = weight.shape[0] + (weight2.shape[0] if weight2 is not None else 0)
    if out is None:
        out = torch.empty((M, N), device=x2d.device, dtype=torch.float16 if out_half else torch.float32)
    epi = EPI_RESID if residual is not None else EPI_NONE
    gemm(x2d, weight, out, M, N, K, lda=lda, ldb=weight.stride(0), ldc=out.stride(0), b_nk=True,
         B2=weight2, n_split=weight.shape[0] if weight2 is not None else 0, bias=bias, bias2=bias2,
         epilogue=epi, R=residual, ldr=residual.stride(0) if residual is not None else 0,
         rowscale=rowscale, rows_per_group=rows_per_group, t5=t5, round_out=round_out)
    return out


def linear_dgrad(dy2d, weight, *, weight2=None, rowscale=None, rows_per_group=0, out=None, dgelu_of=None,
                 mul_by=None, t5=False, round_out=False):
    """dx = (s*dy) [W;W2]   (dy: (M,N), W: (N1,K), W2: (N2,K)); dgelu_of=v multiplies by gelu'(v),
    mul_by=g multiplies elementwise by g (a stored gelu'(v))."""
    M, N = dy2d.shape
    K = weight.shape[1]
    if out is None:
        out = _empty((M, K), dy2d)
    gemm(dy2d, weight, out, M, K, N, lda=dy2d.stride(0), ldb=weight.stride(0), ldc=out.stride(0),
         b_nk=False, B2=weight2, n_split=weight.shape[0] if weight2 is not None else 0,
         rowscale=rowscale, rows_per_group=rows_per_group,
         epilogue=EPI_MUL_DGELU if dgelu_of is not None else (EPI_MUL if mul_by is not None else EPI_NONE),
         R=dgelu_of if dgelu_of is not None else mul_by,
         ldr=(dgelu_of if dgelu_of is not None else mul_by).stride(0) if (dgelu_of is not None or mul_by is not None) else 0,
         t5=t5, round_out=round_out)
    return out


def linear_wgrad(dy2d, x2d, *, want_bias=True, rowscale=None, rows_per_group=0, t5=False, out=None):
    """dW[N,K] = (s*dy)^T x ; db[N] = colsum(s*dy).  dy: (M,N) view, x: (M,K) view.  out: dense (N,K) buffer."""
    M, N = dy2d.shape
    K = x2d.shape[1]
    dW = out.view(N, K) if out is not None else _empty((N, K), dy2d)
    db = _empty((N,), dy2d) if want_bias else None
    gemm(dy2d, x2d, dW, N, K, M, lda=dy2d.stride(0), ldb=x2d.stride(0), ldc=K, a_km=True, b_nk=False,
         rowscale=rowscale, rows_per_group=rows_per_group, colsum=db, t5=t5)
    return dW, db


# --------------------------------------------------------------------------------------------
def scale_round(src2d, cols, rowscale=None, rows_per_group=0, out=None):
    """dense (rows, cols) copy of src2d[:, :cols] scaled per sample and, in tf32 mode, rounded to TF32."""
    rows = src2d.shape[0]
    if out is None:
        out = _empty((rows, cols), src2d)
    _run("uwr_scale_round", f"rows{rows} C{cols}", 8 * rows * cols, 0.0, _ptr(src2d), src2d.stride(0), _ptr(out),
         rows, cols, _ptr(rowscale), rows_per_group, int(_PASSES == 1))
    return out


def scale_round_colsum(src2d, cols, rowscale=None, rows_per_group=0, cs_out=None):
    """scale_round + column sums of the result (bias gradient of the consuming Linear) in one pass."""
    rows = src2d.shape[0]
    if cols % 4 or (cols // 4) > 256 or 256 % (cols // 4):
        out = scale_round(src2d, cols, rowscale, rows_per_group)
        cs = colsum(out, cols)
        if cs_out is not None:
            cs_out.copy_(cs)
            cs = cs_out
        return out, cs
    out = _empty((rows, cols), src2d)
    cs = cs_out if cs_out is not None else _empty((cols,), src2d)
    ws = _ws(1024 * cols * 4, src2d)
    _run("uwr_scale_round_colsum", f"rows{rows} C{cols}", 8 * rows * cols, 0.0, _ptr(src2d), src2d.stride(0),
         _ptr(out), rows, cols, _ptr(rowscale), rows_per_group, int(_PASSES == 1), _ptr(cs), _ptr(ws))
    return out, cs


_DIRECT_WRITES = False


class direct_grad_writes:
    """Context in which backward kernels may OVERWRITE the bucket slot of a parameter instead of returning its
    gradient to autograd.  Only uwr.train.TrainStep enters it, around the single backward that follows its bucket
    zeroing; any other backward (micro-batch accumulation, two forwards before one backward, a plain torch.optim
    loop) gets ordinary accumulate semantics even if the parameters' .grad are bucket views."""

    def __enter__(self):
        global _DIRECT_WRITES
        self._prev, _DIRECT_WRITES = _DIRECT_WRITES, True

    def __exit__(self, *exc):
        global _DIRECT_WRITES
        _DIRECT_WRITES = self._prev


def grad_slot(p):
    """The preallocated gradient buffer of a parameter whose gradients are owned by uwr.train.GradBuckets
    (flag `_uwr_direct`), inside `direct_grad_writes()`: kernels write the gradient straight into it and the autograd
    Function returns None, which saves autograd's `.grad += g` kernel per parameter (~250 tiny launches per AST step).
    Overwrite semantics: only valid because TrainStep zeroes the buckets and every parameter gets one gradient per step."""
    if _DIRECT_WRITES and getattr(p, "_uwr_direct", False):
        g = p.grad
        if g is not None and g.is_cuda and g.is_contiguous() and g.dtype == torch.float32:
            return g
    return None


def fused_grad_slot(p0, p1):
    """One gradient buffer spanning two parameters whose bucket slots are adjacent (to_q | to_kv weights or
    biases, laid out back to back by GradBuckets): the packed-QKV weight gradient is written once, in place."""
    g0, g1 = grad_slot(p0), grad_slot(p1)
    if g0 is None or g1 is None or g0.data_ptr() + g0.numel() * 4 != g1.data_ptr():
        return None
    tail = tuple(g0.shape[1:])
    if tuple(g1.shape[1:]) != tail:
        return None
    rows = g0.shape[0] + g1.shape[0]
    inner = 1
    for d in tail:
        inner *= d
    return torch.as_strided(g0, (rows,) + tail, (inner,) + tuple(g0.stride()[1:]))


def _fresh(w, tag):
    ent = getattr(w, tag, None)
    ver = (w.data_ptr(), w._version, WEIGHT_EPOCH, _PASSES)
    return ent, ver


# Rounded-copy registry: every cached TF32 copy (rounded_weight / packed_qkv) is listed here so that an
# optimizer that rewrites the parameters behind torch's back can refresh ALL of them with two multi-tensor
# launches (refresh_rounded_copies) instead of one lazy launch per weight in the next forward (~100 tiny
# kernels per AST step).
_ROUNDED = {}          # id(w) -> (weakref(w), kind)
_ROUNDED_TABLES = None  # (generation, device tables) of the last refresh
_ROUNDED_GEN = 0


def _register_rounded(w, kind):
    global _ROUNDED_GEN
    import weakref
    if id(w) not in _ROUNDED or _ROUNDED[id(w)][0]() is not w:
        _ROUNDED[id(w)] = (weakref.ref(w), kind)
        _ROUNDED_GEN += 1


def rounded_weight(w):
    """TF32-rounded copy of a weight matrix (cached on the tensor object, refreshed when the parameter
    changes); the weight itself in tf32x3 mode."""
    if _PASSES != 1:
        return w
    base = w._base
    if (base is not None and isinstance(base, torch.nn.Parameter) and w.is_contiguous() and base.is_contiguous()
            and w.numel() == base.numel() and w.data_ptr() == base.data_ptr()):
        # a reshaped view of a parameter (conv.weight.flatten(1), ...): cache on the parameter itself, views are
        # new objects on every call
        return rounded_weight(base).view(w.shape)
    ent, ver = _fresh(w, "_uwr_rounded")
    if ent is None or ent[0] != ver:
        buf = ent[1] if ent is not None and ent[1].shape == w.shape and ent[1].device == w.device else torch.empty_like(w)
        wd = w.detach()
        w2 = wd.reshape(wd.shape[0], -1) if wd.dim() > 1 else wd.reshape(1, -1)   # conv weights: (C0, rest)
        scale_round(w2, w2.shape[1], out=buf.view(w2.shape))
        w._uwr_rounded = ent = (ver, buf)
        if w.is_contiguous() and isinstance(w, torch.nn.Parameter):   # temporaries would grow the registry
            _register_rounded(w, "w")
    return ent[1]


def refresh_rounded_copies():
    """Re-round every registered TF32 weight copy in two multi-tensor launches (weights: rounded; packed
    biases: plain copy) and mark the caches fresh.  Called by FusedClipAdam.step() right after the
    parameter update (also inside a captured CUDA graph)."""
    global _ROUNDED_TABLES
    if _PASSES != 1 or not _ROUNDED:
        return
    live = []
    for key, (ref, kind) in list(_ROUNDED.items()):
        w = ref()
        if w is None or not w.is_cuda:
            del _ROUNDED[key]
            continue
        live.append((w, kind))
    if not live:
        return
    dev = live[0][0].device
    sig = (_ROUNDED_GEN, len(live))
    if _ROUNDED_TABLES is None or _ROUNDED_TABLES[0] != sig:
        wsrc, wdst, wn, bsrc, bdst, bn = [], [], [], [], [], []
        for w, kind in live:
            if kind == "w":
                ent = getattr(w, "_uwr_rounded", None)
                if ent is None:
                    continue
                wsrc.append(w.data_ptr()); wdst.append(ent[1].data_ptr()); wn.append(w.numel())
            else:  # packed qkv: kind = (wkv, bq, bkv)
                ent = getattr(w, "_uwr_qkv", None)
                if ent is None:
                    continue
                wkv, bq, bkv = kind
                Cq, K = w.shape
                wsrc += [w.data_ptr(), wkv.data_ptr()]
                wdst += [ent[1].data_ptr(), ent[1][Cq:].data_ptr()]
                wn += [w.numel(), wkv.numel()]
                if bq is not None:
                    bsrc += [bq.data_ptr(), bkv.data_ptr()]
                    bdst += [ent[2].data_ptr(), ent[2][Cq:].data_ptr()]
                    bn += [bq.numel(), bkv.numel()]

        def tables(src, dst, ns):
            if not src:
                return None
            offs = [0]
            for n in ns:
                offs.append(offs[-1] + n)
            i64 = lambda xs: torch.tensor(xs, dtype=torch.int64, device=dev)
            return (i64(src), i64(dst), i64(offs), len(src), offs[-1])
        _ROUNDED_TABLES = (sig, tables(wsrc, wdst, wn), tables(bsrc, bdst, bn))
    _, tw, tb = _ROUNDED_TABLES
    for t, do_round in ((tw, 1), (tb, 0)):
        if t is not None:
            check(fn["uwr_round_tf32_tensors"](t[0].data_ptr(), t[1].data_ptr(), t[2].data_ptr(), t[3], t[4], do_round,
                                               _stream()), "uwr_round_tf32_tensors")
    for w, kind in live:   # the copies now match the parameters of THIS epoch
        if kind == "w":
            ent = getattr(w, "_uwr_rounded", None)
            if ent is not None:
                w._uwr_rounded = ((w.data_ptr(), w._version, WEIGHT_EPOCH, _PASSES), ent[1])
        else:
            ent = getattr(w, "_uwr_qkv", None)
            if ent is not None:
                wkv, bq, bkv = kind
                ver = (w.data_ptr(), w._version, WEIGHT_EPOCH, _PASSES) + (
                    wkv.data_ptr(), wkv._version, None if bq is None else (bq._version, bkv._version))
                w._uwr_qkv = (ver, ent[1], ent[2])


def packed_qkv(wq, bq, wkv, bkv):
    """[to_q; to_kv] as one (3C, C) TF32-rounded matrix + packed bias (AST.py:47-48,59-60), cached."""
    ent, ver = _fresh(wq, "_uwr_qkv")
    ver = ver + (wkv.data_ptr(), wkv._version, None if bq is None else (bq._version, bkv._version))
    if ent is None or ent[0] != ver:
        Cq, K = wq.shape
        buf = ent[1] if ent is not None else _empty((3 * Cq, K), wq)
        scale_round(wq.detach(), K, out=buf[:Cq])
        scale_round(wkv.detach(), K, out=buf[Cq:])
        bias = None
        if bq is not None:
            bias = ent[2] if ent is not None and ent[2] is not None else _empty((3 * Cq,), wq)
            _run("uwr_scale_round", "bias", 0, 0.0, _ptr(bq.detach()), Cq, _ptr(bias), 1, Cq, None, 0, 0)
            _run("uwr_scale_round", "bias", 0, 0.0, _ptr(bkv.detach()), 2 * Cq, _ptr(bias[Cq:]), 1, 2 * Cq, None, 0, 0)
        wq._uwr_qkv = ent = (ver, buf, bias)
        if wq.is_contiguous() and wkv.is_contiguous() and isinstance(wq, torch.nn.Parameter):
            _register_rounded(wq, (wkv, bq, bkv))
    return ent[1], ent[2]


# --------------------------------------------------------------------------------------------
def layernorm_fwd(x2d, gamma, beta, eps=1e-5, save_stats=True):
    rows, Cc = x2d.shape
    y = torch.empty_like(x2d)
    mean = _empty((rows,), x2d) if save_stats else None
    rstd = _empty((rows,), x2d) if save_stats else None
    _run("uwr_layernorm_fwd", f"rows{rows} C{Cc}", 8 * rows * Cc, 8.0 * rows * Cc,
         _ptr(x2d), _ptr(gamma), _ptr(beta), _ptr(y), _ptr(mean), _ptr(rstd), rows, Cc, eps)
    return y, mean, rstd


def layernorm_bwd_ds(dy2d, x2d, gamma, mean, rstd, dres, rowscale, rows_per_group, dgamma_out=None, dbeta_out=None,
                     cs_out=None):
    """layernorm_bwd that also emits d_s = tf32(rowscale * dx) and its column sums for the next backward function
    (saves the separate scale_round_colsum pass).  Returns None when the shape is not served."""
    rows, Cc = x2d.shape
    if not fn["uwr_layernorm_bwd_ds_supported"](rows, Cc):
        return None
    dx = torch.empty_like(x2d)
    d_s = torch.empty_like(x2d)
    dgamma = dgamma_out if dgamma_out is not None else torch.empty_like(gamma)
    dbeta = dbeta_out if dbeta_out is not None else torch.empty_like(gamma)
    cs = cs_out if cs_out is not None else _empty((Cc,), x2d)
    ws = _ws(fn["uwr_layernorm_bwd_ds_workspace_bytes"](rows, Cc), x2d)
    _run("uwr_layernorm_bwd_ds", f"rows{rows} C{Cc}", (16 + (4 if dres is not None else 0)) * rows * Cc, 16.0 * rows * Cc,
         _ptr(dy2d), _ptr(x2d), _ptr(gamma), _ptr(mean), _ptr(rstd), _ptr(dres), _ptr(dx), _ptr(dgamma), _ptr(dbeta),
         _ptr(rowscale), int(rows_per_group), _ptr(d_s), _ptr(cs), _ptr(ws), rows, Cc)
    return dx, dgamma, dbeta, d_s, cs


def layernorm_bwd(dy2d, x2d, gamma, mean, rstd, dres=None, dgamma_out=None, dbeta_out=None):
    rows, Cc = x2d.shape
    dx = torch.empty_like(x2d)
    dgamma = dgamma_out if dgamma_out is not None else torch.empty_like(gamma)
    dbeta = dbeta_out if dbeta_out is not None else torch.empty_like(gamma)
    ws = _ws(fn["uwr_layernorm_bwd_workspace_bytes"](rows, Cc), x2d)
    _run("uwr_layernorm_bwd", f"rows{rows} C{Cc}", (12 + (4 if dres is not None else 0)) * rows * Cc,
         16.0 * rows * Cc, _ptr(dy2d), _ptr(x2d), _ptr(gamma), _ptr(mean), _ptr(rstd), _ptr(dres), _ptr(dx),
         _ptr(dgamma), _ptr(dbeta), _ptr(ws), rows, Cc)
    return dx, dgamma, dbeta


# --------------------------------------------------------------------------------------------
def _attn_desc(q_buf, q_off, kv_buf, k_off, v_off, table, w_param, B, H, W, heads, head_dim, shift, scale, rounded=False):
    d = AttnDesc()
    d.operands_rounded = int(bool(rounded) and _PASSES == 1)
    d.q, d.ld_q, d.q_off = _ptr(q_buf), q_buf.stride(0), q_off
    d.kv, d.ld_kv, d.k_off, d.v_off = _ptr(kv_buf), kv_buf.stride(0), k_off, v_off
    d.bias_table, d.w_param = _ptr(table), _ptr(w_param)
    d.B, d.H, d.W, d.heads, d.head_dim, d.shift, d.scale = B, H, W, heads, head_dim, shift, scale
    return d


def window_attn_fwd(q_buf, q_off, kv_buf, k_off, v_off, table, w_param, B, H, W, heads, head_dim, shift, scale,
                    rounded=False):
    """q_buf/kv_buf: 2-D token matrices (B*H*W, ld). Returns O (B*H*W, heads*head_dim).
    rounded: q, k, v are exact TF32 values (producing GEMM ran with round_out) -> single exact tensor-core pass."""
    d = _attn_desc(q_buf, q_off, kv_buf, k_off, v_off, table, w_param, B, H, W, heads, head_dim, shift, scale, rounded)
    out = _empty((B * H * W, heads * head_dim), q_buf)
    tiles = B * (H // 8) * (W // 8) * heads
    _run("uwr_window_attn_fwd", f"tiles{tiles} hd{head_dim} shift{shift}", tiles * 4 * 64 * head_dim * 4,
         tiles * 4.0 * 64 * 64 * head_dim, C.byref(d), _ptr(out), out.stride(0))
    return out


def window_attn_bwd(dout, q_buf, q_off, kv_buf, k_off, v_off, table, w_param, B, H, W, heads, head_dim, shift,
                    scale, dq_buf=None, dkv_buf=None, dtable_out=None, dw_out=None, rounded=False, colsum_q=None,
                    colsum_kv=None):
    """rounded: q, k, v AND dout are exact TF32 values.  colsum_q / colsum_kv (both or neither): vectors that receive
    the column sums of dq and of dk | dv (projection bias gradients) at the columns dq_buf / dkv_buf use; the same 3C
    vector twice for a packed q|k|v buffer."""
    d = _attn_desc(q_buf, q_off, kv_buf, k_off, v_off, table, w_param, B, H, W, heads, head_dim, shift, scale, rounded)
    if colsum_q is not None:
        d.dq_colsum, d.dkv_colsum = _ptr(colsum_q), _ptr(colsum_kv)
    if dq_buf is None:
        dq_buf = torch.empty_like(q_buf)
    if dkv_buf is None:
        dkv_buf = dq_buf if kv_buf.data_ptr() == q_buf.data_ptr() else torch.empty_like(kv_buf)
    dtable = dtable_out if dtable_out is not None else torch.empty_like(table)
    dw = dw_out if dw_out is not None else _empty((2,), q_buf)
    ws = _ws(fn["uwr_window_attn_bwd_workspace_bytes"](C.byref(d)), q_buf)
    tiles = B * (H // 8) * (W // 8) * heads
    _run("uwr_window_attn_bwd", f"tiles{tiles} hd{head_dim} shift{shift}", tiles * 7 * 64 * head_dim * 4,
         tiles * 10.0 * 64 * 64 * head_dim, C.byref(d), _ptr(dout), dout.stride(0), _ptr(dq_buf), _ptr(dkv_buf),
         _ptr(dtable), _ptr(dw), _ptr(ws))
    return dq_buf, dkv_buf, dtable, dw


# --------------------------------------------------------------------------------------------
def dwconv_gelu_fwd_half(u_half, weight, bias, B, H, W, Ch):
    """LeFF forward on a float16 u: returns (gelu'(v) as float16, h2 as TF32-rounded float32)."""
    dg = torch.empty((B * H * W, Ch), device=u_half.device, dtype=torch.float16)
    h2 = torch.empty((B * H * W, Ch), device=u_half.device, dtype=torch.float32)
    n = B * H * W * Ch
    _run("uwr_dwconv_gelu_fwd_half", f"B{B} H{H} Ch{Ch} half", n * (2 + 2 + 4), 18.0 * n, _hptr(u_half), _ptr(weight),
         _ptr(bias), _hptr(dg), _ptr(h2), B, H, W, Ch)
    return dg, h2


def dwconv_gelu_fwd(u2d, weight, bias, B, H, W, Ch, mode=0, save_v=True, v_is_dgelu=False):
    """Returns (v or gelu'(v), h2)."""
    v = _empty((B * H * W, Ch), u2d) if save_v else None
    h2 = _empty((B * H * W, Ch), u2d)
    n = B * H * W * Ch
    _run("uwr_dwconv_gelu_fwd", f"B{B} H{H} Ch{Ch} mode{mode}", 4 * n * ((2 if mode else 1) + (2 if save_v else 1)),
         18.0 * n, _ptr(u2d), u2d.stride(0), _ptr(weight), _ptr(bias), _ptr(v), _ptr(h2), B, H, W, Ch, mode, int(v_is_dgelu))
    return v, h2


def gelu_gate_bwd(dh2, u2d, v, Ch, mode, du=None):
    """dv = dh2 * gelu'(v) [* gelu(u2)]; FRFN (mode 1) also fills du[:, Ch:]."""
    rows = dh2.shape[0]
    dv = torch.empty_like(dh2)
    _run("uwr_gelu_gate_bwd", f"rows{rows} Ch{Ch} mode{mode}", 4 * rows * Ch * (3 + 2 * mode), 0.0,
         _ptr(dh2), _ptr(u2d), u2d.stride(0), _ptr(v), _ptr(dv), _ptr(du), rows, Ch, mode)
    return dv


def dwconv_gelu_bwd(dv, u2d, weight, B, H, W, Ch, du=None, plain=False, want_du_colsum=False, dweight_out=None,
                    dbias_out=None, dusum_out=None):
    """dv = dL/d(conv output).  Returns du (same row stride as u; only [:, :Ch] is written), dweight, dbias
    [, column sums of du[:, :Ch]]."""
    half = u2d.dtype == torch.float16
    if du is None:
        du = torch.empty(u2d.shape, device=u2d.device, dtype=torch.float32)
    dweight = dweight_out if dweight_out is not None else torch.empty_like(weight)
    dbias = dbias_out if dbias_out is not None else _empty((Ch,), dv)
    dusum = (dusum_out if dusum_out is not None else _empty((Ch,), dv)) if want_du_colsum else None
    ws = _ws(fn["uwr_dwconv_gelu_bwd_workspace_bytes"](B, H, W, Ch), dv)
    n = B * H * W * Ch
    if half:
        if plain or u2d.stride(0) != Ch:
            raise ValueError("float16 u: LeFF mode on a dense (rows, Ch) matrix only")
        _run("uwr_dwconv_gelu_bwd_half", f"B{B} H{H} Ch{Ch} half", n * (4 + 2 + 4), 36.0 * n, _ptr(dv), _hptr(u2d),
             _ptr(weight), _ptr(du), _ptr(dweight), _ptr(dbias), _ptr(dusum), _ptr(ws), B, H, W, Ch)
        return (du, dweight, dbias, dusum) if want_du_colsum else (du, dweight, dbias)
    _run("uwr_dwconv_gelu_bwd", f"B{B} H{H} Ch{Ch}", 4 * n * 3, 36.0 * n,
         _ptr(dv), _ptr(u2d), u2d.stride(0), _ptr(weight), _ptr(du), _ptr(dweight), _ptr(dbias), _ptr(dusum),
         _ptr(ws), B, H, W, Ch, int(plain))
    if want_du_colsum:
        return du, dweight, dbias, dusum
    return du, dweight, dbias


# --------------------------------------------------------------------------------------------
def input_proj_fwd(img, weight, bias, slope=0.01):
    B, Cin, H, W = img.shape
    Cout = weight.shape[0]
    tokens = _empty((B, H * W, Cout), img)
    _run("uwr_input_proj_fwd", "", 0, 0.0, _ptr(img), _ptr(weight), _ptr(bias), _ptr(tokens), B, H, W, Cin, Cout,
                                   slope)
    return tokens


def input_proj_bwd(dtokens, tokens, img, weight, slope=0.01):
    B, Cin, H, W = img.shape
    Cout = weight.shape[0]
    dweight = torch.empty_like(weight)
    dbias = _empty((Cout,), img)
    ws = _ws(fn["uwr_input_proj_bwd_workspace_bytes"](B, H, W, Cin, Cout), img)
    _run("uwr_input_proj_bwd", "", 0, 0.0, _ptr(dtokens), _ptr(tokens), _ptr(img), _ptr(dweight), _ptr(dbias), _ptr(ws),
                                   B, H, W, Cin, Cout, slope)
    return dweight, dbias


def output_proj_fwd(tokens, weight, bias, residual_img, B, H, W):
    Cin = weight.shape[1]
    out = _empty((B, 3, H, W), tokens)
    _run("uwr_output_proj_fwd", "", 0, 0.0, _ptr(tokens), tokens.stride(-2), _ptr(weight), _ptr(bias),
                                    _ptr(residual_img), _ptr(out), B, H, W, Cin)
    return out


def output_proj_bwd(dout_img, tokens, weight, B, H, W):
    Cin = weight.shape[1]
    dtokens = _empty((B, H * W, Cin), tokens)
    dweight = torch.empty_like(weight)
    dbias = _empty((3,), tokens)
    ws = _ws(fn["uwr_output_proj_bwd_workspace_bytes"](B, H, W, Cin), tokens)
    _run("uwr_output_proj_bwd", "", 0, 0.0, _ptr(dout_img), _ptr(tokens), tokens.stride(-2), _ptr(weight), _ptr(dtokens),
                                    _ptr(dweight), _ptr(dbias), _ptr(ws), B, H, W, Cin)
    return dtokens, dweight, dbias


def _convgemm_desc(x2d, B, H, W, Cin, Cout, kh, kw, stride, pad):
    d = ConvGemmDesc()
    d.x, d.ld_x, d.B, d.H, d.W, d.Cin, d.Cout = _ptr(x2d), x2d.stride(0), B, H, W, Cin, Cout
    d.kh, d.kw, d.stride, d.pad = kh, kw, stride, pad
    return d


def _conv_out(H, W, kh, kw, stride, pad):
    return (H + 2 * pad - kh) // stride + 1, (W + 2 * pad - kw) // stride + 1


def conv_gemm_fwd(x2d, wmat, bias, B, H, W, kh, kw, stride, pad, round_out=False, out=None):
    """y = im2col(x) wmat^T (+ bias) as an implicit GEMM (no im2col buffer); x2d (B*H*W, Cin) and wmat (Cout, kh*kw*Cin,
    K index = (ky, kx, ci)) TF32-rounded.  Returns None when the geometry is not served (caller: im2col + GEMM)."""
    if _PASSES != 1:
        return None
    Cin, Cout = x2d.shape[1], wmat.shape[0]
    d = _convgemm_desc(x2d, B, H, W, Cin, Cout, kh, kw, stride, pad)
    OH, OW = _conv_out(H, W, kh, kw, stride, pad)
    y = out if out is not None else _empty((B * OH * OW, Cout), x2d)
    d.mode, d.w, d.bias, d.y, d.ld_y, d.round_out = 0, _ptr(wmat), _ptr(bias), _ptr(y), y.stride(0), int(round_out)
    if not wmat.is_contiguous() or not fn["uwr_convgemm_tcgen05_supported"](C.byref(d)):
        return None
    rows, K = B * OH * OW, kh * kw * Cin
    _run("uwr_convgemm_tcgen05", f"fwd {kh}x{kw}s{stride} M{rows} N{Cout} K{K}", 4 * (B * H * W * Cin + Cout * K + rows * Cout),
         2.0 * rows * Cout * K, C.byref(d))
    return y


def conv_gemm_wgrad(dy2d, x2d, B, H, W, kh, kw, stride, pad):
    """Weight gradient of the convolution as an implicit GEMM; dy2d (B*OH*OW, Cout) and x2d TF32-rounded.  Returns dwmat
    as a (Cout, kh*kw*Cin) tensor (K index = (ky, kx, ci)) -- a transposed view when the kernel ran in the orientation
    that puts the im2col view on the 128-row operand (thin Cout) -- or None when the geometry is not served."""
    if _PASSES != 1:
        return None
    Cin, Cout = x2d.shape[1], dy2d.shape[1]
    K = kh * kw * Cin
    # MMA cycles ~ padded M x padded N: (Cout -> 128-row tiles, K -> column tiles) vs (K -> rows, Cout -> columns)
    pad_n = lambda n: 32 if n <= 32 else 64 if n <= 64 else 128 if n <= 128 else -(-n // 256) * 256
    transposed = -(-K // 128) * 128 * pad_n(Cout) < -(-Cout // 128) * 128 * pad_n(K)
    d = _convgemm_desc(x2d, B, H, W, Cin, Cout, kh, kw, stride, pad)
    d.mode, d.dy, d.ld_dy = (2 if transposed else 1), _ptr(dy2d), dy2d.stride(0)
    dw = _empty((K, Cout) if transposed else (Cout, K), x2d)
    d.dw = _ptr(dw)
    if not fn["uwr_convgemm_tcgen05_supported"](C.byref(d)):
        return None
    nbytes = fn["uwr_convgemm_tcgen05_workspace_bytes"](C.byref(d))
    ws = _ws(nbytes, x2d) if nbytes else None
    d.workspace, d.workspace_bytes = _ptr(ws), nbytes
    rows = dy2d.shape[0]
    _run("uwr_convgemm_tcgen05", f"wgrad{'T' if transposed else ''} {kh}x{kw}s{stride} M{Cout} N{K} K{rows}",
         4 * (B * H * W * Cin + rows * Cout + Cout * K), 2.0 * rows * Cout * K, C.byref(d))
    return dw.t() if transposed else dw


def im2col_4x4s2(tokens2d, B, H, W, Cc):
    col = _empty((B * (H // 2) * (W // 2), 16 * Cc), tokens2d)
    _run("uwr_im2col_4x4s2", "", 0, 0.0, _ptr(tokens2d), tokens2d.stride(0), _ptr(col), B, H, W, Cc)
    return col


def col2im_4x4s2(dcol, B, H, W, Cc):
    dx = _empty((B * H * W, Cc), dcol)
    _run("uwr_col2im_4x4s2", "", 0, 0.0, _ptr(dcol), _ptr(dx), B, H, W, Cc)
    return dx


def im2col_3x3(tokens2d, B, H, W, Cc):
    """tokens2d: (B*H*W, >=Cc) view (row stride = ld); returns col (B*H*W, 9*Cc), K order (ky,kx,ci)."""
    col = _empty((B * H * W, 9 * Cc), tokens2d)
    _run("uwr_im2col_3x3", f"B{B} H{H} C{Cc}", 4 * B * H * W * Cc * 10, 0.0, _ptr(tokens2d), tokens2d.stride(0),
         _ptr(col), B, H, W, Cc)
    return col


def col2im_3x3(dcol, out2d, B, H, W, Cc, accumulate=False):
    """out2d[:, :Cc] (+)= gathered dcol; out2d is a (B*H*W, >=Cc) view."""
    _run("uwr_col2im_3x3", f"B{B} H{H} C{Cc}", 4 * B * H * W * Cc * 10, 0.0, _ptr(dcol), _ptr(out2d),
         out2d.stride(0), B, H, W, Cc, int(accumulate))
    return out2d


def pixel_scatter_2x2(g, bias, out2d, B, H, W, Cout):
    _run("uwr_pixel_scatter_2x2", "", 0, 0.0, _ptr(g), _ptr(bias), _ptr(out2d), out2d.stride(0), B, H, W, Cout)


def pixel_gather_2x2(dout2d, B, H, W, Cout):
    dg = _empty((B * H * W, 4 * Cout), dout2d)
    _run("uwr_pixel_gather_2x2", "", 0, 0.0, _ptr(dout2d), dout2d.stride(0), _ptr(dg), B, H, W, Cout)
    return dg


def copy2d(src2d, dst2d, cols, accumulate=False):
    rows = src2d.shape[0]
    _run("uwr_copy2d", "", 0, 0.0, _ptr(src2d), src2d.stride(0), _ptr(dst2d), dst2d.stride(0), rows, cols,
                           int(accumulate))


def colsum(x2d, cols, out=None):
    rows = x2d.shape[0]
    if out is None:
        out = _empty((cols,), x2d)
    ws = _ws(1024 * cols * 4, x2d)
    _run("uwr_colsum", f"rows{rows} C{cols}", 4 * rows * cols, 0.0, _ptr(x2d), x2d.stride(0), _ptr(out), _ptr(ws),
         rows, cols)
    return out


# --------------------------------------------------------------------------------------------
def pixel_shuffle2(t2d, B, H, W, Cout, out=None, ocol=0):
    """PixelShuffle(2) on tokens: (B*H*W, 4*Cout) -> (B*2H*2W, Cout) [or columns ocol.. of a wider `out`]."""
    t2d = t2d if t2d.is_contiguous() else t2d.contiguous()
    if out is None:
        out = _empty((B * 4 * H * W, Cout), t2d)
    n = B * H * W * Cout * 4
    _run("uwr_pixel_shuffle2", f"B{B} H{H} C{Cout}", 8 * n, 0.0, _ptr(t2d), _colptr(out, ocol), out.stride(0), B, H, W, Cout)
    return out


def pixel_unshuffle2(t2d, B, H, W, Cout, icol=0):
    """PixelUnshuffle(2) on tokens: (B*2H*2W, ld)[:, icol:icol+Cout] -> (B*H*W, 4*Cout); H, W = the coarse grid."""
    out = _empty((B * H * W, 4 * Cout), t2d)
    n = B * H * W * Cout * 4
    _run("uwr_pixel_unshuffle2", f"B{B} H{H} C{Cout}", 8 * n, 0.0, _colptr(t2d, icol), t2d.stride(0), _ptr(out), B, H, W, Cout)
    return out


def _conv_small_desc(inp, in_tokens, weight, bias, B, H, W, out=None, out_tokens=True, residual=None, round_out=False):
    d = ConvSmallDesc()
    Cout, Cin = weight.shape[0], weight.shape[1]
    d.inp, d.in_tokens, d.ld_in = _ptr(inp), int(in_tokens), (inp.stride(0) if in_tokens else 0)
    d.weight, d.bias, d.residual_img = _ptr(weight), _ptr(bias), _ptr(residual)
    d.out, d.out_tokens, d.ld_out = _ptr(out), int(out_tokens), (out.stride(0) if (out is not None and out_tokens) else 0)
    d.B, d.H, d.W, d.Cin, d.Cout, d.round_out = B, H, W, Cin, Cout, int(round_out)
    return d


def conv3x3_small_fwd(inp, in_tokens, weight, bias, B, H, W, residual=None, round_out=False):
    """thin direct 3x3 conv: NCHW image (3 ch) -> tokens (8|16 ch), or tokens (8 ch) -> NCHW image (3 ch) [+ residual]."""
    Cout, Cin = weight.shape[0], weight.shape[1]
    weight = weight if weight.is_contiguous() else weight.contiguous()
    out_tokens = not in_tokens
    out = _empty((B * H * W, Cout), inp) if out_tokens else _empty((B, Cout, H, W), inp)
    d = _conv_small_desc(inp, in_tokens, weight, bias, B, H, W, out, out_tokens, residual, round_out)
    n = B * H * W
    _run("uwr_conv3x3_small_fwd", f"B{B} H{H} {Cin}->{Cout}", 4 * n * (Cin + Cout), 18.0 * n * Cin * Cout, C.byref(d))
    return out


def conv3x3_small_wgrad(inp, in_tokens, weight, dout, B, H, W, want_bias=True):
    """(dweight, dbias) of conv3x3_small_fwd; dout has the forward output's layout."""
    Cout, Cin = weight.shape[0], weight.shape[1]
    d = _conv_small_desc(inp, in_tokens, weight, None, B, H, W, None, not in_tokens)
    dweight = torch.empty_like(weight, memory_format=torch.contiguous_format)
    dbias = _empty((Cout,), inp) if want_bias else None
    ws = _ws(fn["uwr_conv3x3_small_wgrad_workspace_bytes"](B, H, W, Cin, Cout), inp)
    n = B * H * W
    _run("uwr_conv3x3_small_wgrad", f"B{B} H{H} {Cin}->{Cout}", 4 * n * (Cin + Cout), 18.0 * n * Cin * Cout,
         C.byref(d), _ptr(dout), dout.stride(0) if not in_tokens else 0, _ptr(dweight), _ptr(dbias), _ptr(ws))
    return dweight, dbias


def _ew(name, n, nbytes, *args):
    _run(name, f"n{n}", nbytes, 0.0, *args)


def polar_split_fwd(f):
    """f: (..., 2) interleaved complex -> (abs, angle), SpectralTransformer.py:176-177."""
    n = f.numel() // 2
    mag, pha = _empty(f.shape[:-1], f), _empty(f.shape[:-1], f)
    _ew("uwr_polar_split_fwd", n, 16 * n, _ptr(f), _ptr(mag), _ptr(pha), n)
    return mag, pha


def polar_split_bwd(f, dmag, dpha):
    n = f.numel() // 2
    df = torch.empty_like(f)
    _ew("uwr_polar_split_bwd", n, 24 * n, _ptr(f), _ptr(dmag), _ptr(dpha), _ptr(df), n)
    return df


def polar_join_fwd(mag, pha):
    n = mag.numel()
    z = _empty(tuple(mag.shape) + (2,), mag)
    _ew("uwr_polar_join_fwd", n, 16 * n, _ptr(mag), _ptr(pha), _ptr(z), n)
    return z


def polar_join_bwd(mag, pha, dz):
    n = mag.numel()
    dmag, dpha = torch.empty_like(mag), torch.empty_like(mag)
    _ew("uwr_polar_join_bwd", n, 32 * n, _ptr(mag), _ptr(pha), _ptr(dz), _ptr(dmag), _ptr(dpha), n)
    return dmag, dpha


def cabs_fwd(z):
    n = z.numel() // 2
    a = _empty(z.shape[:-1], z)
    _ew("uwr_cabs_fwd", n, 12 * n, _ptr(z), _ptr(a), n)
    return a


def cabs_bwd(z, da):
    n = z.numel() // 2
    dz = torch.empty_like(z)
    _ew("uwr_cabs_bwd", n, 20 * n, _ptr(z), _ptr(da), _ptr(dz), n)
    return dz


def leaky_relu_fwd(x, slope, round_out=False):
    y = torch.empty_like(x)
    _ew("uwr_leaky_relu_fwd", x.numel(), 8 * x.numel(), _ptr(x), _ptr(y), x.numel(), float(slope), int(round_out))
    return y


def gelu_fwd(x, round_out=False):
    y = torch.empty_like(x)
    _ew("uwr_gelu_fwd", x.numel(), 8 * x.numel(), _ptr(x), _ptr(y), x.numel(), int(round_out))
    return y


def gelu_bwd(x, dy):
    dx = torch.empty_like(x)
    _ew("uwr_gelu_bwd", x.numel(), 12 * x.numel(), _ptr(x), _ptr(dy), _ptr(dx), x.numel())
    return dx


def leaky_relu_bwd(y, dy, slope):
    dx = torch.empty_like(y)
    _ew("uwr_leaky_relu_bwd", y.numel(), 12 * y.numel(), _ptr(y), _ptr(dy), _ptr(dx), y.numel(), float(slope))
    return dx


def even_scatter(y, bias, B, H, W, Cc):
    out = _empty((B * 4 * H * W, Cc), y)
    _run("uwr_even_scatter", f"B{B} H{H} C{Cc}", 20 * B * H * W * Cc, 0.0, _ptr(y), _ptr(bias), _ptr(out), B, H, W, Cc)
    return out


def even_gather(dout, B, H, W, Cc):
    dy = _empty((B * H * W, Cc), dout)
    _run("uwr_even_gather", f"B{B} H{H} C{Cc}", 8 * B * H * W * Cc, 0.0, _ptr(dout), _ptr(dy), B, H, W, Cc)
    return dy


# --------------------------------------------------------------------------------------------
def _colptr(t, col):
    """device pointer of column `col` of a 2-D fp32 token matrix"""
    _ptr(t)
    return t.data_ptr() + 4 * col


def mdta_gram(x, xcol, y, ycol, B, L, heads, c, want_sq=False):
    """G (B, heads, c, c) = per-head X^T Y over the L tokens of each image (+ squared column norms (B, heads*c))."""
    C = heads * c
    G = _empty((B, heads, c, c), x)
    sqx = _empty((B, C), x) if want_sq else None
    sqy = _empty((B, C), x) if want_sq else None
    ws = _ws(fn["uwr_mdta_gram_workspace_bytes"](B, L, heads, c), x)
    _run("uwr_mdta_gram", f"B{B} L{L} h{heads} c{c}", 8 * B * L * C, 2.0 * B * L * C * c,
         _colptr(x, xcol), x.stride(0), _colptr(y, ycol), y.stride(0), B, L, heads, c, _ptr(G), _ptr(sqx), _ptr(sqy),
         _ptr(ws))
    return G, sqx, sqy


def mdta_apply(x, xcol, Mx, B, L, heads, c, transpose=False, yd=None, ycol=0, diag=None, out=None, ocol=0):
    """out[:, ocol:ocol+C] = per-head M (or M^T) applied to the channels of x[:, xcol:xcol+C] (+ diag * yd)."""
    C = heads * c
    if out is None:
        out = _empty((B * L, C), x)
    _run("uwr_mdta_apply", f"B{B} L{L} h{heads} c{c}", (8 + (4 if yd is not None else 0)) * B * L * C,
         2.0 * B * L * C * c, _colptr(x, xcol), x.stride(0), _ptr(Mx), int(transpose),
         _colptr(yd, ycol) if yd is not None else None, yd.stride(0) if yd is not None else 0, _ptr(diag),
         _colptr(out, ocol), out.stride(0), B, L, heads, c)
    return out


# --------------------------------------------------------------------------------------------
LOSS_KINDS = {"L1": 0, "L1withColor": 1, "charbonnier": 2, "L2": 3, "mse01": 4}


def dft_real(x, B, H, W, C, scale, axes):
    """scale * Re(FFT2(x)) of a real (B, H, W, C) token tensor over axes 'hw' (spatial) or 'lc'
    (tokens x channels); symmetric linear map => also its own backward (csrc/fft.cu)."""
    x = x.contiguous()
    y = torch.empty_like(x)
    ws = _ws(fn["uwr_dft_workspace_bytes"](B, H, W, C), x)
    name = "uwr_dft_hw_real" if axes == "hw" else "uwr_dft_lc_real"
    n = x.numel()
    _run(name, f"B{B} H{H} W{W} C{C}", 4 * n * (10 if axes == "lc" else 6), 5.0 * n * math.log2(H * W * (C if axes == "lc" else 1)),
         _ptr(x), _ptr(y), _ptr(ws), B, H, W, C, float(scale))
    return y


def fft2_hw(x, B, H, W, C, in_complex, inverse, scale=1.0):
    """complex FFT2 over (H, W) of (B, H, W, C) [real] or (B, H, W, C, 2) [interleaved complex];
    returns (B, H, W, C, 2) float32 (view_as_complex-compatible)."""
    x = x.contiguous()
    out = torch.empty((B, H, W, C, 2), device=x.device, dtype=torch.float32)
    ws = _ws(fn["uwr_dft_workspace_bytes"](B, H, W, C), x)
    n = B * H * W * C
    _run("uwr_fft2_hw", f"B{B} H{H} W{W} C{C} c{int(in_complex)} i{int(inverse)}", 4 * n * (7 if not in_complex else 8),
         5.0 * n * math.log2(H * W), _ptr(x), _ptr(out), _ptr(ws), B, H, W, C, int(in_complex), int(inverse), float(scale))
    return out


def pixel_loss(pred, truth, kind, batch_divisor=None, want_grad=True):
    B, Cc, H, W = pred.shape
    out = _empty((1,), pred)
    grad = torch.empty_like(pred) if want_grad else None
    ws = _ws(4 * 1024 * 4, pred)
    _run("uwr_pixel_loss", "", 0, 0.0, _ptr(pred), _ptr(truth), _ptr(out), _ptr(grad), _ptr(ws), LOSS_KINDS[kind],
                               B, Cc, H, W, batch_divisor or B)
    return out, grad
