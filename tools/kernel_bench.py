"""Per-kernel timing at the AST B=16 hot shapes (CUDA events, inputs >> L2). Used under ncu too.
usage: python tools/kernel_bench.py [which ...]   which in {t5nt,t5nn,t5tn,legacy,dwf,dwb,attnb,attnf,lnb,colsum}"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "underwater-image-restoration_b200"))
import torch
from uwr import ops

which = sys.argv[1:] or ["t5nt", "t5nn", "t5tn", "legacy", "dwf", "dwb", "attnb", "attnf", "lnb", "colsum"]
dev = "cuda"
M = 16 * 65536


def rnd(*s):
    x = torch.randn(*s, device=dev)
    return ((x.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)


def timeit(name, fn, nbytes, flops, reps=5):
    reps = int(os.environ.get("UWR_KB_REPS", reps))   # UWR_KB_REPS=1: one launch per kernel (ncu --set full captures)
    for _ in range(2 if reps > 1 else 0):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{name:34s} {ms:8.3f} ms  {nbytes / ms / 1e6:8.0f} GB/s  {flops / ms / 1e9:8.1f} TFLOP/s", flush=True)


if "t5nt" in which or "legacy" in which:
    x, w, b = rnd(M, 64), rnd(256, 64), rnd(256)
    out = torch.empty(M, 256, device=dev)
    nb, fl = 4 * (M * 64 + 256 * 64 + M * 256), 2.0 * M * 256 * 64
    if "t5nt" in which:
        timeit("t5 NT M1M N256 K64 bias", lambda: ops.linear(x, w, b, out=out, t5=True), nb, fl)
    if "legacy" in which:
        timeit("legacy NT M1M N256 K64 bias", lambda: ops.linear(x, w, b, out=out, t5=False), nb, fl)
    h2, w2, r = rnd(M, 256), rnd(64, 256), rnd(M, 64)
    out2 = torch.empty(M, 64, device=dev)
    nb2, fl2 = 4 * (M * 256 + 2 * M * 64), 2.0 * M * 256 * 64
    if "t5nt" in which:
        timeit("t5 NT M1M N64 K256 resid", lambda: ops.linear(h2, w2, None, residual=r, out=out2, t5=True), nb2, fl2)
    if "legacy" in which:
        timeit("legacy NT M1M N64 K256 resid", lambda: ops.linear(h2, w2, None, residual=r, out=out2, t5=False), nb2, fl2)
    del x, out, h2, r, out2
if "t5nn" in which:
    d, w2, v = rnd(M, 64), rnd(64, 256), rnd(M, 256)
    out = torch.empty(M, 256, device=dev)
    timeit("t5 NN M1M N256 K64 mul", lambda: ops.linear_dgrad(d, w2, mul_by=v, out=out, t5=True),
           4 * (M * 64 + 2 * M * 256), 2.0 * M * 256 * 64)
    timeit("t5 NN M1M N256 K64 plain", lambda: ops.linear_dgrad(d, w2, out=out, t5=True),
           4 * (M * 64 + M * 256), 2.0 * M * 256 * 64)
    del d, v, out
if "t5h" in which:
    # fp16 storage of the LeFF tensors: linear1 forward with a float16 C, linear2 data gradient with the
    # float16 gelu' multiplier
    x, w, b = rnd(M, 64), rnd(256, 64), rnd(256)
    outh = torch.empty(M, 256, device=dev, dtype=torch.float16)
    timeit("t5 NT M1M N256 K64 bias -> fp16 C", lambda: ops.linear(x, w, b, out=outh, t5=True),
           4 * (M * 64 + 256 * 64) + 2 * M * 256, 2.0 * M * 256 * 64)
    d, w2 = rnd(M, 64), rnd(64, 256)
    vh = torch.rand(M, 256, device=dev).half()
    out = torch.empty(M, 256, device=dev)
    timeit("t5 NN M1M N256 K64 mul fp16 R", lambda: ops.linear_dgrad(d, w2, mul_by=vh, out=out, t5=True),
           4 * (M * 64 + M * 256) + 2 * M * 256, 2.0 * M * 256 * 64)
    del x, outh, d, vh, out
if "dwh" in which:
    B, H, Ch = 16, 256, 256
    uh = torch.randn(B * H * H, Ch, device=dev).half()
    wt, bs = torch.randn(Ch, 1, 3, 3, device=dev) * 0.3, torch.randn(Ch, device=dev)
    n = B * H * H * Ch
    timeit("dwconv fwd fp16 u/v B16 H256 Ch256", lambda: ops.dwconv_gelu_fwd_half(uh, wt, bs, B, H, H, Ch), n * 8, 18.0 * n)
    dv = torch.randn(B * H * H, Ch, device=dev)
    du = torch.empty(B * H * H, Ch, device=dev)
    timeit("dwconv bwd fp16 u B16 H256 Ch256", lambda: ops.dwconv_gelu_bwd(dv, uh, wt, B, H, H, Ch, du=du), n * 10, 36.0 * n)
    del uh, dv, du
if "mdta" in which:
    B, L, heads, c = 8, 65536, 1, 32
    C = heads * c
    qkv = torch.randn(B * L, 3 * C, device=dev)
    A = torch.softmax(torch.randn(B, heads, c, c, device=dev), -1)
    timeit("mdta gram B8 L65536 C32", lambda: ops.mdta_gram(qkv, 0, qkv, C, B, L, heads, c, want_sq=True), 8 * B * L * C, 2.0 * B * L * C * c)
    timeit("mdta apply B8 L65536 C32", lambda: ops.mdta_apply(qkv, 2 * C, A, B, L, heads, c), 8 * B * L * C, 2.0 * B * L * C * c)
    del qkv
if "t5big" in which:
    # tensor-bound shapes of the bottleneck / dec0 stages: cluster pairs with B-tile multicast vs single CTAs
    for (lay, Mm, Nn, Kk) in (("nt", 16384, 2048, 512), ("nt", 65536, 1024, 256), ("nn", 16384, 2048, 512), ("nn", 65536, 256, 1024),
                              ("tn", 512, 2048, 16384), ("tn", 1024, 256, 65536)):
        for on in (False, True):
            ops.set_gemm_cluster(bool(on))
            if lay == "nt":
                a, b_ = rnd(Mm, Kk), rnd(Nn, Kk)
                o = torch.empty(Mm, Nn, device=dev)
                f = lambda: ops.linear(a, b_, None, out=o, t5=True)
            elif lay == "nn":
                a, b_ = rnd(Mm, Kk), rnd(Kk, Nn)
                o = torch.empty(Mm, Nn, device=dev)
                f = lambda: ops.linear_dgrad(a, b_, out=o, t5=True)
            else:
                a, b_ = rnd(Kk, Mm), rnd(Kk, Nn)
                f = lambda: ops.linear_wgrad(a, b_, want_bias=False, t5=True)
            timeit(f"t5 {lay.upper()} M{Mm} N{Nn} K{Kk} {'cluster2+multicast' if on else 'single CTA'}", f,
                   4.0 * (Mm * Kk + Kk * Nn + Mm * Nn), 2.0 * Mm * Nn * Kk, reps=20)
    ops.set_gemm_cluster(True)
if "t5tn" in which:
    du, y = rnd(M, 256), rnd(M, 64)
    timeit("t5 TN M256 N64 K1M", lambda: ops.linear_wgrad(du, y, want_bias=False, t5=True), 4 * (M * 320), 2.0 * M * 256 * 64)
    timeit("legacy TN M256 N64 K1M", lambda: ops.linear_wgrad(du, y, want_bias=False, t5=False), 4 * (M * 320), 2.0 * M * 256 * 64)
    del du, y
if "dwf" in which or "dwb" in which:
    B, H, Ch = 16, 256, 256
    u = torch.randn(B * H * H, Ch, device=dev)
    wt, bs = torch.randn(Ch, 1, 3, 3, device=dev) * 0.3, torch.randn(Ch, device=dev)
    n = B * H * H * Ch
    if "dwf" in which:
        timeit("dwconv fwd B16 H256 Ch256", lambda: ops.dwconv_gelu_fwd(u, wt, bs, B, H, H, Ch, v_is_dgelu=True), 4 * n * 3, 18.0 * n)
    if "dwb" in which:
        dv = torch.randn(B * H * H, Ch, device=dev)
        du = torch.empty_like(u)
        timeit("dwconv bwd B16 H256 Ch256", lambda: ops.dwconv_gelu_bwd(dv, u, wt, B, H, H, Ch, du=du), 4 * n * 3, 36.0 * n)
        del dv, du
    del u
if "attnb" in which or "attnf" in which:
    B, H, heads, hd = 16, 256, 2, 32
    C = heads * hd
    qkv = torch.randn(B * H * H, 3 * C, device=dev)
    table = torch.randn(225, heads, device=dev) * 0.02
    w = torch.ones(2, device=dev)
    tiles = B * (H // 8) ** 2 * heads
    if "attnf" in which:
        for rounded in (False, True):     # True: operands are exact TF32 values -> one tensor-core pass per product
            for t5 in (False, True):
                ops.set_attn_tcgen05(t5)
                timeit(f"attn fwd tiles32768 hd32 {'tcgen05' if t5 else 'mma.sync'}{' exact-tf32 operands' if rounded else ''}",
                       lambda: ops.window_attn_fwd(qkv, 0, qkv, C, 2 * C, table, w, B, H, H, heads, hd, 4, hd ** -0.5, rounded=rounded),
                       tiles * 4 * 64 * hd * 4, tiles * 4.0 * 64 * 64 * hd)
        ops.set_attn_tcgen05("auto")
    if "attnb" in which:
        do = torch.randn(B * H * H, C, device=dev)
        dq = torch.empty_like(qkv)
        for rounded in (False, True):
            timeit(f"attn bwd tiles32768 hd32{' exact-tf32 operands' if rounded else ''}",
                   lambda: ops.window_attn_bwd(do, qkv, 0, qkv, C, 2 * C, table, w, B, H, H, heads, hd, 4, hd ** -0.5, dq_buf=dq, dkv_buf=dq, rounded=rounded),
                   tiles * 7 * 64 * hd * 4, tiles * 10.0 * 64 * 64 * hd)
        del do, dq
    del qkv
if "lnb" in which:
    x, dy, dres = torch.randn(M, 64, device=dev), torch.randn(M, 64, device=dev), torch.randn(M, 64, device=dev)
    g, b = torch.ones(64, device=dev), torch.zeros(64, device=dev)
    y, mean, rstd = ops.layernorm_fwd(x, g, b)
    timeit("LN fwd rows1M C64", lambda: ops.layernorm_fwd(x, g, b), 8 * M * 64, 0.0)
    timeit("LN bwd rows1M C64 (+dres)", lambda: ops.layernorm_bwd(dy, x, g, mean, rstd, dres=dres), 16 * M * 64, 0.0)
if "colsum" in which:
    x = torch.randn(M, 256, device=dev)
    timeit("colsum rows1M C256", lambda: ops.colsum(x, 256), 4 * M * 256, 0.0)
if "fft" in which:
    for (B, H, C) in ((16, 256, 32), (16, 64, 128)):
        x = torch.randn(B, H, H, C, device=dev)
        n = x.numel()
        timeit(f"dft_hw_real B{B} {H}x{H} C{C}", lambda: ops.dft_real(x, B, H, H, C, 1.0, "hw"), 4 * n * 6, 5.0 * n * 2 * (H.bit_length() - 1))
        timeit(f"dft_lc_real B{B} {H}x{H} C{C}", lambda: ops.dft_real(x, B, H, H, C, 1.0, "lc"), 4 * n * 10, 5.0 * n * (2 * (H.bit_length() - 1) + (C.bit_length() - 1)))
        del x

if "oproj" in which:
    # AST boundary convolutions at B = 16, 256 x 256: InputProj 3 -> 32, OutputProj 64 -> 3 (+ residual)
    B_, H_ = 16, 256
    img = torch.randn(B_, 3, H_, H_, device=dev)
    wi, bi = torch.randn(32, 3, 3, 3, device=dev) * 0.1, torch.randn(32, device=dev) * 0.1
    tok64 = torch.randn(B_, H_ * H_, 64, device=dev)
    wo, bo = torch.randn(3, 64, 3, 3, device=dev) * 0.1, torch.randn(3, device=dev) * 0.1
    dimg = torch.randn(B_, 3, H_, H_, device=dev)
    npx = B_ * H_ * H_
    timeit("input_proj fwd 3->32", lambda: ops.input_proj_fwd(img, wi, bi), 4 * npx * (3 + 32), 2.0 * npx * 27 * 32)
    tk = ops.input_proj_fwd(img, wi, bi)
    dtk = torch.randn_like(tk)
    timeit("input_proj bwd 3->32", lambda: ops.input_proj_bwd(dtk, tk, img, wi), 4 * npx * (3 + 64), 2.0 * npx * 27 * 32)
    timeit("output_proj fwd 64->3", lambda: ops.output_proj_fwd(tok64, wo, bo, img, B_, H_, H_), 4 * npx * (64 + 6),
           2.0 * npx * 27 * 64)
    timeit("output_proj bwd 64->3 (tensor core)", lambda: ops.output_proj_bwd(dimg, tok64, wo, B_, H_, H_),
           4 * npx * (128 + 3), 4.0 * npx * 27 * 64)
    ops.set_gemm_precision("tf32x3")
    timeit("output_proj bwd 64->3 (scalar fp32)", lambda: ops.output_proj_bwd(dimg, tok64, wo, B_, H_, H_),
           4 * npx * (128 + 3), 4.0 * npx * 27 * 64)
    ops.set_gemm_precision("tf32")

if "conv" in which:
    # dense convolutions: im2col + GEMM vs the implicit GEMM (TMA-materialised im2col tiles)
    for (B_, H_, Cin, Cout, k, st) in ((16, 256, 32, 64, 4, 2), (16, 128, 64, 128, 4, 2), (16, 64, 128, 256, 4, 2),
                                       (16, 256, 32, 32, 3, 1), (16, 128, 64, 64, 3, 1)):
        x = rnd(B_ * H_ * H_, Cin)
        wm = rnd(Cout, k * k * Cin) * 0.1
        wm = ((wm.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)
        OH = H_ // st
        rows = B_ * OH * OH
        dy = rnd(rows, Cout)
        nb, fl = 4 * (x.numel() + wm.numel() + rows * Cout), 2.0 * rows * Cout * k * k * Cin
        tag = f"{k}x{k}s{st} B{B_} H{H_} {Cin}->{Cout}"
        im2col = (lambda: ops.im2col_4x4s2(x, B_, H_, H_, Cin)) if k == 4 else (lambda: ops.im2col_3x3(x, B_, H_, H_, Cin))
        timeit(f"conv {tag} im2col", im2col, nb, 0.0)
        col = im2col()
        timeit(f"conv {tag} GEMM on col", lambda: ops.linear(col, wm, None, t5=True), nb, fl)
        timeit(f"conv {tag} implicit fwd", lambda: ops.conv_gemm_fwd(x, wm, None, B_, H_, H_, k, k, st, 1), nb, fl)
        timeit(f"conv {tag} wgrad on col", lambda: ops.linear_wgrad(dy, col, want_bias=False, t5=True), nb, fl)
        timeit(f"conv {tag} implicit wgrad", lambda: ops.conv_gemm_wgrad(dy, x, B_, H_, H_, k, k, st, 1), nb, fl)
        del col
