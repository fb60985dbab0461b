"""Does the LeFF chain run faster when it walks the batch in L2-sized chunks?  Its 4C-wide intermediates are produced by one
kernel and consumed by the next (u: linear1 -> dwconv; h2: dwconv -> linear2; dv: dgrad2 -> dwconv_bwd; du: dwconv_bwd ->
dgrad1, wgrad1); over the full batch of 16 every one of them (0.5 - 1 GB) leaves the 126 MB L2 before its consumer starts.
Per image they are 17 - 67 MB.  Forward and backward chains, full batch vs chunks of 1 / 2 / 4 images, each as one CUDA graph.
usage: python tools/micro/l2_chunk_ffn.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "underwater-image-restoration_b200"))
import torch
from uwr import ops

dev = "cuda"
B = 16


def rnd(*s, scale=1.0):
    x = torch.randn(*s, device=dev) * scale
    return ((x.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)


def graph_time(fn, reps=10):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn(); fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            keep = fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    del keep
    return e0.elapsed_time(e1) / reps


for H, C in ((256, 64), (128, 128), (64, 256)):
    Ch, L = 4 * C, H * H
    M = B * L
    x = rnd(M, C)
    n2w, n2b = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    w1, b1 = rnd(Ch, C, scale=0.05), rnd(Ch, scale=0.01)
    dww, dwb = torch.randn(Ch, 1, 3, 3, device=dev) * 0.2, torch.zeros(Ch, device=dev)
    w2, b2 = rnd(C, Ch, scale=0.05), rnd(C, scale=0.01)
    d_s = rnd(M, C, scale=1e-3)
    gw1, gw2 = torch.empty(Ch, C, device=dev), torch.empty(C, Ch, device=dev)

    def fwd(nb):
        saved = []
        out = torch.empty(M, C, device=dev)
        for b0 in range(0, B, nb):
            r = slice(b0 * L, (b0 + nb) * L)
            y2, mean, rstd = ops.layernorm_fwd(x[r], n2w, n2b)
            u = ops.linear(y2, w1, b1, t5=True, out_half=True)
            v, h2 = ops.dwconv_gelu_fwd_half(u, dww, dwb, nb, H, H, Ch)
            ops.linear(h2, w2, b2, residual=x[r], t5=True, out=out[r])
            saved.append((y2, mean, rstd, u, v, h2))
        return saved, out

    saved_full, _ = fwd(B)
    y2f, _, _, uf, vf, h2f = saved_full[0]

    def bwd(nb):
        keep = []
        for b0 in range(0, B, nb):
            r = slice(b0 * L, (b0 + nb) * L)
            dv = ops.linear_dgrad(d_s[r], w2, mul_by=vf[r], t5=True)
            ops.linear_wgrad(d_s[r], h2f[r], want_bias=False, t5=True, out=gw2)
            du, ddww, ddwb, db1 = ops.dwconv_gelu_bwd(dv, uf[r], dww, nb, H, H, Ch, want_du_colsum=True)
            del dv
            dy2 = ops.linear_dgrad(du, w1, t5=True)
            ops.linear_wgrad(du, y2f[r], want_bias=False, t5=True, out=gw1)
            del du
            keep.append(dy2)
        return keep

    for name, fn in (("fwd", fwd), ("bwd", bwd)):
        base = None
        for nb in (16, 4, 2, 1):
            ms = graph_time(lambda: fn(nb))
            base = base or ms
            print(f"H{H} C{C} {name}: {nb:2d} images per chunk  {ms:7.3f} ms  ({ms / base:.2f}x of the full batch)", flush=True)
    del saved_full, y2f, uf, vf, h2f
    torch.cuda.empty_cache()
