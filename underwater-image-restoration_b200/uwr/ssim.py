"""Device versions of the Laplacian gradient loss, MS-SSIM and SSIM (csrc/ssim.cu): the two remaining non-VGG terms of
the reference's "fflMix" loss (src/Losses/losses.py:108-117,162-181) and ModelTrainer.torchSSIM
(src/ModelTrainer.py:23-24).  pytorch_msssim is a third-party package absent from the reference tree; its published
algorithm (VainF/pytorch-msssim 1.0.0) is restated in SURVEY.md Appendix C; the parity tests compare against a CPU
stand-in of that package ("restatement-pinned").

The per-pixel work (separable 11-tap Gaussian blurs of X, Y, X^2, Y^2, XY, the cs / ssim maps, and the blur adjoint
in the backward) runs in the kernels; MS-SSIM's relu / product-of-powers over (planes x 5 scales) scalars is torch
algebra on tiny tensors, differentiated by autograd to get the per-plane coefficient each scale's backward kernel
scales its gradient with.
"""
import torch
from torch.autograd.function import once_differentiable

from . import ops
from ._lib import fn

MS_WEIGHTS = (0.0448, 0.2856, 0.3001, 0.2363, 0.1333)
WIN = 11


_WTS = {}


def _weights(like):
    """MS-SSIM scale weights as a device tensor, created once per device (a host -> device copy is not allowed inside
    a CUDA-graph capture; the eager warm-up steps that precede a capture create it)."""
    if like.device not in _WTS:
        _WTS[like.device] = torch.tensor(MS_WEIGHTS, dtype=torch.float32, device=like.device)
    return _WTS[like.device]


def _c(t):
    return t if t.is_contiguous() else t.contiguous()


class LaplacianL1Fn(torch.autograd.Function):
    """Gradient_Loss (losses.py:162-181): F.l1_loss of the valid 3x3 Laplacians of pred and truth."""

    @staticmethod
    def forward(ctx, pred, truth):
        pred, truth = _c(pred), _c(truth)
        B, Cc, H, W = pred.shape
        out = ops._empty((1,), pred)
        want = ctx.needs_input_grad[0]
        grad = torch.empty_like(pred) if want else None
        ws = ops._ws(fn["uwr_laplacian_l1_workspace_bytes"](B * Cc, H, W), pred)
        ops._run("uwr_laplacian_l1_loss", f"planes{B * Cc} {H}x{W}", 12 * pred.numel(), 0.0, ops._ptr(pred), ops._ptr(truth),
                 ops._ptr(out), ops._ptr(grad), ops._ptr(ws), B * Cc, H, W)
        if want:
            ctx.save_for_backward(grad)
        return out.view(())

    @staticmethod
    @once_differentiable
    def backward(ctx, gout):
        (grad,) = ctx.saved_tensors
        return grad * gout, None


def _scale_fwd(X, Y, planes, H, W, data_range, full, want_maps):
    maps = ops._empty((planes, 3, H - WIN + 1, W - WIN + 1), X) if want_maps else None
    mcs, mss = ops._empty((planes,), X), ops._empty((planes,), X)
    ws = ops._ws(fn["uwr_ssim_workspace_bytes"](planes, H, W), X)
    ops._run("uwr_ssim_scale_fwd", f"planes{planes} {H}x{W}", 4 * planes * H * W * (2 + (3 if want_maps else 0)), 0.0,
             ops._ptr(X), ops._ptr(Y), ops._ptr(maps), ops._ptr(mcs), ops._ptr(mss), ops._ptr(ws), planes, H, W,
             float(data_range), int(full))
    return maps, mcs, mss


class MsSsimFn(torch.autograd.Function):
    """pytorch_msssim.MS_SSIM(data_range, size_average=True, win 11, sigma 1.5, 5 scales) of (X, Y); the gradient goes
    to X only (the prediction, losses.py:114)."""

    @staticmethod
    def forward(ctx, X, Y, data_range):
        X, Y = _c(X), _c(Y)
        B, Cc, H, W = X.shape
        if min(H, W) <= (WIN - 1) * 16:
            raise AssertionError("Image size should be larger than %d due to the 4 downsamplings in ms-ssim"
                                 % ((WIN - 1) * 16))
        if H % 16 or W % 16:
            raise ValueError("uwr ms_ssim needs H and W divisible by 16 (even sides at every scale)")
        planes = B * Cc
        want = ctx.needs_input_grad[0]
        pyr, vals = [], []
        x, y, h, w = X, Y, H, W
        for s in range(5):
            maps, mcs, mss = _scale_fwd(x, y, planes, h, w, data_range, s == 4, want)
            pyr.append((x, y, maps, h, w))
            vals.append(mss if s == 4 else mcs)
            if s < 4:
                xo, yo = ops._empty((planes, h // 2, w // 2), X), ops._empty((planes, h // 2, w // 2), X)
                ops._run("uwr_avgpool2_pair", f"planes{planes} {h}x{w}", 10 * planes * h * w, 0.0, ops._ptr(x), ops._ptr(y),
                         ops._ptr(xo), ops._ptr(yo), planes, h, w)
                x, y, h, w = xo, yo, h // 2, w // 2
        with torch.enable_grad():
            leaves = [v.detach().requires_grad_() for v in vals]
            wts = _weights(X)
            # product of powers written as explicit multiplications: torch.prod's backward counts zeros with .item(),
            # a host sync that invalidates a CUDA-graph capture
            res = None
            for s, v in enumerate(leaves):
                term = torch.relu(v) ** wts[s]
                res = term if res is None else res * term
            res = res.mean()
        if want:
            ctx.graph = (leaves, res)
            ctx.pyr = pyr
            ctx.planes = planes
            ctx.shape = X.shape
        return res.detach()

    @staticmethod
    @once_differentiable
    def backward(ctx, gout):
        leaves, res = ctx.graph
        with torch.enable_grad():
            dvals = torch.autograd.grad(res, leaves, gout)
        dX = None
        for s in (4, 3, 2, 1, 0):
            x, y, maps, h, w = ctx.pyr[s]
            coef = _c(dvals[s] / float((h - WIN + 1) * (w - WIN + 1)))       # value = MEAN over the valid pixels
            out = ops._empty((ctx.planes, h, w), x)
            ops._run("uwr_ssim_scale_bwd", f"planes{ctx.planes} {h}x{w}", 4 * ctx.planes * h * w * 6, 0.0, ops._ptr(x),
                     ops._ptr(y), ops._ptr(maps), ops._ptr(coef), ops._ptr(dX), ops._ptr(out), ctx.planes, h, w)
            dX = out
        ctx.graph = ctx.pyr = None
        return dX.view(ctx.shape), None, None


def ms_ssim(X, Y, data_range=1.0):
    return MsSsimFn.apply(X, Y, data_range)


def gradient_loss(pred, truth):
    return LaplacianL1Fn.apply(pred, truth)


def ssim(X, Y, data_range=1.0, size_average=True):
    """pytorch_msssim.ssim (one scale): ModelTrainer.torchSSIM calls ssim(tar, prd, data_range=1.0, size_average=True)."""
    X, Y = _c(X.detach()), _c(Y.detach())
    B, Cc, H, W = X.shape
    _, _, mss = _scale_fwd(X, Y, B * Cc, H, W, data_range, True, False)
    per = mss.view(B, Cc).mean(1)
    return per.mean() if size_average else per
