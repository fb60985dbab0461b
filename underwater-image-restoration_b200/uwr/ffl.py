"""Focal frequency loss on the shared-memory FFT kernels (uwr_ffl_loss); drop-in for the
focal_frequency_loss.FocalFrequencyLoss(loss_weight=1, alpha=1) the reference constructs at
src/Losses/losses.py:48."""
import torch
from torch.autograd.function import once_differentiable

from . import ops
from ._lib import fn


class FocalFrequencyFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, truth):
        pred = pred.contiguous()
        truth = truth.contiguous()
        B, C, H, W = pred.shape
        if H != W:
            raise ValueError("uwr focal frequency loss needs square power-of-two planes")
        out = ops._empty((1,), pred)
        want = ctx.needs_input_grad[0]
        grad = torch.empty_like(pred) if want else None
        ws = ops._ws(fn["uwr_ffl_workspace_bytes"](B * C, H), pred)
        ops._run("uwr_ffl_loss", f"planes{B * C} S{H}", 12 * pred.numel(), 0.0, ops._ptr(pred), ops._ptr(truth),
                 ops._ptr(out), ops._ptr(grad), ops._ptr(ws), B * C, H)
        if want:
            ctx.save_for_backward(grad)
        return out.view(())

    @staticmethod
    @once_differentiable
    def backward(ctx, gout):
        (grad,) = ctx.saved_tensors
        return grad * gout, None


def ffl_value_and_grad(pred, truth):
    """(loss (1,), dloss/dpred) in one call of uwr_ffl_loss (used by the torch.library op uwr::ffl_loss)."""
    B, C, H, W = pred.shape
    if H != W:
        raise ValueError("uwr focal frequency loss needs square power-of-two planes")
    out = ops._empty((1,), pred)
    grad = torch.empty_like(pred)
    ws = ops._ws(fn["uwr_ffl_workspace_bytes"](B * C, H), pred)
    ops._run("uwr_ffl_loss", f"planes{B * C} S{H}", 12 * pred.numel(), 0.0, ops._ptr(pred), ops._ptr(truth),
             ops._ptr(out), ops._ptr(grad), ops._ptr(ws), B * C, H)
    return out, grad
