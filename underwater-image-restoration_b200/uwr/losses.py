"""Loss registry: drop-in for src/Losses/losses.py::LossFunction (lines 30-160).

"L1", "L2", "L1withColor", "charbonnier" run in one fused value+gradient kernel (uwr_pixel_loss);
"ffl" / "fflCharbonnier" use the shared-memory FFT focal-frequency kernel.  `batch_divisor`
lets a data-parallel caller pass the GLOBAL batch for the reference's "/ (B*C)" (SURVEY.md §8e).
"""
import torch
from torch.autograd.function import once_differentiable

from . import ops


class PixelLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, truth, kind, batch_divisor):
        pred = pred.contiguous()
        truth = truth.contiguous()
        want = ctx.needs_input_grad[0]
        out, grad = ops.pixel_loss(pred, truth, kind, batch_divisor, want_grad=want)
        if want:
            ctx.save_for_backward(grad)
        return out.view(())

    @staticmethod
    @once_differentiable
    def backward(ctx, gout):
        (grad,) = ctx.saved_tensors
        return grad * gout, None, None, None


_PIXEL_KINDS = ("L1", "L2", "L1withColor", "charbonnier")


class LossFunction:
    def __init__(self, loss_name, device=None, batch_divisor=None, vgg_weights=None):
        """vgg_weights ("fflMix" only): None = find torchvision's pretrained VGG16 checkpoint (UWR_VGG16_WEIGHTS or
        the torch hub cache; raises if absent), a path, or "random" for the seeded stand-in (uwr/fflmix.py)."""
        self.loss_name = loss_name
        self.device = device
        self.batch_divisor = batch_divisor
        self.vgg_weights = vgg_weights

    def getloss(self, predicted_data, truth_data):
        name = self.loss_name
        if name in _PIXEL_KINDS:
            return PixelLossFn.apply(predicted_data, truth_data, name, self.batch_divisor)
        if name in ("ffl", "fflCharbonnier"):
            from .ffl import FocalFrequencyFn
            ffl = FocalFrequencyFn.apply(predicted_data, truth_data)
            if name == "ffl":
                return ffl
            return ffl + PixelLossFn.apply(predicted_data, truth_data, "charbonnier", None)
        if name == "fflMix":
            from .fflmix import fflmix_loss
            return fflmix_loss(self, predicted_data, truth_data)
        raise ValueError(f"Unsupported loss: {self.loss_name}")
