"""Validation metrics of the reference trainer on the device.

torchPSNR (src/ModelTrainer.py:17-21, duplicate in src/utils/utils.py:31-35):
    20 * log10(1 / sqrt(mean((clamp01(prd) - clamp01(tar))^2)))   over the WHOLE batch.
The squared-error mean is one pass of the native pixel-loss kernel (kind "mse01").  Under data
parallelism the per-rank means are averaged BEFORE the logarithm (SURVEY.md §8e caveat 2), which
equals the single-process whole-batch value when every rank holds the same number of images.
UIQM (uqim_utils.py) is a CPU metric in the reference and is not part of this module.
"""
import torch

from . import ops


def clamped_mse(tar_img, prd_img):
    """mean((clamp01(prd) - clamp01(tar))^2) as a 1-element device tensor (no host sync)."""
    if not (tar_img.is_cuda and prd_img.is_cuda):
        raise RuntimeError("uwr.metrics runs on CUDA tensors only; there is no CPU fallback")
    if tar_img.shape != prd_img.shape or tar_img.dim() != 4:
        raise ValueError(f"torchPSNR expects two (B, C, H, W) tensors of one shape, got {tuple(tar_img.shape)} "
                         f"and {tuple(prd_img.shape)}")
    mse, _ = ops.pixel_loss(prd_img.detach().float().contiguous(), tar_img.detach().float().contiguous(), "mse01",
                            want_grad=False)
    return mse


def torchPSNR(tar_img, prd_img, group=None):
    """Drop-in for src/ModelTrainer.py:17-21; returns a 0-dim tensor.  `group`: all-reduce the
    squared-error mean over the data-parallel ranks first."""
    mse = clamped_mse(tar_img, prd_img)
    if group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()
                             and torch.distributed.get_world_size() > 1):
        torch.distributed.all_reduce(mse, op=torch.distributed.ReduceOp.AVG, group=group)
    return (20.0 * torch.log10(1.0 / torch.sqrt(mse))).reshape(())


def torchSSIM(tar_img, prd_img):
    """Drop-in for src/ModelTrainer.py:23-24: pytorch_msssim.ssim(tar, prd, data_range=1.0, size_average=True) on the
    separable-Gaussian SSIM kernel (csrc/ssim.cu); returns a 0-dim device tensor."""
    from .ssim import ssim
    if not (tar_img.is_cuda and prd_img.is_cuda):
        raise RuntimeError("uwr.metrics runs on CUDA tensors only; there is no CPU fallback")
    return ssim(tar_img.float(), prd_img.float(), data_range=1.0, size_average=True)
