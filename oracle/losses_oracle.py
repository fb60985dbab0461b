"""CPU restatement of the reference losses and metrics (TEST INFRASTRUCTURE, see oracle/__init__.py).

Follows src/Losses/losses.py:54-81,182-213, src/Losses/luminanceLoss.py:5-21,
src/ModelTrainer.py:17-21 and the published algorithm of focal_frequency_loss 0.3.0
(absent third-party dependency, restated per SURVEY.md Appendix C).
"""
import torch


def l1(pred, truth, batch_divisor=None):
    """losses.py:55-57."""
    b = batch_divisor or truth.shape[0]
    return (pred - truth).abs().mean() / (b * truth.shape[1])


def l2(pred, truth, batch_divisor=None):
    """losses.py:76-78."""
    b = batch_divisor or truth.shape[0]
    return ((pred - truth) ** 2).mean() / (b * truth.shape[1])


def charbonnier(pred, truth, eps=1e-3):
    """losses.py:182-193."""
    d = pred - truth
    return torch.sqrt(d * d + eps * eps).mean()


def color(pred, truth):
    """ColorLoss, losses.py:195-213."""
    return ((pred - truth) ** 2).mean(dim=(2, 3)).mean()


def luminance(pred, truth):
    """LuminanceLoss, luminanceLoss.py:5-21."""
    y = torch.tensor([0.299, 0.587, 0.114], dtype=pred.dtype, device=pred.device).view(1, 3, 1, 1)
    return (((pred - truth) * y).sum(1, keepdim=True) ** 2).mean()


def l1_with_color(pred, truth, batch_divisor=None):
    """losses.py:58-66 with the documented patch P2 (LuminanceLoss attached)."""
    b = batch_divisor or truth.shape[0]
    loss = 0.5 * color(pred, truth) + 0.25 * (pred - truth).abs().mean() + 0.25 * luminance(pred, truth)
    return loss / (b * truth.shape[1])


def focal_frequency(pred, truth):
    """FocalFrequencyLoss(loss_weight=1, alpha=1) as constructed at losses.py:48."""
    fp = torch.fft.fft2(pred, norm="ortho")
    ft = torch.fft.fft2(truth, norm="ortho")
    d2 = (fp.real - ft.real) ** 2 + (fp.imag - ft.imag) ** 2
    w = torch.sqrt(d2)
    w = w / w.amax(dim=(-2, -1), keepdim=True)
    w = torch.nan_to_num(w, nan=0.0).clamp(0.0, 1.0).detach()
    return (w * d2).mean()


def torch_psnr(tar, prd):
    """torchPSNR, ModelTrainer.py:17-21."""
    d = prd.clamp(0, 1) - tar.clamp(0, 1)
    return 20 * torch.log10(1 / (d ** 2).mean().sqrt())
