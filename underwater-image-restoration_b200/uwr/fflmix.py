"""The "fflMix" loss of the reference (src/Losses/losses.py:108-117):

    0.03*Charbonnier + 0.025*VGGPerceptual + 0.01*Gradient + 0.005*FFL + 0.1*(1 - MS_SSIM)

returned as the 6-tuple (loss, charb, perc, grad, ffl, ssim) that ModelTrainer.py:82-85 unpacks.
Charbonnier and the focal frequency term run on the uwr kernels; the VGG16 perceptual term, the
Laplacian gradient term and MS-SSIM stay on PyTorch/cuDNN ops in this round (SURVEY.md §8a row 35,
§8f rank 3).  Patch P3: torchvision's ImageNet VGG16 weights cannot be downloaded
offline, so the perceptual network is a seeded random-init VGG16 (`VGG_SEED`); with real weights
present, load them into `.vgg` before training.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

VGG_SEED = 777


class VGGPerceptual(nn.Module):
    """VGGPerceptualLoss (losses.py:215-255): features[:4], [4:9], [9:16], [16:23] of VGG16, inputs
    normalised with the ImageNet mean/std and resized to 224x224, L1 between the four feature taps."""

    def __init__(self):
        super().__init__()
        import torchvision
        state = torch.random.get_rng_state()
        torch.manual_seed(VGG_SEED)
        feats = torchvision.models.vgg16(weights=None).features
        torch.random.set_rng_state(state)
        self.blocks = nn.ModuleList([feats[:4].eval(), feats[4:9].eval(), feats[9:16].eval(), feats[16:23].eval()])
        for p in self.parameters():
            p.requires_grad = False
        self.register_buffer("mean", torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1))
        self.register_buffer("std", torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1))

    def forward(self, inp, target):
        x = F.interpolate((inp - self.mean) / self.std, mode="bilinear", size=(224, 224), align_corners=False)
        y = F.interpolate((target - self.mean) / self.std, mode="bilinear", size=(224, 224), align_corners=False)
        loss = 0.0
        for blk in self.blocks:
            x, y = blk(x), blk(y)
            loss = loss + F.l1_loss(x, y)
        return loss


def gradient_loss(x, y):
    """Gradient_Loss (losses.py:162-181): 3x3 Laplacian per channel, valid convolution, L1."""
    k = x.new_tensor([[0.0, 1.0, 0.0], [1.0, -4.0, 1.0], [0.0, 1.0, 0.0]]).view(1, 1, 3, 3).repeat(3, 1, 1, 1)
    return F.l1_loss(F.conv2d(x, k, groups=3), F.conv2d(y, k, groups=3))


def _gauss(size, sigma, like):
    c = torch.arange(size, dtype=like.dtype, device=like.device) - size // 2
    g = torch.exp(-(c ** 2) / (2 * sigma ** 2))
    return g / g.sum()


def ms_ssim(X, Y, data_range=1.0, win_size=11, sigma=1.5):
    """pytorch_msssim.MS_SSIM(win 11, sigma 1.5, 5 scales, size_average) as constructed at losses.py:46."""
    ch = X.shape[1]
    g = _gauss(win_size, sigma, X)
    wh, ww = g.view(1, 1, -1, 1).repeat(ch, 1, 1, 1), g.view(1, 1, 1, -1).repeat(ch, 1, 1, 1)
    blur = lambda t: F.conv2d(F.conv2d(t, wh, groups=ch), ww, groups=ch)
    c1, c2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    weights = X.new_tensor([0.0448, 0.2856, 0.3001, 0.2363, 0.1333])
    vals = []
    for i in range(5):
        mu1, mu2 = blur(X), blur(Y)
        s1, s2, s12 = blur(X * X) - mu1 * mu1, blur(Y * Y) - mu2 * mu2, blur(X * Y) - mu1 * mu2
        cs_map = (2 * s12 + c2) / (s1 + s2 + c2)
        ssim_map = ((2 * mu1 * mu2 + c1) / (mu1 * mu1 + mu2 * mu2 + c1)) * cs_map
        cs, ss = cs_map.flatten(2).mean(-1), ssim_map.flatten(2).mean(-1)
        if i < 4:
            vals.append(torch.relu(cs))
            pad = [d % 2 for d in X.shape[2:]]
            X, Y = F.avg_pool2d(X, 2, padding=pad), F.avg_pool2d(Y, 2, padding=pad)
        else:
            vals.append(torch.relu(ss))
    stack = torch.stack(vals, 0)
    return torch.prod(stack ** weights.view(-1, 1, 1), 0).mean()


_VGG = {}


def fflmix_loss(lossfn, pred, truth):
    from .ffl import FocalFrequencyFn
    from .losses import PixelLossFn
    dev = pred.device
    if dev not in _VGG:
        _VGG[dev] = VGGPerceptual().to(dev)
    charb = PixelLossFn.apply(pred, truth, "charbonnier", None)
    perc = _VGG[dev](pred, truth)
    grad = gradient_loss(pred, truth)
    ffl = FocalFrequencyFn.apply(pred, truth)
    ssim = 1 - ms_ssim(pred, truth)
    loss = 0.03 * charb + 0.025 * perc + 0.01 * grad + 0.005 * ffl + 0.1 * ssim
    return loss, charb, perc, grad, ffl, ssim
