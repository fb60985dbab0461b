"""Restatement of pytorch_msssim (VainF/pytorch-msssim 1.0.0; unpinned in requirements.txt).
Separable 11-tap Gaussian, valid convolutions, 5-scale MS-SSIM."""
import torch
import torch.nn as nn
import torch.nn.functional as F


def _gauss_1d(size, sigma):
    c = torch.arange(size, dtype=torch.float32) - size // 2
    g = torch.exp(-(c ** 2) / (2 * sigma ** 2))
    return (g / g.sum()).view(1, 1, -1)


def _blur(x, win):
    ch = x.shape[1]
    out = x
    for i, s in enumerate(x.shape[2:]):
        if s >= win.shape[-1]:
            w = win.transpose(2 + i, -1) if i == 0 else win
            out = F.conv2d(out, w, stride=1, padding=0, groups=ch)
    return out


def _ssim(x, y, data_range, win, K=(0.01, 0.03)):
    c1 = (K[0] * data_range) ** 2
    c2 = (K[1] * data_range) ** 2
    win = win.to(x.device, dtype=x.dtype)
    mu1, mu2 = _blur(x, win), _blur(y, win)
    mu1_sq, mu2_sq, mu12 = mu1 * mu1, mu2 * mu2, mu1 * mu2
    s1 = _blur(x * x, win) - mu1_sq
    s2 = _blur(y * y, win) - mu2_sq
    s12 = _blur(x * y, win) - mu12
    cs_map = (2 * s12 + c2) / (s1 + s2 + c2)
    ssim_map = ((2 * mu12 + c1) / (mu1_sq + mu2_sq + c1)) * cs_map
    return torch.flatten(ssim_map, 2).mean(-1), torch.flatten(cs_map, 2).mean(-1)


def _win4d(win_size, win_sigma, channels):
    w = _gauss_1d(win_size, win_sigma)           # (1,1,K)
    return w.repeat(channels, 1, 1).unsqueeze(2)  # (C,1,1,K)


def ssim(X, Y, data_range=255, size_average=True, win_size=11, win_sigma=1.5, win=None,
         K=(0.01, 0.03), nonnegative_ssim=False):
    if win is None:
        win = _win4d(win_size, win_sigma, X.shape[1])
    s, _ = _ssim(X, Y, data_range, win, K)
    if nonnegative_ssim:
        s = torch.relu(s)
    return s.mean() if size_average else s.mean(1)


def ms_ssim(X, Y, data_range=255, size_average=True, win_size=11, win_sigma=1.5, win=None,
            weights=None, K=(0.01, 0.03)):
    if win is None:
        win = _win4d(win_size, win_sigma, X.shape[1])
    smaller = min(X.shape[-2:])
    assert smaller > (win_size - 1) * (2 ** 4)
    if weights is None:
        weights = [0.0448, 0.2856, 0.3001, 0.2363, 0.1333]
    wts = X.new_tensor(weights)
    mcs = []
    levels = wts.shape[0]
    for i in range(levels):
        s, cs = _ssim(X, Y, data_range, win, K)
        if i < levels - 1:
            mcs.append(torch.relu(cs))
            pad = [d % 2 for d in X.shape[2:]]
            X = F.avg_pool2d(X, kernel_size=2, padding=pad)
            Y = F.avg_pool2d(Y, kernel_size=2, padding=pad)
    s = torch.relu(s)
    stack = torch.stack(mcs + [s], 0)
    val = torch.prod(stack ** wts.view(-1, 1, 1), 0)
    return val.mean() if size_average else val.mean(1)


class MS_SSIM(nn.Module):
    def __init__(self, data_range=255, size_average=True, win_size=11, win_sigma=1.5, channel=3,
                 spatial_dims=2, weights=None, K=(0.01, 0.03)):
        super().__init__()
        self.win_size = win_size
        self.register_buffer("win", _win4d(win_size, win_sigma, channel), persistent=False)
        self.size_average = size_average
        self.data_range = data_range
        self.weights = weights
        self.K = K

    def forward(self, X, Y):
        return ms_ssim(X, Y, data_range=self.data_range, size_average=self.size_average,
                       win=self.win, weights=self.weights, K=self.K)
