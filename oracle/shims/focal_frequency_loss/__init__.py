"""Restatement of focal_frequency_loss.FocalFrequencyLoss (EndlessSora/focal-frequency-loss 0.3.0,
unpinned in the reference's requirements.txt).  Only the configuration the reference constructs
(losses.py:48: loss_weight=1.0, alpha=1.0, patch_factor=1, no ave_spectrum/log/batch matrix)
is exercised; the other switches are restated from the published algorithm."""
import torch
import torch.nn as nn


class FocalFrequencyLoss(nn.Module):
    def __init__(self, loss_weight=1.0, alpha=1.0, patch_factor=1, ave_spectrum=False,
                 log_matrix=False, batch_matrix=False):
        super().__init__()
        self.loss_weight = loss_weight
        self.alpha = alpha
        self.patch_factor = patch_factor
        self.ave_spectrum = ave_spectrum
        self.log_matrix = log_matrix
        self.batch_matrix = batch_matrix

    def tensor2freq(self, x):
        pf = self.patch_factor
        _, _, h, w = x.shape
        assert h % pf == 0 and w % pf == 0
        ph, pw = h // pf, w // pf
        patches = [x[:, :, i * ph:(i + 1) * ph, j * pw:(j + 1) * pw]
                   for i in range(pf) for j in range(pf)]
        y = torch.stack(patches, 1)
        freq = torch.fft.fft2(y, norm="ortho")
        return torch.stack([freq.real, freq.imag], -1)

    def loss_formulation(self, recon_freq, real_freq, matrix=None):
        if matrix is not None:
            weight = matrix.detach()
        else:
            d = (recon_freq - real_freq) ** 2
            m = torch.sqrt(d[..., 0] + d[..., 1]) ** self.alpha
            if self.log_matrix:
                m = torch.log(m + 1.0)
            if self.batch_matrix:
                m = m / m.max()
            else:
                m = m / m.max(-1).values.max(-1).values[:, :, :, None, None]
            m[torch.isnan(m)] = 0.0
            m = torch.clamp(m, min=0.0, max=1.0)
            weight = m.clone().detach()
        d = (recon_freq - real_freq) ** 2
        dist = d[..., 0] + d[..., 1]
        return torch.mean(weight * dist)

    def forward(self, pred, target, matrix=None, **kwargs):
        pf = self.tensor2freq(pred)
        tf = self.tensor2freq(target)
        if self.ave_spectrum:
            pf = torch.mean(pf, 0, keepdim=True)
            tf = torch.mean(tf, 0, keepdim=True)
        return self.loss_formulation(pf, tf, matrix) * self.loss_weight
