"""CPU checks of the identities the FFT kernels are built on (oracle/fft_oracle.py) against numpy / torch FFTs,
and of the batched MDTA attention algebra (uwr.fn._mdta_matrices) against the reference formulation."""
import numpy as np
import torch
import torch.nn.functional as F

from oracle import fft_oracle as fo


def _rng(*s, seed=0):
    return np.random.default_rng(seed).standard_normal(s)


def test_real_part_dft_is_its_own_inverse_and_adjoint():
    x, y = _rng(2, 8, 16, 3, seed=1), _rng(2, 8, 16, 3, seed=2)
    f = fo.dft_hw_real(x)
    assert np.allclose(f, np.fft.fftn(x, axes=(1, 2)).real)
    assert np.allclose(fo.dft_hw_real(x, 1.0 / (8 * 16)), np.fft.ifftn(x, axes=(1, 2)).real)
    assert np.isclose((fo.dft_hw_real(x) * y).sum(), (x * fo.dft_hw_real(y)).sum())   # symmetric => own backward


def test_four_step_token_axis_fft_matches_fftn():
    for (B, H, W, C) in ((2, 8, 8, 4), (1, 16, 8, 8), (1, 4, 32, 2)):
        x = _rng(B, H, W, C, seed=3)
        ref = np.fft.fftn(x.reshape(B, H * W, C), axes=(1, 2)).real.reshape(B, H, W, C)
        assert np.allclose(fo.dft_lc_real_four_step(x), ref, atol=1e-9)
        refi = np.fft.ifftn(x.reshape(B, H * W, C), axes=(1, 2)).real.reshape(B, H, W, C)
        assert np.allclose(fo.dft_lc_real_four_step(x, 1.0 / (H * W * C)), refi, atol=1e-12)


def test_tiled_spectrum_inverse_is_scatter_of_small_inverse():
    spec = _rng(8, 12, seed=4) + 1j * _rng(8, 12, seed=5)
    ref = np.fft.ifft2(np.tile(spec, (2, 2)))
    assert np.allclose(fo.upsample_tiled_ifft2(spec), ref, atol=1e-12)


def test_batched_mdta_matrices_match_reference_formulation():
    """uwr.fn._mdta_matrices (per-head Gram blocks + squared norms -> normalise -> softmax) vs
    SpectralTransformer.py:97-101 written out with F.normalize / softmax on (b, heads, c, L) tensors."""
    from uwr.fn import _mdta_matrices
    B, L, C, heads = 3, 40, 8, 2
    g = torch.Generator().manual_seed(0)
    qk = torch.randn(B, L, 2 * C, generator=g, dtype=torch.float64)
    temp = torch.tensor([0.7, 1.9], dtype=torch.float64).view(1, heads, 1, 1)
    c = C // heads
    q = qk[:, :, :C].transpose(1, 2).reshape(B, heads, c, L)
    k = qk[:, :, C:].transpose(1, 2).reshape(B, heads, c, L)
    G = q @ k.transpose(-2, -1)                                      # what uwr_mdta_gram returns: (B, heads, c, c)
    sq_q, sq_k = (q * q).sum(-1).reshape(B, C), (k * k).sum(-1).reshape(B, C)
    A = _mdta_matrices(G, sq_q, sq_k, temp, heads)
    ref = torch.softmax(F.normalize(q, dim=-1) @ F.normalize(k, dim=-1).transpose(-2, -1) * temp, dim=-1)
    assert torch.allclose(A, ref, atol=1e-12)
