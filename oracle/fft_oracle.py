"""CPU restatement (numpy, fp64) of the frequency-domain transforms of the reference and of the pass
structure the CUDA kernels use for them.  TEST INFRASTRUCTURE ONLY (tests/ import it; the product never does).

Reference semantics
  * FDFP (src/model/block.py:532-556):           f = fftn(x, dim=(H, W)).real ... ifftn(f, dim=(H, W)).real
  * EncoderBlock (src/model/model.py:72-88):     f = fftn(a, dim=(-2, -1)).real on (B, L, C) tokens; ifftn(...).real
  * torch / numpy conventions: unnormalised forward, 1/N inverse.

Identities the kernels rely on (each is asserted in tests/test_fft_oracle_cpu.py):
  1. For a REAL input, Re(ifftn(x)) = Re(fftn(x)) / N, and x -> Re(fftn(x)) is a symmetric linear map, so one
     kernel (with a scale) is the forward, the inverse and both backward passes.
  2. Four-step FFT of the token axis (csrc/fft.cu, uwr_dft_lc_real): with L = H*W, l = y*W + x and
     k = k1 + H*k2,   X[k] = sum_x  W_L^(x*k1) * [ sum_y x[y, x] W_H^(y*k1) ] * W_W^(x*k2),
     i.e. length-H FFTs down the columns, a twiddle, length-W FFTs along the rows, transposed store.
  3. SpectralTransformer.UpSample (src/Models/SpectralTransformer.py:174-188): ifft2 of the (2, 2)-tiled spectrum
     at 2H x 2W equals ifft2 of the spectrum at H x W written to the even pixels, zeros elsewhere.
"""
import numpy as np


def dft_hw_real(x, scale=1.0):
    """x: (B, H, W, C) real -> scale * Re(FFT2 over (H, W)); pass structure of uwr_dft_hw_real."""
    t = np.fft.fft(x.astype(np.float64), axis=2)      # pass 1: along W, real -> complex
    t = np.fft.fft(t, axis=1)                         # pass 2: along H, complex -> real part
    return scale * t.real


def dft_lc_real_four_step(x, scale=1.0):
    """x: (B, H, W, C) real tokens (l = y*W + x) -> scale * Re(FFT2 over (L = H*W, C)), computed exactly as
    uwr_dft_lc_real does: channel-axis FFT, length-H FFTs over y with the twiddle exp(-2 pi i x k1 / L),
    length-W FFTs over x, result for frequency k = k1 + H*k2 stored at token index k."""
    B, H, W, C = x.shape
    L = H * W
    t = np.fft.fft(x.astype(np.float64), axis=3)      # rows pass (channels)
    t = np.fft.fft(t, axis=1)                         # step 1: over y -> index k1 (still laid out [k1][x])
    k1 = np.arange(H).reshape(1, H, 1, 1)
    xx = np.arange(W).reshape(1, 1, W, 1)
    t = t * np.exp(-2j * np.pi * (xx * k1 % L) / L)   # twiddle W_L^(x*k1)
    t = np.fft.fft(t, axis=2)                         # step 2: over x -> index k2 (layout [k1][k2])
    out = np.transpose(t, (0, 2, 1, 3)).reshape(B, L, C)   # token k = k2*H + k1: transposed store
    return (scale * out.real).reshape(B, H, W, C)


def upsample_tiled_ifft2(spec):
    """spec: (H, W) complex.  ifft2(tile(spec, (2, 2))) as the H x W inverse scattered to the even pixels."""
    H, W = spec.shape
    out = np.zeros((2 * H, 2 * W), dtype=np.complex128)
    out[::2, ::2] = np.fft.ifft2(spec)
    return out
