"""Shared-memory FFT passes (csrc/fft.cu) against torch.fft in fp64 on the same inputs: the
real-part 2-D DFTs of FDFP (block.py:532-556) and EncoderBlock (model.py:72-88), their inverses,
their gradients (the map is symmetric), and the complex FFT2 used by SpectralTransformer.UpSample."""
import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu
TOL = 2e-5  # pure fp32 butterflies; length-65536 transforms stay below 1e-5


def _r(*shape, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return torch.randn(*shape, generator=g).cuda()


@pytest.mark.parametrize("B,H,W,C", [(2, 16, 16, 32), (1, 64, 64, 48), (1, 32, 64, 8), (1, 256, 256, 32), (2, 128, 128, 20)])
def test_dft_hw_real(B, H, W, C):
    from uwr import ops
    x = _r(B, H, W, C, seed=1)
    y = ops.dft_real(x, B, H, W, C, 1.0, "hw")
    ref = torch.fft.fftn(x.double(), dim=(1, 2)).real
    assert rel_l2(y, ref) < TOL
    # Re(ifftn) of a real tensor = the same transform scaled by 1/(H*W)
    yi = ops.dft_real(x, B, H, W, C, 1.0 / (H * W), "hw")
    assert rel_l2(yi, torch.fft.ifftn(x.double(), dim=(1, 2)).real) < TOL


@pytest.mark.parametrize("B,H,W,C", [(2, 16, 16, 32), (1, 32, 32, 256), (1, 64, 64, 128), (1, 256, 256, 32), (1, 128, 128, 64)])
def test_dft_lc_real(B, H, W, C):
    from uwr import ops
    x = _r(B, H, W, C, seed=2)
    y = ops.dft_real(x, B, H, W, C, 1.0, "lc")
    ref = torch.fft.fftn(x.double().view(B, H * W, C), dim=(-2, -1)).real.view(B, H, W, C)
    assert rel_l2(y, ref) < TOL
    yi = ops.dft_real(x, B, H, W, C, 1.0 / (H * W * C), "lc")
    refi = torch.fft.ifftn(x.double().view(B, H * W, C), dim=(-2, -1)).real.view(B, H, W, C)
    assert rel_l2(yi, refi) < TOL


@pytest.mark.parametrize("axes", ["hw", "lc"])
def test_dft_real_autograd(axes):
    from uwr import fn
    B, H, W, C = 2, 32, 32, 64
    x = _r(B, H, W, C, seed=3).requires_grad_()
    g = _r(B, H, W, C, seed=4)
    y = fn.dft_real(x, B, H, W, C, 0.5, axes)
    y.backward(g)
    xd = x.detach().double().requires_grad_()
    if axes == "hw":
        ref = 0.5 * torch.fft.fftn(xd, dim=(1, 2)).real
    else:
        ref = 0.5 * torch.fft.fftn(xd.view(B, H * W, C), dim=(-2, -1)).real.view(B, H, W, C)
    ref.backward(g.double())
    assert rel_l2(y, ref) < TOL
    assert rel_l2(x.grad, xd.grad) < TOL


@pytest.mark.parametrize("B,H,W,C", [(2, 32, 32, 128), (1, 64, 64, 64), (1, 128, 128, 32), (2, 12, 24, 16), (1, 8, 12, 128)])
def test_fft2_hw_complex_roundtrip(B, H, W, C):
    from uwr import ops
    x = _r(B, H, W, C, seed=5)
    f = ops.fft2_hw(x, B, H, W, C, in_complex=False, inverse=False)
    ref = torch.fft.fft2(x.double(), dim=(1, 2))
    assert rel_l2(f, torch.view_as_real(ref)) < TOL
    back = ops.fft2_hw(f, B, H, W, C, in_complex=True, inverse=True, scale=1.0 / (H * W))
    assert rel_l2(back, torch.view_as_real(x.double().to(torch.complex128))) < TOL
