"""Model-level parity with the "fflMix" loss (BASELINE config 3's loss) driving the backward: NewBigFRFNModel on the uwr
kernels + uwr LossFunction("fflMix") vs the CPU oracle + the reference's loss formula (losses.py:108-117) restated
with the stand-ins of oracle/shims (focal_frequency_loss, pytorch_msssim) and the P3 seeded VGG16.
FFL / MS-SSIM are restatement-pinned, the VGG term is unpinned against real pretrained weights."""
import os
import sys

import pytest
import torch
import torch.nn.functional as F

from conftest import ROOT, rel_l2

pytestmark = pytest.mark.gpu


def _reference_fflmix(pred, truth):
    """losses.py:108-117 on CPU tensors."""
    sys.path.insert(0, os.path.join(ROOT, "oracle", "shims"))
    try:
        from focal_frequency_loss import FocalFrequencyLoss
        from pytorch_msssim import MS_SSIM
    finally:
        sys.path.pop(0)
    import torchvision
    from oracle import losses_oracle
    st = torch.random.get_rng_state()
    torch.manual_seed(777)                                         # patch P3 (uwr.fflmix.VGG_SEED)
    feats = torchvision.models.vgg16(weights=None).features.eval()
    torch.random.set_rng_state(st)
    for p in feats.parameters():
        p.requires_grad = False
    mean = torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1)
    x = F.interpolate((pred - mean) / std, mode="bilinear", size=(224, 224), align_corners=False)   # losses.py:236-248
    y = F.interpolate((truth - mean) / std, mode="bilinear", size=(224, 224), align_corners=False)
    perc = 0.0
    for blk in (feats[:4], feats[4:9], feats[9:16], feats[16:23]):
        x, y = blk(x), blk(y)
        perc = perc + F.l1_loss(x, y)
    k = torch.tensor([[0.0, 1.0, 0.0], [1.0, -4.0, 1.0], [0.0, 1.0, 0.0]]).view(1, 1, 3, 3).repeat(3, 1, 1, 1)
    grad = F.l1_loss(F.conv2d(pred, k, groups=3), F.conv2d(truth, k, groups=3))                      # losses.py:162-181
    charb = losses_oracle.charbonnier(pred, truth)
    ffl = FocalFrequencyLoss(loss_weight=1.0, alpha=1.0)(pred, truth)
    ssim = 1 - MS_SSIM(data_range=1.0, size_average=True, channel=3)(pred, truth)
    return 0.03 * charb + 0.025 * perc + 0.01 * grad + 0.005 * ffl + 0.1 * ssim


@pytest.mark.parametrize("precision", ["tf32", "tf32x3"])
def test_newbigfrfn_backward_driven_by_fflmix(precision):
    from uwr import ops
    ops.set_gemm_precision(precision)
    try:
        _run(precision)
    finally:
        ops.set_gemm_precision("tf32")


def _run(precision):
    from oracle import newbig_oracle
    from uwr.losses import LossFunction
    from uwr.newbig import MyBigFRFNModel
    B, S = 1, 256                       # MS-SSIM needs sides > 160
    torch.manual_seed(1234)
    model = MyBigFRFNModel()
    sd_cpu = {k: v.detach().clone() for k, v in model.state_dict().items()}
    model = model.cuda().eval()
    g = torch.Generator().manual_seed(2024)
    raw = torch.rand(B, 3, S, S, generator=g)
    ref = (raw * torch.tensor([0.9, 1.05, 1.1]).view(1, 3, 1, 1) + 0.03 * torch.randn(B, 3, S, S, generator=g)).clamp(0, 1)
    out = model(raw.cuda())
    tup = LossFunction("fflMix", "cuda", vgg_weights="random").getloss(out, ref.cuda())
    assert len(tup) == 6
    tup[0].backward()
    sd_o = {k: (v.clone().requires_grad_() if v.is_floating_point() and v.dim() > 0 and "dwt" not in k else v)
            for k, v in sd_cpu.items()}
    out_o = newbig_oracle.newbig_frfn_forward(sd_o, raw)
    loss_o = _reference_fflmix(out_o, ref)
    loss_o.backward()
    named = dict(model.named_parameters())
    live = [n for n, v in sd_o.items() if getattr(v, "grad", None) is not None]
    gnorm = torch.sqrt(sum((sd_o[n].grad.double() ** 2).sum() for n in live)).item()
    tot = sum(((named[n].grad.double().cpu() - sd_o[n].grad.double()) ** 2).sum().item() for n in live)
    e_loss = abs(tup[0].item() - loss_o.item()) / abs(loss_o.item())
    print(f"fflMix-driven NewBigFRFN 256 [{precision}]: loss {tup[0].item():.6f} vs {loss_o.item():.6f} ({e_loss:.1e}), out "
          f"{rel_l2(out, out_o):.2e}, grads {tot ** 0.5 / gnorm:.2e}")
    # The loss cotangent is computed from each side's OWN prediction here (no same-cotangent protocol), and VGG's ReLUs /
    # L1 feature taps are non-smooth: a TF32-level difference of the prediction flips signs and is amplified.  With
    # error-compensated products the two sides agree to the stated 1e-3; single-pass TF32 is bounded looser.
    exact = precision == "tf32x3"
    assert e_loss < 1e-3 and rel_l2(out, out_o) < (1e-3 if exact else 2e-3)
    assert tot ** 0.5 / gnorm < (1e-3 if exact else 2e-2)
