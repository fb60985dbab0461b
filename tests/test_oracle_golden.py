"""CPU: pin the oracle (oracle/*.py) against golden vectors produced by the UNMODIFIED reference
(oracle/make_golden.py), and against the live reference when /root/reference is mounted."""
import hashlib
import os
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT, rel_l2

GOLD = os.path.join(ROOT, "tests", "golden")


def _sha(t):
    return hashlib.sha1(t.detach().cpu().contiguous().numpy().tobytes()).hexdigest()


def _load(name):
    return torch.load(os.path.join(GOLD, name), weights_only=False)


def test_ast_oracle_matches_reference_golden():
    """seeded weights (our module's init is RNG-identical to the reference) -> same output/loss/grads"""
    from oracle import ast_oracle, losses_oracle
    from uwr.ast import AST
    g = _load("ast_128.pt")
    torch.manual_seed(g["seed_weights"])
    model = AST(img_size=128)
    got = [(k, list(v.shape), str(v.dtype), _sha(v)) for k, v in model.state_dict().items()]
    assert got == g["state_dict_sha1"], "state_dict keys/shapes/values differ from the reference (AST.py:680-872)"
    gen = torch.Generator().manual_seed(g["seed_data"])
    raw = torch.rand(1, 3, 128, 128, generator=gen) * 2 - 1
    ref = torch.rand(1, 3, 128, 128, generator=gen) * 2 - 1
    sd = {k: (v.detach().clone().requires_grad_() if v.is_floating_point() else v) for k, v in model.state_dict().items()}
    out = ast_oracle.ast_forward(sd, raw, img_size=128)
    assert rel_l2(out, g["out"]) < 1e-6
    loss = losses_oracle.l1(out, ref)
    assert abs(loss.item() - g["loss_l1"]) < 1e-7 * abs(g["loss_l1"]) + 1e-10
    loss.backward()
    pg = torch.Generator().manual_seed(g["proj_seed"])
    for n, p in model.named_parameters():
        r = torch.randn(p.shape, generator=pg)
        norm, proj = g["grad_norm_proj"][n]
        gr = sd[n].grad
        assert abs(gr.norm().item() - norm) <= 2e-4 * norm + 1e-12, n
        assert abs((gr * r).sum().item() - proj) <= 2e-3 * norm * (p.numel() ** 0.5) * 1e-2 + 1e-10, n


def test_ast_256_default_state_dict_matches_reference():
    from uwr.ast import AST
    g = _load("ast_256_state_sha1.pt")
    torch.manual_seed(1234)
    model = AST()
    got = [(k, list(v.shape), str(v.dtype), _sha(v)) for k, v in model.state_dict().items()]
    assert len(got) == 274
    assert got == g["state_dict_sha1"]
    assert sum(p.numel() for p in model.parameters()) == 19919507


def test_block_oracle_matches_reference_golden():
    from oracle import ast_oracle
    g = _load("block_shift.pt")
    y = ast_oracle.transformer_block(g["state"], "", g["x"], 2, 4, True, "leff")
    assert rel_l2(y, g["y"]) < 1e-6
    from uwr.ast import _relative_position_index
    assert torch.equal(_relative_position_index(8), g["rel_index"])


def test_losses_and_metrics_oracle_match_reference_golden():
    from oracle import losses_oracle as lo, uiqm_oracle
    g = _load("losses_metrics.pt")
    torch.manual_seed(0)
    p = torch.rand(2, 3, 256, 256)
    t = torch.rand(2, 3, 256, 256)
    close = lambda a, b: abs(a - b) <= 2e-6 * abs(b) + 1e-9
    assert close(lo.l1(p, t).item(), g["L1"])
    assert close(lo.l2(p, t).item(), g["L2"])
    assert close(lo.charbonnier(p, t).item(), g["charbonnier"])
    assert close(lo.l1_with_color(p, t).item(), g["L1withColor"])
    assert close(lo.luminance(p, t).item(), g["Luminance"])
    assert close((lo.focal_frequency(p, t) + lo.charbonnier(p, t)).item(), g["fflCharbonnier"])
    assert close(lo.charbonnier(p, p).item(), g["charbonnier_identical"])   # src/Loss.ipynb:45 -> 0.0010
    assert abs(g["charbonnier_identical"] - 1e-3) < 1e-8
    assert lo.focal_frequency(p, p).item() == 0.0                           # src/Loss.ipynb:49
    assert close(lo.torch_psnr(t, p).item(), g["psnr"])
    # SURVEY.md §8a row 38 golden: UIQM of default_rng(0) noise = 2.964016389614244
    img = (np.random.default_rng(0).random((256, 256, 3)) * 255).astype(np.uint8)
    got = uiqm_oracle.get_uiqm(img)
    for a, b in zip(got, g["uiqm"]):
        assert abs(float(a) - b) <= 1e-5 * abs(b)
    assert abs(g["uiqm"][0] - 2.964016389614244) < 1e-9


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="reference not mounted")
def test_oracle_vs_live_reference_train_mode_masks():
    """live check incl. DropPath: the reference's own timm-semantics masks are captured and replayed"""
    sys.path[:0] = [os.path.join(ROOT, "oracle", "shims"), "/root/reference"]
    from src.Models.AST import AST as RefAST
    from oracle import ast_oracle
    import timm.layers as tl
    torch.manual_seed(1234)
    ref = RefAST(img_size=128)
    ref.train()
    masks = {}
    orig = tl.DropPath.forward

    def spy(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1.0 - self.drop_prob
        m = x.new_empty((x.shape[0], 1, 1)).bernoulli_(keep) / keep
        masks.setdefault(id(self), []).append(m.view(-1))
        return x * m
    tl.DropPath.forward = spy
    try:
        gen = torch.Generator().manual_seed(3)
        x = torch.rand(2, 3, 128, 128, generator=gen) * 2 - 1
        torch.manual_seed(11)
        y = ref(x)
    finally:
        tl.DropPath.forward = orig
    sd = ref.state_dict()
    dsc = {}
    for name, mod in ref.named_modules():
        if id(mod) in masks:
            pre = name[: -len("drop_path")]
            ms = masks[id(mod)]
            dsc[pre] = (ms[0], ms[1]) if len(ms) == 2 else (None, ms[0])
    yo = ast_oracle.ast_forward(sd, x, img_size=128, drop_scales=dsc)
    assert rel_l2(yo, y) < 1e-6


def test_newbigfrfn_state_dict_and_oracle_match_reference_golden():
    """632-entry state_dict (incl. Haar buffers) is RNG-identical to the reference; the oracle
    (with patch P1) reproduces the patched reference output."""
    from oracle import newbig_oracle
    from uwr.newbig import MyBigFRFNModel
    g = _load("newbigfrfn_128.pt")
    torch.manual_seed(g["seed_weights"])
    model = MyBigFRFNModel()
    got = [(k, list(v.shape), str(v.dtype), _sha(v)) for k, v in model.state_dict().items()]
    assert len(got) == 632
    assert got == g["state_dict_sha1"]
    assert sum(p.numel() for p in model.parameters()) == 35949007
    gen = torch.Generator().manual_seed(g["seed_data"])
    x = torch.rand(1, 3, 128, 128, generator=gen) * 2 - 1
    with torch.no_grad():
        y = newbig_oracle.newbig_frfn_forward(model.state_dict(), x)
    assert rel_l2(y, g["out"]) < 1e-6


def test_broken_reference_models_construct_and_fail_like_the_reference():
    """NewModel / NewBigModel (model.py:162-463): same seeded state_dict as the reference module, and forward raises
    the exception the reference's forward raises (golden written by oracle/make_golden.py from the reference)."""
    import uwr
    g = _load("newmodels_state_sha1.pt")
    for name, count in (("NewModel", 339), ("NewBigModel", 607)):
        torch.manual_seed(1234)
        model = uwr.init_model(name)
        got = [(k, list(v.shape), str(v.dtype), _sha(v)) for k, v in model.state_dict().items()]
        assert len(got) == count and got == g[name]["state_dict_sha1"]
        kind, msg = g[name]["error"]
        with pytest.raises(Exception) as ei:
            model(torch.zeros(1, 3, 128, 128))
        assert type(ei.value).__name__ == kind and str(ei.value) == msg


def test_spectral_state_dict_and_oracle_match_reference_golden():
    from oracle import spectral_oracle
    from uwr.spectral import SpectralTransformer
    g = _load("spectral_64x96.pt")
    torch.manual_seed(g["seed_weights"])
    model = SpectralTransformer()
    got = [(k, list(v.shape), str(v.dtype), _sha(v)) for k, v in model.state_dict().items()]
    assert len(got) == 443 and got == g["state_dict_sha1"]
    assert sum(p.numel() for p in model.parameters()) == 2430709
    gen = torch.Generator().manual_seed(g["seed_data"])
    x = torch.rand(1, 3, 64, 96, generator=gen) * 2 - 1
    with torch.no_grad():
        y = spectral_oracle.spectral_forward(model.state_dict(), x)
    assert rel_l2(y, g["out"]) < 1e-5


def test_fflmix_torch_terms_match_reference_golden():
    """The PyTorch-op terms of "fflMix" (VGG perceptual with the P3 seeded weights, Laplacian gradient,
    MS-SSIM) reproduce the reference LossFunction("fflMix") components (losses.py:108-117)."""
    from uwr import fflmix
    g = _load("losses_metrics.pt")["fflMix"]          # [loss, charb, perc, grad, ffl, 1 - ms_ssim]
    torch.manual_seed(0)
    p = torch.rand(2, 3, 256, 256)
    t = torch.rand(2, 3, 256, 256)
    close = lambda a, b: abs(a - b) <= 1e-5 * abs(b) + 1e-8
    assert close(fflmix.VGGPerceptual()(p, t).item(), g[2])
    assert close(fflmix.gradient_loss(p, t).item(), g[3])
    assert close(1 - fflmix.ms_ssim(p, t).item(), g[5])
    assert close(0.03 * g[1] + 0.025 * g[2] + 0.01 * g[3] + 0.005 * g[4] + 0.1 * g[5], g[0])


def test_wavelet_oracle_matches_reference_modules_golden_free():
    """oracle/wavelet_oracle.py (closed forms of DWT_2D / IDWT_2D forward AND the reference's hand-written backward,
    wave_modules.py:9-181) is self-consistent on the identities that hold by construction: every sub-band of the DWT is
    replicated over C/4 channels, the DWT 'gradient' is channel-independent, IDWT maps groups of 4 channels."""
    from oracle import wavelet_oracle as wo
    g = torch.Generator().manual_seed(3)
    B, C, h, w = 2, 16, 8, 12
    x = torch.randn(B, 4 * h * w, C, generator=g, dtype=torch.float64)
    y = wo.dwt_fwd(x, B, h, w).view(B, h * w, 4, C // 4)
    assert torch.equal(y, y[..., :1].expand_as(y))
    # ll sub-band = half the sum of the 2x2 block of the channel-summed image
    S = x.view(B, h, 2, w, 2, C).sum((-1, 2, 4)).reshape(B, h * w)
    assert torch.allclose(y[:, :, 0, 0], 0.5 * S, atol=1e-12)
    d = wo.dwt_bwd(torch.randn(B, h * w, C, generator=g, dtype=torch.float64), B, h, w)
    assert torch.equal(d, d[..., :1].expand_as(d))
    z = wo.idwt_fwd(torch.randn(B, h * w, C, generator=g, dtype=torch.float64), B, h, w)
    assert z.shape == (B, 4 * h * w, C)
    assert wo.idwt_bwd(z, B, h, w).shape == (B, h * w, C)
