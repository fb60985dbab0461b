"""Generate tests/golden/*.pt from the UNMODIFIED reference (TEST INFRASTRUCTURE).

Run in the authoring container only (needs /root/reference; the stand-ins under oracle/shims/ supply
the third-party packages that are absent from the image):

    python oracle/make_golden.py

The vectors pin the oracle (oracle/*.py) and, through it, the CUDA path.  Documented reference
patches used here: P2 = LuminanceLoss attached for "L1withColor".
"""
import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path[:0] = [os.path.join(ROOT, "oracle", "shims"), REF]
OUT = os.path.join(ROOT, "tests", "golden")


def sha(t):
    return hashlib.sha1(t.detach().cpu().contiguous().numpy().tobytes()).hexdigest()


def main():
    os.makedirs(OUT, exist_ok=True)
    from src.Models.AST import AST, TransformerBlock, WindowAttention_sparse
    from src.Losses.losses import LossFunction
    from src.Losses.luminanceLoss import LuminanceLoss
    from src.ModelTrainer import torchPSNR
    import uqim_utils

    # ---- AST: seeded weights, seeded input, output + loss + per-tensor gradient norms/projections
    torch.manual_seed(1234)
    model = AST(img_size=128)
    keys = [(k, list(v.shape), str(v.dtype), sha(v)) for k, v in model.state_dict().items()]
    model.eval()
    g = torch.Generator().manual_seed(2024)
    raw = torch.rand(1, 3, 128, 128, generator=g) * 2 - 1
    ref = torch.rand(1, 3, 128, 128, generator=g) * 2 - 1
    out = model(raw)
    lf = LossFunction("L1", "cpu")
    loss = lf.getloss(out, ref)
    loss.backward()
    pg = torch.Generator().manual_seed(99)
    grads = {}
    for n, p in model.named_parameters():
        r = torch.randn(p.shape, generator=pg)
        grads[n] = (p.grad.norm().item(), (p.grad * r).sum().item())
    torch.save({"state_dict_sha1": keys, "out": out.detach(), "loss_l1": loss.item(), "grad_norm_proj": grads,
                "seed_weights": 1234, "seed_data": 2024, "proj_seed": 99}, os.path.join(OUT, "ast_128.pt"))

    # full-size default model: only the hashes (weights are regenerated from the seed on the GPU box)
    torch.manual_seed(1234)
    m256 = AST()
    torch.save({"state_dict_sha1": [(k, list(v.shape), str(v.dtype), sha(v)) for k, v in m256.state_dict().items()]},
               os.path.join(OUT, "ast_256_state_sha1.pt"))

    # ---- building blocks: one shifted sparse-attention block and its mask / index buffers
    torch.manual_seed(5)
    blk = TransformerBlock(64, (16, 16), 2, win_size=8, shift_size=4, att=True, sparseAtt=True)
    for p in blk.parameters():
        torch.nn.init.normal_(p, std=0.1)
    blk.eval()
    x = torch.randn(2, 256, 64)
    y = blk(x)
    torch.save({"state": {k: v.clone() for k, v in blk.state_dict().items()}, "x": x, "y": y.detach(),
                "rel_index": blk.attn.relative_position_index.clone()}, os.path.join(OUT, "block_shift.pt"))

    # ---- NewBigFRFNModel (patch P1: token->NCHW transpose before output_proj, model.py:435-437 vs 637)
    from src.model.model import MyBigFRFNModel
    torch.manual_seed(1234)
    nb = MyBigFRFNModel()
    nkeys = [(k, list(v.shape), str(v.dtype), sha(v)) for k, v in nb.state_dict().items()]
    orig = nb.output_proj.forward

    def patched(tk):
        b, l, c = tk.shape
        h = int(l ** 0.5)
        return orig(tk.transpose(1, 2).reshape(b, c, h, h).contiguous())
    nb.output_proj.forward = patched
    nb.eval()
    gg = torch.Generator().manual_seed(2024)
    xr = torch.rand(1, 3, 128, 128, generator=gg) * 2 - 1
    import io, contextlib
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
        yr = nb(xr)
    torch.save({"state_dict_sha1": nkeys, "out": yr, "seed_weights": 1234, "seed_data": 2024},
               os.path.join(OUT, "newbigfrfn_128.pt"))

    # ---- NewBigFRFNModel in Haar-wavelet mode (use_dwt="Wavelet"; unreachable from the CLI):
    # output + loss gradients of the reference itself, INCLUDING its hand-written (non-adjoint) DWT / IDWT backward
    torch.manual_seed(1234)
    nw = MyBigFRFNModel(use_dwt="Wavelet")
    origw = nw.output_proj.forward
    nw.output_proj.forward = lambda tk: origw(tk.transpose(1, 2).reshape(tk.shape[0], tk.shape[2], 128, 128).contiguous())
    nw.eval()
    gw = torch.Generator().manual_seed(2024)
    xw = torch.rand(1, 3, 128, 128, generator=gw) * 2 - 1
    tw = torch.rand(1, 3, 128, 128, generator=gw) * 2 - 1
    with contextlib.redirect_stdout(io.StringIO()):
        yw = nw(xw)
    ((yw - tw) ** 2).mean().backward()
    pgw = torch.Generator().manual_seed(99)
    gradsw = {}
    for n, p in nw.named_parameters():
        r = torch.randn(p.shape, generator=pgw)
        if p.grad is not None:
            gradsw[n] = (p.grad.norm().item(), (p.grad * r).sum().item())
    torch.save({"out": yw.detach(), "grad_norm_proj": gradsw, "seed_weights": 1234, "seed_data": 2024, "proj_seed": 99},
               os.path.join(OUT, "newbigfrfn_wavelet_128.pt"))

    # ---- NewModel / NewBigModel: constructible, forward raises ("keep them registry-constructible and
    # failing identically") -> seeded state_dict hashes + the exception each forward raises
    from src.model.model import MyBigModel, MyModel
    broken = {}
    for nm, cls in (("NewModel", MyModel), ("NewBigModel", MyBigModel)):
        torch.manual_seed(1234)
        m = cls()
        ent = {"state_dict_sha1": [(k, list(v.shape), str(v.dtype), sha(v)) for k, v in m.state_dict().items()]}
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                m(torch.zeros(1, 3, 128, 128))
        except Exception as e:  # noqa: BLE001
            ent["error"] = (type(e).__name__, str(e))
        broken[nm] = ent
    torch.save(broken, os.path.join(OUT, "newmodels_state_sha1.pt"))

    # ---- SpectralTransformer (runs as shipped)
    from src.Models.SpectralTransformer import SpectralTransformer
    torch.manual_seed(1234)
    sp = SpectralTransformer()
    sp.eval()
    skeys = [(k, list(v.shape), str(v.dtype), sha(v)) for k, v in sp.state_dict().items()]
    gs = torch.Generator().manual_seed(2024)
    xs = torch.rand(1, 3, 64, 96, generator=gs) * 2 - 1          # non-square is legal (H, W % 8 == 0)
    with torch.no_grad():
        ys = sp(xs)
    torch.save({"state_dict_sha1": skeys, "out": ys, "seed_weights": 1234, "seed_data": 2024},
               os.path.join(OUT, "spectral_64x96.pt"))

    # ---- losses / metrics known answers
    torch.manual_seed(0)
    p = torch.rand(2, 3, 256, 256)
    t = torch.rand(2, 3, 256, 256)
    vals = {}
    for name in ("L1", "L2", "charbonnier", "fflCharbonnier"):
        vals[name] = LossFunction(name, "cpu").getloss(p, t).item()
    lfc = LossFunction("L1withColor", "cpu")
    lfc.luminanceLoss = LuminanceLoss()  # patch P2
    import io, contextlib
    with contextlib.redirect_stdout(io.StringIO()):
        vals["L1withColor"] = lfc.getloss(p, t).item()
    vals["Luminance"] = LuminanceLoss()(p, t).item()
    vals["charbonnier_identical"] = LossFunction("charbonnier", "cpu").getloss(p, p).item()  # Loss.ipynb:45
    vals["psnr"] = torchPSNR(t, p).item()
    # "fflMix" with patch P3: pretrained VGG16 weights are not downloadable offline -> the reference's four
    # vgg16(pretrained=True) calls are redirected to a re-seeded random-init vgg16 (identical each call)
    import torchvision
    orig_vgg = torchvision.models.vgg16

    def seeded_vgg(*a, **k):
        st = torch.random.get_rng_state()
        torch.manual_seed(777)
        m = orig_vgg(weights=None)
        torch.random.set_rng_state(st)
        return m
    torchvision.models.vgg16 = seeded_vgg
    try:
        mix = LossFunction("fflMix", "cpu").getloss(p, t)
    finally:
        torchvision.models.vgg16 = orig_vgg
    vals["fflMix"] = [v.item() for v in mix]
    img = (np.random.default_rng(0).random((256, 256, 3)) * 255).astype(np.uint8)
    vals["uiqm"] = [float(v) for v in uqim_utils.getUIQM(img)]
    torch.save(vals, os.path.join(OUT, "losses_metrics.pt"))
    print({k: v for k, v in vals.items()})


if __name__ == "__main__":
    main()
