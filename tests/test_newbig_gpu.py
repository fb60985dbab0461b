"""NewBigFRFNModel (BASELINE config 3's architecture) on the uwr kernels vs the CPU oracle (patch P1),
same-cotangent protocol, eval and train mode (injected encoder DropPath masks)."""
import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu


# Tolerance.  North star: outputs and gradients within 1e-3 relative.  Outputs meet it in every mode.  Gradients: this
# U-Net chains ~23 NON-residual projections (mlp_proj of every DecoderBlock, the Down/Upsample and in/out convolutions,
# model.py:96-160, block.py:42-153), each adding ~4e-4 of independent TF32 operand-rounding noise to the whole
# gradient stream, so single-pass TF32 lands at 0.9-1.4e-3 (the bf16 storage the north star allows gives 6e-3 on the
# reference itself).  With error-compensated products (`tf32x3`) the same kernels are at ~1e-5, i.e.
# the difference is operand rounding, not algorithm: the 1e-3 bound is asserted there, 2e-3 in the default mode.
@pytest.mark.parametrize("train,S,precision", [(False, 128, "tf32"), (True, 128, "tf32"), (True, 256, "tf32"),
                                               (True, 128, "tf32x3")])
def test_newbigfrfn_vs_oracle(train, S, precision):
    """128x128 (the smallest legal input) and BASELINE config 3's own resolution, 256x256 (B = 2)."""
    from uwr import ops
    ops.set_gemm_precision(precision)
    try:
        _newbig_parity(train, S, 1e-3 if precision == "tf32x3" else 2e-3)
    finally:
        ops.set_gemm_precision("tf32")


def _newbig_parity(train, S, grad_tol):
    from oracle import newbig_oracle
    from uwr.ast import DropPath
    from uwr.newbig import MyBigFRFNModel
    B = 1 if not train else 2
    torch.manual_seed(1234)
    model = MyBigFRFNModel()
    sd_cpu = {k: v.detach().clone() for k, v in model.state_dict().items()}
    model = model.cuda()
    g = torch.Generator().manual_seed(2024)
    raw = torch.rand(B, 3, S, S, generator=g) * 2 - 1
    cot = torch.randn(B, 3, S, S, generator=g) * 1e-3
    drop = {}
    if train:
        model.train()
        gm = torch.Generator().manual_seed(7)
        for name, mod in model.named_modules():
            if isinstance(mod, DropPath) and name.endswith("drop_path2"):
                pre = name[: -len("drop_path2")]
                keep = 1.0 - mod.drop_prob
                mf = torch.bernoulli(torch.full((B,), keep), generator=gm) / keep
                ms = torch.bernoulli(torch.full((B,), keep), generator=gm) / keep
                drop[pre] = (mf, ms)
                blk = dict(model.named_modules())[pre[:-1]]
                blk.drop_path2.scale = (lambda batch, device, _m=mf: _m.to(device).float())
                blk.drop_path.scale = (lambda batch, device, _m=ms: _m.to(device).float())
    else:
        model.eval()
    out = model(raw.cuda())
    out.backward(cot.cuda())
    sd_o = {k: (v.clone().requires_grad_() if v.is_floating_point() and v.dim() > 0 and "dwt" not in k else v)
            for k, v in sd_cpu.items()}
    out_o = newbig_oracle.newbig_frfn_forward(sd_o, raw, drop_scales=drop)
    out_o.backward(cot)
    e_out = rel_l2(out, out_o)
    e_res = rel_l2(out - raw.cuda(), out_o - raw)
    named = dict(model.named_parameters())
    live = [n for n, v in sd_o.items() if getattr(v, "grad", None) is not None]
    gnorm = torch.sqrt(sum((sd_o[n].grad.double() ** 2).sum() for n in live)).item()
    tot, worst = 0.0, (0.0, "")
    for n in live:
        assert named[n].grad is not None, n
        d = (named[n].grad.double().cpu() - sd_o[n].grad.double()).norm().item()
        tot += d * d
        r = d / max(sd_o[n].grad.norm().item(), 1e-3 * gnorm)
        worst = max(worst, (r, n))
    dead = [n for n in named if n not in live]
    print(f"NewBigFRFN parity train={train} S={S}: out {e_out:.2e} residual {e_res:.2e} grads {tot ** 0.5 / gnorm:.2e} "
          f"worst {worst[1]} {worst[0]:.2e}; {len(dead)} dead tensors")
    assert e_out < 1e-3 and e_res < 3e-3      # e_res: the network's own contribution out - x (diagnostic, 2.2e-3 at 256)
    assert tot ** 0.5 / gnorm < grad_tol
    # dead parameters of the Fourier mode get no gradient, as in the reference
    for n in dead:
        assert named[n].grad is None or float(named[n].grad.abs().max()) == 0.0, n


def test_haar_wavelet_kernels_vs_oracle():
    """csrc/wavelet.cu (DWT / IDWT forward and the reference's own backward formulas) vs oracle/wavelet_oracle.py, which
    is pinned against the reference's DWT_2D / IDWT_2D modules (tests/golden via oracle/make_golden.py)."""
    from oracle import wavelet_oracle as wo
    from uwr import fn
    g = torch.Generator().manual_seed(4)
    for B, C, h, w in ((2, 32, 8, 12), (1, 64, 16, 16), (2, 8, 4, 4)):
        x = torch.randn(B, 4 * h * w, C, generator=g)
        xc = x.cuda().requires_grad_()
        y = fn.HaarDWTFn.apply(xc, B, h, w)
        gy = torch.randn(B, h * w, C, generator=g)
        y.backward(gy.cuda())
        assert rel_l2(y, wo.dwt_fwd(x.double(), B, h, w)) < 1e-6
        assert rel_l2(xc.grad, wo.dwt_bwd(gy.double(), B, h, w)) < 1e-6
        z = torch.randn(B, h * w, C, generator=g)
        zc = z.cuda().requires_grad_()
        o = fn.HaarIDWTFn.apply(zc, B, h, w)
        go = torch.randn(B, 4 * h * w, C, generator=g)
        o.backward(go.cuda())
        assert rel_l2(o, wo.idwt_fwd(z.double(), B, h, w)) < 1e-6
        assert rel_l2(zc.grad, wo.idwt_bwd(go.double(), B, h, w)) < 1e-6


def test_newbigfrfn_wavelet_mode_vs_reference_golden():
    """MyBigFRFNModel(use_dwt="Wavelet") (model.py:64-88, block.py:532-552) on the uwr kernels vs the output and the
    loss gradients of the REFERENCE ITSELF (tests/golden/newbigfrfn_wavelet_128.pt, written by oracle/make_golden.py:
    per-tensor gradient norms and random projections, including the reference's non-adjoint DWT / IDWT backward)."""
    import os
    from conftest import ROOT
    from uwr.newbig import MyBigFRFNModel
    gold = torch.load(os.path.join(ROOT, "tests", "golden", "newbigfrfn_wavelet_128.pt"), weights_only=False)
    torch.manual_seed(gold["seed_weights"])
    model = MyBigFRFNModel(use_dwt="Wavelet").cuda().eval()
    g = torch.Generator().manual_seed(gold["seed_data"])
    x = torch.rand(1, 3, 128, 128, generator=g) * 2 - 1
    t = torch.rand(1, 3, 128, 128, generator=g) * 2 - 1
    y = model(x.cuda())
    ((y - t.cuda()) ** 2).mean().backward()
    e_out = rel_l2(y, gold["out"])
    pg = torch.Generator().manual_seed(gold["proj_seed"])
    num = den = 0.0
    missing = []
    for n, p in model.named_parameters():
        r = torch.randn(p.shape, generator=pg)
        if n not in gold["grad_norm_proj"]:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, n
            continue
        if p.grad is None:
            missing.append(n)
            continue
        norm_ref, proj_ref = gold["grad_norm_proj"][n]
        gcpu = p.grad.detach().cpu()
        num += (gcpu.norm().item() - norm_ref) ** 2 + ((gcpu * r).sum().item() - proj_ref) ** 2
        den += norm_ref ** 2 + proj_ref ** 2
    assert not missing, missing
    print(f"NewBigFRFN wavelet mode vs reference golden: out {e_out:.2e}, gradient norms/projections {(num / den) ** 0.5:.2e}")
    assert e_out < 1e-3
    assert (num / den) ** 0.5 < 5e-3
