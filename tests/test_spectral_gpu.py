"""SpectralTransformer (BASELINE config 2's architecture, with the "L1withColor" loss) on the uwr kernels
vs the CPU oracle."""
import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,H,W", [(2, 64, 64), (1, 64, 96), (2, 256, 256)])   # last: BASELINE config 2's resolution
def test_spectral_vs_oracle(B, H, W):
    from oracle import spectral_oracle, losses_oracle
    from uwr.spectral import SpectralTransformer
    from uwr.losses import LossFunction
    torch.manual_seed(1234)
    model = SpectralTransformer()
    sd_cpu = {k: v.detach().clone() for k, v in model.state_dict().items()}
    model = model.cuda().train()
    g = torch.Generator().manual_seed(2024)
    raw = torch.rand(B, 3, H, W, generator=g) * 2 - 1
    ref = torch.rand(B, 3, H, W, generator=g) * 2 - 1
    sd_o = {k: v.clone().requires_grad_() for k, v in sd_cpu.items()}
    out_o = spectral_oracle.spectral_forward(sd_o, raw)
    loss_o = losses_oracle.l1_with_color(out_o, ref)
    (cot,) = torch.autograd.grad(loss_o, out_o, retain_graph=True)
    out_o.backward(cot)
    out = model(raw.cuda())
    loss = LossFunction("L1withColor", "cuda").getloss(out.detach(), ref.cuda())
    out.backward(cot.cuda())
    assert abs(loss.item() - loss_o.item()) < 1e-4 * abs(loss_o.item())
    e_out = rel_l2(out, out_o)
    named = dict(model.named_parameters())
    live = [n for n, v in sd_o.items() if v.grad is not None]
    gnorm = torch.sqrt(sum((sd_o[n].grad.double() ** 2).sum() for n in live)).item()
    tot, worst = 0.0, (0.0, "")
    for n in live:
        assert named[n].grad is not None, n
        d = (named[n].grad.double().cpu() - sd_o[n].grad.double()).norm().item()
        tot += d * d
        worst = max(worst, (d / max(sd_o[n].grad.norm().item(), 1e-3 * gnorm), n))
    dead = [n for n in named if n not in live]
    print(f"Spectral parity B={B} {H}x{W}: out {e_out:.2e} grads {tot ** 0.5 / gnorm:.2e} worst {worst[1]} "
          f"{worst[0]:.2e}; dead params {sum(named[n].numel() for n in dead)}")
    assert e_out < 1e-3
    assert tot ** 0.5 / gnorm < 1e-3          # north star: gradients within 1e-3 relative
    assert sum(named[n].numel() for n in dead) == 230233
    for n in dead:
        assert named[n].grad is None, n
