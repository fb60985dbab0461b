"""Which parameter tensors carry the whole-model gradient error of NewBigFRFN (vs the CPU oracle, same cotangent)?
Prints, per tensor, its share of the squared error and of the squared gradient norm.  usage: python tools/grad_error_by_tensor.py"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'underwater-image-restoration_b200'))
import torch
from oracle import newbig_oracle
from uwr.newbig import MyBigFRFNModel
from uwr.ast import DropPath
B, S = 2, 128
torch.manual_seed(1234)
model = MyBigFRFNModel()
sd_cpu = {k: v.detach().clone() for k, v in model.state_dict().items()}
model = model.cuda().eval()
g = torch.Generator().manual_seed(2024)
raw = torch.rand(B, 3, S, S, generator=g) * 2 - 1
cot = torch.randn(B, 3, S, S, generator=g) * 1e-3
out = model(raw.cuda()); out.backward(cot.cuda())
sd_o = {k: (v.clone().requires_grad_() if v.is_floating_point() and v.dim() > 0 and "dwt" not in k else v) for k, v in sd_cpu.items()}
out_o = newbig_oracle.newbig_frfn_forward(sd_o, raw, drop_scales={}); out_o.backward(cot)
named = dict(model.named_parameters())
live = [n for n, v in sd_o.items() if getattr(v, "grad", None) is not None]
gn2 = sum((sd_o[n].grad.double() ** 2).sum().item() for n in live)
rows = []
for n in live:
    d2 = ((named[n].grad.double().cpu() - sd_o[n].grad.double()) ** 2).sum().item()
    rows.append((d2 / gn2, (sd_o[n].grad.double() ** 2).sum().item() / gn2, n))
rows.sort(reverse=True)
print("global rel err", (sum(r[0] for r in rows)) ** 0.5)
for c, share, n in rows[:25]:
    print(f"{n:60s} err^2 share {c / sum(r[0] for r in rows):6.3f}  grad-norm^2 share {share:6.3f}  rel {(c / max(share, 1e-30)) ** 0.5:.2e}")
