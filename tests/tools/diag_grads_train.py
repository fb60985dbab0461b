"""Diagnostic: train-mode (injected DropPath masks) per-tensor gradient error vs fp64 oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "underwater-image-restoration_b200"))
import torch
from oracle import ast_oracle, losses_oracle
from uwr.ast import AST, DropPath
from uwr import ops

S, B = 128, 2
torch.manual_seed(1234)
model = AST(img_size=S)
sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
g = torch.Generator().manual_seed(2024)
raw = torch.rand(B, 3, S, S, generator=g) * 2 - 1
ref = torch.rand(B, 3, S, S, generator=g) * 2 - 1
gm = torch.Generator().manual_seed(7)
drop = {}
for name, mod in model.named_modules():
    if isinstance(mod, DropPath):
        keep = 1.0 - mod.drop_prob
        pre = name[: -len("drop_path")]
        ma = torch.bernoulli(torch.full((B,), keep), generator=gm) / keep
        mm = torch.bernoulli(torch.full((B,), keep), generator=gm) / keep
        drop[pre] = (ma, mm)
print({k: (v[0].tolist(), v[1].tolist()) for k, v in drop.items() if (v[0].min() == 0 or v[1].min() == 0)})
s64 = {k: (v.double().clone().requires_grad_() if v.is_floating_point() else v) for k, v in sd.items()}
dsc = {k: ((a.double() if (k + "attn.w") in sd else None), m.double()) for k, (a, m) in drop.items()}
o64 = ast_oracle.ast_forward(s64, raw.double(), img_size=S, drop_scales=dsc)
losses_oracle.l1(o64, ref.double()).backward()
gn = torch.sqrt(sum((v.grad ** 2).sum() for v in s64.values() if v.is_floating_point())).item()
model = model.cuda().train()
ops.set_gemm_precision("tf32x3")
for name, mod in model.named_modules():
    if isinstance(mod, DropPath):
        pre = name[: -len("drop_path")]
        mod._seq = list(drop[pre]) if (pre + "attn.w") in sd else [drop[pre][1]]
        mod.scale = (lambda batch, device, _m=mod: _m._seq.pop(0).to(device).float().contiguous())
out = model(raw.cuda())
loss = (out - ref.cuda()).abs().mean() / (B * 3)
loss.backward()
rows = []
for n, p in model.named_parameters():
    go = s64[n].grad
    d = (p.grad.double().cpu() - go).norm().item()
    rows.append((d / gn, d / max(go.norm().item(), 1e-30), go.norm().item() / gn, n))
rows.sort(reverse=True)
print(f"train tf32x3: out err {((out.double().cpu()-o64).norm()/o64.norm()).item():.2e} global grad err {sum(r[0]**2 for r in rows)**0.5:.2e}")
for r in rows[:25]:
    print("   contrib %.2e rel %.2e normshare %.2e %s" % r)
