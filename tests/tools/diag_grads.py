"""Diagnostic: per-tensor gradient error of the uwr AST vs the fp64 CPU oracle (top contributors)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "underwater-image-restoration_b200"))
import torch
from oracle import ast_oracle, losses_oracle
from uwr.ast import AST
from uwr import ops

S = int(sys.argv[1]) if len(sys.argv) > 1 else 128
B = 2
torch.manual_seed(1234)
model = AST(img_size=S)
sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
g = torch.Generator().manual_seed(2024)
raw = torch.rand(B, 3, S, S, generator=g) * 2 - 1
ref = torch.rand(B, 3, S, S, generator=g) * 2 - 1
s64 = {k: (v.double().clone().requires_grad_() if v.is_floating_point() else v) for k, v in sd.items()}
o64 = ast_oracle.ast_forward(s64, raw.double(), img_size=S)
losses_oracle.l1(o64, ref.double()).backward()
gn = torch.sqrt(sum((v.grad ** 2).sum() for v in s64.values() if v.is_floating_point())).item()
model = model.cuda().eval()
for mode in ("tf32", "tf32x3"):
    ops.set_gemm_precision(mode)
    model.zero_grad(set_to_none=True)
    out = model(raw.cuda())
    loss = (out - ref.cuda()).abs().mean() / (B * 3)
    loss.backward()
    rows = []
    for n, p in model.named_parameters():
        go = s64[n].grad
        d = (p.grad.double().cpu() - go).norm().item()
        rows.append((d / gn, d / max(go.norm().item(), 1e-30), go.norm().item() / gn, n))
    rows.sort(reverse=True)
    print(f"== {mode}: out err {((out.double().cpu()-o64).norm()/o64.norm()).item():.2e} global grad err {sum(r[0]**2 for r in rows)**0.5:.2e}")
    for r in rows[:14]:
        print("   contrib %.2e rel %.2e normshare %.2e %s" % r)
