"""Experiment: AST parity with the attention products single-pass TF32 vs error-compensated 3xTF32
(UWR_ATTN_X3=0/1, read once per process).  argv: size qk_scale — qk_scale multiplies to_q / to_kv weights so that
the scores leave the near-zero regime of a freshly initialised model."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "underwater-image-restoration_b200"))
import torch
from oracle import ast_oracle, losses_oracle
from uwr.ast import AST

S = int(sys.argv[1]) if len(sys.argv) > 1 else 128
QK = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
B = 2
torch.manual_seed(1234)
model = AST(img_size=S)
with torch.no_grad():
    for n, p in model.named_parameters():
        if "to_q.weight" in n or "to_kv.weight" in n:
            p.mul_(QK)
sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
g = torch.Generator().manual_seed(2024)
raw = torch.rand(B, 3, S, S, generator=g) * 2 - 1
ref = torch.rand(B, 3, S, S, generator=g) * 2 - 1
s64 = {k: (v.double().clone().requires_grad_() if v.is_floating_point() else v) for k, v in sd.items()}
o64 = ast_oracle.ast_forward(s64, raw.double(), img_size=S)
loss = losses_oracle.l1(o64, ref.double())
(cot,) = torch.autograd.grad(loss, o64, retain_graph=True)
o64.backward(cot)
gn = torch.sqrt(sum((v.grad ** 2).sum() for v in s64.values() if v.is_floating_point())).item()
model = model.cuda().eval()
out = model(raw.cuda())
out.backward(cot.float().cuda())
rows = []
for n, p in model.named_parameters():
    go = s64[n].grad
    d = (p.grad.double().cpu() - go).norm().item()
    rows.append((d / gn, d / max(go.norm().item(), 1e-30), go.norm().item() / gn, n))
rows.sort(reverse=True)
print(f"== X3={os.environ.get('UWR_ATTN_X3', '1')} S={S} qk x{QK}: out err {((out.double().cpu()-o64).norm()/o64.norm()).item():.2e} "
      f"global grad err {sum(r[0]**2 for r in rows)**0.5:.2e}")
for r in rows[:6]:
    print("   contrib %.2e rel %.2e normshare %.2e %s" % r)
