import sys
sys.path.insert(0, '/root/repo/underwater-image-restoration_b200')
import torch
from uwr.optim import FusedClipAdam
torch.manual_seed(0)
ps = [torch.nn.Parameter(torch.randn(64, 32).cuda())]
qs = [torch.nn.Parameter(p.detach().clone()) for p in ps]
opt = FusedClipAdam(ps, lr=1e-3, max_norm=1.0); ref = torch.optim.Adam(qs, lr=1e-3)
for it in range(3):
    g = torch.randn_like(ps[0]); ps[0].grad = g.clone(); qs[0].grad = g.clone()
    opt.step(); torch.nn.utils.clip_grad_norm_(qs, 1.0); ref.step()
print("live diff", (ps[0]-qs[0]).abs().max().item())
sd = opt.state_dict(); sdr = ref.state_dict()
print("step", sd["state"][0]["step"], sdr["state"][0]["step"])
print("m diff", (sd["state"][0]["exp_avg"]-sdr["state"][0]["exp_avg"]).abs().max().item(), "v diff", (sd["state"][0]["exp_avg_sq"]-sdr["state"][0]["exp_avg_sq"]).abs().max().item())
print({k:v for k,v in sd["param_groups"][0].items() if k!="params"})
g = torch.randn_like(ps[0])
for name, src in (("mine->mine", sd), ("mine->torch", sd), ("torch->mine", sdr), ("torch->torch", sdr)):
    p2 = torch.nn.Parameter(ps[0].detach().clone())
    o = (FusedClipAdam([p2], lr=5.0, max_norm=1.0) if name.endswith("mine") else torch.optim.Adam([p2], lr=5.0))
    o.load_state_dict(src)
    p2.grad = g.clone()
    if not name.endswith("mine"): torch.nn.utils.clip_grad_norm_([p2], 1.0)
    o.step()
    print(name, "delta norm", (p2.detach()-ps[0].detach()).norm().item(), "step_count", getattr(o, "step_count", None))
# manual fp64 Adam step from the saved state
m = sd["state"][0]["exp_avg"].double(); v = sd["state"][0]["exp_avg_sq"].double()
gn = g.double().norm(); gc = g.double() * min(1.0, 1.0 / (gn.item() + 1e-6))
m2 = 0.9 * m + 0.1 * gc; v2 = 0.999 * v + 0.001 * gc * gc
for st in (3, 4, 5):
    upd = 1e-3 / (1 - 0.9 ** st) * m2 / ((v2 / (1 - 0.999 ** st)).sqrt() + 1e-8)
    print("manual step", st, "delta norm", upd.norm().item())
# a 4th live step on both
g4 = g
ps[0].grad = g4.clone(); qs[0].grad = g4.clone()
b0, c0 = ps[0].detach().clone(), qs[0].detach().clone()
opt.step(); torch.nn.utils.clip_grad_norm_(qs, 1.0); ref.step()
print("live 4th: mine", (ps[0].detach() - b0).norm().item(), "torch", (qs[0].detach() - c0).norm().item())
