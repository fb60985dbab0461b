"""BASELINE config 5: AST inference sweep (eval + no_grad), batch 1..64 at 256^2 and 512^2 on one
B200, eager launches and CUDA-graph replay, plus PSNR / UIQM parity against the CPU oracle on the
structured synthetic pair of SURVEY.md §8d.  Writes a JSON summary (profiles/r1_inference_sweep.json)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "underwater-image-restoration_b200"))
import numpy as np
import torch
import torch.nn.functional as F
import uwr
from uwr.graph import GraphedForward
from oracle import ast_oracle, losses_oracle, uiqm_oracle


def structured_pair(B, S):
    g7, g8 = torch.Generator().manual_seed(7), torch.Generator().manual_seed(8)
    ref01 = F.interpolate(torch.rand(B, 3, S // 16, S // 16, generator=g7), scale_factor=16, mode="bilinear")
    att = torch.tensor([0.35, 0.80, 0.90]).view(1, 3, 1, 1)
    raw01 = (ref01 * att + 0.10 + 0.02 * torch.randn(B, 3, S, S, generator=g8)).clamp(0, 1)
    return (raw01 - 0.5) / 0.5, (ref01 - 0.5) / 0.5


def timed(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


out = {"sweep": [], "parity": {}}
torch.manual_seed(1234)
model = uwr.AST(img_size=256)
sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
model = model.cuda().eval()

# ---- parity (PSNR within 0.01 dB, UIQM within 0.01) ----
raw, ref = structured_pair(1, 256)
with torch.no_grad():
    y = model(raw.cuda()).cpu()
    yo = ast_oracle.ast_forward(sd, raw, img_size=256)
psnr, psnr_o = losses_oracle.torch_psnr(ref, y).item(), losses_oracle.torch_psnr(ref, yo).item()
to_img = lambda t: (np.clip(t[0].permute(1, 2, 0).numpy(), 0, 1) * 255).astype(np.uint8)   # Visualiser.py:53-54
uiqm, uiqm_o = uiqm_oracle.get_uiqm(to_img(y))[0], uiqm_oracle.get_uiqm(to_img(yo))[0]
out["parity"] = {"psnr_uwr": psnr, "psnr_oracle": psnr_o, "uiqm_uwr": float(uiqm), "uiqm_oracle": float(uiqm_o),
                 "out_rel_l2": ((y - yo).norm() / yo.norm()).item()}
print("parity", out["parity"], flush=True)
assert abs(psnr - psnr_o) < 0.01 and abs(uiqm - uiqm_o) < 0.01

# ---- throughput sweep ----
for S in (256, 512):
    for B in (1, 2, 4, 8, 16, 32, 64):
        if S == 512 and B > 16:
            continue
        x = (torch.rand(B, 3, S, S) * 2 - 1).cuda()
        with torch.no_grad():
            ms_eager = timed(lambda: model(x), 10 if B <= 8 else 4)
            gf = GraphedForward(model, x)
            ms_graph = timed(lambda: gf(x), 10 if B <= 8 else 4)
            del gf
        row = {"size": S, "batch": B, "ms_per_image_eager": ms_eager / B, "ms_per_image_graph": ms_graph / B,
               "images_per_s_graph": 1000.0 * B / ms_graph}
        out["sweep"].append(row)
        print(row, flush=True)
        torch.cuda.empty_cache()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r1_inference_sweep.json"), "w"), indent=1)
