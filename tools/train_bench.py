"""Training-step throughput of any registry architecture through uwr.train.TrainStep
(BASELINE configs 2 and 3: SpectralTransformer B=8 "L1withColor"; NewBigFRFNModel B=16).
usage: python tools/train_bench.py ARCH LOSS BATCH [SIZE] [STEPS]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "underwater-image-restoration_b200"))
import torch
import uwr
from uwr import ops
from uwr.train import TrainStep

arch, loss, B = sys.argv[1], sys.argv[2], int(sys.argv[3])
S = int(sys.argv[4]) if len(sys.argv) > 4 else 256
steps = int(sys.argv[5]) if len(sys.argv) > 5 else 5
torch.manual_seed(1234)
model = uwr.init_model(arch).cuda().train()
step = TrainStep(model, loss, lr=1e-3, local_batch=B, vgg_weights="random" if loss == "fflMix" else None)  # P3: synthetic bench
g = torch.Generator().manual_seed(2024)
raw = (torch.rand(B, 3, S, S, generator=g) * 2 - 1).cuda()
ref = (torch.rand(B, 3, S, S, generator=g) * 2 - 1).cuda()
first = None
for _ in range(3):
    l, n = step(raw, ref)
    if first is None:   # the seeded first step: comparable across kernel versions (later steps follow a chaotic trajectory at
        first = (l.item(), n[0].item())   # lr 1e-3 on random targets, their loss / norm are only a liveness check)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n0 = ops.launch_count()
e0.record()
for _ in range(steps):
    l, n = step(raw, ref)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
ms_graph = None
try:
    from uwr.graph import GraphedTrainStep
    gs = GraphedTrainStep(step, raw, ref, warmup=1)
    gs(raw, ref)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        gs(raw, ref)
    e1.record()
    torch.cuda.synchronize()
    ms_graph = e0.elapsed_time(e1) / steps
    del gs
except Exception as ex:  # report, do not hide
    print("graph capture failed:", repr(ex)[:300])
with ops.KernelProfile() as prof:
    step(raw, ref)
kernel_table = prof.table()
fam = {}
for r in kernel_table:
    f = fam.setdefault(r["kernel"], [0.0, 0])
    f[0] += r["ms_total"]; f[1] += r["launches"]
res = {"arch": arch, "loss": loss, "batch": B, "size": S, "ms_per_step": ms, "images_per_s": 1000.0 * B / ms,
       "ms_per_step_graph": ms_graph, "images_per_s_graph": (1000.0 * B / ms_graph) if ms_graph else None,
       "uwr_launches_per_step": (ops.launch_count() - n0) // (steps + 1), "loss_first_step": first[0], "grad_norm_first_step": first[1], "loss_value": l.item(), "grad_norm": n[0].item(),
       "max_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30,
       "uwr_kernel_ms": {k: round(v[0], 3) for k, v in sorted(fam.items(), key=lambda kv: -kv[1][0])[:8]},
       "uwr_kernel_ms_total": round(sum(v[0] for v in fam.values()), 2)}
if os.environ.get("UWR_EAGER_REF"):
    # the kernel bar for this config: the reference modules' math (oracle port: plain torch ops ->
    # cuBLAS / cuDNN / cuFFT / ATen) with torch autograd, clip_grad_norm_ and torch.optim.Adam on the SAME GPU, fp32
    from oracle import losses_oracle, newbig_oracle, spectral_oracle
    from uwr import fflmix
    del step, model
    torch.cuda.empty_cache()
    torch.manual_seed(1234)
    sd0 = uwr.init_model(arch).state_dict()
    fwd = {"SpectralTransformer": spectral_oracle.spectral_forward, "NewBigFRFNModel": newbig_oracle.newbig_frfn_forward}[arch]

    def eager_loss(out, tgt):
        if loss == "L1withColor":
            return losses_oracle.l1_with_color(out, tgt)
        if loss == "L1":
            return losses_oracle.l1(out, tgt)
        if loss == "fflCharbonnier":
            return losses_oracle.focal_frequency(out, tgt) + losses_oracle.charbonnier(out, tgt)
        if loss == "fflMix":       # losses.py:108-117 with torch ops (VGG on cuDNN, patch P3 weights)
            vgg = fflmix.load_vgg_weights("random", out.device)
            return (0.03 * losses_oracle.charbonnier(out, tgt) + 0.025 * vgg(out, tgt) + 0.01 * fflmix.gradient_loss(out, tgt)
                    + 0.005 * losses_oracle.focal_frequency(out, tgt) + 0.1 * (1 - fflmix.ms_ssim(out, tgt)))
        raise ValueError(loss)

    res["eager"] = {}
    for tag, tf32 in (("fp32", False), ("tf32", True)):
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.allow_tf32 = tf32
        sd = {k: v.detach().clone().cuda() for k, v in sd0.items()}
        params = {k: v.requires_grad_() for k, v in sd.items() if v.is_floating_point() and v.dim() > 0 and "dwt" not in k}
        plist = list(params.values())
        opt = torch.optim.Adam(plist, lr=1e-3)

        def one():
            opt.zero_grad(set_to_none=True)
            eager_loss(fwd(sd, raw), ref).backward()
            torch.nn.utils.clip_grad_norm_([p for p in plist if p.grad is not None], 1.0)
            opt.step()
        for _ in range(3):
            one()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(steps):
            one()
        e1.record()
        torch.cuda.synchronize()
        ems = e0.elapsed_time(e1) / steps
        res["eager"][tag] = {"ms_per_step": ems, "images_per_s": 1000.0 * B / ems,
                             "speedup_eager_step": ems / ms,
                             "speedup_graphed_step": (ems / ms_graph) if ms_graph else None}
        del opt, plist, params, sd
        torch.cuda.empty_cache()
print(json.dumps(res))
if os.environ.get("UWR_PROFILE_OUT"):
    json.dump({"result": res, "kernels": kernel_table}, open(os.environ["UWR_PROFILE_OUT"], "w"), indent=1)
if os.environ.get("UWR_TORCHPROF") and not os.environ.get("UWR_EAGER_REF"):
    # every CUDA kernel of one step (ours + ATen/cuDNN/cuFFT), to see what is still library code
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as tp:
        step(raw, ref)
        torch.cuda.synchronize()
    agg = {}
    for ev in tp.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            a = agg.setdefault(ev.name[:90], [0.0, 0])
            a[0] += ev.device_time / 1e3; a[1] += 1
    tot = sum(v[0] for v in agg.values())
    print(f"all CUDA kernels of one step: {tot:.2f} ms")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(os.environ["UWR_TORCHPROF"])]:
        print(f"  {v[0]:8.3f} ms {v[1]:5d}x  {k}")
